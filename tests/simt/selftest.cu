/* simt-check self test: small kernels whose results are known, one per interpreter feature,
 * including the failures it must detect (a read before the wait of an asynchronous copy,
 * a warp whose lanes wait for different collectives). Built and run by
 * tests/test_simt_check.py; exits 0 when every check holds. Test infrastructure only. */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "simt_async.h"

#define FULL 0xffffffffu

extern "C" void simt_check_reset(void);

static __global__ void
k_scan(const int *in, int *out)
{
	const int lane = threadIdx.x & 31;
	int v = in[threadIdx.x];
	for(int o = 1; o < 32; o <<= 1)
	{
		const int t = __shfl_up_sync(FULL, v, o);
		if(lane >= o) v += t;
	}
	out[threadIdx.x] = v;
}

/* stream compaction of the even values, ballots only */
static __global__ void
k_compact(const int *in, int *out, int *count)
{
	const int lane = threadIdx.x & 31;
	const bool keep = (in[lane] & 1) == 0;
	const unsigned m = __ballot_sync(FULL, keep);
	if(keep) out[__popc(m & ((1u << lane) - 1))] = in[lane];
	if(lane == 0) *count = __popc(m);
}

/* lanes that hold the same key rank themselves; only the lanes of a divergent branch take part,
 * the others wait at the full-warp barrier behind it */
static __global__ void
k_match(const int *key, int *rank, int *groups)
{
	const int lane = threadIdx.x & 31;
	const bool in = key[lane] >= 0;
	const unsigned ml = __ballot_sync(FULL, in);
	int r = -1;
	if(in)
	{
		const unsigned peers = __match_any_sync(ml, key[lane]);
		r = __popc(peers & ((1u << lane) - 1));
		__syncwarp(ml);
	}
	__syncwarp();
	rank[lane] = r;
	const int mx = __reduce_max_sync(FULL, r);
	if(lane == 0) *groups = mx;
}

static __global__ void
k_block_sum(const double *in, double *out)
{
	__shared__ double part[256];
	part[threadIdx.x] = in[blockIdx.x * blockDim.x + threadIdx.x];
	__syncthreads();
	for(int s = blockDim.x / 2; s > 0; s >>= 1)
	{
		if((int) threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
		__syncthreads();
	}
	if(threadIdx.x == 0) out[blockIdx.x] = part[0];
}

/* cp.async: the value is not there before the wait (early[] must hold NaNs), it is after */
static __global__ void
k_async(const double *in, double *early, double *late)
{
	extern __shared__ __align__(128) unsigned char smem[];
	double *st = (double *) smem;
	cp_async8(st + threadIdx.x, in + threadIdx.x);
	cp_async_commit();
	cp_async8(st + 32 + threadIdx.x, in + 32 + threadIdx.x);
	cp_async_commit();
	early[threadIdx.x] = st[threadIdx.x];
	cp_async_wait<1>();
	late[threadIdx.x] = st[threadIdx.x];
	early[32 + threadIdx.x] = st[32 + threadIdx.x];
	cp_async_wait<0>();
	late[32 + threadIdx.x] = st[32 + threadIdx.x];
}

/* a two-stage ring of bulk copies on mbarriers, phases alternating with every reuse */
static __global__ void
k_ring(const double *in, double *out, int nbatch, int *err)
{
	extern __shared__ __align__(128) unsigned char smem[];
	uint64_t *bar = (uint64_t *) smem;
	double *ring = (double *) (smem + 128);
	const int lane = threadIdx.x;
	if(lane < 2) mbar_init(bar + lane, 1);
	__syncwarp();
	if(lane == 0) { mbar_expect_tx(bar, 256); tma_bulk_load(ring, in, 256, bar); }
	for(int b = 0; b < nbatch; b++)
	{
		__syncwarp();        /* the stage refilled below was read in the previous iteration */
		if(lane == 0 && b + 1 < nbatch)
		{
			uint64_t *mb = bar + (b + 1) % 2;
			mbar_expect_tx(mb, 256);
			tma_bulk_load(ring + ((b + 1) % 2) * 32, in + (b + 1) * 32, 256, mb);
		}
		if(mbar_wait(bar + b % 2, (b / 2) & 1)) { *err = 1; return; }
		out[b * 32 + lane] = 2.0 * ring[(b % 2) * 32 + lane];
	}
}

/* 2-D tensor tile with rows and columns beyond the array: zero filled */
static __global__ void
k_tile(const __grid_constant__ CUtensorMap map, double *out, int c0, int c1, int n, int *err)
{
	extern __shared__ __align__(128) unsigned char smem[];
	uint64_t *bar = (uint64_t *) smem;
	double *tile = (double *) (smem + 128);
	if(threadIdx.x == 0)
	{
		mbar_init(bar, 1);
		mbar_expect_tx(bar, n * 8);
		tma_load_2d(tile, &map, c0, c1, bar);
	}
	__syncthreads();
	if(mbar_wait(bar, 0)) { *err = 1; return; }
	for(int i = threadIdx.x; i < n; i += blockDim.x) out[i] = tile[i];
}

/* lanes 0 and 1 wait for each other in different collectives: must be reported, not hang */
static __global__ void
k_deadlock(int *out)
{
	if(threadIdx.x == 0) out[0] = __shfl_sync(0x3u, 1, 0);
	else out[1] = (int) __ballot_sync(0x3u, 1);
}

static int failures;
#define EXPECT(cond, ...) do { if(!(cond)) { failures++; printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } while(0)

int
main()
{
	int *in, *out, *cnt;
	cudaMalloc(&in, 64 * sizeof(int));
	cudaMalloc(&out, 64 * sizeof(int));
	cudaMalloc(&cnt, 4 * sizeof(int));

	for(int i = 0; i < 64; i++) in[i] = i % 7;
	k_scan<<<1, 64>>>(in, out);
	for(int w = 0; w < 2; w++)
		for(int l = 0, s = 0; l < 32; l++) { s += in[w * 32 + l]; EXPECT(out[w * 32 + l] == s, "scan lane %d: %d != %d", l, out[w * 32 + l], s); }

	for(int i = 0; i < 32; i++) in[i] = (i * 5 + 3) % 11;
	k_compact<<<1, 32>>>(in, out, cnt);
	{
		int k = 0;
		for(int i = 0; i < 32; i++) if(in[i] % 2 == 0) { EXPECT(out[k] == in[i], "compaction slot %d", k); k++; }
		EXPECT(*cnt == k, "compaction count %d != %d", *cnt, k);
	}

	for(int i = 0; i < 32; i++) in[i] = i % 3 == 0 ? -1 : i % 4;
	k_match<<<1, 32>>>(in, out, cnt);
	{
		int seen[4] = { 0, 0, 0, 0 }, mx = -1;
		for(int i = 0; i < 32; i++)
		{
			const int want = in[i] < 0 ? -1 : seen[in[i]]++;
			EXPECT(out[i] == want, "match rank lane %d: %d != %d", i, out[i], want);
			if(want > mx) mx = want;
		}
		EXPECT(*cnt == mx, "reduce_max %d != %d", *cnt, mx);
	}

	double *din, *dout, *dout2;
	cudaMalloc(&din, 1024 * sizeof(double));
	cudaMalloc(&dout, 1024 * sizeof(double));
	cudaMalloc(&dout2, 1024 * sizeof(double));
	for(int i = 0; i < 1024; i++) din[i] = i + 0.25;
	k_block_sum<<<4, 256>>>(din, dout);
	for(int b = 0; b < 4; b++)
	{
		double s = 0;
		for(int i = 0; i < 256; i++) s += din[b * 256 + i];
		EXPECT(dout[b] == s, "block sum %d: %g != %g", b, dout[b], s);
	}

	k_async<<<1, 32, 1024>>>(din, dout, dout2);
	for(int i = 0; i < 64; i++)
	{
		EXPECT(isnan(dout[i]), "cp.async value %d visible before its wait (%g)", i, dout[i]);
		EXPECT(dout2[i] == din[i], "cp.async value %d wrong after the wait", i);
	}

	*cnt = 0;
	k_ring<<<1, 32, 128 + 2 * 256>>>(din, dout, 9, cnt);
	EXPECT(*cnt == 0, "mbarrier ring: a wait failed");
	for(int i = 0; i < 9 * 32; i++) EXPECT(dout[i] == 2.0 * din[i], "ring element %d: %g", i, dout[i]);

	{
		/* 8 x 6 array (row stride 8 doubles), box 4 x 3 */
		typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
				const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
				CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
		void *fn = NULL;
		cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault);
		CUtensorMap map;
		cuuint64_t dims[2] = { 8, 6 }, strides[1] = { 64 };
		cuuint32_t box[2] = { 4, 3 }, estr[2] = { 1, 1 };
		CUresult r = ((encode_fn) fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, din, dims, strides, box, estr,
				CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
				CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		EXPECT(r == CUDA_SUCCESS, "tensor map refused");
		box[0] = 3;       /* 24-byte rows: the driver refuses boxes whose inner extent is not a multiple of 16 B */
		CUtensorMap bad;
		EXPECT(((encode_fn) fn)(&bad, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, din, dims, strides, box, estr,
				CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
				CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS, "a misaligned box was accepted");
		*cnt = 0;
		k_tile<<<1, 32, 128 + 12 * 8>>>(map, dout, 6, 4, 12, cnt);
		EXPECT(*cnt == 0, "tensor load: wait failed");
		for(int rr = 0; rr < 3; rr++)
			for(int c = 0; c < 4; c++)
			{
				const int gx = 6 + c, gy = 4 + rr;
				const double want = (gx < 8 && gy < 6) ? din[gy * 8 + gx] : 0.0;
				EXPECT(dout[rr * 4 + c] == want, "tile (%d,%d): %g != %g", rr, c, dout[rr * 4 + c], want);
			}
	}

	EXPECT(cudaGetLastError() == cudaSuccess, "an error was pending before the deadlock test");
	fprintf(stderr, "(the next message is expected)\n");
	k_deadlock<<<1, 2>>>(out);
	EXPECT(cudaGetLastError() != cudaSuccess, "lanes waiting in different collectives were not reported");
	simt_check_reset();
	EXPECT(cudaGetLastError() == cudaSuccess, "reset did not clear the fault");
	for(int i = 0; i < 64; i++) in[i] = 1;
	k_scan<<<1, 64>>>(in, out);
	EXPECT(out[63] == 32, "launch after reset");

	printf(failures ? "simt-check selftest: %d FAILED\n" : "simt-check selftest: ok\n", failures);
	return failures ? 1 : 0;
}
