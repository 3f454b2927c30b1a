/* cpic_b200: simulation context, stage orchestration and the C ABI (include/cpic_b200.h).
 *
 * Stage order and semantics follow the reference's sim_step (src/sim.c:481-581); every
 * entry point names the reference function it replaces. Device work goes to one stream;
 * the pushes of the odd species run on a second one so that the last, partly filled wave
 * of one species' CTAs overlaps the first wave of the next.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <cufft.h>

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "cpic_b200.h"
#include "kernels.cuh"
#include "comm.h"
#include "host/glibc_rand.h"

#ifndef CPIC_B200_SIMT_CHECK
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#endif

/* ------------------------------------------------------------------ errors */

static thread_local char g_err[512] = "";

static int
fail(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}

#define CK(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) \
	return fail(CPIC_B200_ECUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while(0)
#define CKFFT(call) do { cufftResult r_ = (call); if(r_ != CUFFT_SUCCESS) \
	return fail(CPIC_B200_ECUDA, "%s:%d: %s: cufft error %d", __FILE__, __LINE__, #call, (int) r_); } while(0)

extern "C" const char *cpic_b200_last_error(void) { return g_err; }
extern "C" void cpic_b200_set_error_(const char *msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }
#ifdef CPIC_B200_SIMT_CHECK
/* tests/simt: this translation unit compiled for the CPU test interpreter says so, and the Python
 * view refuses it outside the test suite */
extern "C" const char *cpic_b200_version(void) { return "cpic_b200 0.1 (simt-check interpreter build: tests only)"; }
#else
extern "C" const char *cpic_b200_version(void) { return "cpic_b200 0.1 (sm_100a)"; }
#endif

/* ----------------------------------------------------------------- context */

struct SpeciesHost {
	SpeciesDev d;
	CUtensorMap segmap[2];   /* the segment arrays as a 2D tensor (slot, array): boxes of 32 x 5 and 32 x 6 */
	void *block;             /* one allocation: x y ux uy uz id count (the "image") */
	size_t block_bytes;
	void *oblock;            /* the two outboxes */
	void *fblock;            /* far-mover list */
	int reserve;             /* floor for the block capacity (cpic_b200_reserve) */
	std::vector<int> *tally; /* per-block counts of a streamed initialisation (count_particles / add_particles) */
	double *pE;              /* optional per-particle E (segment and outboxes) */
	int arr;                 /* outbox that holds the pending arrivals */
	int *hpos;               /* id -> place in the host's own order (cpic_b200_set_host_order) */
	long long hpos_n, horder_n;
	long long n;
	double q, m;
};

enum { T_FIELD_E, T_PUSH, T_EXCHANGE, T_RHO, T_SOLVER, T_GATHER, T_COUNT };

struct cpic_b200_sim {
	cpic_b200_params_t p;
	Geom g;
	int nb;                  /* particle blocks of the slab */
	int nob;                 /* outbox blocks: nb (+ 2*nbx ghost rows with several ranks) */
	long long iter;
	double umax[3];
	cudaStream_t stream;
	cudaStream_t stream2;    /* pushes of the odd species (forked from / joined into `stream`) */
	cudaStream_t stream_fields;              /* cpic_b200_get_fields_begin: grid downloads next to the compute stream */
	cudaEvent_t ev_fields_from, ev_fields_done;
	bool fields_pending;
	cudaStream_t stream_in, stream_out;      /* cpic_b200_step_host: uploads / downloads next to the compute stream */
	cudaEvent_t ev_up[CPIC_B200_MAX_SPECIES], ev_packed[CPIC_B200_MAX_SPECIES];
	double *hstage[2][CPIC_B200_MAX_SPECIES];        /* device staging of the species' images: [up / down] */
	size_t hstage_cap[2][CPIC_B200_MAX_SPECIES];
	long long *hoff[2][CPIC_B200_MAX_SPECIES];       /* block offsets of those images (nb + 1) */
	int *hcnt[CPIC_B200_MAX_SPECIES];                /* uploaded block counts */
	long long *h_band;                               /* pinned: new totals of the bands, far-mover counts */
	cudaEvent_t ev_band[CPIC_B200_MAX_SPECIES][2 * 64];      /* banded step: [species][uploaded j | packed 64 + j] */
	cudaEvent_t ev_fork, ev_join;
	bool overlap_species;
	int device;

	double *rho, *phi, *phi_raw, *Ex, *Ey, *G;
	cufftDoubleComplex *gk;
	cufftHandle plan_fwd, plan_inv;
	bool have_plans;
	double *tiles;           /* deposit: node sums per particle block, nb x (BX+1)(BY+1) */
	double *red;             /* reduction scratch */
	double *img;             /* staging of the compact particle image */
	long long *img_off;
	size_t img_cap;
	int *errflag;
	int *h_err;              /* pinned */
	CUtensorMap mapEx, mapEy;
	size_t smem_push, smem_dep;
	int dep_cols;            /* accumulator columns per node of the deposit (kernels.cuh: k_deposit) */
	int dep_ctas;            /* CTAs of the deposit that are resident at once on this device */

	SpeciesHost sp[CPIC_B200_MAX_SPECIES];

	Comm *comm;              /* NULL on one rank */
	bool p2p_stale;          /* the peers do not know this rank's current species storage (comm.h: peer memory) */

	bool timing;
	cudaEvent_t ev[2];
	double ms[T_COUNT];
	long long launches;
};

typedef cpic_b200_sim sim_t_;

static int
pick_div(long long n, int maxd)
{
	int d = maxd;
	while(d > 1 && n % d) d >>= 1;
	return d;
}

typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
		const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
		CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
		CUtensorMapFloatOOBfill);

static int
make_tensor_map(CUtensorMap *map, double *base, const Geom &g)
{
	static encode_fn encode = NULL;
	if(!encode)
	{
		void *fn = NULL;
		cudaDriverEntryPointQueryResult qres;
		CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
		if(!fn || qres != cudaDriverEntryPointSuccess)
			return fail(CPIC_B200_ECUDA, "cuTensorMapEncodeTiled is not available in this driver");
		encode = (encode_fn) fn;
	}
	/* dim 0 = columns (contiguous), dim 1 = rows */
	cuuint64_t dims[2] = { (cuuint64_t) g.SE, (cuuint64_t) (g.ny + 1) };
	cuuint64_t strides[1] = { (cuuint64_t) g.SE * sizeof(double) };
	cuuint32_t box[2] = { (cuuint32_t) g.TW, (cuuint32_t) g.TH };
	cuuint32_t estr[2] = { 1, 1 };
	CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr,
			CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
			CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if(r != CUDA_SUCCESS)
		return fail(CPIC_B200_ECUDA, "cuTensorMapEncodeTiled failed (%d) for E tile %dx%d of %dx%d",
				(int) r, g.TW, g.TH, g.SE, g.ny + 1);
	return 0;
}

/* The six segment arrays of a species (one allocation, `astride` doubles apart) as a 2D tensor:
 * dim 0 = slot, dim 1 = array; a box is one batch of 32 slots of the first `rows` arrays */
static int
make_segment_map(CUtensorMap *map, double *base, size_t nslot, size_t astride, int rows)
{
	static encode_fn encode = NULL;
	if(!encode)
	{
		void *fn = NULL;
		cudaDriverEntryPointQueryResult qres;
		CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
		if(!fn || qres != cudaDriverEntryPointSuccess)
			return fail(CPIC_B200_ECUDA, "cuTensorMapEncodeTiled is not available in this driver");
		encode = (encode_fn) fn;
	}
	cuuint64_t dims[2] = { (cuuint64_t) nslot, 6 };
	cuuint64_t strides[1] = { (cuuint64_t) astride * sizeof(double) };
	cuuint32_t box[2] = { 32, (cuuint32_t) rows };
	cuuint32_t estr[2] = { 1, 1 };
	CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr,
			CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
			CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if(r != CUDA_SUCCESS)
		return fail(CPIC_B200_ECUDA, "cuTensorMapEncodeTiled failed (%d) for the segments (%zu slots, stride %zu)", (int) r, nslot, astride);
	return 0;
}

static size_t
tile_bytes(const Geom &g)
{
	size_t b = (size_t) g.TH * g.TW * sizeof(double);
	return (b + 127) & ~(size_t) 127;
}

/* dynamic shared memory of k_gather_push<MODE>: header, scratch, two E tiles, per-warp rings */
template <int MODE>
static size_t
push_smem_bytes(const cpic_b200_sim *s)
{
	return s->smem_push + (size_t) s->g.WPC * PIPE_STAGES * PipeArrays<MODE>::N * 32 * sizeof(double);
}

extern "C" int
cpic_b200_create(const cpic_b200_params_t *pp, cpic_b200_sim_t **out)
{
	if(!pp || !out) return fail(CPIC_B200_EINVAL, "null argument");
	const cpic_b200_params_t &p = *pp;
	if(p.nx < 1 || p.ny < 1 || p.nx > (1 << 30) || p.ny > (1 << 30))
		return fail(CPIC_B200_EINVAL, "grid.points out of range");
	if(p.nranks < 1 || p.rank < 0 || p.rank >= p.nranks)
		return fail(CPIC_B200_EINVAL, "bad rank %d of %d", p.rank, p.nranks);
	/* reference src/sim.c:116-130 */
	if(p.ny % p.nranks)
		return fail(CPIC_B200_EINVAL, "The number of grid points in Y %lld cannot be divided by the number of processes %d",
				(long long) p.ny, p.nranks);
	if(p.plasma_chunks < 1 || p.nx % p.plasma_chunks)
		return fail(CPIC_B200_EINVAL, "The number of grid points in X %lld cannot be divided by the number of plasma chunks %lld",
				(long long) p.nx, (long long) p.plasma_chunks);
	if(p.nspecies < 1 || p.nspecies > CPIC_B200_MAX_SPECIES)
		return fail(CPIC_B200_EINVAL, "between 1 and %d species are supported", CPIC_B200_MAX_SPECIES);
	if(!(p.dt > 0) || !(p.Lx > 0) || !(p.Ly > 0) || !(p.e0 > 0))
		return fail(CPIC_B200_EINVAL, "time_step, space_length and vacuum_permittivity must be positive");
	if(p.nranks > 1 && (p.nx % 2))
		return fail(CPIC_B200_EINVAL, "several ranks need an even number of grid points in X");

	int ndev = 0;
	if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
		return fail(CPIC_B200_ECUDA, "no CUDA device: cpic_b200 has no CPU path");
	int dev = p.device;
	if(dev < 0) CK(cudaGetDevice(&dev));
	CK(cudaSetDevice(dev));

	sim_t_ *s = new sim_t_();
	memset(s, 0, sizeof(*s));
	s->p = p;
	s->device = dev;
	if(s->p.capacity_factor <= 1.0) s->p.capacity_factor = 1.5;

	Geom &g = s->g;
	g.nx = (int) p.nx;
	g.ny_glob = (int) p.ny;
	g.ny = (int) (p.ny / p.nranks);
	g.row0 = p.rank * g.ny;
	g.S = 2 * (g.nx / 2 + 1);
	int bc = p.block_cells > 0 ? p.block_cells : 8;
	if(getenv("CPIC_B200_BLOCK_CELLS")) bc = atoi(getenv("CPIC_B200_BLOCK_CELLS"));
	if(bc != 1 && bc != 2 && bc != 4 && bc != 8 && bc != 16 && bc != 32)
	{
		delete s;
		return fail(CPIC_B200_EINVAL, "block_cells must be a power of two between 1 and 32");
	}
	g.BX = pick_div(g.nx, bc);
	g.BY = pick_div(g.ny, bc);
	g.lBX = 0; while((1 << g.lBX) < g.BX) g.lBX++;
	g.lBY = 0; while((1 << g.lBY) < g.BY) g.lBY++;
	g.nbx = g.nx / g.BX;
	g.nby = g.ny / g.BY;
	g.nby_glob = g.nby * p.nranks;
	g.brow0 = p.rank * g.nby;
	g.WPC = pick_div(g.nbx, g.BX > 8 ? MAX_WPC / 2 : MAX_WPC);
	/* a TMA box holds at most 256 elements per dimension; the pair tile has 2*(WPC*BX+2) */
	while(g.WPC > 1 && 2 * (g.WPC * g.BX + 2) > 256) g.WPC >>= 1;
	g.TW = g.WPC * g.BX + 2;
	if(g.TW & 1) g.TW++;
	g.TH = g.BY + 1;
	g.tile_dbl = (int) ((((size_t) g.TH * g.TW * sizeof(double) + 127) & ~(size_t) 127) / sizeof(double));
	g.SE = g.nx + 2;
	if(g.SE & 1) g.SE++;
	if(g.SE < g.TW) g.SE = g.TW;
	g.Lx = p.Lx;
	g.Ly = p.Ly;
	g.dx = p.Lx / (double) p.nx;          /* reference src/sim.c:172-176 */
	g.dy = p.Ly / (double) p.ny;
	g.idx = 1.0 / g.dx;
	g.idy = 1.0 / g.dy;
	g.y0 = p.rank * (g.dy * g.ny);        /* reference src/field.c:37,149: rank * (dx[Y]*shape[Y]) */
	s->nb = g.nbx * g.nby;
	s->nob = s->nb + (p.nranks > 1 ? 2 * g.nbx : 0);
	s->iter = -1;
	/* reference src/sim.c:198-200 */
	s->umax[0] = (double) (p.nx / p.plasma_chunks) * g.dx / p.dt;
	s->umax[1] = (double) g.ny * g.dy / p.dt;
	s->umax[2] = 1.0 * 0.0 / p.dt;

#define CKD(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) { \
	int rc_ = fail(CPIC_B200_ECUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
	cpic_b200_destroy(s); return rc_; } } while(0)

	CKD(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
	CKD(cudaStreamCreateWithFlags(&s->stream2, cudaStreamNonBlocking));
	CKD(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
	CKD(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
	{
		const char *e = getenv("CPIC_B200_OVERLAP_SPECIES");
		s->overlap_species = !(e && atoi(e) == 0);
	}
	CKD(cudaEventCreate(&s->ev[0]));
	CKD(cudaEventCreate(&s->ev[1]));

	const size_t nc = (size_t) g.nx / 2 + 1;
	CKD(cudaMalloc(&s->rho, (size_t) (g.ny + 1) * g.S * sizeof(double)));
	CKD(cudaMalloc(&s->phi, (size_t) (g.ny + 3) * g.S * sizeof(double)));
	CKD(cudaMalloc(&s->phi_raw, (size_t) g.ny * g.S * sizeof(double)));
	CKD(cudaMalloc(&s->Ex, (size_t) (g.ny + 1) * g.SE * sizeof(double)));
	CKD(cudaMalloc(&s->Ey, (size_t) (g.ny + 1) * g.SE * sizeof(double)));
	CKD(cudaMemset(s->rho, 0, (size_t) (g.ny + 1) * g.S * sizeof(double)));
	CKD(cudaMemset(s->phi, 0, (size_t) (g.ny + 3) * g.S * sizeof(double)));
	/* cuFFT leaves the padding columns [nx, S) of its real output alone; they are copied into phi */
	CKD(cudaMemset(s->phi_raw, 0, (size_t) g.ny * g.S * sizeof(double)));
	CKD(cudaMemset(s->Ex, 0, (size_t) (g.ny + 1) * g.SE * sizeof(double)));
	CKD(cudaMemset(s->Ey, 0, (size_t) (g.ny + 1) * g.SE * sizeof(double)));
	CKD(cudaMalloc(&s->tiles, (size_t) s->nb * (g.BX + 1) * (g.BY + 1) * sizeof(double)));
	CKD(cudaMalloc(&s->red, ((size_t) std::max(s->nb, g.ny) + 16) * sizeof(double)));
	/* [0] deferred error bits, [1 .. MAX_SPECIES] capacity requests (agreed over the ranks) */
	CKD(cudaMalloc(&s->errflag, 16 * sizeof(int)));
	CKD(cudaMemset(s->errflag, 0, 16 * sizeof(int)));
	CKD(cudaMallocHost(&s->h_err, 16 * sizeof(int)));

	if(p.nranks == 1)
	{
		/* MFT_init, reference src/solver.c:207-335: the G table for all rows (one rank) */
		std::vector<double> G((size_t) g.ny * nc);
		const double cx = 2.0 * M_PI / (double) g.nx, cy = 2.0 * M_PI / (double) g.ny_glob;
		for(int iy = 0; iy < g.ny; iy++)
			for(size_t ix = 0; ix < nc; ix++)
				G[(size_t) iy * nc + ix] = (ix == 0 && iy == 0) ? 0.0 :
					1.0 / (2.0 * (cos(cx * (double) ix) + cos(cy * (double) iy)) - 4.0);
		CKD(cudaMalloc(&s->G, G.size() * sizeof(double)));
		CKD(cudaMemcpy(s->G, G.data(), G.size() * sizeof(double), cudaMemcpyHostToDevice));
		CKD(cudaMalloc(&s->gk, (size_t) g.ny * nc * sizeof(cufftDoubleComplex)));

		/* padded real rows of 2*(nx/2+1) doubles, as FFTW's r2c/c2r layout
		 * (reference src/solver.c:314-330) */
		int n[2] = { g.ny, g.nx };
		int rembed[2] = { g.ny, g.S };
		int cembed[2] = { g.ny, (int) nc };
#define CKFD(call) do { cufftResult r_ = (call); if(r_ != CUFFT_SUCCESS) { \
	int rc_ = fail(CPIC_B200_ECUDA, "%s:%d: %s: cufft error %d", __FILE__, __LINE__, #call, (int) r_); \
	cpic_b200_destroy(s); return rc_; } } while(0)
		CKFD(cufftPlanMany(&s->plan_fwd, 2, n, rembed, 1, g.ny * g.S, cembed, 1, g.ny * (int) nc, CUFFT_D2Z, 1));
		CKFD(cufftPlanMany(&s->plan_inv, 2, n, cembed, 1, g.ny * (int) nc, rembed, 1, g.ny * g.S, CUFFT_Z2D, 1));
		CKFD(cufftSetStream(s->plan_fwd, s->stream));
		CKFD(cufftSetStream(s->plan_inv, s->stream));
		s->have_plans = true;
	}

	int rc = make_tensor_map(&s->mapEx, s->Ex, g);
	if(!rc) rc = make_tensor_map(&s->mapEy, s->Ey, g);
	if(rc) { cpic_b200_destroy(s); return rc; }

	/* barrier + per-warp scratch + two E tiles + per-warp prefetch rings (sized for the
	 * widest mode: 8 arrays) */
	s->smem_push = PUSH_SMEM_HEADER + MAX_WPC * 32 * sizeof(int) + 2 * tile_bytes(g);
	/* deposit: [warp of the CTA][node][column]; 16 columns when a few CTAs still fit an SM */
	s->dep_cols = DEP_MAX_COLS;
	if(getenv("CPIC_B200_DEP_COLS"))       /* experiments: 1, 2, 4, 8 or 16 */
	{
		const int c = atoi(getenv("CPIC_B200_DEP_COLS"));
		if(c == 1 || c == 2 || c == 4 || c == 8 || c == 16) s->dep_cols = c;
	}
	while(s->dep_cols > 1 && (size_t) DEP_WARPS * (g.BX + 1) * (g.BY + 1) * s->dep_cols * sizeof(double) > 100 * 1024)
		s->dep_cols >>= 1;
	s->smem_dep = (size_t) DEP_WARPS * (g.BX + 1) * (g.BY + 1) * s->dep_cols * sizeof(double);
	/* the opt-in shared-memory limits are per function AND per device: set them for this
	 * simulation's device, whatever another simulation of the process did on another one */
	CKD((cudaFuncSetAttribute(k_deposit<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) s->smem_dep)));
	CKD((cudaFuncSetAttribute(k_deposit<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) s->smem_dep)));
	CKD((cudaFuncSetAttribute(k_deposit<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) s->smem_dep)));
	CKD((cudaFuncSetAttribute(k_deposit<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) s->smem_dep)));
	{
		int per_sm = 0, sms = 0;
		CKD((cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_deposit<true, 16>, 32 * DEP_WARPS, s->smem_dep)));
		CKD(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
		s->dep_ctas = std::max(1, per_sm) * std::max(1, sms);
	}
	CKD(cudaFuncSetAttribute(k_gather_push<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) push_smem_bytes<0>(s)));
	CKD(cudaFuncSetAttribute(k_gather_push<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) push_smem_bytes<1>(s)));
	CKD(cudaFuncSetAttribute(k_gather_push<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) push_smem_bytes<2>(s)));

	for(int i = 0; i < p.nspecies; i++) { s->sp[i].q = p.q[i]; s->sp[i].m = p.m[i]; }

	*out = s;
	return 0;
}

static void
free_species(SpeciesHost &h)
{
	cudaFree(h.block);
	cudaFree(h.oblock);
	cudaFree(h.fblock);
	cudaFree(h.pE);
	double q = h.q, m = h.m;
	int reserve = h.reserve;
	std::vector<int> *tally = h.tally;
	int *hpos = h.hpos;          /* the host's order outlives a change of the device layout */
	long long hpos_n = h.hpos_n, horder_n = h.horder_n;
	memset(&h, 0, sizeof(h));
	h.q = q; h.m = m; h.reserve = reserve; h.tally = tally;
	h.hpos = hpos; h.hpos_n = hpos_n; h.horder_n = horder_n;
}

extern "C" void
cpic_b200_destroy(cpic_b200_sim_t *s)
{
	if(!s) return;
	cudaSetDevice(s->device);
	if(s->stream) cudaStreamSynchronize(s->stream);
	for(int i = 0; i < CPIC_B200_MAX_SPECIES; i++)
	{
		free_species(s->sp[i]);
		delete s->sp[i].tally;
		s->sp[i].tally = NULL;
		cudaFree(s->sp[i].hpos);
		s->sp[i].hpos = NULL;
	}
	if(s->comm) comm_destroy(s->comm);
	if(s->have_plans) { cufftDestroy(s->plan_fwd); cufftDestroy(s->plan_inv); }
	cudaFree(s->rho); cudaFree(s->phi); cudaFree(s->phi_raw); cudaFree(s->Ex); cudaFree(s->Ey);
	cudaFree(s->G); cudaFree(s->gk); cudaFree(s->tiles);
	cudaFree(s->red); cudaFree(s->errflag); cudaFree(s->img); cudaFree(s->img_off);
	if(s->h_err) cudaFreeHost(s->h_err);
	if(s->ev[0]) cudaEventDestroy(s->ev[0]);
	if(s->ev[1]) cudaEventDestroy(s->ev[1]);
	if(s->stream) cudaStreamDestroy(s->stream);
	if(s->stream2) cudaStreamDestroy(s->stream2);
	if(s->stream_fields) { cudaStreamSynchronize(s->stream_fields); cudaStreamDestroy(s->stream_fields); }
	if(s->ev_fields_from) cudaEventDestroy(s->ev_fields_from);
	if(s->ev_fields_done) cudaEventDestroy(s->ev_fields_done);
	if(s->stream_in) cudaStreamDestroy(s->stream_in);
	if(s->stream_out) cudaStreamDestroy(s->stream_out);
	for(int i = 0; i < CPIC_B200_MAX_SPECIES; i++)
	{
		if(s->ev_up[i]) cudaEventDestroy(s->ev_up[i]);
		if(s->ev_packed[i]) cudaEventDestroy(s->ev_packed[i]);
		for(int k = 0; k < 2; k++) { cudaFree(s->hstage[k][i]); cudaFree(s->hoff[k][i]); }
		cudaFree(s->hcnt[i]);
		for(int k = 0; k < 128; k++) if(s->ev_band[i][k]) cudaEventDestroy(s->ev_band[i][k]);
	}
	if(s->h_band) cudaFreeHost(s->h_band);
	if(s->ev_fork) cudaEventDestroy(s->ev_fork);
	if(s->ev_join) cudaEventDestroy(s->ev_join);
	delete s;
}

/* --------------------------------------------------------------- particles */

static size_t
align256(size_t v) { return (v + 255) & ~(size_t) 255; }

/* Host lists may hold a particle exactly on the upper edge of the domain; the reference wraps it
 * when it places the particles (particle_comm_initial -> periodic_boundary, src/comm_plasma.c:725-747:
 * r >= L becomes r - L) */
static inline double wrap_upper(double r, double L) { return r >= L ? r - L : r; }

/* Several ranks: a particle is this rank's when its ROW is (the way every kernel and the slab filter
 * of the front end place it); the row at the slab's upper edge belongs to the next rank, and y == Ly
 * would have to be wrapped to rank 0 by the caller */
static inline bool
in_slab_rows(const Geom &g, double y)
{
	if(!(y >= 0.0 && y < g.Ly)) return false;
	const int row = global_row(g, y);
	return row >= g.row0 && row < g.row0 + g.ny;
}

static int ensure_particle_E(sim_t_ *s, int is);
static int image_staging(sim_t_ *s, size_t doubles);
static int p2p_attach(sim_t_ *s);
static int fields_guard(sim_t_ *s);

static int alloc_species_storage(sim_t_ *s, int is, int cap);

/* (Re)allocate one species with `cap` slots per block */
static int
alloc_species(sim_t_ *s, int is, int cap)
{
	SpeciesHost &h = s->sp[is];
	free_species(h);
	/* sizes first: nothing is allocated when a limit is exceeded */
	{
		const double frac = s->p.outbox_fraction > 0 ? s->p.outbox_fraction : 0.3;
		int ocs = (((int) ceil(cap * frac) + 31) / 32) * 32;
		if(ocs < 32) ocs = 32;
		const int occ = ((ocs / 4 + 31) / 32) * 32;
		/* slot numbers are 32-bit, and signed where they are TMA coordinates */
		if((double) s->nb * cap >= 2147483648.0 || (double) s->nob * (4.0 * ocs + 4.0 * occ) >= 4294967296.0)
			return fail(CPIC_B200_EINVAL, "species %d needs more than 2^31 particle slots on one GPU; lower capacity_factor / outbox_fraction or use more ranks", is);
	}
	int rc = alloc_species_storage(s, is, cap);
	if(rc) free_species(h);      /* never leave a half-built species behind */
	return rc;
}

static int
alloc_species_storage(sim_t_ *s, int is, int cap)
{
	SpeciesHost &h = s->sp[is];
	s->p2p_stale = true;
	const size_t nslot = (size_t) s->nb * cap;
	const size_t arr = align256(nslot * sizeof(double));
	const size_t cnt = align256((size_t) s->nob * sizeof(int));
	h.block_bytes = 6 * arr + cnt;
	CK(cudaMalloc(&h.block, h.block_bytes));
	CK(cudaMemsetAsync(h.block, 0, h.block_bytes, s->stream));
	char *b = (char *) h.block;
	/* plain layout: six arrays `arr` bytes apart; batch-major (SEG_AOSOA): 32 doubles apart */
	const size_t step = SEG_AOSOA ? 32 * sizeof(double) : arr;
	h.d.x = (double *) (b + 0 * step);
	h.d.y = (double *) (b + 1 * step);
	h.d.ux = (double *) (b + 2 * step);
	h.d.uy = (double *) (b + 3 * step);
	h.d.uz = (double *) (b + 4 * step);
	h.d.id = (long long *) (b + 5 * step);
	h.d.count = (int *) (b + 6 * arr);
	h.d.cap = cap;
	h.d.astride = (unsigned) (arr / sizeof(double));
	if(!SEG_AOSOA)
	{
		int rc = make_segment_map(&h.segmap[0], h.d.x, nslot, arr / sizeof(double), 5);
		if(!rc) rc = make_segment_map(&h.segmap[1], h.d.x, nslot, arr / sizeof(double), 6);
		if(rc) return rc;
	}

	/* outbox regions: four sides of ocs slots, four corners of occ, stored code-major */
	double frac = s->p.outbox_fraction > 0 ? s->p.outbox_fraction : 0.3;
	int ocs = (((int) ceil(cap * frac) + 31) / 32) * 32;
	if(ocs < 32) ocs = 32;
	int occ = ((ocs / 4 + 31) / 32) * 32;
	h.d.ocs = ocs;
	h.d.occ = occ;
	h.d.nob = s->nob;
	unsigned off = 0;
	for(int c = 0; c < 9; c++)
	{
		h.d.rcap[c] = c == DEST_STAY ? 0 : ((c & 1) ? ocs : occ);
		h.d.roff[c] = off;
		off += (unsigned) s->nob * (unsigned) h.d.rcap[c];
	}
	const size_t oslot = off;
	const size_t orec = align256(oslot * OREC * sizeof(double));
	const size_t ocnt = align256((size_t) s->nob * 9 * sizeof(int));
	const size_t one = orec + ocnt;
	CK(cudaMalloc(&h.oblock, 2 * one));
	CK(cudaMemsetAsync(h.oblock, 0, 2 * one, s->stream));
	for(int k = 0; k < 2; k++)
	{
		char *o = (char *) h.oblock + k * one;
		Outbox &ob = h.d.ob[k];
		ob.rec = (double *) o;
		ob.count = (int *) (o + orec);
		ob.recE = NULL;
	}
	h.arr = 0;

	const size_t fbytes = ((size_t) FAR_CAP * 10 + (size_t) FAR_FACE * 16) * sizeof(double) + 256;
	CK(cudaMalloc(&h.fblock, fbytes));
	CK(cudaMemsetAsync(h.fblock, 0, fbytes, s->stream));
	{
		double *f = (double *) h.fblock;
		h.d.fx = f; h.d.fy = f + FAR_CAP; h.d.fux = f + 2 * FAR_CAP; h.d.fuy = f + 3 * FAR_CAP;
		h.d.fuz = f + 4 * FAR_CAP; h.d.fEx = f + 5 * FAR_CAP; h.d.fEy = f + 6 * FAR_CAP;
		h.d.fid = (long long *) (f + 7 * FAR_CAP);
		h.d.fkey = (long long *) (f + 8 * FAR_CAP);
		h.d.fidx = (int *) (f + 9 * FAR_CAP);
		h.d.rfar[0] = f + 10 * FAR_CAP;
		h.d.rfar[1] = f + 10 * FAR_CAP + 8 * FAR_FACE;
		h.d.fcount = (int *) (f + 10 * FAR_CAP + 16 * FAR_FACE);
		h.d.rfcount = h.d.fcount + 4;
	}

	if(s->p.keep_particle_E) return ensure_particle_E(s, is);
	return 0;
}

static int
ensure_particle_E(sim_t_ *s, int is)
{
	SpeciesHost &h = s->sp[is];
	if(h.d.pEx || !h.block) return 0;
	s->p2p_stale = true;
	const size_t arr = align256((size_t) s->nb * h.d.cap * sizeof(double));
	/* (E_x, E_y) per outbox slot, next to the records */
	const size_t oarr = align256(((size_t) h.d.roff[8] + (size_t) s->nob * h.d.rcap[8]) * 2 * sizeof(double));
	CK(cudaMalloc(&h.pE, 2 * arr + 2 * oarr));
	CK(cudaMemsetAsync(h.pE, 0, 2 * arr + 2 * oarr, s->stream));
	h.d.pEx = h.pE;
	h.d.pEy = (double *) ((char *) h.pE + arr);
	for(int k = 0; k < 2; k++)
		h.d.ob[k].recE = (double *) ((char *) h.pE + 2 * arr + k * oarr);
	return 0;
}

static int
cap_for(const sim_t_ *s, long long maxcount)
{
	long long c = (long long) ceil((double) maxcount * s->p.capacity_factor) + 64;
	c = (c + 31) / 32 * 32;
	if(c > (1LL << 30)) c = 1LL << 30;
	return (int) c;
}

/* plasma_init + particle_comm_initial (reference src/plasma.c:292-316 and
 * src/comm_plasma.c:1122-1142 in global mode): the host sorts the particles into their
 * blocks with a stable counting sort (input order is kept inside a block) and uploads. */
extern "C" int
cpic_b200_set_particles(cpic_b200_sim_t *s, int is, int64_t n, const int64_t *id,
		const double *x, const double *y, const double *ux, const double *uy, const double *uz)
{
	if(!s || is < 0 || is >= s->p.nspecies || n < 0) return fail(CPIC_B200_EINVAL, "bad species or count");
	if(n > 0 && (!x || !y || !ux || !uy)) return fail(CPIC_B200_EINVAL, "null particle array");
	CK(cudaSetDevice(s->device));
	const Geom &g = s->g;
	const double y1 = g.y0 + g.dy * g.ny;

	std::vector<int> blk((size_t) n);
	std::vector<int> cnt((size_t) s->nb, 0);
	const bool one = s->p.nranks == 1;
	for(int64_t i = 0; i < n; i++)
	{
		/* several ranks: the row at the slab's upper edge belongs to the next rank */
		if(!(x[i] >= 0.0 && x[i] <= g.Lx) || !(one ? (y[i] >= 0.0 && y[i] <= y1) : in_slab_rows(g, y[i])))
			return fail(CPIC_B200_EINVAL, "particle %lld at (%g, %g) is outside this rank's slab [0,%g]x[%g,%g%c",
					(long long) i, x[i], y[i], g.Lx, g.y0, y1, one ? ']' : ')');
		int b = block_of(g, wrap_upper(x[i], g.Lx), one ? wrap_upper(y[i], g.Ly) : y[i]);
		blk[(size_t) i] = b;
		cnt[(size_t) b]++;
	}
	int maxc = 0;
	for(int b = 0; b < s->nb; b++) maxc = std::max(maxc, cnt[(size_t) b]);
	/* leave room for the mean as well: a nearly empty start must not pin the capacity */
	long long mean = s->nb ? (n + s->nb - 1) / s->nb : 0;
	int cap = cap_for(s, std::max<long long>(maxc, mean));

	SpeciesHost &h = s->sp[is];
	if(cap < h.reserve) cap = h.reserve;
	if(!h.block || h.d.cap != cap)
	{
		int rc = alloc_species(s, is, cap);
		if(rc) return rc;
	}
	cap = h.d.cap;

	/* the host image of the six segment arrays, in the device's layout: array a starts astep[a]
	 * doubles after x, element i of block b sits seg_slot() further */
	const size_t seg_doubles = (h.block_bytes - align256((size_t) s->nob * sizeof(int))) / sizeof(double);
	const size_t astep = ((char *) h.d.y - (char *) h.d.x) / sizeof(double);
	std::vector<double> hseg(seg_doubles, 0.0);
	std::vector<int> fill((size_t) s->nb, 0);
	for(int64_t i = 0; i < n; i++)
	{
		int b = blk[(size_t) i];
		const size_t k = seg_slot(cap, b, fill[(size_t) b]++);
		hseg[k] = wrap_upper(x[i], g.Lx); hseg[k + astep] = one ? wrap_upper(y[i], g.Ly) : y[i];
		hseg[k + 2 * astep] = ux[i]; hseg[k + 3 * astep] = uy[i]; hseg[k + 4 * astep] = uz ? uz[i] : 0.0;
		const long long pid = id ? id[i] : i;
		memcpy(&hseg[k + 5 * astep], &pid, sizeof(pid));
	}
	CK(cudaMemcpyAsync(h.d.x, hseg.data(), seg_doubles * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	CK(cudaMemcpyAsync(h.d.count, cnt.data(), (size_t) s->nb * sizeof(int), cudaMemcpyHostToDevice, s->stream));
	for(int k = 0; k < 2; k++) CK(cudaMemsetAsync(h.d.ob[k].count, 0, (size_t) s->nob * 9 * sizeof(int), s->stream));
	CK(cudaStreamSynchronize(s->stream));
	h.n = n;
	return 0;
}

/* Streamed initialisation for populations whose host arrays do not fit at once (1e9 particles,
 * SURVEY 8d): the caller generates the population twice, batch by batch --
 *   1. cpic_b200_count_particles(): per-block tallies, on the host;
 *   2. cpic_b200_reserve_counted(): segments sized from the tallies;
 *   3. cpic_b200_add_particles(): every batch is binned (stable) and appended to its blocks.
 * Inside a block the particles end up in batch order, then input order. */
extern "C" int
cpic_b200_count_particles(cpic_b200_sim_t *s, int is, int64_t n, const double *x, const double *y)
{
	if(!s || is < 0 || is >= s->p.nspecies || n < 0 || (n && (!x || !y)))
		return fail(CPIC_B200_EINVAL, "bad species or arrays");
	SpeciesHost &h = s->sp[is];
	const Geom &g = s->g;
	/* tallies are kept for the blocks of every rank's slab: each rank sees the whole stream, so all
	 * of them arrive at the same capacity without talking to each other */
	if(!h.tally) h.tally = new std::vector<int>((size_t) g.nbx * (size_t) g.nby_glob, 0);
	for(int64_t i = 0; i < n; i++)
	{
		if(!(x[i] >= 0.0 && x[i] <= g.Lx) || !(y[i] >= 0.0 && y[i] <= g.Ly))
			return fail(CPIC_B200_EINVAL, "species %d: particle %lld at (%g, %g) lies outside the domain", is,
					(long long) i, x[i], y[i]);
		const size_t b = (size_t) (global_row(g, wrap_upper(y[i], g.Ly)) >> g.lBY) * (size_t) g.nbx
			+ (size_t) (cell_ix(g, wrap_upper(x[i], g.Lx)) >> g.lBX);
		(*h.tally)[b]++;
	}
	return 0;
}

extern "C" int
cpic_b200_reserve_counted(cpic_b200_sim_t *s, int is)
{
	if(!s || is < 0 || is >= s->p.nspecies) return fail(CPIC_B200_EINVAL, "bad species");
	CK(cudaSetDevice(s->device));
	SpeciesHost &h = s->sp[is];
	if(!h.tally) return fail(CPIC_B200_EINVAL, "species %d: cpic_b200_count_particles was not called", is);
	long long total = 0;
	int maxc = 0;
	for(int c : *h.tally) { total += c; maxc = std::max(maxc, c); }
	const long long nbg = (long long) h.tally->size();
	const long long mean = nbg ? (total + nbg - 1) / nbg : 0;
	int cap = cap_for(s, std::max<long long>(maxc, mean));
	if(cap < h.reserve) cap = h.reserve;
	delete h.tally;
	h.tally = NULL;
	int rc = alloc_species(s, is, cap);       /* zero counts, empty outboxes */
	if(rc) return rc;
	h.n = 0;
	CK(cudaStreamSynchronize(s->stream));
	return 0;
}

extern "C" int
cpic_b200_add_particles(cpic_b200_sim_t *s, int is, int64_t n, const int64_t *id,
		const double *x, const double *y, const double *ux, const double *uy, const double *uz)
{
	if(!s || is < 0 || is >= s->p.nspecies || n < 0 || (n && (!id || !x || !y || !ux || !uy)))
		return fail(CPIC_B200_EINVAL, "bad species or arrays");
	CK(cudaSetDevice(s->device));
	SpeciesHost &h = s->sp[is];
	if(!h.block) return fail(CPIC_B200_EINVAL, "species %d has no storage: reserve_counted or set_particles first", is);
	if(n == 0) return 0;
	const Geom &g = s->g;
	/* stable binning of the batch, then the compact image the append kernel scatters */
	std::vector<int> blk((size_t) n), cnt((size_t) s->nb, 0), have((size_t) s->nb);
	for(int64_t i = 0; i < n; i++)
	{
		const double y1 = g.y0 + g.dy * g.ny;
		if(!(x[i] >= 0.0 && x[i] <= g.Lx) || !(s->p.nranks == 1 ? (y[i] >= 0.0 && y[i] <= y1) : in_slab_rows(g, y[i])))
			return fail(CPIC_B200_EINVAL, "species %d: particle %lld at (%g, %g) lies outside this rank's slab", is,
					(long long) i, x[i], y[i]);
		const int b = block_of(g, wrap_upper(x[i], g.Lx), s->p.nranks == 1 ? wrap_upper(y[i], g.Ly) : y[i]);
		blk[(size_t) i] = b;
		cnt[(size_t) b]++;
	}
	CK(cudaMemcpyAsync(have.data(), h.d.count, have.size() * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaStreamSynchronize(s->stream));
	std::vector<long long> off((size_t) s->nb);
	long long m = 0;
	for(int b = 0; b < s->nb; b++)
	{
		if(have[(size_t) b] + cnt[(size_t) b] > h.d.cap)
			return fail(CPIC_B200_ECAPACITY, "species %d: block %d would hold %d particles, capacity %d", is, b,
					have[(size_t) b] + cnt[(size_t) b], h.d.cap);
		off[(size_t) b] = m;
		m += cnt[(size_t) b];
	}
	std::vector<double> img((size_t) 6 * (size_t) n);
	std::vector<long long> at(off);
	for(int64_t i = 0; i < n; i++)
	{
		const size_t k = (size_t) at[(size_t) blk[(size_t) i]]++;
		img[k] = wrap_upper(x[i], g.Lx); img[(size_t) n + k] = s->p.nranks == 1 ? wrap_upper(y[i], g.Ly) : y[i];
		img[2 * (size_t) n + k] = ux[i]; img[3 * (size_t) n + k] = uy[i]; img[4 * (size_t) n + k] = uz ? uz[i] : 0.0;
		memcpy(&img[5 * (size_t) n + k], &id[i], sizeof(int64_t));
	}
	int rc = image_staging(s, (size_t) 6 * (size_t) n);
	if(rc) return rc;
	int *dcnt = (int *) (s->img_off + s->nb + 1);
	CK(cudaMemcpyAsync(s->img_off, off.data(), off.size() * sizeof(long long), cudaMemcpyHostToDevice, s->stream));
	CK(cudaMemcpyAsync(dcnt, cnt.data(), cnt.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream));
	CK(cudaMemcpyAsync(s->img, img.data(), img.size() * sizeof(double), cudaMemcpyHostToDevice, s->stream));
	k_image_copy<<<(s->nb + 7) / 8, 256, 0, s->stream>>>(h.d, s->nb, dcnt, s->img_off, s->img, n, -1);
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(s->stream));     /* the host vectors go out of scope */
	h.n += n;
	return 0;
}

extern "C" int
cpic_b200_init_uniform(cpic_b200_sim_t *s, int is, int64_t n, int64_t id0, double vx, double vy, uint64_t seed)
{
	return cpic_b200_init_beam(s, is, n, id0, 0.0, 0.0, vx, vy, seed);
}

extern "C" int
cpic_b200_init_beam(cpic_b200_sim_t *s, int is, int64_t n, int64_t id0, double dux, double duy,
		double vx, double vy, uint64_t seed)
{
	if(!s || is < 0 || is >= s->p.nspecies || n < 0) return fail(CPIC_B200_EINVAL, "bad species or count");
	CK(cudaSetDevice(s->device));
	long long per = (n + s->nb - 1) / s->nb;
	int cap = cap_for(s, per);
	SpeciesHost &h = s->sp[is];
	if(cap < h.reserve) cap = h.reserve;
	if(!h.block || h.d.cap != cap)
	{
		int rc = alloc_species(s, is, cap);
		if(rc) return rc;
	}
	k_init_uniform<<<(s->nb + 7) / 8, 256, 0, s->stream>>>(h.d, s->g, s->nb, n, id0, vx, vy, dux, duy, seed);
	CK(cudaGetLastError());
	for(int k = 0; k < 2; k++) CK(cudaMemsetAsync(h.d.ob[k].count, 0, (size_t) s->nob * 9 * sizeof(int), s->stream));
	h.n = n;
	return 0;
}

/* plasma_init with the reference's own initial conditions drawn on the device (the host initialiser
 * makes four serial rand() calls per particle: minutes for 1e9 particles). Two passes over the runs:
 * the first tallies every particle per particle block (all slabs: every rank arrives at the same
 * capacity), the second keeps this rank's slab, sorts every batch by block (stable radix sort: a
 * block holds its particles in the order they were drawn) and appends it to the segments. */
struct RefInitScratch {
	double *x, *y, *ux, *uy;
	long long *id;
	unsigned *states;
	int *key, *val, *key2, *val2, *cnt, *tally;
	long long *off;
	void *tmp;
	size_t tmp_bytes;
};

static void
refinit_free(RefInitScratch &w)
{
	cudaFree(w.x); cudaFree(w.id); cudaFree(w.states); cudaFree(w.key); cudaFree(w.cnt); cudaFree(w.tally);
	cudaFree(w.off); cudaFree(w.tmp);
	memset(&w, 0, sizeof(w));
}

/* One batch of one run into w.x .. w.id: thread states from `st` (advanced past the batch on return) */
static int
refinit_batch(sim_t_ *s, RefInitScratch &w, const cpic_b200_init_run_t &r, long long k0, long long n,
		uint32_t st[31], const uint32_t stepm[31][31], std::vector<uint32_t> &hstates)
{
	const long long nthreads = (n + REFINIT_K - 1) / REFINIT_K;
	RefInitRun run;
	run.method = r.method; run.first = r.first; run.step = r.step;
	for(int d = 0; d < 2; d++) { run.v[d] = r.v[d]; run.dr[d] = r.dr[d]; run.r0[d] = r.r0[d]; }
	run.Lx = s->g.Lx; run.Ly = s->g.Ly;
	if(r.method == 0)
	{
		hstates.resize((size_t) nthreads * 31);
		for(long long t = 0; t < nthreads; t++)
		{
			memcpy(&hstates[(size_t) t * 31], st, 31 * sizeof(uint32_t));
			GlibcRandJump::apply(stepm, st);          /* 4 * REFINIT_K draws further */
		}
		CK(cudaMemcpyAsync(w.states, hstates.data(), hstates.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
	}
	k_refinit_generate<<<(unsigned) ((nthreads + REFINIT_THREADS - 1) / REFINIT_THREADS), REFINIT_THREADS, 0, s->stream>>>(
			run, k0, n, w.states, w.x, w.y, w.ux, w.uy, w.id);
	CK(cudaGetLastError());
	/* hstates is overwritten by the next batch */
	CK(cudaStreamSynchronize(s->stream));
	return 0;
}

extern "C" int
cpic_b200_init_reference(cpic_b200_sim_t *s, int nruns, const cpic_b200_init_run_t *runs, int64_t batch)
{
	if(!s || nruns < 0 || (nruns && !runs)) return fail(CPIC_B200_EINVAL, "bad runs");
	CK(cudaSetDevice(s->device));
	if(batch < REFINIT_K) batch = 1 << 22;
	batch = (batch + REFINIT_K - 1) / REFINIT_K * REFINIT_K;
	if(batch > (1LL << 30)) batch = 1LL << 30;
	static const GlibcRandJump *J = new GlibcRandJump();
	uint32_t stepm[31][31];
	J->power(4ull * REFINIT_K, stepm);
	const Geom &g = s->g;
	const size_t nbg = (size_t) g.nbx * g.nby_glob;
	RefInitScratch w;
	memset(&w, 0, sizeof(w));
#define CKW(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) { refinit_free(w); \
	return fail(CPIC_B200_ECUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } } while(0)
	CKW(cudaMalloc(&w.x, (size_t) batch * 4 * sizeof(double)));
	w.y = w.x + batch; w.ux = w.y + batch; w.uy = w.ux + batch;
	CKW(cudaMalloc(&w.id, (size_t) batch * sizeof(long long)));
	CKW(cudaMalloc(&w.states, (size_t) (batch / REFINIT_K) * 31 * sizeof(unsigned)));
	CKW(cudaMalloc(&w.key, (size_t) batch * 4 * sizeof(int)));
	w.val = w.key + batch; w.key2 = w.val + batch; w.val2 = w.key2 + batch;
	CKW(cudaMalloc(&w.cnt, ((size_t) s->nb + 1) * sizeof(int)));
	CKW(cudaMalloc(&w.off, ((size_t) s->nb + 1) * sizeof(long long)));
	std::vector<uint32_t> hstates;
	int rc = 0;
	for(int pass = 0; pass < 2 && !rc; pass++)
	{
		if(pass == 0)
		{
			/* per species: tallies over the blocks of every slab */
			CKW(cudaMalloc(&w.tally, (size_t) s->p.nspecies * nbg * sizeof(int)));
			CKW(cudaMemsetAsync(w.tally, 0, (size_t) s->p.nspecies * nbg * sizeof(int), s->stream));
		}
		else
		{
			for(int is = 0; is < s->p.nspecies && !rc; is++)
			{
				SpeciesHost &h = s->sp[is];
				delete h.tally;
				h.tally = new std::vector<int>(nbg);
				CKW(cudaMemcpy(h.tally->data(), w.tally + (size_t) is * nbg, nbg * sizeof(int), cudaMemcpyDeviceToHost));
				long long total = 0;
				for(int c : *h.tally) total += c;
				if(total == 0) { delete h.tally; h.tally = NULL; continue; }
				rc = cpic_b200_reserve_counted(s, is);
			}
			if(rc) break;
#ifndef CPIC_B200_SIMT_CHECK
			size_t t1 = 0, t2 = 0;
			cub::DeviceRadixSort::SortPairs(NULL, t1, w.key, w.key2, w.val, w.val2, (int) batch, 0, 32, s->stream);
			cub::DeviceScan::ExclusiveSum(NULL, t2, w.cnt, w.off, s->nb + 1, s->stream);
			w.tmp_bytes = std::max(t1, t2);
			CKW(cudaMalloc(&w.tmp, w.tmp_bytes));
#endif
		}
		for(int ir = 0; ir < nruns && !rc; ir++)
		{
			const cpic_b200_init_run_t &r = runs[ir];
			if(r.species < 0 || r.species >= s->p.nspecies || r.count < 0 || r.step < 1)
			{ rc = fail(CPIC_B200_EINVAL, "run %d: bad species, count or step", ir); break; }
			uint32_t st[31];
			glibc_rand_seed(r.seed, st);
			J->jump(st, (uint64_t) r.draw0);
			SpeciesHost &h = s->sp[r.species];
			for(long long k0 = 0; k0 < r.count && !rc; k0 += batch)
			{
				const long long n = std::min<long long>(batch, r.count - k0);
				if((rc = refinit_batch(s, w, r, k0, n, st, stepm, hstates))) break;
				const unsigned grid = (unsigned) ((n + 255) / 256);
				if(pass == 0)
				{
					k_refinit_keys<<<grid, 256, 0, s->stream>>>(w.x, w.y, n, g, s->nb, w.tally + (size_t) r.species * nbg, NULL, NULL, NULL);
					CKW(cudaGetLastError());
					continue;
				}
#ifdef CPIC_B200_SIMT_CHECK
				/* the CPU test build has no device sort: the batch is binned by the host path */
				std::vector<double> hx((size_t) n * 4);
				std::vector<int64_t> hid((size_t) n);
				CKW(cudaMemcpy(hx.data(), w.x, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost));
				CKW(cudaMemcpy(hx.data() + n, w.y, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost));
				CKW(cudaMemcpy(hx.data() + 2 * n, w.ux, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost));
				CKW(cudaMemcpy(hx.data() + 3 * n, w.uy, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost));
				CKW(cudaMemcpy(hid.data(), w.id, (size_t) n * sizeof(int64_t), cudaMemcpyDeviceToHost));
				std::vector<double> kx, ky, kux, kuy;
				std::vector<int64_t> kid;
				for(long long j = 0; j < n; j++)
				{
					const double py = hx[(size_t) (n + j)] >= g.Ly ? hx[(size_t) (n + j)] - g.Ly : hx[(size_t) (n + j)];
					const int row = global_row(g, py);
					if(row < g.row0 || row >= g.row0 + g.ny) continue;
					kx.push_back(hx[(size_t) j]); ky.push_back(py);
					kux.push_back(hx[(size_t) (2 * n + j)]); kuy.push_back(hx[(size_t) (3 * n + j)]);
					kid.push_back(hid[(size_t) j]);
				}
				rc = cpic_b200_add_particles(s, r.species, (int64_t) kid.size(), kid.data(), kx.data(), ky.data(), kux.data(), kuy.data(), NULL);
#else
				CKW(cudaMemsetAsync(w.cnt, 0, ((size_t) s->nb + 1) * sizeof(int), s->stream));
				k_refinit_keys<<<grid, 256, 0, s->stream>>>(w.x, w.y, n, g, s->nb, NULL, w.key, w.val, w.cnt);
				CKW(cudaGetLastError());
				int bits = 1;
				while((1LL << bits) <= s->nb) bits++;
				size_t tb = w.tmp_bytes;
				CKW(cub::DeviceRadixSort::SortPairs(w.tmp, tb, w.key, w.key2, w.val, w.val2, (int) n, 0, bits, s->stream));
				tb = w.tmp_bytes;
				CKW(cub::DeviceScan::ExclusiveSum(w.tmp, tb, w.cnt, w.off, s->nb + 1, s->stream));
				long long n_in = 0;
				CKW(cudaMemcpyAsync(&n_in, w.off + s->nb, sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
				CKW(cudaStreamSynchronize(s->stream));
				if(n_in == 0) continue;
				if((rc = image_staging(s, (size_t) 6 * (size_t) n_in))) break;
				k_refinit_gather<<<(unsigned) ((n_in + 255) / 256), 256, 0, s->stream>>>(w.val2, n_in, w.x, w.y, w.ux, w.uy, w.id, s->img, g);
				k_image_copy<<<(s->nb + 7) / 8, 256, 0, s->stream>>>(h.d, s->nb, w.cnt, w.off, s->img, n_in, -1);
				CKW(cudaGetLastError());
				CKW(cudaStreamSynchronize(s->stream));
				h.n += n_in;
#endif
			}
		}
	}
#undef CKW
	/* a block that ended up fuller than its capacity would have been refused by the append */
	refinit_free(w);
	return rc;
}

/* Host view of the fill of one species: counts of the segments and of the regions that hold
 * pending arrivals */
static int
occupancy(sim_t_ *s, int is, int64_t out[6])
{
	SpeciesHost &h = s->sp[is];
	memset(out, 0, 6 * sizeof(int64_t));
	if(!h.block) return 0;
	std::vector<int> cnt((size_t) s->nb), oc((size_t) s->nob * 9);
	CK(cudaMemcpyAsync(cnt.data(), h.d.count, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaMemcpyAsync(oc.data(), h.d.ob[h.arr].count, oc.size() * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaStreamSynchronize(s->stream));
	/* a block's next segment must hold everything it has now plus what is on its way from its
	 * eight neighbours (the same neighbour walk as find_arrivals, kernels.cuh) */
	const Geom &g = s->g;
	int64_t mb = 0, ms = 0, mc = 0;
	for(int b = 0; b < s->nb; b++)
	{
		const int bx = b % g.nbx, by = b / g.nbx;
		int64_t need = cnt[(size_t) b];
		for(int k = 0; k < 9; k++)
		{
			if(k == DEST_STAY) continue;
			int nx_ = bx + k % 3 - 1, ny_ = by + k / 3 - 1, src;
			if(nx_ < 0) nx_ += g.nbx; else if(nx_ >= g.nbx) nx_ -= g.nbx;
			if(g.nby_glob == g.nby)
			{
				if(ny_ < 0) ny_ += g.nby; else if(ny_ >= g.nby) ny_ -= g.nby;
				src = ny_ * g.nbx + nx_;
			}
			else if(ny_ < 0) src = s->nb + nx_;                /* north ghost row */
			else if(ny_ >= g.nby) src = s->nb + g.nbx + nx_;   /* south ghost row */
			else src = ny_ * g.nbx + nx_;
			need += oc[(size_t) (8 - k) * s->nob + (size_t) src];
		}
		mb = std::max(mb, need);
	}
	for(int c = 0; c < 9; c++)
	{
		if(c == DEST_STAY) continue;
		for(int b = 0; b < s->nob; b++)
		{
			const int v = oc[(size_t) c * s->nob + b];
			if(c & 1) ms = std::max<int64_t>(ms, v); else mc = std::max<int64_t>(mc, v);
		}
	}
	out[0] = mb; out[1] = h.d.cap;
	out[2] = ms; out[3] = h.d.ocs;
	out[4] = mc; out[5] = h.d.occ;
	return 0;
}

extern "C" int
cpic_b200_occupancy(cpic_b200_sim_t *s, int is, int64_t out[6])
{
	if(!s || !out || is < 0 || is >= s->p.nspecies) return fail(CPIC_B200_EINVAL, "bad argument");
	CK(cudaSetDevice(s->device));
	return occupancy(s, is, out);
}

static int alloc_species(sim_t_ *s, int is, int cap);

/* Re-lays a species out with `newcap` slots per block (and exchange regions in proportion) */
static int
regrow(sim_t_ *s, int is, int newcap)
{
	SpeciesHost old = s->sp[is];
	SpeciesHost &h = s->sp[is];
	const bool hadE = old.d.pEx != NULL;
	/* detach the old storage so that alloc_species does not free it */
	h.block = h.oblock = h.fblock = NULL;
	h.pE = NULL;
	memset(&h.d, 0, sizeof(h.d));
	int rc = alloc_species(s, is, newcap);
	if(!rc && hadE && !h.d.pEx) rc = ensure_particle_E(s, is);
	if(rc)
	{
		/* the species keeps its old storage (and the error of the failed allocation) */
		free_species(h);
		h = old;
		return rc;
	}
	k_regrow<<<(s->nb + 7) / 8, 256, 0, s->stream>>>(old.d, h.d, s->g, s->nb, old.arr);
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(s->stream));
	h.n = old.n;
	cudaFree(old.block); cudaFree(old.oblock); cudaFree(old.fblock); cudaFree(old.pE);
	return 0;
}

/* Grows the storage of any species that runs above 80 % of its block or region capacity.
 * Collective decisions are not needed on one rank; with several ranks the capacities must
 * stay equal, so growth there is left to the caller (cpic_b200_reserve). */
static int
check_capacity(sim_t_ *s)
{
	int want[CPIC_B200_MAX_SPECIES] = { 0 };
	bool any = false;
	for(int is = 0; is < s->p.nspecies; is++)
	{
		int64_t o[6];
		int rc = occupancy(s, is, o);
		if(rc) return rc;
		if(!o[1]) continue;
		const bool tight = o[0] * 10 > o[1] * 8 || o[2] * 10 > o[3] * 8 || o[4] * 10 > o[5] * 8;
		if(!tight) continue;
		int64_t w = std::max<int64_t>(o[1] * 3 / 2, o[0] * 2);
		/* regions scale with the block capacity; make sure they clear the observed peaks */
		const double frac = s->p.outbox_fraction > 0 ? s->p.outbox_fraction : 0.3;
		w = std::max<int64_t>(w, (int64_t) (2.0 * o[2] / frac));
		w = std::max<int64_t>(w, (int64_t) (8.0 * o[4] / frac));
		w = (w + 31) / 32 * 32;
		want[is] = (int) std::min<int64_t>(w, 1 << 30);
		any = true;
	}
	if(s->comm)
	{
		/* every rank must keep the same capacities (the exchange buffers are sized from them):
		 * the largest request wins. All ranks reach this point at the same steps. */
		for(int is = 0; is < s->p.nspecies; is++) s->h_err[1 + is] = want[is];
		CK(cudaMemcpyAsync(s->errflag + 1, s->h_err + 1, (size_t) s->p.nspecies * sizeof(int), cudaMemcpyHostToDevice, s->stream));
		if(comm_allreduce_max(s->comm, s->errflag + 1, s->p.nspecies, s->stream)) return CPIC_B200_ECUDA;
		CK(cudaMemcpyAsync(s->h_err + 1, s->errflag + 1, (size_t) s->p.nspecies * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
		CK(cudaStreamSynchronize(s->stream));
		any = false;
		for(int is = 0; is < s->p.nspecies; is++) { want[is] = s->h_err[1 + is]; any = any || want[is] > 0; }
	}
	if(!any) return 0;
	if(s->comm && comm_p2p(s->comm))
	{
		/* peer memory: every rank closes its mappings of the outboxes that are about to be replaced,
		 * and only when all have (one more reduction as a barrier) are they freed */
		for(int is = 0; is < s->p.nspecies; is++)
		{
			if(want[is] <= 0 || !s->sp[is].block || want[is] <= s->sp[is].d.cap) continue;
			if(comm_p2p_unmap(s->comm, comm_export_species(is, 0)) || comm_p2p_unmap(s->comm, comm_export_species(is, 1)))
				return CPIC_B200_ECUDA;
		}
		if(comm_allreduce_max(s->comm, s->errflag + 1, 1, s->stream)) return CPIC_B200_ECUDA;
		CK(cudaStreamSynchronize(s->stream));
	}
	for(int is = 0; is < s->p.nspecies; is++)
	{
		if(want[is] <= 0 || !s->sp[is].block || want[is] <= s->sp[is].d.cap) continue;
		int rc = regrow(s, is, want[is]);
		if(rc) return rc;
	}
	if(s->comm && s->p2p_stale) return p2p_attach(s);
	return 0;
}

extern "C" int64_t
cpic_b200_capacity(cpic_b200_sim_t *s, int is)
{
	if(!s || is < 0 || is >= s->p.nspecies) return -1;
	return s->sp[is].block ? s->sp[is].d.cap : 0;
}

extern "C" int
cpic_b200_reserve(cpic_b200_sim_t *s, int is, int64_t capacity)
{
	if(!s || is < 0 || is >= s->p.nspecies || capacity < 0 || capacity > (1 << 30))
		return fail(CPIC_B200_EINVAL, "bad species or capacity");
	s->sp[is].reserve = (int) ((capacity + 31) / 32 * 32);
	return 0;
}

/* Fold the pending arrivals of one species into the block segments (see k_absorb) */
static int
absorb(sim_t_ *s, int is)
{
	SpeciesHost &h = s->sp[is];
	if(!h.block) return 0;
	/* k_absorb leaves a block alone when its arrivals would not fit and says so in its own flag word
	 * (errflag[15]); a larger segment then takes them (regrow merges them on the way). With several
	 * ranks the capacity is a collective decision, so it is an error here. */
	int *flag = s->errflag + 15;
	CK(cudaMemsetAsync(flag, 0, sizeof(int), s->stream));
	k_absorb<<<(s->nb + 7) / 8, 256, 0, s->stream>>>(h.d, s->g, s->nb, h.arr, flag, 0, s->nb);
	{
		cudaError_t e = cudaGetLastError();
		if(e != cudaSuccess) return fail(CPIC_B200_ECUDA, "kernel launch failed: %s", cudaGetErrorString(e));
	}
	CK(cudaMemcpyAsync(s->h_err + 15, flag, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaStreamSynchronize(s->stream));
	if(s->h_err[15])
	{
		int64_t o[6];
		int rc = occupancy(s, is, o);
		if(rc) return rc;
		if(s->comm)
			return fail(CPIC_B200_ECAPACITY, "species %d: a block holds %lld particles with its arrivals, capacity %lld: "
					"raise capacity_factor (now %g) or call cpic_b200_sync more often", is, (long long) o[0], (long long) o[1],
					s->p.capacity_factor);
		return regrow(s, is, (int) std::min<int64_t>(((std::max(o[0], o[1] + 1) * 3 / 2 + 31) / 32) * 32, 1 << 30));
	}
	return 0;
}

extern "C" int64_t
cpic_b200_num_particles(cpic_b200_sim_t *s, int is)
{
	if(!s || is < 0 || is >= s->p.nspecies) return -1;
	SpeciesHost &h = s->sp[is];
	if(!h.block) return 0;
	cudaSetDevice(s->device);
	if(absorb(s, is)) return -1;
	std::vector<int> cnt((size_t) s->nb);
	if(cudaMemcpyAsync(cnt.data(), h.d.count, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess) return -1;
	if(cudaStreamSynchronize(s->stream) != cudaSuccess) return -1;
	long long n = 0;
	for(int c : cnt) n += c;
	return n;
}

extern "C" int64_t
cpic_b200_get_particles(cpic_b200_sim_t *s, int is, int64_t capn, int64_t *id, double *x, double *y,
		double *ux, double *uy, double *uz, double *Ex, double *Ey)
{
	if(!s || is < 0 || is >= s->p.nspecies) { fail(CPIC_B200_EINVAL, "bad species"); return -1; }
	SpeciesHost &h = s->sp[is];
	if(!h.block) return 0;
	cudaSetDevice(s->device);
	if(absorb(s, is)) return -1;
	const int cap = h.d.cap;
	std::vector<int> cnt((size_t) s->nb);
	if(cudaMemcpyAsync(cnt.data(), h.d.count, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess
			|| cudaStreamSynchronize(s->stream) != cudaSuccess)
	{ fail(CPIC_B200_ECUDA, "copy of block counts failed"); return -1; }
	long long n = 0;
	for(int c : cnt) n += c;
	if(n > capn) return n;

	const size_t nslot = (size_t) s->nb * cap;
	/* the six segment arrays in one copy (they are one allocation, in either layout) */
	const size_t seg_doubles = (h.block_bytes - align256((size_t) s->nob * sizeof(int))) / sizeof(double);
	const size_t astep = ((char *) h.d.y - (char *) h.d.x) / sizeof(double);
	std::vector<double> seg(seg_doubles);
	if(cudaMemcpy(seg.data(), h.d.x, seg_doubles * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
	{ fail(CPIC_B200_ECUDA, "particle download failed"); return -1; }
	double *outs[6] = { x, y, ux, uy, uz, (double *) id };
	for(int a = 0; a < 6; a++)
	{
		if(!outs[a]) continue;
		size_t k = 0;
		for(int b = 0; b < s->nb; b++)
			for(int i = 0; i < cnt[(size_t) b]; i++)
				outs[a][k++] = seg[seg_slot(cap, b, i) + (size_t) a * astep];
	}
	std::vector<double> tmp(nslot);
	struct { const void *src; double *dst; } arrs[] = { { h.d.pEx, Ex }, { h.d.pEy, Ey } };
	for(auto &a : arrs)
	{
		if(!a.dst) continue;
		if(!a.src)
		{
			for(long long i = 0; i < n; i++) a.dst[i] = 0.0;
			continue;
		}
		if(cudaMemcpy(tmp.data(), a.src, nslot * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
		{ fail(CPIC_B200_ECUDA, "particle download failed"); return -1; }
		size_t k = 0;
		for(int b = 0; b < s->nb; b++)
		{
			memcpy(a.dst + k, tmp.data() + (size_t) b * cap, (size_t) cnt[(size_t) b] * sizeof(double));
			k += (size_t) cnt[(size_t) b];
		}
	}
	return n;
}

/* Host lists that keep their own order (the drop-in binding, dropin/cpic_b200_stages.c): `ids` are the
 * particle ids in the order the host walks its lists */
extern "C" int
cpic_b200_set_host_order(cpic_b200_sim_t *s, int is, int64_t n, const int64_t *ids)
{
	if(!s || is < 0 || is >= s->p.nspecies || n < 0 || (n && !ids)) return fail(CPIC_B200_EINVAL, "bad species or ids");
	CK(cudaSetDevice(s->device));
	SpeciesHost &h = s->sp[is];
	int64_t maxid = -1;
	for(int64_t k = 0; k < n; k++)
	{
		if(ids[k] < 0 || ids[k] >= (1LL << 31)) return fail(CPIC_B200_EINVAL, "particle id %lld out of range", (long long) ids[k]);
		maxid = std::max(maxid, ids[k]);
	}
	std::vector<int> pos((size_t) (maxid + 1), -1);
	for(int64_t k = 0; k < n; k++) pos[(size_t) ids[k]] = (int) k;
	cudaFree(h.hpos);
	h.hpos = NULL;
	h.hpos_n = maxid + 1;
	h.horder_n = n;
	if(maxid >= 0)
	{
		CK(cudaMalloc(&h.hpos, pos.size() * sizeof(int)));
		CK(cudaMemcpy(h.hpos, pos.data(), pos.size() * sizeof(int), cudaMemcpyHostToDevice));
	}
	return 0;
}

/* The particles of one species in the order given to cpic_b200_set_host_order: entry k of every array
 * belongs to ids[k]. Any pointer may be NULL; pinned arrays (cpic_b200_host_alloc) are filled by DMA. */
extern "C" int
cpic_b200_get_particles_ordered(cpic_b200_sim_t *s, int is, int64_t n, double *x, double *y,
		double *ux, double *uy, double *uz, double *Ex, double *Ey)
{
	if(!s || is < 0 || is >= s->p.nspecies) return fail(CPIC_B200_EINVAL, "bad species");
	CK(cudaSetDevice(s->device));
	SpeciesHost &h = s->sp[is];
	if(n != h.horder_n) return fail(CPIC_B200_EINVAL, "species %d: %lld particles asked for, the host order holds %lld", is,
			(long long) n, (long long) h.horder_n);
	if(n == 0 || !h.block) return 0;
	int rc = absorb(s, is);
	if(rc) return rc;
	if((rc = image_staging(s, (size_t) 7 * (size_t) n))) return rc;
	int *flag = s->errflag + 14;
	CK(cudaMemsetAsync(flag, 0, sizeof(int), s->stream));
	k_gather_ordered<<<(s->nb + 7) / 8, 256, 0, s->stream>>>(h.d, s->nb, h.hpos, h.hpos_n, s->img, n, flag);
	CK(cudaGetLastError());
	double *outs[7] = { x, y, ux, uy, uz, Ex, Ey };
	for(int a = 0; a < 7; a++)
		if(outs[a]) CK(cudaMemcpyAsync(outs[a], s->img + (size_t) a * (size_t) n, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaMemcpyAsync(s->h_err + 14, flag, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaStreamSynchronize(s->stream));
	if(s->h_err[14]) return fail(CPIC_B200_EINVAL, "species %d holds a particle whose id is not in the host order", is);
	return 0;
}

/* ------------------------------------------------------------------ timing */

struct StageTimer {
	sim_t_ *s; int which;
	StageTimer(sim_t_ *s_, int w) : s(s_), which(w)
	{
		if(s->timing) cudaEventRecord(s->ev[0], s->stream);
	}
	~StageTimer()
	{
		if(!s->timing) return;
		float ms = 0;
		cudaEventRecord(s->ev[1], s->stream);
		cudaEventSynchronize(s->ev[1]);
		cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]);
		s->ms[which] += ms;
	}
};

extern "C" int
cpic_b200_timing(cpic_b200_sim_t *s, int enable)
{
	if(!s) return fail(CPIC_B200_EINVAL, "null sim");
	s->timing = enable != 0;
	memset(s->ms, 0, sizeof(s->ms));
	s->launches = 0;
	return 0;
}

extern "C" int
cpic_b200_get_timing(cpic_b200_sim_t *s, double ms[6], int64_t launches[1])
{
	if(!s) return fail(CPIC_B200_EINVAL, "null sim");
	if(ms) for(int i = 0; i < T_COUNT; i++) ms[i] = s->ms[i];
	if(launches) launches[0] = s->launches;
	return 0;
}

/* ------------------------------------------------------------------ stages */

static int
check_launch(sim_t_ *s, int n = 1)
{
	s->launches += n;
	cudaError_t e = cudaGetLastError();
	if(e != cudaSuccess) return fail(CPIC_B200_ECUDA, "kernel launch failed: %s", cudaGetErrorString(e));
	return 0;
}

/* MFT_solve, reference src/solver.c:465-509 (one rank: cuFFT D2Z, G multiply, Z2D;
 * several ranks: comm.cu) */
static int
solve(sim_t_ *s)
{
	const Geom &g = s->g;
	if(s->comm) return comm_solve(s->comm, s->rho, s->phi_raw, s->stream, s->errflag, &s->launches);
	const size_t n = (size_t) g.ny * (g.nx / 2 + 1);
	CKFFT(cufftExecD2Z(s->plan_fwd, s->rho, s->gk));
	int blocks = (int) std::min<size_t>((n + 255) / 256, 148 * 16);
	k_green<<<blocks, 256, 0, s->stream>>>(s->gk, s->G, n);
	CKFFT(cufftExecZ2D(s->plan_inv, s->gk, s->phi_raw));
	return check_launch(s, 1);      /* own kernels only; cuFFT's are not counted */
}

extern "C" int
cpic_b200_solve(cpic_b200_sim_t *s)
{
	if(!s) return fail(CPIC_B200_EINVAL, "null sim");
	CK(cudaSetDevice(s->device));
	int rc = solve(s);
	if(rc) return rc;
	const Geom &g = s->g;
	dim3 grid((g.S + 127) / 128, g.ny + 3);
	k_phi_finish<<<grid, 128, 0, s->stream>>>(s->phi_raw, s->phi, g, (double) (int) ((long long) g.nx * g.ny_glob), 0);
	return check_launch(s);
}

/* stage_field_E, reference src/field.c:450-501: solve, phi halo, E = -grad phi */
extern "C" int
cpic_b200_stage_field_E(cpic_b200_sim_t *s)
{
	if(!s) return fail(CPIC_B200_EINVAL, "null sim");
	if(s->p.nranks > 1 && !s->comm) return fail(CPIC_B200_EINVAL, "rank %d of %d has no communicator: call cpic_b200_comm_init first", s->p.rank, s->p.nranks);
	CK(cudaSetDevice(s->device));
	const Geom &g = s->g;
	int rc;
	if((rc = fields_guard(s))) return rc;
	{
		StageTimer ts(s, T_SOLVER);
		rc = solve(s);
		if(rc) return rc;
	}
	StageTimer t(s, T_FIELD_E);
	dim3 grid((g.S + 127) / 128, g.ny + 3);
	/* N is an int in the reference (src/solver.c:366, :494) */
	const double N = (double) (int) ((long long) g.nx * g.ny_glob);
	if(!s->comm)
	{
		/* one rank: normalisation, ghost rows and the differences in one pass */
		dim3 gridPE((std::max(g.S, g.SE) + 127) / 128, g.ny + 3);
		k_phi_E<<<gridPE, 128, 0, s->stream>>>(s->phi_raw, s->phi, s->Ex, s->Ey, g, N);
		return check_launch(s);
	}
	k_phi_finish<<<grid, 128, 0, s->stream>>>(s->phi_raw, s->phi, g, N, 0);
	if((rc = check_launch(s))) return rc;
	if((rc = comm_phi_halo(s->comm, s->phi, s->stream, s->errflag, &s->launches))) return rc;
	dim3 gridE((g.SE + 127) / 128, g.ny + 1);
	k_field_E<<<gridE, 128, 0, s->stream>>>(s->phi, s->Ex, s->Ey, g);
	return check_launch(s);
}

static PushParams
push_params(const sim_t_ *s, int is)
{
	PushParams pp;
	const double q = s->sp[is].q, m = s->sp[is].m;
	/* reference src/mover.c:204-225: iteration 0 rewinds the velocities half a step */
	if(s->iter == 0) { pp.dt = -s->p.dt / 2; pp.set_r = 0; }
	else { pp.dt = s->p.dt; pp.set_r = 1; }
	pp.dtqm2 = 0.5 * pp.dt * q / m;
	const double t[3] = { s->p.B[0] * pp.dtqm2, s->p.B[1] * pp.dtqm2, s->p.B[2] * pp.dtqm2 };
	pp.tx = t[0]; pp.ty = t[1]; pp.tz = t[2];
	pp.sx = 2.0 * t[0] / fma(t[0], t[0], 1.0);
	pp.sy = 2.0 * t[1] / fma(t[1], t[1], 1.0);
	pp.sz = 2.0 * t[2] / fma(t[2], t[2], 1.0);
	pp.umax_x = s->umax[0]; pp.umax_y = s->umax[1]; pp.umax_z = s->umax[2];
	return pp;
}

/* row0, rows >= 0: only the block rows [row0, row0 + rows) (a band); the caller flips h.arr once all
 * bands of the step are out */
template <int MODE>
static int
launch_gather_push(sim_t_ *s, int is, cudaStream_t stream, int row0 = -1, int rows = 0)
{
	SpeciesHost &h = s->sp[is];
	if(!h.block) return 0;
	const Geom &g = s->g;
	const size_t smem = push_smem_bytes<MODE>(s);
	const int ncx = g.nbx / g.WPC;
	const int ctas = row0 < 0 ? s->nb / g.WPC : rows * ncx;
	/* a push reads the pending arrivals and fills the other outbox */
	const int cur = MODE == 0 ? h.arr : h.arr ^ 1;
	k_gather_push<MODE><<<ctas, 32 * g.WPC, smem, stream>>>(h.d, g, push_params(s, is),
			s->mapEx, s->mapEy, h.segmap[0], h.segmap[1], s->nb, cur, row0 < 0 ? 0 : row0 * ncx, s->errflag);
	if(MODE != 0 && row0 < 0) h.arr = cur;
	return check_launch(s);
}

/* stage_plasma_E, reference src/particle.c:232-248 */
extern "C" int
cpic_b200_stage_plasma_E(cpic_b200_sim_t *s)
{
	if(!s) return fail(CPIC_B200_EINVAL, "null sim");
	if(s->p.nranks > 1 && !s->comm) return fail(CPIC_B200_EINVAL, "rank %d of %d has no communicator: call cpic_b200_comm_init first", s->p.rank, s->p.nranks);
	CK(cudaSetDevice(s->device));
	StageTimer t(s, T_GATHER);
	for(int is = 0; is < s->p.nspecies; is++)
	{
		int rc = ensure_particle_E(s, is);
		if(!rc) rc = launch_gather_push<0>(s, is, s->stream);
		if(rc) return rc;
	}
	return 0;
}

/* comm_plasma, reference src/comm_plasma.c:1122-1142. Inside one rank the exchange is
 * part of the push kernel (leavers go straight to the outbox region their new block
 * reads); what is left is the Y pass between ranks (src/comm_plasma.c:1086-1120): the
 * regions of the edge block rows that point across the slab face travel to the
 * neighbour's ghost outbox rows. */
static int
exchange(sim_t_ *s)
{
	StageTimer t(s, T_EXCHANGE);
	static_assert(SET_MAX_SPECIES >= CPIC_B200_MAX_SPECIES, "species set too small");
	SpeciesDev *sps[CPIC_B200_MAX_SPECIES];
	int arrs[CPIC_B200_MAX_SPECIES], nsp = 0;
	SpeciesSet set;
	for(int is = 0; is < s->p.nspecies; is++)
	{
		SpeciesHost &h = s->sp[is];
		if(!h.block) continue;
		set.sp[nsp] = h.d;
		set.arr[nsp] = h.arr;
		sps[nsp] = &h.d;
		arrs[nsp++] = h.arr;
	}
	set.n = nsp;
	if(!nsp) return 0;
	/* particles that jumped further than a neighbouring block (normally none): one CTA per species */
	k_far_insert<<<nsp, 1024, 0, s->stream>>>(set, s->g, s->errflag);
	int rc = check_launch(s);
	if(rc) return rc;
	if(s->comm)
	{
		rc = comm_particles(s->comm, sps, arrs, nsp, s->g, s->nb, s->stream, s->errflag, &s->launches);
		if(rc) return rc;
		/* far movers received from the neighbour ranks */
		k_far_insert<<<nsp, 1024, 0, s->stream>>>(set, s->g, s->errflag);
		if((rc = check_launch(s))) return rc;
	}
	return 0;
}

/* stage_plasma_r, reference src/mover.c:331-362: plasma_mover then comm_plasma */
template <int MODE>
static int
stage_plasma_r(sim_t_ *s)
{
	if(MODE == 1)
		for(int is = 0; is < s->p.nspecies; is++) { int rc = ensure_particle_E(s, is); if(rc) return rc; }
	/* several ranks over peer memory: the neighbours must know where this rank's outboxes are */
	if(s->comm && s->p2p_stale) { int rc = p2p_attach(s); if(rc) return rc; }
	{
		StageTimer t(s, T_PUSH);
		/* species are independent until the exchange: odd ones go to the second stream */
		const bool fork = s->overlap_species && s->p.nspecies > 1;
		if(fork)
		{
			CK(cudaEventRecord(s->ev_fork, s->stream));
			CK(cudaStreamWaitEvent(s->stream2, s->ev_fork, 0));
		}
		for(int is = 0; is < s->p.nspecies; is++)
		{
			int rc = launch_gather_push<MODE>(s, is, (fork && (is & 1)) ? s->stream2 : s->stream);
			if(rc) return rc;
		}
		if(fork)
		{
			CK(cudaEventRecord(s->ev_join, s->stream2));
			CK(cudaStreamWaitEvent(s->stream, s->ev_join, 0));
		}
	}
	return exchange(s);
}

extern "C" int
cpic_b200_stage_plasma_r(cpic_b200_sim_t *s)
{
	if(!s) return fail(CPIC_B200_EINVAL, "null sim");
	if(s->p.nranks > 1 && !s->comm) return fail(CPIC_B200_EINVAL, "rank %d of %d has no communicator: call cpic_b200_comm_init first", s->p.rank, s->p.nranks);
	CK(cudaSetDevice(s->device));
	return stage_plasma_r<1>(s);
}

/* stage_field_rho, reference src/field.c:268-356 */
extern "C" int
cpic_b200_stage_field_rho(cpic_b200_sim_t *s)
{
	if(!s) return fail(CPIC_B200_EINVAL, "null sim");
	if(s->p.nranks > 1 && !s->comm) return fail(CPIC_B200_EINVAL, "rank %d of %d has no communicator: call cpic_b200_comm_init first", s->p.rank, s->p.nranks);
	CK(cudaSetDevice(s->device));
	{ int rc = fields_guard(s); if(rc) return rc; }
	StageTimer t(s, T_RHO);
	const Geom &g = s->g;
	static_assert(DEP_MAX_SPECIES >= CPIC_B200_MAX_SPECIES, "deposit set too small");
	DepositSet set;
	set.n = 0;
	for(int is = 0; is < s->p.nspecies; is++)
	{
		SpeciesHost &h = s->sp[is];
		if(!h.block) continue;
		DepositSpecies &d = set.s[set.n++];
		const Outbox &in = h.d.ob[h.arr];
		d.x = h.d.x; d.y = h.d.y; d.count = h.d.count;
		d.arec = in.rec; d.acount = in.count;
		d.vq = -h.q / s->p.e0;       /* reference src/interpolate.c:307 */
		d.cap = h.d.cap; d.nob = h.d.nob;
		memcpy(d.roff, h.d.roff, sizeof(d.roff));
		memcpy(d.rcap, h.d.rcap, sizeof(d.rcap));
	}
	if(!set.n)
	{
		/* no particles at all: rho_reset only */
		CK(cudaMemsetAsync(s->rho, 0, (size_t) (g.ny + 1) * g.S * sizeof(double), s->stream));
	}
	else
	{
		/* a persistent grid: as many CTAs as fit the device at once, blocks handed out round-robin */
		const int ctas = std::min((s->nb + DEP_WARPS - 1) / DEP_WARPS, s->dep_ctas);
		if(DEP_FUSED)
		{
			/* every species in one pass over the blocks */
			if(s->dep_cols == 16) k_deposit<true, 16><<<ctas, 32 * DEP_WARPS, s->smem_dep, s->stream>>>(set, g, s->nb, 16, s->tiles);
			else k_deposit<true, 0><<<ctas, 32 * DEP_WARPS, s->smem_dep, s->stream>>>(set, g, s->nb, s->dep_cols, s->tiles);
			int rc = check_launch(s);
			if(rc) return rc;
		}
		else
			for(int k = 0; k < set.n; k++)
			{
				DepositSet one;
				one.s[0] = set.s[k];
				one.n = 1;
				if(k == 0) k_deposit<true, 0><<<ctas, 32 * DEP_WARPS, s->smem_dep, s->stream>>>(one, g, s->nb, s->dep_cols, s->tiles);
				else k_deposit<false, 0><<<ctas, 32 * DEP_WARPS, s->smem_dep, s->stream>>>(one, g, s->nb, s->dep_cols, s->tiles);
				int rc = check_launch(s);
				if(rc) return rc;
			}
		/* one rank: the ghost row goes to ourselves (comm_send_ghost_rho / comm_recv_ghost_rho, reference
		 * src/comm_field.c:51-136) and is folded into row 0 by the same pass */
		k_rho_assemble<<<dim3((g.nx + 127) / 128, g.ny + 1), 128, 0, s->stream>>>(s->tiles, s->rho, g, s->comm ? 0 : 1);
		int rc = check_launch(s);
		if(rc) return rc;
	}
	if(s->comm)
	{
		int rc = comm_rho_halo(s->comm, s->rho, s->stream, s->errflag, &s->launches);
		if(rc) return rc;
	}
	else if(!set.n)
	{
		k_rho_fold<<<(g.nx + 127) / 128, 128, 0, s->stream>>>(s->rho, s->rho + (size_t) g.ny * g.S, g);
		int rc = check_launch(s);
		if(rc) return rc;
	}
	return 0;
}

/* sim_pre_step, reference src/sim.c:208-236 */
extern "C" int
cpic_b200_pre_step(cpic_b200_sim_t *s)
{
	if(!s) return fail(CPIC_B200_EINVAL, "null sim");
	s->iter = -1;
	int rc = cpic_b200_stage_field_rho(s);
	if(!rc) rc = cpic_b200_stage_field_E(s);
	s->iter = 0;
	return rc;
}

/* sim_step, reference src/sim.c:481-581, gather and push fused */
extern "C" int
cpic_b200_step(cpic_b200_sim_t *s)
{
	if(!s) return fail(CPIC_B200_EINVAL, "null sim");
	if(s->iter < 0) return fail(CPIC_B200_EINVAL, "call cpic_b200_pre_step first");
	CK(cudaSetDevice(s->device));
	int rc = cpic_b200_stage_field_E(s);
	if(!rc) rc = stage_plasma_r<2>(s);
	if(!rc) rc = cpic_b200_stage_field_rho(s);
	if(rc) return rc;
	s->iter++;
	return 0;
}

extern "C" int
cpic_b200_run(cpic_b200_sim_t *s, int64_t steps)
{
	for(int64_t i = 0; i < steps; i++)
	{
		int rc = cpic_b200_step(s);
		if(!rc && (i & 31) == 31) rc = check_capacity(s);
		if(rc) return rc;
	}
	return cpic_b200_sync(s);
}

/* `steps` sim_steps bracketed by CUDA events on the simulation's stream */
extern "C" int
cpic_b200_run_timed(cpic_b200_sim_t *s, int64_t steps, double *ms)
{
	if(!s || !ms) return fail(CPIC_B200_EINVAL, "null argument");
	CK(cudaSetDevice(s->device));
	cudaEvent_t a, b;
	CK(cudaEventCreate(&a));
	CK(cudaEventCreate(&b));
	CK(cudaStreamSynchronize(s->stream));
	CK(cudaEventRecord(a, s->stream));
	int rc = 0;
	for(int64_t i = 0; i < steps && !rc; i++)
	{
		rc = cpic_b200_step(s);
		if(!rc && (i & 31) == 31) rc = check_capacity(s);
	}
	CK(cudaEventRecord(b, s->stream));
	CK(cudaEventSynchronize(b));
	float t = 0;
	CK(cudaEventElapsedTime(&t, a, b));
	cudaEventDestroy(a);
	cudaEventDestroy(b);
	*ms = t;
	if(rc) return rc;
	return cpic_b200_sync(s);
}

extern "C" int64_t cpic_b200_iter(cpic_b200_sim_t *s) { return s ? s->iter : -1; }

extern "C" int
cpic_b200_set_iter(cpic_b200_sim_t *s, int64_t iter)
{
	if(!s) return fail(CPIC_B200_EINVAL, "null sim");
	s->iter = iter;
	return 0;
}

extern "C" int
cpic_b200_sync(cpic_b200_sim_t *s)
{
	if(!s) return fail(CPIC_B200_EINVAL, "null sim");
	CK(cudaSetDevice(s->device));
	/* with several ranks an error anywhere is an error everywhere (no rank is left waiting) */
	if(s->comm && comm_allreduce_max(s->comm, s->errflag, 1, s->stream)) return CPIC_B200_ECUDA;
	CK(cudaMemcpyAsync(s->h_err, s->errflag, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaStreamSynchronize(s->stream));
	const int e = *s->h_err;
	if(!e) return check_capacity(s);
	CK(cudaMemsetAsync(s->errflag, 0, sizeof(int), s->stream));
	if(e & ERRBIT_TMA) return fail(CPIC_B200_ECUDA, "a TMA tile load did not complete (tensor map rejected)");
	if(e & ERRBIT_PEER) return fail(CPIC_B200_ECUDA, "a neighbour rank did not arrive at an exchange (peer-memory handshake timed out)");
	if(e & ERRBIT_VELOCITY) return fail(CPIC_B200_EVELOCITY, "Max velocity exceeded (umax = %g %g %g)", s->umax[0], s->umax[1], s->umax[2]);
	if(e & ERRBIT_FAR) return fail(CPIC_B200_EFAR, "a particle crossed a slab face by more than one particle block row (%d cells) in one step", s->g.BY);
	return fail(CPIC_B200_ECAPACITY, "capacity exceeded (%s%s%s%s): raise capacity_factor (now %g) / outbox_fraction",
			(e & ERRBIT_CAPACITY) ? "particle block segment " : "", (e & ERRBIT_REGION) ? "exchange region " : "",
			(e & ERRBIT_FARLIST) ? "far-mover list " : "", (e & ERRBIT_ABSORB) ? "arrivals do not fit " : "",
			s->p.capacity_factor);
}

/* ------------------------------------------------------------------ fields */

static int
field_geom(const sim_t_ *s, int f, double **base, int64_t *rows, int64_t *stride, int64_t *dstride)
{
	const Geom &g = s->g;
	switch(f)
	{
		case CPIC_B200_RHO: *base = s->rho; *rows = g.ny + 1; *stride = g.S; *dstride = g.S; return 0;
		case CPIC_B200_PHI: *base = s->phi; *rows = g.ny + 3; *stride = g.S; *dstride = g.S; return 0;
		case CPIC_B200_EX: *base = s->Ex; *rows = g.ny + 1; *stride = g.nx; *dstride = g.SE; return 0;
		case CPIC_B200_EY: *base = s->Ey; *rows = g.ny + 1; *stride = g.nx; *dstride = g.SE; return 0;
	}
	return fail(CPIC_B200_EINVAL, "unknown field %d", f);
}

extern "C" int
cpic_b200_field_shape(cpic_b200_sim_t *s, int f, int64_t *rows, int64_t *stride)
{
	double *base = NULL; int64_t r = 0, st = 0, ds = 0;
	if(!s) return fail(CPIC_B200_EINVAL, "null sim");
	int rc = field_geom(s, f, &base, &r, &st, &ds);
	if(rc) return rc;
	if(rows) *rows = r;
	if(stride) *stride = st;
	return 0;
}

extern "C" int
cpic_b200_get_field(cpic_b200_sim_t *s, int f, double *host)
{
	double *base = NULL; int64_t r = 0, st = 0, ds = 0;
	if(!s || !host) return fail(CPIC_B200_EINVAL, "null argument");
	CK(cudaSetDevice(s->device));
	int rc = field_geom(s, f, &base, &r, &st, &ds);
	if(rc) return rc;
	CK(cudaMemcpy2DAsync(host, (size_t) st * sizeof(double), base, (size_t) ds * sizeof(double),
				(size_t) st * sizeof(double), (size_t) r, cudaMemcpyDeviceToHost, s->stream));
	CK(cudaStreamSynchronize(s->stream));
	return 0;
}

/* output_fields without stalling the step (reference src/output.c:594-635 writes the grids right
 * after stage_field_E, src/sim.c:507): the four arrays start their way to (pinned) host buffers on a
 * copy stream of their own, behind everything issued so far; the stages that overwrite a grid
 * (stage_field_rho: rho, the next stage_field_E: phi and E) wait for the copy on the device, the host
 * does not wait until cpic_b200_get_fields_end. host[f] == NULL skips grid f. */
extern "C" int
cpic_b200_get_fields_begin(cpic_b200_sim_t *s, double *const host[4])
{
	if(!s || !host) return fail(CPIC_B200_EINVAL, "null argument");
	CK(cudaSetDevice(s->device));
	if(!s->stream_fields)
	{
		CK(cudaStreamCreateWithFlags(&s->stream_fields, cudaStreamNonBlocking));
		CK(cudaEventCreateWithFlags(&s->ev_fields_from, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&s->ev_fields_done, cudaEventDisableTiming));
	}
	CK(cudaEventRecord(s->ev_fields_from, s->stream));
	CK(cudaStreamWaitEvent(s->stream_fields, s->ev_fields_from, 0));
	for(int f = 0; f < 4; f++)
	{
		if(!host[f]) continue;
		double *base = NULL; int64_t r = 0, st = 0, ds = 0;
		int rc = field_geom(s, f, &base, &r, &st, &ds);
		if(rc) return rc;
		CK(cudaMemcpy2DAsync(host[f], (size_t) st * sizeof(double), base, (size_t) ds * sizeof(double),
					(size_t) st * sizeof(double), (size_t) r, cudaMemcpyDeviceToHost, s->stream_fields));
	}
	CK(cudaEventRecord(s->ev_fields_done, s->stream_fields));
	s->fields_pending = true;
	return 0;
}

extern "C" int
cpic_b200_get_fields_end(cpic_b200_sim_t *s)
{
	if(!s) return fail(CPIC_B200_EINVAL, "null sim");
	if(!s->stream_fields) return 0;
	CK(cudaSetDevice(s->device));
	CK(cudaEventSynchronize(s->ev_fields_done));
	return 0;
}

/* the grids are about to be overwritten: a download in flight goes first (on the device) */
static int
fields_guard(sim_t_ *s)
{
	if(!s->fields_pending) return 0;
	CK(cudaStreamWaitEvent(s->stream, s->ev_fields_done, 0));
	s->fields_pending = false;
	return 0;
}

extern "C" int
cpic_b200_set_field(cpic_b200_sim_t *s, int f, const double *host)
{
	double *base = NULL; int64_t r = 0, st = 0, ds = 0;
	if(!s || !host) return fail(CPIC_B200_EINVAL, "null argument");
	CK(cudaSetDevice(s->device));
	int rc = field_geom(s, f, &base, &r, &st, &ds);
	if(rc) return rc;
	CK(cudaMemcpy2DAsync(base, (size_t) ds * sizeof(double), host, (size_t) st * sizeof(double),
				(size_t) st * sizeof(double), (size_t) r, cudaMemcpyHostToDevice, s->stream));
	if(f == CPIC_B200_EX || f == CPIC_B200_EY)
	{
		/* wrap columns [nx, SE) repeat columns [0, SE-nx) */
		const Geom &g = s->g;
		for(int c = g.nx; c < g.SE; c++)
			CK(cudaMemcpy2DAsync(base + c, (size_t) ds * sizeof(double), host + (c - g.nx) % g.nx,
						(size_t) st * sizeof(double), sizeof(double), (size_t) r, cudaMemcpyHostToDevice, s->stream));
	}
	CK(cudaStreamSynchronize(s->stream));
	return 0;
}

/* conservation_energy, reference src/sim.c:332-405 (compiled out there):
 * KE = sum_s m_s/2 sum(ux^2+uy^2), PE = sum rho*phi over the slab */
extern "C" int
cpic_b200_energy(cpic_b200_sim_t *s, double *kinetic, double *potential)
{
	if(!s) return fail(CPIC_B200_EINVAL, "null sim");
	CK(cudaSetDevice(s->device));
	const Geom &g = s->g;
	double ke = 0.0, pe = 0.0, v;
	double *res = s->red + std::max(s->nb, g.ny);
	for(int is = 0; is < s->p.nspecies; is++)
	{
		SpeciesHost &h = s->sp[is];
		if(!h.block) continue;
		if(absorb(s, is)) return CPIC_B200_ECUDA;
		k_kinetic<<<(s->nb + 7) / 8, 256, 0, s->stream>>>(h.d, s->nb, s->red);
		k_sum<<<1, 1024, 0, s->stream>>>(s->red, s->nb, res);
		CK(cudaMemcpyAsync(&v, res, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
		CK(cudaStreamSynchronize(s->stream));
		ke += v * h.m / 2.0;
	}
	k_potential<<<g.ny, 256, 0, s->stream>>>(s->rho, s->phi, g, s->red);
	k_sum<<<1, 1024, 0, s->stream>>>(s->red, g.ny, res);
	CK(cudaMemcpyAsync(&v, res, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaStreamSynchronize(s->stream));
	pe = v;
	if(kinetic) *kinetic = ke;
	if(potential) *potential = pe;
	return 0;
}

/* -------------------------------------------------- raw images (bench e2e) */

/* Image layout (host): per species { int64 n; int32 count[nb] (padded to 8 bytes);
 * double x[n], y[n], ux[n], uy[n], uz[n]; int64 id[n] }. Sized for the particle numbers
 * at the time of the call plus 12.5 % (the population of a rank changes with several ranks). */
static int64_t
image_species_bytes(const sim_t_ *s, int64_t n)
{
	return 8 + (((int64_t) s->nb * 4 + 7) & ~7LL) + 6 * 8 * n;
}

extern "C" int64_t
cpic_b200_image_bytes(cpic_b200_sim_t *s)
{
	if(!s) return -1;
	int64_t bytes = 0;
	for(int is = 0; is < s->p.nspecies; is++)
	{
		if(!s->sp[is].block) continue;
		int64_t n = cpic_b200_num_particles(s, is);
		if(n < 0) return -1;
		bytes += image_species_bytes(s, n + n / 8 + 1024);
	}
	return bytes;
}

static int
image_staging(sim_t_ *s, size_t doubles)
{
	if(doubles <= s->img_cap) return 0;
	cudaFree(s->img);
	cudaFree(s->img_off);
	s->img = NULL; s->img_off = NULL; s->img_cap = 0;
	CK(cudaMalloc(&s->img, doubles * sizeof(double)));
	CK(cudaMalloc(&s->img_off, ((size_t) s->nb + 1) * (sizeof(long long) + sizeof(int))));
	s->img_cap = doubles;
	return 0;
}

extern "C" int
cpic_b200_image_download(cpic_b200_sim_t *s, void *host, int64_t bytes)
{
	if(!s || !host) return fail(CPIC_B200_EINVAL, "null argument");
	CK(cudaSetDevice(s->device));
	char *p = (char *) host, *end = p + bytes;
	std::vector<int> cnt((size_t) s->nb);
	std::vector<long long> off((size_t) s->nb);
	for(int is = 0; is < s->p.nspecies; is++)
	{
		SpeciesHost &h = s->sp[is];
		if(!h.block) continue;
		int rc = absorb(s, is);
		if(rc) return rc;
		CK(cudaMemcpyAsync(cnt.data(), h.d.count, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
		CK(cudaStreamSynchronize(s->stream));
		long long n = 0;
		for(int b = 0; b < s->nb; b++) { off[(size_t) b] = n; n += cnt[(size_t) b]; }
		if(p + image_species_bytes(s, n) > end) return fail(CPIC_B200_EINVAL, "image buffer too small");
		if((rc = image_staging(s, (size_t) 6 * n))) return rc;
		CK(cudaMemcpyAsync(s->img_off, off.data(), off.size() * sizeof(long long), cudaMemcpyHostToDevice, s->stream));
		k_image_copy<<<(s->nb + 7) / 8, 256, 0, s->stream>>>(h.d, s->nb, h.d.count, s->img_off, s->img, n, 1);
		CK(cudaGetLastError());
		*(int64_t *) p = n; p += 8;
		memcpy(p, cnt.data(), cnt.size() * sizeof(int)); p += ((int64_t) s->nb * 4 + 7) & ~7LL;
		CK(cudaMemcpyAsync(p, s->img, (size_t) 6 * n * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
		p += 6 * 8 * n;
		CK(cudaStreamSynchronize(s->stream));
	}
	return 0;
}

extern "C" int
cpic_b200_image_upload(cpic_b200_sim_t *s, const void *host, int64_t bytes)
{
	if(!s || !host) return fail(CPIC_B200_EINVAL, "null argument");
	CK(cudaSetDevice(s->device));
	const char *p = (const char *) host, *end = p + bytes;
	std::vector<long long> off((size_t) s->nb);
	for(int is = 0; is < s->p.nspecies; is++)
	{
		SpeciesHost &h = s->sp[is];
		if(!h.block) continue;
		if(p + 8 > end) return fail(CPIC_B200_EINVAL, "truncated image");
		const long long n = *(const int64_t *) p; p += 8;
		const int *cnt = (const int *) p; p += ((int64_t) s->nb * 4 + 7) & ~7LL;
		if(n < 0 || p + 6 * 8 * n > end) return fail(CPIC_B200_EINVAL, "truncated image");
		long long m = 0;
		for(int b = 0; b < s->nb; b++)
		{
			if(cnt[b] < 0 || cnt[b] > h.d.cap) return fail(CPIC_B200_ECAPACITY, "image block %d holds %d particles, capacity %d", b, cnt[b], h.d.cap);
			off[(size_t) b] = m; m += cnt[b];
		}
		if(m != n) return fail(CPIC_B200_EINVAL, "image counts do not add up");
		int rc = image_staging(s, (size_t) 6 * n);
		if(rc) return rc;
		int *dcnt = (int *) (s->img_off + s->nb + 1);
		/* the image holds every particle: no arrivals are pending afterwards */
		for(int k = 0; k < 2; k++) CK(cudaMemsetAsync(h.d.ob[k].count, 0, (size_t) s->nob * 9 * sizeof(int), s->stream));
		CK(cudaMemcpyAsync(s->img_off, off.data(), off.size() * sizeof(long long), cudaMemcpyHostToDevice, s->stream));
		CK(cudaMemcpyAsync(dcnt, cnt, (size_t) s->nb * sizeof(int), cudaMemcpyHostToDevice, s->stream));
		CK(cudaMemcpyAsync(s->img, p, (size_t) 6 * n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
		k_image_copy<<<(s->nb + 7) / 8, 256, 0, s->stream>>>(h.d, s->nb, dcnt, s->img_off, s->img, n, 0);
		CK(cudaGetLastError());
		/* `off` is reused by the next species: the copies above must have left the host */
		CK(cudaStreamSynchronize(s->stream));
		p += 6 * 8 * n;
	}
	return 0;
}

/* One sim_step with the particle state living in pinned HOST memory (the image of
 * cpic_b200_image_download): every species is uploaded, pushed and downloaded in turn on three
 * streams, so that the upload of one species runs while the previous one is pushed and downloaded
 * -- PCIe is full duplex, the serial upload / step / download uses one direction at a time. The
 * image comes back with the new block counts and particles. One rank (the population of a rank
 * changes with several); the capacities must hold the step (no growth in flight). */
extern "C" int
cpic_b200_step_host(cpic_b200_sim_t *s, void *host, int64_t bytes)
{
	if(!s || !host) return fail(CPIC_B200_EINVAL, "null argument");
	if(s->iter < 0) return fail(CPIC_B200_EINVAL, "call cpic_b200_pre_step first");
	if(s->comm) return fail(CPIC_B200_EINVAL, "cpic_b200_step_host runs on one rank");
	CK(cudaSetDevice(s->device));
	if(!s->stream_in)
	{
		CK(cudaStreamCreateWithFlags(&s->stream_in, cudaStreamNonBlocking));
		CK(cudaStreamCreateWithFlags(&s->stream_out, cudaStreamNonBlocking));
		for(int i = 0; i < CPIC_B200_MAX_SPECIES; i++)
		{
			CK(cudaEventCreateWithFlags(&s->ev_up[i], cudaEventDisableTiming));
			CK(cudaEventCreateWithFlags(&s->ev_packed[i], cudaEventDisableTiming));
		}
	}
	char *p = (char *) host, *end = p + bytes;
	const size_t cbytes = ((size_t) s->nb * 4 + 7) & ~(size_t) 7;
	struct Sub { long long n; int *cnt; double *data; } sub[CPIC_B200_MAX_SPECIES];
	std::vector<long long> off[CPIC_B200_MAX_SPECIES];
	/* whatever ran before (a download that filled the image) is complete */
	CK(cudaStreamSynchronize(s->stream));
	for(int is = 0; is < s->p.nspecies; is++)
	{
		SpeciesHost &h = s->sp[is];
		sub[is].n = -1;
		if(!h.block) continue;
		if(p + 8 > end) return fail(CPIC_B200_EINVAL, "truncated image");
		Sub &u = sub[is];
		u.n = *(const int64_t *) p; p += 8;
		u.cnt = (int *) p; p += cbytes;
		u.data = (double *) p;
		if(u.n < 0 || p + 6 * 8 * u.n > end) return fail(CPIC_B200_EINVAL, "truncated image");
		p += 6 * 8 * u.n;
		off[is].resize((size_t) s->nb + 1);
		long long m = 0;
		for(int b = 0; b < s->nb; b++)
		{
			if(u.cnt[b] < 0 || u.cnt[b] > h.d.cap) return fail(CPIC_B200_ECAPACITY, "image block %d holds %d particles, capacity %d", b, u.cnt[b], h.d.cap);
			off[is][(size_t) b] = m; m += u.cnt[b];
		}
		off[is][(size_t) s->nb] = m;
		if(m != u.n) return fail(CPIC_B200_EINVAL, "image counts do not add up");
		for(int k = 0; k < 2; k++)
		{
			if(s->hstage_cap[k][is] < (size_t) 6 * (size_t) u.n)
			{
				cudaFree(s->hstage[k][is]);
				s->hstage[k][is] = NULL; s->hstage_cap[k][is] = 0;
				CK(cudaMalloc(&s->hstage[k][is], (size_t) 6 * (size_t) u.n * sizeof(double)));
				s->hstage_cap[k][is] = (size_t) 6 * (size_t) u.n;
			}
			if(!s->hoff[k][is]) CK(cudaMalloc(&s->hoff[k][is], ((size_t) s->nb + 64 + 1) * sizeof(long long)));
		}
		if(!s->hcnt[is]) CK(cudaMalloc(&s->hcnt[is], (size_t) s->nb * sizeof(int)));
		/* upload: counts, offsets, the six arrays */
		CK(cudaMemcpyAsync(s->hcnt[is], u.cnt, (size_t) s->nb * sizeof(int), cudaMemcpyHostToDevice, s->stream_in));
		CK(cudaMemcpyAsync(s->hoff[0][is], off[is].data(), ((size_t) s->nb + 1) * sizeof(long long), cudaMemcpyHostToDevice, s->stream_in));
		CK(cudaMemcpyAsync(s->hstage[0][is], u.data, (size_t) 6 * (size_t) u.n * sizeof(double), cudaMemcpyHostToDevice, s->stream_in));
		CK(cudaEventRecord(s->ev_up[is], s->stream_in));
	}
	/* the fields of this step do not depend on the upload */
	int rc = cpic_b200_stage_field_E(s);
	if(rc) return rc;
	int *flag = s->errflag + 15;
	CK(cudaMemsetAsync(flag, 0, sizeof(int), s->stream));
	for(int is = 0; is < s->p.nspecies; is++)
	{
		SpeciesHost &h = s->sp[is];
		const Sub &u = sub[is];
		if(u.n < 0) continue;
		CK(cudaStreamWaitEvent(s->stream, s->ev_up[is], 0));
		/* the image holds every particle: no arrivals are pending afterwards */
		for(int k = 0; k < 2; k++) CK(cudaMemsetAsync(h.d.ob[k].count, 0, (size_t) s->nob * 9 * sizeof(int), s->stream));
		k_image_copy<<<(s->nb + 7) / 8, 256, 0, s->stream>>>(h.d, s->nb, s->hcnt[is], s->hoff[0][is], s->hstage[0][is], u.n, 0);
		if((rc = check_launch(s))) return rc;
		if((rc = launch_gather_push<2>(s, is, s->stream))) return rc;
		SpeciesSet set;
		set.sp[0] = h.d; set.arr[0] = h.arr; set.n = 1;
		k_far_insert<<<1, 1024, 0, s->stream>>>(set, s->g, s->errflag);
		k_absorb<<<(s->nb + 7) / 8, 256, 0, s->stream>>>(h.d, s->g, s->nb, h.arr, flag, 0, s->nb);
		k_count_offsets<<<1, 1024, 0, s->stream>>>(h.d.count, s->nb, s->hoff[1][is]);
		k_image_copy<<<(s->nb + 7) / 8, 256, 0, s->stream>>>(h.d, s->nb, h.d.count, s->hoff[1][is], s->hstage[1][is], u.n, 1);
		if((rc = check_launch(s, 4))) return rc;
		CK(cudaEventRecord(s->ev_packed[is], s->stream));
		CK(cudaStreamWaitEvent(s->stream_out, s->ev_packed[is], 0));
		CK(cudaMemcpyAsync(u.cnt, h.d.count, (size_t) s->nb * sizeof(int), cudaMemcpyDeviceToHost, s->stream_out));
		CK(cudaMemcpyAsync(u.data, s->hstage[1][is], (size_t) 6 * (size_t) u.n * sizeof(double), cudaMemcpyDeviceToHost, s->stream_out));
	}
	rc = cpic_b200_stage_field_rho(s);
	if(rc) return rc;
	s->iter++;
	CK(cudaMemcpyAsync(s->h_err + 15, flag, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaStreamSynchronize(s->stream_in));
	CK(cudaStreamSynchronize(s->stream_out));
	CK(cudaStreamSynchronize(s->stream));
	if(s->h_err[15])
		return fail(CPIC_B200_ECAPACITY, "a particle block could not take its arrivals during cpic_b200_step_host: "
				"raise capacity_factor (now %g)", s->p.capacity_factor);
	return cpic_b200_sync(s);
}

/* ---- the host image in bands of block rows (cpic_b200_step_host_banded) ----
 * Layout: int64 magic, bands, species, blocks; then per species that holds particles, per band j:
 * int64 n, int64 cap, int32 counts[blocks of the band] (padded to 8 bytes), double arrays[6][cap]
 * (x y ux uy uz id; `cap` >= n leaves room for the band's population to change). */
#define BAND_MAGIC 0x31444e4243495043LL      /* "CPICBND1" */
#define BAND_MAX 64

struct BandView {
	int64_t *n, *cap;
	int *cnt;
	double *data;
	int b0, nblk;
	size_t stage;            /* first double of the band in the species' device staging */
};

static void
band_rows(const sim_t_ *s, int bands, int j, int *b0, int *nblk)
{
	const int r0 = (int) ((long long) s->g.nby * j / bands), r1 = (int) ((long long) s->g.nby * (j + 1) / bands);
	*b0 = r0 * s->g.nbx;
	*nblk = (r1 - r0) * s->g.nbx;
}

/* Walks the image: fills view[species][band]; with `caps` (from cpic_b200_banded_image_bytes) lays it out */
static int
band_views(sim_t_ *s, char *host, int64_t bytes, int bands, BandView view[][BAND_MAX], size_t stage_total[], bool create,
		const std::vector<std::vector<int64_t>> *caps)
{
	char *p = host, *end = host + bytes;
	if(p + 32 > end) return fail(CPIC_B200_EINVAL, "truncated banded image");
	int64_t *hdr = (int64_t *) p;
	if(create) { hdr[0] = BAND_MAGIC; hdr[1] = bands; hdr[2] = s->p.nspecies; hdr[3] = s->nb; }
	else if(hdr[0] != BAND_MAGIC || hdr[1] != bands || hdr[2] != s->p.nspecies || hdr[3] != s->nb)
		return fail(CPIC_B200_EINVAL, "not a banded image of this simulation");
	p += 32;
	for(int is = 0; is < s->p.nspecies; is++)
	{
		stage_total[is] = 0;
		if(!s->sp[is].block) continue;
		for(int j = 0; j < bands; j++)
		{
			BandView &v = view[is][j];
			band_rows(s, bands, j, &v.b0, &v.nblk);
			if(p + 16 > end) return fail(CPIC_B200_EINVAL, "truncated banded image");
			v.n = (int64_t *) p; v.cap = v.n + 1; p += 16;
			if(create) { *v.cap = (*caps)[(size_t) is][(size_t) j]; *v.n = 0; }
			v.cnt = (int *) p; p += ((size_t) v.nblk * 4 + 7) & ~(size_t) 7;
			v.data = (double *) p;
			if(*v.cap < 0 || *v.n < 0 || *v.n > *v.cap || p + 6 * 8 * *v.cap > end) return fail(CPIC_B200_EINVAL, "truncated banded image");
			p += 6 * 8 * *v.cap;
			v.stage = stage_total[is];
			stage_total[is] += (size_t) 6 * (size_t) *v.cap;
		}
	}
	return 0;
}

static int
band_choose(const sim_t_ *s, int bands)
{
	if(bands < 1) bands = 8;
	if(bands > BAND_MAX) bands = BAND_MAX;
	if(bands > s->g.nby) bands = s->g.nby;
	return bands;
}

/* per-band capacities from the present block counts: a quarter of slack */
static int
band_caps(sim_t_ *s, int bands, std::vector<std::vector<int64_t>> &caps)
{
	caps.assign((size_t) s->p.nspecies, std::vector<int64_t>());
	std::vector<int> cnt((size_t) s->nb);
	for(int is = 0; is < s->p.nspecies; is++)
	{
		SpeciesHost &h = s->sp[is];
		if(!h.block) continue;
		int rc = absorb(s, is);
		if(rc) return rc;
		CK(cudaMemcpyAsync(cnt.data(), h.d.count, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
		CK(cudaStreamSynchronize(s->stream));
		for(int j = 0; j < bands; j++)
		{
			int b0, nblk;
			band_rows(s, bands, j, &b0, &nblk);
			int64_t n = 0;
			for(int b = b0; b < b0 + nblk; b++) n += cnt[(size_t) b];
			caps[(size_t) is].push_back((n + n / 4 + 4096 + 31) / 32 * 32);
		}
	}
	return 0;
}

extern "C" int64_t
cpic_b200_banded_image_bytes(cpic_b200_sim_t *s, int bands)
{
	if(!s) return -1;
	cudaSetDevice(s->device);
	bands = band_choose(s, bands);
	std::vector<std::vector<int64_t>> caps;
	if(band_caps(s, bands, caps)) return -1;
	int64_t bytes = 32;
	for(int is = 0; is < s->p.nspecies; is++)
		for(size_t j = 0; j < caps[(size_t) is].size(); j++)
		{
			int b0, nblk;
			band_rows(s, bands, (int) j, &b0, &nblk);
			bytes += 16 + (((int64_t) nblk * 4 + 7) & ~7LL) + 6 * 8 * caps[(size_t) is][j];
		}
	return bytes;
}

static int
band_staging(sim_t_ *s, const size_t stage_total[])
{
	for(int is = 0; is < s->p.nspecies; is++)
	{
		if(!s->sp[is].block) continue;
		for(int k = 0; k < 2; k++)
		{
			if(s->hstage_cap[k][is] < stage_total[is])
			{
				cudaFree(s->hstage[k][is]);
				s->hstage[k][is] = NULL; s->hstage_cap[k][is] = 0;
				CK(cudaMalloc(&s->hstage[k][is], stage_total[is] * sizeof(double)));
				s->hstage_cap[k][is] = stage_total[is];
			}
			/* band-local offsets: nblk + 1 entries per band */
			if(!s->hoff[k][is]) CK(cudaMalloc(&s->hoff[k][is], ((size_t) s->nb + BAND_MAX + 1) * sizeof(long long)));
		}
		if(!s->hcnt[is]) CK(cudaMalloc(&s->hcnt[is], (size_t) s->nb * sizeof(int)));
	}
	if(!s->h_band) CK(cudaMallocHost(&s->h_band, (size_t) CPIC_B200_MAX_SPECIES * (BAND_MAX + 1) * sizeof(long long)));
	if(!s->stream_in)
	{
		CK(cudaStreamCreateWithFlags(&s->stream_in, cudaStreamNonBlocking));
		CK(cudaStreamCreateWithFlags(&s->stream_out, cudaStreamNonBlocking));
		for(int i = 0; i < CPIC_B200_MAX_SPECIES; i++)
		{
			CK(cudaEventCreateWithFlags(&s->ev_up[i], cudaEventDisableTiming));
			CK(cudaEventCreateWithFlags(&s->ev_packed[i], cudaEventDisableTiming));
		}
	}
	for(int i = 0; i < CPIC_B200_MAX_SPECIES; i++)
		for(int k = 0; k < 2 * BAND_MAX; k++)
			if(!s->ev_band[i][k]) CK(cudaEventCreateWithFlags(&s->ev_band[i][k], cudaEventDisableTiming));
	return 0;
}

/* absorb + offsets + pack of band j of one species on the compute stream; the band's new total lands in
 * the pinned word `total`; records ev_band[is][64 + j] */
static int
band_pack(sim_t_ *s, int is, int j, const BandView &v, int *flag)
{
	SpeciesHost &h = s->sp[is];
	long long *off = s->hoff[1][is] + v.b0 + j;
	k_absorb<<<(v.nblk + 7) / 8, 256, 0, s->stream>>>(h.d, s->g, s->nb, h.arr, flag, v.b0, v.b0 + v.nblk);
	k_count_offsets<<<1, 1024, 0, s->stream>>>(h.d.count + v.b0, v.nblk, off);
	k_band_copy<<<(v.nblk + 7) / 8, 256, 0, s->stream>>>(h.d, v.b0, v.nblk, h.d.count + v.b0, off,
			s->hstage[1][is] + v.stage, *v.cap, 1, flag);
	int rc = check_launch(s, 3);
	if(rc) return rc;
	CK(cudaMemcpyAsync(s->h_band + (size_t) is * (BAND_MAX + 1) + j, off + v.nblk, sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaEventRecord(s->ev_band[is][BAND_MAX + j], s->stream));
	return 0;
}

/* the packed band goes to the host image (after its event): counts and the six arrays */
static int
band_download(sim_t_ *s, int is, int j, const BandView &v)
{
	SpeciesHost &h = s->sp[is];
	CK(cudaEventSynchronize(s->ev_band[is][BAND_MAX + j]));
	const long long n = s->h_band[(size_t) is * (BAND_MAX + 1) + j];
	if(n < 0 || n > *v.cap)
		return fail(CPIC_B200_ECAPACITY, "species %d: band %d holds %lld particles, its image %lld: make a new image", is, j, n, (long long) *v.cap);
	*v.n = n;
	CK(cudaStreamWaitEvent(s->stream_out, s->ev_band[is][BAND_MAX + j], 0));
	CK(cudaMemcpyAsync(v.cnt, h.d.count + v.b0, (size_t) v.nblk * sizeof(int), cudaMemcpyDeviceToHost, s->stream_out));
	for(int a = 0; a < 6; a++)
		CK(cudaMemcpyAsync(v.data + (size_t) a * *v.cap, s->hstage[1][is] + v.stage + (size_t) a * *v.cap, (size_t) n * sizeof(double),
					cudaMemcpyDeviceToHost, s->stream_out));
	return 0;
}

/* The present state as a banded image (bands <= 0: 8) */
extern "C" int
cpic_b200_banded_image_download(cpic_b200_sim_t *s, void *host, int64_t bytes, int bands)
{
	if(!s || !host) return fail(CPIC_B200_EINVAL, "null argument");
	CK(cudaSetDevice(s->device));
	bands = band_choose(s, bands);
	std::vector<std::vector<int64_t>> caps;
	int rc = band_caps(s, bands, caps);
	if(rc) return rc;
	static BandView view[CPIC_B200_MAX_SPECIES][BAND_MAX];
	size_t stage_total[CPIC_B200_MAX_SPECIES];
	if((rc = band_views(s, (char *) host, bytes, bands, view, stage_total, true, &caps))) return rc;
	if((rc = band_staging(s, stage_total))) return rc;
	int *flag = s->errflag + 15;
	CK(cudaMemsetAsync(flag, 0, sizeof(int), s->stream));
	for(int is = 0; is < s->p.nspecies; is++)
	{
		if(!s->sp[is].block) continue;
		for(int j = 0; j < bands; j++)
		{
			if((rc = band_pack(s, is, j, view[is][j], flag))) return rc;
			if((rc = band_download(s, is, j, view[is][j]))) return rc;
		}
	}
	CK(cudaMemcpyAsync(s->h_err + 15, flag, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaStreamSynchronize(s->stream_out));
	CK(cudaStreamSynchronize(s->stream));
	if(s->h_err[15]) return fail(CPIC_B200_ECAPACITY, "a band outgrew its image");
	return 0;
}

/* One sim_step with the particle state in pinned host memory, in bands of block rows: a band is
 * uploaded, unpacked and pushed while the next ones are still on their way, and as soon as its
 * neighbour bands are pushed too its arrivals are absorbed and it is packed and downloaded -- uploads
 * and downloads overlap all along (cpic_b200_step_host overlaps species only). Band 0 closes the ring
 * with the last band and goes last. Far movers (rare) are placed when every band is pushed: if there
 * were any, the species is packed and downloaded again. One rank. */
extern "C" int
cpic_b200_step_host_banded(cpic_b200_sim_t *s, void *host, int64_t bytes)
{
	if(!s || !host) return fail(CPIC_B200_EINVAL, "null argument");
	if(s->iter < 0) return fail(CPIC_B200_EINVAL, "call cpic_b200_pre_step first");
	if(s->comm) return fail(CPIC_B200_EINVAL, "cpic_b200_step_host_banded runs on one rank");
	CK(cudaSetDevice(s->device));
	if(bytes < 32 || ((int64_t *) host)[0] != BAND_MAGIC) return fail(CPIC_B200_EINVAL, "not a banded image");
	const int bands = (int) ((int64_t *) host)[1];
	if(bands < 1 || bands > BAND_MAX || bands > s->g.nby) return fail(CPIC_B200_EINVAL, "bad number of bands");
	static BandView view[CPIC_B200_MAX_SPECIES][BAND_MAX];
	size_t stage_total[CPIC_B200_MAX_SPECIES];
	int rc = band_views(s, (char *) host, bytes, bands, view, stage_total, false, NULL);
	if(rc) return rc;
	if((rc = band_staging(s, stage_total))) return rc;
	CK(cudaStreamSynchronize(s->stream));
	const Geom &g = s->g;
	std::vector<std::vector<long long>> off((size_t) s->p.nspecies * bands);

	/* uploads: every band of every species, in the order they are used */
	for(int is = 0; is < s->p.nspecies; is++)
	{
		SpeciesHost &h = s->sp[is];
		if(!h.block) continue;
		for(int j = 0; j < bands; j++)
		{
			const BandView &v = view[is][j];
			std::vector<long long> &o = off[(size_t) is * bands + j];
			o.resize((size_t) v.nblk + 1);
			long long m = 0;
			for(int k = 0; k < v.nblk; k++)
			{
				if(v.cnt[k] < 0 || v.cnt[k] > h.d.cap) return fail(CPIC_B200_ECAPACITY, "image block %d holds %d particles, capacity %d", v.b0 + k, v.cnt[k], h.d.cap);
				o[(size_t) k] = m; m += v.cnt[k];
			}
			o[(size_t) v.nblk] = m;
			if(m != *v.n) return fail(CPIC_B200_EINVAL, "image counts of band %d do not add up", j);
			CK(cudaMemcpyAsync(s->hcnt[is] + v.b0, v.cnt, (size_t) v.nblk * sizeof(int), cudaMemcpyHostToDevice, s->stream_in));
			CK(cudaMemcpyAsync(s->hoff[0][is] + v.b0 + j, o.data(), o.size() * sizeof(long long), cudaMemcpyHostToDevice, s->stream_in));
			for(int a = 0; a < 6; a++)
				CK(cudaMemcpyAsync(s->hstage[0][is] + v.stage + (size_t) a * *v.cap, v.data + (size_t) a * *v.cap, (size_t) *v.n * sizeof(double),
							cudaMemcpyHostToDevice, s->stream_in));
			CK(cudaEventRecord(s->ev_band[is][j], s->stream_in));
		}
	}
	/* the fields of this step do not depend on the upload */
	if((rc = cpic_b200_stage_field_E(s))) return rc;
	int *flag = s->errflag + 15;
	CK(cudaMemsetAsync(flag, 0, sizeof(int), s->stream));
	for(int is = 0; is < s->p.nspecies; is++)
	{
		SpeciesHost &h = s->sp[is];
		if(!h.block) continue;
		/* the image holds every particle: no arrivals are pending */
		for(int k = 0; k < 2; k++) CK(cudaMemsetAsync(h.d.ob[k].count, 0, (size_t) s->nob * 9 * sizeof(int), s->stream));
		const int cur = h.arr ^ 1;
		for(int j = 0; j < bands; j++)
		{
			const BandView &v = view[is][j];
			CK(cudaStreamWaitEvent(s->stream, s->ev_band[is][j], 0));
			k_band_copy<<<(v.nblk + 7) / 8, 256, 0, s->stream>>>(h.d, v.b0, v.nblk, s->hcnt[is] + v.b0, s->hoff[0][is] + v.b0 + j,
					s->hstage[0][is] + v.stage, *v.cap, 0, flag);
			if((rc = check_launch(s))) return rc;
			if((rc = launch_gather_push<2>(s, is, s->stream, v.b0 / g.nbx, v.nblk / g.nbx))) return rc;
			/* band j-1 has all its neighbours pushed now (band 0 waits for the last one) */
			if(j >= 2)
			{
				h.arr = cur;         /* k_absorb reads the outbox this push fills */
				if((rc = band_pack(s, is, j - 1, view[is][j - 1], flag))) return rc;
				h.arr = cur ^ 1;
			}
		}
		h.arr = cur;
		/* far movers of the whole species: their number first (pinned), then their placement */
		CK(cudaMemcpyAsync(s->h_band + (size_t) is * (BAND_MAX + 1) + BAND_MAX, h.d.fcount, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
		SpeciesSet set;
		set.sp[0] = h.d; set.arr[0] = h.arr; set.n = 1;
		k_far_insert<<<1, 1024, 0, s->stream>>>(set, g, s->errflag);
		if((rc = check_launch(s))) return rc;
		if(bands >= 2) { if((rc = band_pack(s, is, bands - 1, view[is][bands - 1], flag))) return rc; }
		if((rc = band_pack(s, is, 0, view[is][0], flag))) return rc;
	}
	if((rc = cpic_b200_stage_field_rho(s))) return rc;
	s->iter++;
	CK(cudaMemcpyAsync(s->h_err + 15, flag, sizeof(int), cudaMemcpyDeviceToHost, s->stream));

	/* downloads, band by band as they are packed: 1 .. bands-1, then 0 */
	for(int is = 0; is < s->p.nspecies; is++)
	{
		SpeciesHost &h = s->sp[is];
		if(!h.block) continue;
		for(int k = 1; k <= bands; k++)
			if((rc = band_download(s, is, k % bands, view[is][k % bands]))) return rc;
		/* the far-mover count was copied before the last bands were packed */
		const int nfar = *(const int *) (s->h_band + (size_t) is * (BAND_MAX + 1) + BAND_MAX);
		if(nfar > 0)
		{
			/* k_far_insert appended particles to blocks that were packed before: once more, in order */
			CK(cudaStreamSynchronize(s->stream_out));
			for(int j = 0; j < bands; j++)
			{
				if((rc = band_pack(s, is, j, view[is][j], flag))) return rc;
				if((rc = band_download(s, is, j, view[is][j]))) return rc;
			}
		}
	}
	CK(cudaStreamSynchronize(s->stream_in));
	CK(cudaStreamSynchronize(s->stream_out));
	CK(cudaStreamSynchronize(s->stream));
	if(s->h_err[15])
		return fail(CPIC_B200_ECAPACITY, "a particle block or a band of the image overflowed during cpic_b200_step_host_banded: "
				"raise capacity_factor (now %g) or make a new image", s->p.capacity_factor);
	return cpic_b200_sync(s);
}

extern "C" void *
cpic_b200_host_alloc(size_t bytes)
{
	void *p = NULL;
	if(cudaMallocHost(&p, bytes) != cudaSuccess) return NULL;
	return p;
}

extern "C" void cpic_b200_host_free(void *p) { if(p) cudaFreeHost(p); }

/* ---------------------------------------------------------------- multi-GPU */

/* Peer memory (comm.h): tells the neighbour ranks where this rank's outboxes live now and maps
 * theirs: pob[north/south][buffer] of every species. Collective -- every rank calls it at the same
 * points: when the communicator is created and after species storage has changed (capacity growth
 * agreed over the ranks, the per-particle E arrays at the first staged step). */
static int
p2p_attach(sim_t_ *s)
{
	Comm *c = s->comm;
	s->p2p_stale = false;
	if(!comm_p2p(c)) return 0;
	for(int is = 0; is < s->p.nspecies; is++)
	{
		SpeciesHost &h = s->sp[is];
		/* mapped again from scratch: a new allocation can sit at the address of the one it replaces */
		if(comm_p2p_unmap(c, comm_export_species(is, 0)) || comm_p2p_unmap(c, comm_export_species(is, 1)))
			return CPIC_B200_ECUDA;
		/* the outboxes are one allocation (two buffers of records + counters); the (E_x, E_y) of the
		 * records live in the per-particle E allocation */
		if(comm_p2p_export(c, comm_export_species(is, 0), h.oblock, h.oblock ? 1 : 0)
				|| comm_p2p_export(c, comm_export_species(is, 1), h.pE, h.pE ? 1 : 0))
			return CPIC_B200_ECUDA;
	}
	if(comm_p2p_refresh(c, s->stream)) return CPIC_B200_ECUDA;
	const int peer[2] = { comm_rank_north(c), comm_rank_south(c) };
	for(int is = 0; is < s->p.nspecies; is++)
	{
		SpeciesHost &h = s->sp[is];
		memset(h.d.pob, 0, sizeof(h.d.pob));
		if(!h.oblock) continue;
		for(int d = 0; d < 2; d++)
		{
			char *ro = (char *) comm_p2p_remote(c, comm_export_species(is, 0), peer[d]);
			char *re = (char *) comm_p2p_remote(c, comm_export_species(is, 1), peer[d]);
			if(!ro) return fail(CPIC_B200_EINVAL, "rank %d has no storage for species %d: every rank must hold every species", peer[d], is);
			if(h.pE && !re) return fail(CPIC_B200_EINVAL, "rank %d keeps no per-particle E for species %d", peer[d], is);
			for(int k = 0; k < 2; k++)
			{
				/* equal capacities on all ranks (cpic_b200_capacity): the neighbours' layout is this rank's */
				h.d.pob[d][k].rec = (double *) (ro + ((char *) h.d.ob[k].rec - (char *) h.oblock));
				h.d.pob[d][k].count = (int *) (ro + ((char *) h.d.ob[k].count - (char *) h.oblock));
				h.d.pob[d][k].recE = h.d.ob[k].recE ? (double *) (re + ((char *) h.d.ob[k].recE - (char *) h.pE)) : NULL;
			}
		}
	}
	return 0;
}

extern "C" int
cpic_b200_comm_id(void *id128)
{
	if(!id128) return fail(CPIC_B200_EINVAL, "null id");
	return comm_unique_id(id128, g_err, sizeof(g_err));
}

extern "C" int
cpic_b200_comm_init(cpic_b200_sim_t *s, const void *id128)
{
	if(!s || !id128) return fail(CPIC_B200_EINVAL, "null argument");
	if(s->p.nranks == 1) return 0;
	CK(cudaSetDevice(s->device));
	if(s->comm) return fail(CPIC_B200_EINVAL, "communicator already initialised");
	s->comm = comm_create(id128, s->p.rank, s->p.nranks, s->g, s->stream, g_err, sizeof(g_err));
	if(!s->comm) return CPIC_B200_ECUDA;
	if(comm_p2p(s->comm))
	{
		if(comm_p2p_export(s->comm, COMM_EXPORT_PHI, s->phi, (size_t) (s->g.ny + 3) * s->g.S * sizeof(double))) return CPIC_B200_ECUDA;
		return p2p_attach(s);
	}
	return 0;
}
