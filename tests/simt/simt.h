/* simt-check: a lockstep SIMT interpreter for the CPU test suite.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under tests/simt/ is part of the product: the package
 * never builds, loads or falls back to it, and libcpic_b200.so does not contain it. Its
 * purpose is to execute the *same kernel source* (cpic_b200/csrc/kernels.cuh, sim.cu, comm.cu)
 * on the CPU in the `-m "not gpu"` tests, so that index arithmetic, compaction order,
 * capacities, barrier protocols and pipeline hazards of the device code are checked against
 * the oracle where no GPU exists. It says nothing about speed.
 *
 * Model: one CTA at a time; every CUDA thread is a fiber (own stack, hand-written context
 * switch); warp collectives (__ballot_sync, __shfl_*_sync, __match_any_sync,
 * __reduce_max_sync, __syncwarp) and __syncthreads suspend the fiber until every
 * participant has arrived, exactly the convergence the *_sync primitives guarantee.
 * Asynchronous copies (cp.async groups, TMA bulk / tensor loads on an mbarrier) are
 * completed as LATE as the programming model allows -- at the matching wait -- so that a
 * read before the wait sees the poison the buffers are filled with.
 */
#ifndef SIMT_CHECK_H
#define SIMT_CHECK_H

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>

struct uint3 { unsigned x, y, z; };
struct dim3 {
	unsigned x, y, z;
	dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 v = { x, y }; return v; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) alignas(n)
#define __shared__ static

using std::min;
using std::max;

namespace simt {

extern uint3 tid, bid;
extern dim3 bdim, gdim;

/* kernel<<<grid, block, smem, stream>>>(args) is rewritten by translate.py into
 * simt::launch(grid, block, smem, [=]() { kernel(args); }) */
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
unsigned char *dyn_smem();

enum { OP_BALLOT, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_MATCH_ANY, OP_REDUCE_MAX, OP_SYNC };
uint64_t collective(int op, unsigned mask, uint64_t v, int aux);
void sync_cta();
void yield();

/* asynchronous copies, completed at the wait */
void async_copy8(void *dst, const void *src);
void async_commit();
void misuse(const char *what);
void async_wait(int keep_newest);
void bar_init(uint64_t *bar, int count);
void bar_expect_tx(uint64_t *bar, uint32_t bytes);
int bar_try_wait(uint64_t *bar, uint32_t parity);
void bar_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar);
void bar_tensor_load_2d(void *dst, const void *map, int c0, int c1, uint64_t *bar);

/* statistics of the last launch, for the tests */
extern long long n_launches, n_switches;

template <typename T> static inline uint64_t bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> static inline T unbits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

}  /* namespace simt */

#define threadIdx simt::tid
#define blockIdx simt::bid
#define blockDim simt::bdim
#define gridDim simt::gdim

/* ---- warp and CTA primitives ---- */
static inline unsigned __ballot_sync(unsigned m, int p) { return (unsigned) simt::collective(simt::OP_BALLOT, m, p != 0, 0); }
template <typename T> static inline T __shfl_sync(unsigned m, T v, int src) { return simt::unbits<T>(simt::collective(simt::OP_SHFL, m, simt::bits(v), src)); }
template <typename T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d) { return simt::unbits<T>(simt::collective(simt::OP_SHFL_UP, m, simt::bits(v), (int) d)); }
template <typename T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d) { return simt::unbits<T>(simt::collective(simt::OP_SHFL_DOWN, m, simt::bits(v), (int) d)); }
template <typename T> static inline unsigned __match_any_sync(unsigned m, T v) { return (unsigned) simt::collective(simt::OP_MATCH_ANY, m, simt::bits(v), 0); }
static inline int __reduce_max_sync(unsigned m, int v) { return (int) (int64_t) simt::collective(simt::OP_REDUCE_MAX, m, (uint64_t) (int64_t) v, 0); }
static inline void __syncwarp(unsigned m = 0xffffffffu) { simt::collective(simt::OP_SYNC, m, 0, 0); }
static inline void __syncthreads() { simt::sync_cta(); }

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned) v) : 32; }
static inline long long __double_as_longlong(double v) { return simt::unbits<long long>(simt::bits(v)); }
static inline double __longlong_as_double(long long v) { return simt::unbits<double>(simt::bits(v)); }

/* one fiber runs at a time: plain read-modify-write is atomic */
template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; if(v > o) *p = v; return o; }

/* ---- the part of the CUDA runtime API the library uses ---- */
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
typedef struct simt_stream_ *cudaStream_t;
typedef struct simt_event_ *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEnableDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };

cudaError_t simt_malloc(void **p, size_t n);
template <typename T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return simt_malloc((void **) p, n); }
template <typename T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { return simt_malloc((void **) p, n); }
cudaError_t cudaFree(void *p);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t st = 0);
cudaError_t cudaMemcpy2DAsync(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height,
		cudaMemcpyKind k, cudaStream_t st = 0);
cudaError_t cudaMemset(void *d, int v, size_t n);
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st = 0);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetLastError(void);
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaDeviceSynchronize(void);
enum { cudaEventDefault = 0, cudaEventDisableTiming = 2 };
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = 0);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaGetDriverEntryPoint(const char *name, void **fn, unsigned long long flags, cudaDriverEntryPointQueryResult *res = NULL);
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F *, cudaFuncAttribute, int bytes)
{
	return bytes <= 227 * 1024 ? cudaSuccess : cudaErrorInvalidValue;
}

/* a small "device": 4 SMs with 2 resident CTAs each, so that persistent grids loop over their work */
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }
template <typename F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F *, int, size_t)
{
	*n = 2;
	return cudaSuccess;
}

/* no peer memory between the processes of the CPU test suite: the multi-rank tests take the NCCL path */
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return cudaErrorInvalidValue; }
static inline cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return cudaErrorInvalidValue; }
static inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaErrorInvalidValue; }
static inline void __threadfence_system() {}
static inline long long clock64() { return 0; }
static inline void __nanosleep(unsigned) {}

/* ---- driver API: tensor maps ---- */
typedef int CUresult;
enum { CUDA_SUCCESS = 0, CUDA_ERROR_INVALID_VALUE = 1 };
typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_FLOAT64 = 10 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0, CU_TENSOR_MAP_L2_PROMOTION_L2_128B = 2 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };
struct alignas(64) CUtensorMap {
	void *base;
	uint64_t dims[2];
	uint64_t row_bytes;
	uint32_t box[2];
	uint32_t elem;
	char pad[128 - 8 - 16 - 8 - 8 - 4];
};
typedef CUtensorMap CUtensorMap_st;

/* ---- cuFFT: the two 2-D plans of the single-rank solver are executed (radix-2 / plain DFT);
 * the batched 1-D plans of the distributed solver are accepted and refused at execution ---- */
typedef int cufftResult;
enum { CUFFT_SUCCESS = 0, CUFFT_INVALID_PLAN = 1, CUFFT_EXEC_FAILED = 6, CUFFT_NOT_SUPPORTED = 16 };
typedef int cufftHandle;
typedef double2 cufftDoubleComplex;
typedef double cufftDoubleReal;
enum cufftType { CUFFT_D2Z = 0x6a, CUFFT_Z2D = 0x6c, CUFFT_Z2Z = 0x69 };
enum { CUFFT_FORWARD = -1, CUFFT_INVERSE = 1 };
cufftResult cufftPlanMany(cufftHandle *plan, int rank, int *n, int *inembed, int istride, int idist,
		int *onembed, int ostride, int odist, cufftType type, int batch);
cufftResult cufftSetStream(cufftHandle plan, cudaStream_t s);
cufftResult cufftDestroy(cufftHandle plan);
cufftResult cufftExecD2Z(cufftHandle plan, cufftDoubleReal *in, cufftDoubleComplex *out);
cufftResult cufftExecZ2D(cufftHandle plan, cufftDoubleComplex *in, cufftDoubleReal *out);
cufftResult cufftExecZ2Z(cufftHandle plan, cufftDoubleComplex *in, cufftDoubleComplex *out, int dir);

#endif
