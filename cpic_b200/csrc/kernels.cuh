/* Device kernels of the cpic_b200 hot path (sm_100a).
 *
 * Particle storage ("particle blocks"): every species is a set of fixed-capacity SoA
 * segments, one per block of BX x BY grid cells. The particles whose cell lies in block
 * b are its own segment plus the runs that its eight neighbours left for it in their
 * outboxes during the last push ("arrivals"); they are never copied just to be moved:
 * the next push streams over segment and arrivals alike, writes the particles that stay
 * back into the segment (compacted, in place) and the ones that leave into its own
 * outbox. One warp owns one particle block for the duration of a kernel and walks it
 * sequentially in batches of 32, so every running count, compaction and floating point
 * sum has a fixed order; blocks only meet through fixed-order reads of neighbour
 * outboxes / halo arrays. No floating point atomics are used anywhere.
 */
#ifndef CPIC_B200_KERNELS_CUH
#define CPIC_B200_KERNELS_CUH

#include <cuda.h>
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>

#include "geom.h"

#define FULL 0xffffffffu

#ifndef MAX_WPC
#define MAX_WPC 8
#endif
/* OWN_BULK 1: a batch of the own segment is fetched by one lane with one TMA bulk copy per array
 * (UBLKCP, completion on a per-stage mbarrier); 0: by all lanes with 16-byte cp.async copies,
 * three instructions per batch, in the same async-copy groups as the arrivals */
#ifndef OWN_BULK
#define OWN_BULK 1
#endif

/* SEG_AOSOA 0: a species' segments are six arrays (x y ux uy uz id) of nb*cap slots each.
 * 1: batch-major -- the 32 slots of a batch hold x[32] y[32] ux[32] uy[32] uz[32] id[32] in
 * 1536 contiguous bytes, so that a batch is one TMA copy and one address; the six pointers of
 * SpeciesDev then point 32 doubles apart into one array. Either way element i of block b is
 * ptr[seg_slot(cap, b, i)]. (The optional per-particle E arrays stay plain: b*cap + i.) */
#ifndef SEG_AOSOA
#define SEG_AOSOA 0
#endif

/* SEG_TENSOR 1 (plain layout only): the fused kernel fetches a batch of the own segment -- 32 slots of
 * x y ux uy uz (id) -- with ONE tensor copy (cp.async.bulk.tensor.2d over the species' storage seen as
 * a 2D tensor slot x array) instead of five or six 1D bulk copies: the elected lane's issue sequence
 * was a fifth of the kernel's instructions */
#ifndef SEG_TENSOR
#define SEG_TENSOR (!SEG_AOSOA)
#endif

/* doubles from a batch's x values to its y values (and so on): SEG_STEP = 32 batch-major;
 * in the plain layout the arrays are SpeciesDev::astride apart */
#if SEG_AOSOA
#define SEG_STEP ((size_t) 32)
#else
#define SEG_STEP ((size_t) sp.astride)
#endif

/* Slot numbers of the plain layout fit 32 bits (checked when a species is allocated): one
 * IMAD.WIDE per access instead of 64-bit address chains */
#if SEG_AOSOA
typedef size_t seg_index_t;
#else
typedef unsigned seg_index_t;
#endif

__host__ __device__ __forceinline__ seg_index_t
seg_slot(int cap, int b, int i)
{
#if SEG_AOSOA
	return ((size_t) b * cap + (size_t) (i & ~31)) * 6 + (size_t) (i & 31);
#else
	return (unsigned) b * (unsigned) cap + (unsigned) i;
#endif
}

/* Outbox: the particles that left their block in one push, one region per block and
 * destination code (geom.h), so that the receiving block finds its arrivals as
 * contiguous runs. A slot is one 48-byte record (x y ux uy uz id): a leaver is written
 * with three 16-byte stores and the leavers of a block for one destination fill
 * consecutive records, so the few dozen particles a region receives per step are one
 * contiguous run in memory instead of one short run in each of six arrays. Regions are
 * stored code-major -- all blocks' code-0 regions, then all code-1 regions ... -- so that
 * what one block row sends across a slab face (codes 0,1,2 of row 0, codes 6,7,8 of the last
 * row) is three contiguous chunks. Side regions (codes 1,3,5,7) hold `ocs` slots, corner
 * regions (0,2,6,8) `occ`. Two outboxes alternate: a push reads the arrivals of the
 * previous push from one and fills the other. */
#define OREC 6                   /* doubles per outbox record */
struct Outbox {
	double *rec;             /* [slot][x y ux uy uz id] */
	double *recE;            /* [slot][Ex Ey]: travels with the particle when the per-particle E is kept */
	int *count;              /* [code][block]: leavers of that block with that destination */
};

/* Device view of one species */
struct SpeciesDev {
	double *x, *y, *ux, *uy, *uz;
	double *pEx, *pEy;       /* gathered field per particle, optional (ppack.E, reference src/def.h:96) */
	long long *id;
	int *count;              /* particles in the block's own segment */
	Outbox ob[2];
	/* The outboxes of the neighbour ranks, mapped over NVLink (CUDA IPC): pob[0] the north
	 * rank's, pob[1] the south rank's, both buffers. With them a leaver that crosses a slab face is
	 * written straight into the GHOST OUTBOX ROW of its new rank (rec == NULL: no peer mapping, the
	 * face regions travel by NCCL, comm.cu) */
	Outbox pob[2][2];
	int cap;                 /* slots per block segment */
	unsigned astride;        /* doubles between the segment arrays x y ux uy uz id (one allocation) */
	int ocs, occ;            /* slots per side / corner region */
	int nob;                 /* blocks that own an outbox: the slab's, plus two ghost rows with several ranks */
	unsigned roff[9];        /* first slot of the regions of code c */
	int rcap[9];             /* slots per region of code c (0 for code 4) */
	/* particles that jumped further than a neighbouring block in one step (rare): a small
	 * list, inserted by k_far_insert in id order */
	double *fx, *fy, *fux, *fuy, *fuz, *fEx, *fEy;
	long long *fid;
	long long *fkey;         /* sort scratch: (destination block, id) */
	int *fidx;
	int *fcount;
	/* far movers that belong to the north [0] / south [1] neighbour rank: FAR_FACE entries of
	 * eight values (x y ux uy uz id Ex Ey, array-major), sent with the particle faces */
	double *rfar[2];
	int *rfcount;            /* [2] */
};

/* Every species of a simulation, for the kernels that take one CTA (or grid row) per species */
#define SET_MAX_SPECIES 8
struct SpeciesSet {
	SpeciesDev sp[SET_MAX_SPECIES];
	int arr[SET_MAX_SPECIES];    /* the outbox that holds the pending arrivals */
	int n;
};

#define FAR_CAP 65536
#define FAR_FACE 2048

/* Slot of entry `pos` in the region (block b, destination code c) */
__device__ __forceinline__ unsigned
region_slot(const SpeciesDev &sp, int c, int b, int pos)
{
	return sp.roff[c] + (unsigned) b * (unsigned) sp.rcap[c] + (unsigned) pos;
}

/* Everything the mover needs besides the particle (reference src/mover.c:191-226) */
struct PushParams {
	double dt;               /* -dt/2 at iteration 0 */
	double dtqm2;            /* 0.5 * dt * q / m */
	double tx, ty, tz;       /* t = B * dtqm2 (reference src/mover.c:45) */
	double sx, sy, sz;       /* s = 2t / (1 + t^2), per component (src/mover.c:46,51) */
	double umax_x, umax_y, umax_z;
	int set_r;               /* 0 at iteration 0: velocities only */
};

/* ------------------------------------------------------------------------ TMA */

#ifdef CPIC_B200_SIMT_CHECK
/* tests/simt (CPU test suite): the same six primitives run by a lockstep SIMT interpreter */
#include "simt_async.h"
#else

__device__ __forceinline__ uint32_t
smem_u32(const void *p)
{
	return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void
mbar_init(uint64_t *bar, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void
mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
			:: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

/* Returns 0 on success, 1 when the barrier did not complete (descriptor error): the
 * caller raises an error flag instead of hanging the device. */
__device__ __forceinline__ int
mbar_wait(uint64_t *bar, uint32_t parity)
{
	uint32_t done = 0;
	for(int spin = 0; spin < (1 << 22); spin++)
	{
		asm volatile("{\n\t.reg .pred p;\n\t"
				"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
				"selp.u32 %0, 1, 0, p;\n\t}"
				: "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
		if(done) return 0;
	}
	return 1;
}

/* 2D tile load: box (TW x TH doubles) of a row-major array whose first coordinate is
 * the column. Out-of-range elements are zero-filled by the TMA unit. */
__device__ __forceinline__ void
tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
			" [%0], [%1, {%2, %3}], [%4];"
			:: "r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
			: "memory");
}

/* 1D bulk copy global->shared through the TMA unit (UBLKCP): `bytes` multiple of 16, both
 * addresses 16 B aligned; completes on the mbarrier */
__device__ __forceinline__ void
tma_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			:: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

/* 16-byte asynchronous global->shared copy (LDGSTS.128), L2 only: the particle pipeline's
 * prefetch of arrival records (and of segment batches with OWN_BULK=0); streamed once */
__device__ __forceinline__ void
cp_async16(void *dst, const void *src)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst)), "l"(src) : "memory");
}

__device__ __forceinline__ void
cp_async_commit()
{
	asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int N>
__device__ __forceinline__ void
cp_async_wait()
{
	asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}

#endif  /* CPIC_B200_SIMT_CHECK */

/* ------------------------------------------------------------ field kernels */

/* MFT_kernel, reference src/solver.c:337-363: g[l][k] *= G[l][k]. 16 B per element
 * read + 8 B of G + 16 B written: HBM-bound streaming. */
static __global__ void
k_green(cufftDoubleComplex *__restrict__ g, const double *__restrict__ G, size_t n)
{
	size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	size_t stride = (size_t) gridDim.x * blockDim.x;
	for(; i < n; i += stride)
	{
		cufftDoubleComplex v = g[i];
		double c = G[i];
		v.x *= c;
		v.y *= c;
		g[i] = v;
	}
}

/* MFT_normalize (reference src/solver.c:365-379: phi /= nx*ny on the nx live columns)
 * fused with the single-rank phi halo (src/comm_field.c:139-201: slab rows 0,1 become
 * the south ghosts, row ny-1 the north ghost). `raw` is the Z2D output (ny x S), `phi`
 * the ny+3 row array. With several ranks only the slab rows are written here and the
 * ghosts arrive over NCCL. */
static __global__ void
k_phi_finish(const double *__restrict__ raw, double *__restrict__ phi, Geom g, double N,
		int fill_ghosts)
{
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	int r = blockIdx.y;               /* row of the ny+3 row array */
	int src;
	if(c >= g.S) return;
	if(r >= 1 && r <= g.ny) src = r - 1;
	else if(!fill_ghosts) return;
	else if(r == 0) src = g.ny - 1;
	else src = r - 1 - g.ny;          /* rows ny+1, ny+2 <- slab rows 0, 1 */
	double v = raw[(size_t) src * g.S + c];
	if(c < g.nx) v /= N;
	phi[(size_t) r * g.S + c] = v;
}

/* field_E_compute, reference src/field.c:358-416: centred differences of phi on rows
 * [0, ny], X periodic, Y through the ghost rows. Also fills the wrap columns
 * [nx, SE) of the device E arrays (column nx+j repeats column j) so that a particle
 * block at the right edge can fetch its tile with one TMA box. */
static __global__ void
k_field_E(const double *__restrict__ phi, double *__restrict__ Ex, double *__restrict__ Ey, Geom g)
{
	int ix = blockIdx.x * blockDim.x + threadIdx.x;
	int iy = blockIdx.y;              /* 0 .. ny */
	if(ix >= g.SE) return;
	int c = ix % g.nx;
	int x0 = c == 0 ? g.nx - 1 : c - 1;
	int x1 = c == g.nx - 1 ? 0 : c + 1;
	const double *p = phi + (size_t) (iy + 1) * g.S;   /* slab row iy is array row iy+1 */
	double dx2 = 2 * g.dx, dy2 = 2 * g.dy;
	const double ey = (p[c - g.S] - p[c + g.S]) / dy2;
	const double ex = (p[x0] - p[x1]) / dx2;
	Ey[(size_t) iy * g.SE + ix] = ey;
	Ex[(size_t) iy * g.SE + ix] = ex;
}

/* k_phi_finish and k_field_E in one pass (one rank: the phi ghost rows are images of the slab's
 * own rows, so nothing has to travel in between). Thread (c, r) writes phi[r][c] and, for r <= ny,
 * E at (c, r) from the same normalised values k_phi_finish would have stored: raw / N first, then
 * the differences -- bit for bit what the two kernels produce. */
static __global__ void
k_phi_E(const double *__restrict__ raw, double *__restrict__ phi, double *__restrict__ Ex,
		double *__restrict__ Ey, Geom g, double N)
{
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	const int r = blockIdx.y;               /* row of the ny+3 row phi array */
	/* slab row behind array row q: 0 <- ny-1, 1..ny <- 0..ny-1, ny+1, ny+2 <- 0, 1 */
	auto src = [&](int q) { return q >= 1 && q <= g.ny ? q - 1 : q == 0 ? g.ny - 1 : q - 1 - g.ny; };
	auto val = [&](int q, int col) { return raw[(size_t) src(q) * g.S + col] / N; };      /* col < nx */
	if(c < g.S)
	{
		double v = raw[(size_t) src(r) * g.S + c];
		if(c < g.nx) v /= N;
		phi[(size_t) r * g.S + c] = v;
	}
	if(r <= g.ny && c < g.SE)
	{
		const int iy = r, cc = c % g.nx;
		const int x0 = cc == 0 ? g.nx - 1 : cc - 1;
		const int x1 = cc == g.nx - 1 ? 0 : cc + 1;
		const double dx2 = 2 * g.dx, dy2 = 2 * g.dy;
		/* slab row iy is array row iy+1: north neighbour array row iy, south iy+2 */
		Ey[(size_t) iy * g.SE + c] = (val(iy, cc) - val(iy + 2, cc)) / dy2;
		Ex[(size_t) iy * g.SE + c] = (val(iy + 1, x0) - val(iy + 1, x1)) / dx2;
	}
}

/* comm_recv_ghost_rho, reference src/comm_field.c:99-136: the ghost row received from rank-1
 * is added to row 0 (on one rank k_rho_assemble folds the own ghost row on its way). */
static __global__ void
k_rho_fold(double *__restrict__ rho, const double *__restrict__ recv, Geom g)
{
	int x = blockIdx.x * blockDim.x + threadIdx.x;
	if(x >= g.nx) return;
	rho[x] += recv[x];
}

/* ------------------------------------------------ peer synchronisation (several ranks) */

/* Ranks that write into each other's memory over NVLink meet through flag words: after the
 * kernels that produced the data, k_peer_signal stores the sequence number of the exchange into
 * flag `slot` of every listed peer (system-scope fence first: the data is visible before the
 * flag); k_peer_wait spins until its own flags, written by those peers, have reached the number.
 * A peer that never arrives raises ERRBIT_PEER instead of hanging the device. */
#define PEER_MAX 16
struct PeerJob {
	int *flag[PEER_MAX];     /* signal: the peers' flag arrays; wait: n times the own array */
	int slot[PEER_MAX];      /* the word of each */
	int n;
	int value;
};

static __global__ void
k_peer_signal(const __grid_constant__ PeerJob job)
{
	__threadfence_system();
	if((int) threadIdx.x < job.n)
		*(volatile int *) (job.flag[threadIdx.x] + job.slot[threadIdx.x]) = job.value;
}

static __global__ void
k_peer_wait(const __grid_constant__ PeerJob job, int *__restrict__ errflag)
{
	if((int) threadIdx.x >= job.n) return;
	const volatile int *f = job.flag[threadIdx.x] + job.slot[threadIdx.x];
#ifdef CPIC_B200_SIMT_CHECK
	if(*f - job.value < 0) atomicOr(errflag, ERRBIT_PEER);       /* one process at a time: nobody to wait for */
#else
	const long long t0 = clock64();
	while(*f - job.value < 0)
	{
		if(clock64() - t0 > 20000000000LL) { atomicOr(errflag, ERRBIT_PEER); break; }      /* ~10 s */
		__nanosleep(200);
	}
	__threadfence_system();
#endif
}

/* ---------------------------------------------------------- particle kernels */

/* Shared memory carve-up of the particle kernels (dynamic):
 *   [0, 16)                      mbarrier
 *   [128, 128 + 2*TH*TW*8)       E_x tile, E_y tile (TMA destinations, 128 B aligned)
 * The deposit kernel uses WPC*(BY+1)*(BX+1) doubles instead. */

__device__ __forceinline__ void
boris(const PushParams &pp, double Ex, double Ey, double &ux, double &uy, double &uz)
{
	/* reference src/mover.c:22-70; per-component s denominator as in the reference.
	 * t and s depend on the species only (B is uniform): formed once on the host. */
	const double k = pp.dtqm2;
	const double tx = pp.tx, ty = pp.ty, tz = pp.tz, sx = pp.sx, sy = pp.sy, sz = pp.sz;
	const double mx = FMA(k, Ex, ux), my = FMA(k, Ey, uy), mz = uz;      /* E_z = 0 */
	double qx, qy, qz;
	if(tx == 0.0 && ty == 0.0)
	{
		/* B along Z (every shipped configuration): the products with t_x, t_y, s_x, s_y
		 * are exact zeros, so dropping them leaves every bit as in the general branch */
		const double px = ADD(MUL(my, tz), mx);
		const double py = ADD(-MUL(mx, tz), my);
		qx = ADD(MUL(py, sz), mx);
		qy = ADD(-MUL(px, sz), my);
		qz = mz;
	}
	else
	{
		const double px = ADD(SUB(MUL(my, tz), MUL(mz, ty)), mx);
		const double py = ADD(SUB(MUL(mz, tx), MUL(mx, tz)), my);
		const double pz = ADD(SUB(MUL(mx, ty), MUL(my, tx)), mz);
		qx = ADD(SUB(MUL(py, sz), MUL(pz, sy)), mx);
		qy = ADD(SUB(MUL(pz, sx), MUL(px, sz)), my);
		qz = ADD(SUB(MUL(px, sy), MUL(py, sx)), mz);
	}
	ux = FMA(k, Ex, qx);
	uy = FMA(k, Ey, qy);
	uz = qz;
}

/* Bilinear gather from the CTA's E tile (reference src/interpolate.c:102-155). lx, ly
 * are the cell's coordinates inside the tile. */
__device__ __forceinline__ double
tile_gather(const double *t, int TW, int lx, int ly, double w00, double w01, double w10, double w11)
{
	const double *p = t + ly * TW + lx;
	double v = MUL(w00, p[0]);
	v = FMA(w01, p[TW], v);
	v = FMA(w10, p[1], v);
	v = FMA(w11, p[TW + 1], v);
	return v;
}

/* The arrivals of block b: lane k < 9 (k != 4) looks at neighbour k, which sits at
 * (k%3-1, k/3-1) and addressed b with code 8-k. With several ranks the block rows -1 and
 * nby are ghost outboxes filled from the neighbour ranks. The nine run starts (exclusive
 * prefix) and source blocks go to the warp's scratch in shared memory (18 ints);
 * returns the total. */
struct Arrivals {
	const int *start;        /* [9] */
	const int *src;          /* [9] */
	int total;
};

__device__ __forceinline__ Arrivals
find_arrivals(const int *__restrict__ in_count, int sp_nob, const Geom &g, int nb, int b, int lane, int *scratch)
{
	const int bx = b % g.nbx, by = b / g.nbx;
	int src = 0, a_k = 0;
	if(lane < 9 && lane != DEST_STAY)
	{
		const int ndx = lane % 3 - 1, ndy = lane / 3 - 1;
		int nbx_ = bx + ndx, nby_ = by + ndy;
		if(nbx_ < 0) nbx_ += g.nbx; else if(nbx_ >= g.nbx) nbx_ -= g.nbx;
		if(g.nby_glob == g.nby)
		{
			if(nby_ < 0) nby_ += g.nby; else if(nby_ >= g.nby) nby_ -= g.nby;
			src = nby_ * g.nbx + nbx_;
		}
		else if(nby_ < 0) src = nb + nbx_;                 /* north ghost row */
		else if(nby_ >= g.nby) src = nb + g.nbx + nbx_;    /* south ghost row */
		else src = nby_ * g.nbx + nbx_;
		a_k = in_count[(size_t) (8 - lane) * sp_nob + src];
	}
	int apre = a_k;
	for(int o = 1; o < 16; o <<= 1)
	{
		const int ta = __shfl_up_sync(FULL, apre, o);
		if(lane >= o) apre += ta;
	}
	Arrivals A;
	A.total = __shfl_sync(FULL, apre, 8);
	if(lane < 9) { scratch[lane] = apre - a_k; scratch[9 + lane] = src; }
	__syncwarp();
	A.start = scratch;
	A.src = scratch + 9;
	return A;
}

__device__ __forceinline__ Arrivals
find_arrivals(const Outbox &in, int sp_nob, const Geom &g, int nb, int b, int lane, int *scratch)
{
	return find_arrivals(in.count, sp_nob, g, nb, b, lane, scratch);
}

/* Outbox slot of arrival f (0 <= f < A.total) */
__device__ __forceinline__ unsigned
arrival_slot(const Arrivals &A, const unsigned *roff, const int *rcap, int f)
{
	int k = 0;
#pragma unroll
	for(int q = 1; q < 9; q++) if(f >= A.start[q]) k = q;
	return roff[8 - k] + (unsigned) A.src[k] * (unsigned) rcap[8 - k] + (unsigned) (f - A.start[k]);
}

__device__ __forceinline__ unsigned
arrival_slot(const Arrivals &A, const SpeciesDev &sp, int f)
{
	return arrival_slot(A, sp.roff, sp.rcap, f);
}

/* MODE 0: stage_plasma_E alone   (gather, store E per particle)
 * MODE 1: stage_plasma_r alone   (push from the stored per-particle E, then exchange)
 * MODE 2: both fused             (gather + push + exchange; E kept when pEx != NULL)
 *
 * plasma_mover and comm_plasma (reference src/mover.c:229-246, src/comm_plasma.c:1122-1142)
 * in one pass. One warp per particle block, WPC blocks of one block row per CTA sharing
 * one E tile fetched by TMA. The warp streams over the block's own segment and then
 * over its arrivals (outbox `cur^1`), in batches of 32 with the next batch's loads
 * already in flight. A particle that stays in the block is written back to the
 * segment at the running write cursor (in place: the cursor never passes the read
 * position); one that leaves goes to the region of its destination in outbox `cur`.
 * Ranks inside a batch come from ballots, so the order is: batches in order, lanes in
 * order. MODE 0 moves nothing. */
#ifndef PIPE_STAGES
#define PIPE_STAGES 3
#endif
#define PUSH_SMEM_HEADER 384      /* tile barrier + MAX_WPC * PIPE_STAGES stage barriers */

/* arrays staged per batch: x y (ux uy uz id) (Ex Ey) */
template <int MODE> struct PipeArrays { static const int N = MODE == 0 ? 2 : MODE == 2 ? 6 : 8; };

#ifndef PUSH_MIN_CTAS
#define PUSH_MIN_CTAS 4
#endif

template <int MODE>
__global__ void __launch_bounds__(32 * MAX_WPC, PUSH_MIN_CTAS)
k_gather_push(SpeciesDev sp, Geom g, PushParams pp,
		const __grid_constant__ CUtensorMap mapEx, const __grid_constant__ CUtensorMap mapEy,
		const __grid_constant__ CUtensorMap mapS5, const __grid_constant__ CUtensorMap mapS6,
		int nb, int cur, int cta0, int *__restrict__ errflag)
{
	constexpr int NARR = PipeArrays<MODE>::N;
	extern __shared__ __align__(128) unsigned char smem[];
	uint64_t *bar = (uint64_t *) smem;                                  /* E tile barrier */
	uint64_t *sbar = (uint64_t *) (smem + 16) + (threadIdx.x >> 5) * PIPE_STAGES;   /* per warp, per stage */
	int *wscratch = (int *) (smem + PUSH_SMEM_HEADER) + (threadIdx.x >> 5) * 32;   /* 32 ints per warp */
	double *tEx = (double *) (smem + PUSH_SMEM_HEADER + MAX_WPC * 32 * sizeof(int));
	double *tEy = tEx + g.tile_dbl;
	/* per-warp ring of PIPE_STAGES batches x NARR arrays x 32 lanes */
	double *ring = tEy + g.tile_dbl + (threadIdx.x >> 5) * (PIPE_STAGES * NARR * 32);

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned lt = (1u << lane) - 1;
	const int ncx = g.nbx / g.WPC;
	/* cta0: a launch can take a band of block rows (cpic_b200_step_host_banded) */
	const int cta = (int) blockIdx.x + cta0;
	const int by = cta / ncx, cx = cta % ncx;
	const int bx = cx * g.WPC + warp;
	const int b = by * g.nbx + bx;
	int *ocnt = wscratch + 18;       /* leavers per destination code so far */
	if(lane < 9) ocnt[lane] = 0;
	if(lane < PIPE_STAGES) mbar_init(sbar + lane, 1);
	__syncwarp();

	if(MODE != 1 && threadIdx.x == 0)
	{
		mbar_init(bar, 1);
		uint32_t bytes = 2u * (uint32_t) (g.TH * g.TW) * 8u;
		mbar_expect_tx(bar, bytes);
		tma_load_2d(tEx, &mapEx, cx * g.WPC * g.BX, by * g.BY, bar);
		tma_load_2d(tEy, &mapEy, cx * g.WPC * g.BX, by * g.BY, bar);
	}

	/* MODE 0 leaves the arrivals where they are (outbox `cur`); a push consumes the
	 * previous push's outbox and fills `cur` */
	const Outbox &in = sp.ob[MODE == 0 ? cur : cur ^ 1];
	const Outbox &out = sp.ob[cur];
	const int cnt = sp.count[b];
	const unsigned base = (unsigned) b * (unsigned) sp.cap;      /* slot indices fit 32 bits (checked on the host) */
	const int tx0 = cx * g.WPC * g.BX, ty0 = by * g.BY;   /* tile origin in cells */
	const int gby = g.brow0 + by;
	/* the walk: nbo batches over the own segment, then nba over the arrivals */
	const int nbo = (cnt + 31) >> 5;
	Arrivals A;                      /* filled below, once the first loads are in flight */
	A.total = 0; A.start = wscratch; A.src = wscratch + 9;
	int w = 0;                       /* write cursor of the segment */
	int bad = 0;
	unsigned idmask = 0;             /* bit (bi % PIPE_STAGES): the ids of that staged batch were fetched */

	/* Stage batch bi: every lane copies its own element of every array into the ring
	 * (asynchronously, no registers held); with_id also fetches the ids. */
#define ISSUE_BATCH(bi, with_id) do { \
	const bool own_ = (bi) < nbo; \
	double *st0_ = ring + ((bi) % PIPE_STAGES) * (NARR * 32); \
	if(with_id) idmask |= 1u << ((bi) % PIPE_STAGES); else idmask &= ~(1u << ((bi) % PIPE_STAGES)); \
	if(own_ && !OWN_BULK) { \
		/* a whole batch of the segment is 256 contiguous, aligned bytes per array = 16 chunks of \
		 * 16 B; chunk c (array c/16) is copied by lane c%32: x,y | ux,uy | uz,(id) | (E_x,E_y) */ \
		const unsigned q_ = base + (bi) * 32 + (lane & 15) * 2; \
		double *d_ = st0_ + lane * 2; \
		const unsigned h_ = lane >> 4; \
		/* array a of the batch starts a*SEG_STEP doubles after its x values */ \
		const double *g_ = sp.x + seg_slot(sp.cap, b, (bi) * 32) + (lane & 15) * 2 + (size_t) h_ * SEG_STEP; \
		cp_async16(d_, g_); \
		if(MODE != 0) { \
			cp_async16(d_ + 64, g_ + 2 * SEG_STEP); \
			if(h_ == 0 || (with_id)) cp_async16(d_ + 128, g_ + 4 * SEG_STEP); \
		} \
		if(MODE == 1) cp_async16(d_ + 192, (h_ ? sp.pEy : sp.pEx) + q_); \
		cp_async_commit(); \
	} else if(own_) { \
		/* a whole batch of the segment is 256 contiguous, aligned bytes per array: one TMA \
		 * bulk copy each, issued by one lane, landing on the stage's mbarrier */ \
		if(lane == 0) { \
			const unsigned q_ = base + (bi) * 32; \
			uint64_t *mb_ = sbar + (bi) % PIPE_STAGES; \
			const int na_ = (MODE == 0 ? 2 : 5) + ((MODE != 0 && (with_id)) ? 1 : 0) + (MODE == 1 ? 2 : 0); \
			mbar_expect_tx(mb_, (uint32_t) na_ * 256u); \
			const seg_index_t g_ = seg_slot(sp.cap, b, (bi) * 32); \
			if(SEG_TENSOR && MODE == 2) { \
				/* the segment arrays as a 2D tensor (slot, array): the batch's five or six arrays are \
				 * one box, one TMA instruction */ \
				tma_load_2d(st0_, (with_id) ? &mapS6 : &mapS5, (int) g_, 0, mb_); \
			} else if(SEG_AOSOA) { \
				/* the batch's arrays are contiguous: one copy */ \
				tma_bulk_load(st0_, sp.x + g_, (MODE == 0 ? 2u : (with_id) ? 6u : 5u) * 256u, mb_); \
			} else { \
			tma_bulk_load(st0_ + 0 * 32, sp.x + g_, 256, mb_); tma_bulk_load(st0_ + 1 * 32, sp.y + g_, 256, mb_); \
			if(MODE != 0) { tma_bulk_load(st0_ + 2 * 32, sp.ux + g_, 256, mb_); tma_bulk_load(st0_ + 3 * 32, sp.uy + g_, 256, mb_); \
				tma_bulk_load(st0_ + 4 * 32, sp.uz + g_, 256, mb_); if(with_id) tma_bulk_load(st0_ + 5 * 32, sp.id + g_, 256, mb_); } \
			} \
			if(MODE == 1) { tma_bulk_load(st0_ + 6 * 32, sp.pEx + q_, 256, mb_); tma_bulk_load(st0_ + 7 * 32, sp.pEy + q_, 256, mb_); } \
		} \
	} else { \
		/* arrivals are scattered over up to eight runs of records: per lane three 16-byte async \
		 * copies, landing as (x,y) (ux,uy) (uz,id) pairs, lane l at doubles 2l, 2l+1 of each array pair */ \
		const int t_ = ((bi) - nbo) * 32 + lane; \
		double *st_ = st0_ + 2 * lane; \
		if(t_ < A.total) { \
			const size_t q_ = arrival_slot(A, sp, t_); \
			const double *r_ = in.rec + q_ * OREC; \
			cp_async16(st_, r_); \
			if(MODE != 0) { cp_async16(st_ + 64, r_ + 2); cp_async16(st_ + 128, r_ + 4); } \
			if(MODE == 1) cp_async16(st_ + 192, in.recE + q_ * 2); \
		} \
		cp_async_commit(); \
	} } while(0)

	/* Prologue: the first batches of the own segment go out before anything else is known;
	 * the neighbours' counters (the arrivals) are fetched while those loads and the tile are
	 * in flight. Ids are only needed by particles that change slot: every arrival, and
	 * segment particles once a leaver has opened a gap. A block that receives particles
	 * almost surely loses some as well, so such a block fetches its ids from the first batch
	 * on; a quiet block fetches them from the point where the write cursor lags (one late
	 * fetch covers the batches already in flight). */
#pragma unroll
	for(int k = 0; k < PIPE_STAGES - 1; k++)
		if(k < nbo) ISSUE_BATCH(k, true);
	A = find_arrivals(in, sp.nob, g, nb, b, lane, wscratch);
	const bool busy = A.total > 0;
	const int nba = (A.total + 31) >> 5, nbt = nbo + nba;
#pragma unroll
	for(int k = 0; k < PIPE_STAGES - 1; k++)
		if(k >= nbo)
		{
			if(k < nbt) ISSUE_BATCH(k, true);
			else cp_async_commit();
		}

	if(MODE != 1)
	{
		__syncthreads();
		if(mbar_wait(bar, 0))
		{
			if(threadIdx.x == 0) atomicOr(errflag, ERRBIT_TMA);
			cp_async_wait<0>();
			return;
		}
	}

	for(int bi = 0; bi < nbt; bi++)
	{
		const bool own = bi < nbo;                       /* warp-uniform */
		const int t = (own ? bi : bi - nbo) * 32 + lane; /* index inside the segment / the arrivals */
		const bool valid = t < (own ? cnt : A.total);

		/* the stage refilled below was read in the previous iteration: MODE 0 has no ballot in
		 * between that would order those reads before lane 0's refill */
		if(MODE == 0) __syncwarp();
		/* keep the pipeline full: batch bi + PIPE_STAGES - 1. The ids are needed only by
		 * particles that change slot: every arrival, and segment particles once the write
		 * cursor lags the read position (a late fetch covers the batch where that starts). */
		{
			const int bn = bi + PIPE_STAGES - 1;
			if(bn < nbt) ISSUE_BATCH(bn, busy || bn >= nbo || w != bi * 32);
			else cp_async_commit();
		}
		/* batch bi has landed: segment batches on their stage barrier (its parity flips every
		 * reuse), arrival batches when their own async-copy group is the oldest but
		 * PIPE_STAGES-1 (every batch index >= nbo, real or not, commits exactly one group) */
		if(own && OWN_BULK)
		{
			if(mbar_wait(sbar + bi % PIPE_STAGES, (bi / PIPE_STAGES) & 1)) { atomicOr(errflag, ERRBIT_TMA); break; }
		}
		else
		{
			cp_async_wait<PIPE_STAGES - 1>();
			/* a lane waits for its own copies only: the chunks of a segment batch were fetched by other lanes */
			if(own) __syncwarp();
		}

		/* a segment batch is staged array by array (lane l at l of each), an arrival batch
		 * record by record (lane l at 2l, 2l+1 of each array pair) */
		const double *st = ring + (bi % PIPE_STAGES) * (NARR * 32);
		const int e0 = own ? lane : 2 * lane, e1 = own ? 32 + lane : 2 * lane + 1;
		const bool have_id = (idmask >> (bi % PIPE_STAGES)) & 1;
		/* source slot: needed for the E store of MODE 0 and the late id fetch (segment only) */
		const unsigned s = own ? base + t : ((MODE == 0 && valid) ? (unsigned) arrival_slot(A, sp, t) : 0u);
		double x = 0, y = 0, ux = 0, uy = 0, uz = 0, Ex = 0, Ey = 0;
		long long pid = 0;
		if(valid)
		{
			x = st[e0]; y = st[e1];
			if(MODE != 0) { ux = st[64 + e0]; uy = st[64 + e1]; uz = st[128 + e0]; if(have_id) pid = __double_as_longlong(st[128 + e1]); }
			if(MODE == 1) { Ex = st[192 + e0]; Ey = st[192 + e1]; }
		}

		if(MODE != 1 && valid)
		{
			int i0x, i0y;
			double w00, w01, w10, w11;
			cic_weights(g, x, y, i0x, i0y, w00, w01, w10, w11);
			Ex = tile_gather(tEx, g.TW, i0x - tx0, i0y - ty0, w00, w01, w10, w11);
			Ey = tile_gather(tEy, g.TW, i0x - tx0, i0y - ty0, w00, w01, w10, w11);
			if(MODE == 0)
			{
				if(own) { sp.pEx[s] = Ex; sp.pEy[s] = Ey; }
				else *(double2 *) (in.recE + (size_t) s * 2) = make_double2(Ex, Ey);
			}
		}

		if(MODE == 0) continue;

		int dest = DEST_STAY;
		if(valid)
		{
			boris(pp, Ex, Ey, ux, uy, uz);
			/* reference src/mover.c:97-137 aborts; here the flag is raised and the
			 * host reports it at the next synchronisation */
			if(fabs(ux) > pp.umax_x || fabs(uy) > pp.umax_y || fabs(uz) > pp.umax_z) bad = 1;

			if(pp.set_r)
			{
				x = FMA(ux, pp.dt, x);
				y = FMA(uy, pp.dt, y);
				/* periodic_boundary_ppack, reference src/comm_plasma.c:725-747 */
				if(x >= g.Lx) x = SUB(x, g.Lx); else if(x < 0.0) x = ADD(x, g.Lx);
				if(y >= g.Ly) y = SUB(y, g.Ly); else if(y < 0.0) y = ADD(y, g.Ly);

				const int ncxb = cell_ix(g, x) >> g.lBX;
				const int ngby = global_row(g, y) >> g.lBY;
				if(ncxb != bx || ngby != gby)
				{
					const int ddx = ring_delta(ncxb, bx, g.nbx);
					const int ddy = ring_delta(ngby, gby, g.nby_glob);
					if(ddx < -1 || ddx > 1 || ddy < -1 || ddy > 1) dest = DEST_FAR;
					else dest = (ddy + 1) * 3 + (ddx + 1);
				}
			}
		}

		const bool stay = valid && dest == DEST_STAY;
		const bool leave = valid && !stay;
		const unsigned ms = __ballot_sync(FULL, stay);
		const int dpos = w + __popc(ms & lt);
		/* the id (and the kept E) only move when the particle changes slot */
		const bool moved = !own || dpos != t;
		if(!have_id && valid && (moved || leave)) pid = sp.id[seg_slot(sp.cap, b, t)];      /* late fetch: first shifted batch only (segment) */
		__syncwarp();                /* every id is read before a neighbour lane may overwrite its slot */

		if(stay)
		{
			if(dpos < sp.cap)
			{
				const unsigned d = base + dpos;
				const seg_index_t ds = seg_slot(sp.cap, b, dpos);
				if(pp.set_r || moved) { sp.x[ds] = x; sp.y[ds] = y; }
				sp.ux[ds] = ux; sp.uy[ds] = uy; sp.uz[ds] = uz;
				if(moved) sp.id[ds] = pid;
				if(sp.pEx) { sp.pEx[d] = Ex; sp.pEy[d] = Ey; }
			}
			else bad |= 2;
		}
		w += __popc(ms);

		const unsigned ml = __ballot_sync(FULL, leave);
		if(ml)
		{
			/* the leavers that share my destination: one ballot per bit of the code (0..9) */
			unsigned peers = ml;
#pragma unroll
			for(int k = 0; k < 4; k++)
			{
				const unsigned bk = __ballot_sync(FULL, leave && ((dest >> k) & 1));
				peers &= ((dest >> k) & 1) ? bk : ~bk;
			}
			/* rank inside the destination region: batches in order, lanes in order */
			const int pos = leave ? ocnt[dest] + __popc(peers & lt) : 0;
			__syncwarp();
			if(leave)
			{
				if((peers & lt) == 0) ocnt[dest] = pos + __popc(peers);
				/* a leaver whose region is full takes the far movers' road: listed now, placed in its
				 * new block by k_far_insert (in id order, so the result stays deterministic) */
				if(dest == DEST_FAR || pos >= sp.rcap[dest])
				{
					const int k = atomicAdd(sp.fcount, 1);
					if(k < FAR_CAP)
					{
						sp.fx[k] = x; sp.fy[k] = y;
						sp.fux[k] = ux; sp.fuy[k] = uy; sp.fuz[k] = uz;
						sp.fid[k] = pid;
						sp.fEx[k] = Ex; sp.fEy[k] = Ey;
					}
					else bad |= 16;
				}
				else
				{
					/* a leaver that crosses a slab face goes straight to the ghost outbox row of the
					 * neighbour rank (peer memory): what row 0 sends north is the north rank's south ghost
					 * row, what the last row sends south the south rank's north ghost row */
					const Outbox *po = &out;
					int rb = b;
					if(dest < 3 && by == 0 && sp.pob[0][cur].rec) { po = &sp.pob[0][cur]; rb = nb + g.nbx + bx; }
					else if(dest > 5 && by == g.nby - 1 && sp.pob[1][cur].rec) { po = &sp.pob[1][cur]; rb = nb + bx; }
					const size_t o = region_slot(sp, dest, rb, pos);
					double2 *r = (double2 *) (po->rec + o * OREC);
					r[0] = make_double2(x, y);
					r[1] = make_double2(ux, uy);
					r[2] = make_double2(uz, __longlong_as_double(pid));
					if(po->recE) *(double2 *) (po->recE + o * 2) = make_double2(Ex, Ey);
				}
			}
			__syncwarp();
		}
	}
#undef ISSUE_BATCH
	cp_async_wait<0>();

	if(MODE != 0)
	{
		if(lane == 0) sp.count[b] = w < sp.cap ? w : sp.cap;
		if(lane < 9)
		{
			const int v = ocnt[lane];
			const int rc = sp.rcap[lane];
			int n = v < rc ? v : rc;
			/* the counters of the regions that were written into a neighbour rank's ghost row */
			if(lane < 3 && by == 0 && sp.pob[0][cur].rec)
			{
				sp.pob[0][cur].count[(size_t) lane * sp.nob + nb + g.nbx + bx] = n;
				n = 0;
			}
			else if(lane > 5 && by == g.nby - 1 && sp.pob[1][cur].rec)
			{
				sp.pob[1][cur].count[(size_t) lane * sp.nob + nb + bx] = n;
				n = 0;
			}
			out.count[(size_t) lane * sp.nob + b] = n;
		}
		if(bad)
		{
			int bits = 0;
			if(bad & 1) bits |= ERRBIT_VELOCITY;
			if(bad & 2) bits |= ERRBIT_CAPACITY;
			if(bad & 4) bits |= ERRBIT_FAR;
			if(bad & 8) bits |= ERRBIT_REGION;
			if(bad & 16) bits |= ERRBIT_FARLIST;
			atomicOr(errflag, bits);
		}
	}
}

/* Far movers: the list filled by the push (in arbitrary order) is sorted by (destination
 * block, particle id) by one CTA -- a bitonic sort, in shared memory when the list is
 * short, in global scratch otherwise -- and every block's run is appended to its segment
 * in that order, so the result does not depend on the order of arrival. A destination
 * outside this rank's slab is an error (the reference bounds a step to one chunk,
 * src/sim.c:198-200; here the bound across a slab face is one block row). */
#define FAR_SMEM 2048

template <typename K, typename I>
__device__ __forceinline__ void
bitonic_sort(K *key, I *idx, int m)
{
	for(int k = 2; k <= m; k <<= 1)
		for(int j = k >> 1; j > 0; j >>= 1)
		{
			for(int i = threadIdx.x; i < m; i += blockDim.x)
			{
				const int l = i ^ j;
				if(l > i)
				{
					const bool up = (i & k) == 0;
					const K a = key[i], b = key[l];
					if((a > b) == up)
					{
						key[i] = b; key[l] = a;
						const I t = idx[i]; idx[i] = idx[l]; idx[l] = t;
					}
				}
			}
			__syncthreads();
		}
}

static __global__ void __launch_bounds__(1024)
k_far_insert(const __grid_constant__ SpeciesSet set, Geom g, int *__restrict__ errflag)
{
	__shared__ long long skey[FAR_SMEM];
	__shared__ int sidx[FAR_SMEM];
	const SpeciesDev &sp = set.sp[blockIdx.x];       /* one CTA per species */
	int n = *sp.fcount;
	if(n == 0) return;
	if(n > FAR_CAP) n = FAR_CAP;
	int m = 1;
	while(m < n) m <<= 1;
	long long *key = m <= FAR_SMEM ? skey : sp.fkey;
	int *idx = m <= FAR_SMEM ? sidx : sp.fidx;
	/* keys: local destinations (block << 40 | id) < REMOTE <= north rank < south rank < OUT */
	const long long REMOTE = 1LL << 62, OUT = 0x7ffffffffffffffeLL, PAD = 0x7fffffffffffffffLL;
	for(int i = threadIdx.x; i < m; i += blockDim.x)
	{
		long long k = PAD;
		if(i < n)
		{
			const double x = sp.fx[i], y = sp.fy[i];
			const int row = global_row(g, y);
			const long long pid = sp.fid[i] & 0xffffffffffLL;
			if(row >= g.row0 && row < g.row0 + g.ny)
				k = ((long long) block_of(g, x, y) << 40) | pid;
			else
			{
				/* rows to walk north from the slab's first row / south from its last row */
				const int dn = (g.row0 - row + g.ny_glob) % g.ny_glob;
				const int ds = (row - (g.row0 + g.ny - 1) + g.ny_glob) % g.ny_glob;
				const int dir = dn <= ds ? 0 : 1;
				if((dir == 0 ? dn : ds) > g.ny || g.ny_glob == g.ny) { k = OUT; atomicOr(errflag, ERRBIT_FAR); }
				else k = REMOTE | ((long long) dir << 60) | pid;
			}
		}
		key[i] = k;
		idx[i] = i;
	}
	__syncthreads();
	bitonic_sort(key, idx, m);
	/* the first entry of every block's run appends the whole run */
	for(int i = threadIdx.x; i < n; i += blockDim.x)
	{
		const long long k = key[i];
		if(k >= REMOTE) continue;
		const int b = (int) (k >> 40);
		if(i > 0 && (int) (key[i - 1] >> 40) == b) continue;
		int pos = sp.count[b];
		for(int j = i; j < n && key[j] < REMOTE && (int) (key[j] >> 40) == b; j++)
		{
			if(pos >= sp.cap) { atomicOr(errflag, ERRBIT_ABSORB); break; }
			const int e = idx[j];
			const size_t d = (size_t) b * sp.cap + pos, ds = seg_slot(sp.cap, b, pos);
			sp.x[ds] = sp.fx[e]; sp.y[ds] = sp.fy[e];
			sp.ux[ds] = sp.fux[e]; sp.uy[ds] = sp.fuy[e]; sp.uz[ds] = sp.fuz[e];
			sp.id[ds] = sp.fid[e];
			if(sp.pEx) { sp.pEx[d] = sp.fEx[e]; sp.pEy[d] = sp.fEy[e]; }
			pos++;
		}
		sp.count[b] = pos;
	}
	/* particles for the neighbour ranks, in id order, into the lists that travel with the
	 * particle faces (comm.cu) */
	if(threadIdx.x == 0)
	{
		int cn[2] = { sp.rfcount[0], sp.rfcount[1] };
		for(int j = 0; j < n; j++)
		{
			const long long k = key[j];
			if(k < REMOTE || k >= OUT) continue;
			const int dir = (int) ((k >> 60) & 1);
			if(cn[dir] >= FAR_FACE) { atomicOr(errflag, ERRBIT_FARLIST); continue; }
			const int e = idx[j];
			double *o = sp.rfar[dir] + cn[dir];
			o[0 * FAR_FACE] = sp.fx[e]; o[1 * FAR_FACE] = sp.fy[e];
			o[2 * FAR_FACE] = sp.fux[e]; o[3 * FAR_FACE] = sp.fuy[e]; o[4 * FAR_FACE] = sp.fuz[e];
			o[5 * FAR_FACE] = __longlong_as_double(sp.fid[e]);
			o[6 * FAR_FACE] = sp.fEx[e]; o[7 * FAR_FACE] = sp.fEy[e];
			cn[dir]++;
		}
		sp.rfcount[0] = cn[0];
		sp.rfcount[1] = cn[1];
	}
	__syncthreads();
	if(threadIdx.x == 0) *sp.fcount = 0;
}

/* Far movers across a slab face. pack != 0: the list for direction `dir` goes into the
 * face buffer section `buf` ([count as double][8 x FAR_FACE values]) and is emptied; else the
 * section received from a neighbour is appended to the local far-mover list, which the next
 * k_far_insert places. */
__device__ __forceinline__ void
far_face(const SpeciesDev &sp, int dir, double *__restrict__ buf, int pack, int *__restrict__ errflag)
{
	__shared__ int base;
	if(pack)
	{
		const int n = sp.rfcount[dir];
		for(int i = threadIdx.x; i < 8 * FAR_FACE; i += blockDim.x)
			if(i % FAR_FACE < n) buf[1 + i] = sp.rfar[dir][i];
		__syncthreads();
		if(threadIdx.x == 0) { buf[0] = (double) n; sp.rfcount[dir] = 0; }
		return;
	}
	const int n = (int) buf[0];
	if(threadIdx.x == 0) base = atomicAdd(sp.fcount, n);
	__syncthreads();
	if(base + n > FAR_CAP)
	{
		if(threadIdx.x == 0) atomicOr(errflag, ERRBIT_FARLIST);
		return;
	}
	for(int i = threadIdx.x; i < n; i += blockDim.x)
	{
		const double *e = buf + 1 + i;
		const int k = base + i;
		sp.fx[k] = e[0 * FAR_FACE]; sp.fy[k] = e[1 * FAR_FACE];
		sp.fux[k] = e[2 * FAR_FACE]; sp.fuy[k] = e[3 * FAR_FACE]; sp.fuz[k] = e[4 * FAR_FACE];
		sp.fid[k] = __double_as_longlong(e[5 * FAR_FACE]);
		sp.fEx[k] = e[6 * FAR_FACE]; sp.fEy[k] = e[7 * FAR_FACE];
	}
}

/* Moves every block's particles (own segment, then pending arrivals) into a species laid
 * out with a larger capacity `dst`. Not on the per-step path: capacity management only. */
static __global__ void __launch_bounds__(256)
k_regrow(SpeciesDev sp, SpeciesDev dst, Geom g, int nb, int arr)
{
	__shared__ int scratch[8][18];
	const int lane = threadIdx.x & 31;
	const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if(b >= nb) return;
	const Outbox &in = sp.ob[arr];
	const Arrivals A = find_arrivals(in, sp.nob, g, nb, b, lane, scratch[threadIdx.x >> 5]);
	const int cnt = sp.count[b];
	const size_t base = (size_t) b * sp.cap, dbase = (size_t) b * dst.cap;
	for(int i = lane; i < cnt + A.total; i += 32)
	{
		const size_t d = dbase + i;
		if(i >= dst.cap) break;
		const size_t dd = seg_slot(dst.cap, b, i);
		if(i < cnt)
		{
			const size_t s = base + i, ss = seg_slot(sp.cap, b, i);
			dst.x[dd] = sp.x[ss]; dst.y[dd] = sp.y[ss];
			dst.ux[dd] = sp.ux[ss]; dst.uy[dd] = sp.uy[ss]; dst.uz[dd] = sp.uz[ss];
			dst.id[dd] = sp.id[ss];
			if(dst.pEx && sp.pEx) { dst.pEx[d] = sp.pEx[s]; dst.pEy[d] = sp.pEy[s]; }
		}
		else
		{
			const size_t s = arrival_slot(A, sp, i - cnt);
			const double *r = in.rec + s * OREC;
			dst.x[dd] = r[0]; dst.y[dd] = r[1];
			dst.ux[dd] = r[2]; dst.uy[dd] = r[3]; dst.uz[dd] = r[4];
			dst.id[dd] = __double_as_longlong(r[5]);
			if(dst.pEx && in.recE) { dst.pEx[d] = in.recE[s * 2]; dst.pEy[d] = in.recE[s * 2 + 1]; }
		}
	}
	if(lane == 0) dst.count[b] = min(cnt + A.total, dst.cap);
}

/* Appends the pending arrivals of every block (outbox `arr`) to its own segment and
 * clears the runs it consumed. Not on the per-step path: used before particles are
 * handed to the host (downloads, diagnostics). One warp per block. */
static __global__ void __launch_bounds__(256)
k_absorb(SpeciesDev sp, Geom g, int nb, int arr, int *__restrict__ errflag, int b0, int b1)
{
	const int lane = threadIdx.x & 31;
	const int b = b0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     /* blocks [b0, b1) */
	if(b >= b1) return;
	__shared__ int scratch[8][18];
	const Outbox &in = sp.ob[arr];
	const Arrivals A = find_arrivals(in, sp.nob, g, nb, b, lane, scratch[threadIdx.x >> 5]);
	if(A.total == 0) return;
	const int cnt = sp.count[b];
	const size_t base = (size_t) b * sp.cap;
	if(cnt + A.total > sp.cap)
	{
		if(lane == 0) atomicOr(errflag, ERRBIT_ABSORB);
		return;
	}
	for(int f = lane; f < A.total; f += 32)
	{
		const size_t so = arrival_slot(A, sp, f), d = base + cnt + f, ds = seg_slot(sp.cap, b, cnt + f);
		const double *r = in.rec + so * OREC;
		sp.x[ds] = r[0]; sp.y[ds] = r[1];
		sp.ux[ds] = r[2]; sp.uy[ds] = r[3]; sp.uz[ds] = r[4];
		sp.id[ds] = __double_as_longlong(r[5]);
		if(sp.pEx) { sp.pEx[d] = in.recE[so * 2]; sp.pEy[d] = in.recE[so * 2 + 1]; }
	}
	__syncwarp();                /* every lane has read the old count */
	if(lane == 0) sp.count[b] = cnt + A.total;
	/* the runs are consumed: lane k clears the counter it read */
	if(lane < 9 && lane != DEST_STAY)
	{
		in.count[(size_t) (8 - lane) * sp.nob + A.src[lane]] = 0;
	}
}

/* interpolate_p2f_rho, reference src/interpolate.c:282-346 / :161-276, accumulate-correct.
 *
 * k_deposit: one warp sums one particle block at a time -- species after species, own segment,
 * then the arrivals pending in the outbox -- into private accumulators in shared memory,
 * indexed by NODE of the block's (BX+1) x (BY+1) tile and replicated in `ncol` COLUMNS (16 when
 * the tile allows): lane l adds to column l % ncol of the four nodes around its particle.
 * A column is 8 bytes wide and the columns of a node are adjacent, so the 16 lanes of a
 * half-warp hit 16 different bank pairs whatever cells their particles sit in: every
 * shared-memory access of the accumulation is conflict free (one wavefront per half-warp), and
 * no two lanes of a half-warp share an address. Lanes l and l + ncol do share a column: the
 * 32 / ncol groups take turns (warp barrier in between). No votes, no shuffles, no atomics;
 * every accumulator has a fixed order of additions (species, batch, group, the lane's two
 * particles), and the columns of a node are added up in a fixed (rotated, again conflict-free)
 * order. The node sums of the block go to `tiles` (block-major); warps are independent of each
 * other (no CTA barrier, blocks handed out round-robin to a persistent grid), and a block's sums
 * do not depend on which warp formed them.
 *
 * k_rho_assemble: every node of the slab from the (up to four) block tiles that share it, in a
 * fixed order; rho_reset (src/field.c:163-210) is implicit, and on one rank the ghost row is
 * folded into row 0 on the way (src/comm_field.c:51-136: the send to oneself). */
#define DEP_MAX_SPECIES 8
#define DEP_MAX_COLS 16
#ifndef DEP_FUSED
#define DEP_FUSED 1
#endif
#ifndef DEP_WARPS
#define DEP_WARPS 4              /* warps per CTA: shared memory per CTA = DEP_WARPS tiles */
#endif

/* What the deposit reads of one species */
struct DepositSpecies {
	const double *x, *y;     /* segments */
	const int *count;
	const double *arec;      /* the outbox that holds the pending arrivals (records, x y first) */
	const int *acount;
	double vq;               /* -q / e0, reference src/interpolate.c:307 */
	int cap, nob;
	unsigned roff[9];
	int rcap[9];
};
struct DepositSet {
	DepositSpecies s[DEP_MAX_SPECIES];
	int n;
};

/* Weights and first accumulator of a particle, then the lane's turn at the accumulators */
struct DepContribution {
	double a00, a01, a10, a11;
	int n0;                  /* accumulator of node (lx, ly), this lane's column; -1: no particle */
};

__device__ __forceinline__ DepContribution
dep_contribution(const Geom &g, double x, double y, double vq, bool valid, int cx0, int cy0, int NW, int ncol, int col)
{
	DepContribution c;
	int i0x, i0y;
	double w00, w01, w10, w11;
	cic_weights(g, x, y, i0x, i0y, w00, w01, w10, w11);
	c.a00 = MUL(w00, vq); c.a01 = MUL(w01, vq); c.a10 = MUL(w10, vq); c.a11 = MUL(w11, vq);
	c.n0 = valid ? ((i0y - cy0) * NW + (i0x - cx0)) * ncol + col : -1;
	return c;
}

/* corners: 00 = (x, y), 10 = (x+1, y), 01 = (x, y+1), 11 = (x+1, y+1) */
__device__ __forceinline__ void
dep_add(double *t, const DepContribution &c, int ncol, int rowstep)
{
	if(c.n0 < 0) return;
	double *q0 = t + c.n0, *q1 = q0 + rowstep;
	const double v00 = q0[0], v10 = q0[ncol], v01 = q1[0], v11 = q1[ncol];
	q0[0] = ADD(v00, c.a00);
	q0[ncol] = ADD(v10, c.a10);
	q1[0] = ADD(v01, c.a01);
	q1[ncol] = ADD(v11, c.a11);
}

/* NCOL > 0: the number of columns is a compile-time constant (the shipped 16: two groups of a
 * half-warp each); NCOL == 0: `ncol_rt` columns, any power of two up to 16 (large particle blocks) */
template <bool FIRST, int NCOL>
__global__ void __launch_bounds__(32 * DEP_WARPS)
k_deposit(const __grid_constant__ DepositSet set, Geom g, int nb, int ncol_rt, double *__restrict__ tiles)
{
	extern __shared__ __align__(128) unsigned char smem[];
	const int ncol = NCOL > 0 ? NCOL : ncol_rt;
	const int NW = g.BX + 1;                     /* nodes per tile row */
	const int NN = NW * (g.BY + 1);              /* nodes per block tile */
	const int wsz = NN * ncol;                   /* accumulators per warp: [node][column] */
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	double *t = (double *) smem + warp * wsz;
	const int col = lane & (ncol - 1), grp = lane / ncol, ngrp = 32 / ncol;
	const int rowstep = NW * ncol;
	__shared__ int scratch_[DEP_WARPS][20];
	int *scratch = scratch_[warp];

	/* the lane's turn: groups one after the other, a warp barrier in between */
#define DEP_TURN(BODY) do { \
	if(NCOL == 16) { \
		if(grp == 0) { BODY } __syncwarp(); if(grp == 1) { BODY } __syncwarp(); \
	} else for(int p_ = 0; p_ < ngrp; p_++) { if(grp == p_) { BODY } __syncwarp(); } } while(0)

	for(int b = blockIdx.x * DEP_WARPS + warp; b < nb; b += gridDim.x * DEP_WARPS)
	{
		const int bx = b % g.nbx, by = b / g.nbx;
		const int cx0 = bx * g.BX, cy0 = by * g.BY;
		for(int k = lane; k < wsz; k += 32) t[k] = 0.0;
		__syncwarp();

		for(int is = 0; is < set.n; is++)
		{
			const DepositSpecies &sp = set.s[is];
			const double vq = sp.vq;
			const int cnt = sp.count[b];
			/* element i of the block's segment: plain layout: base + i (measured 6 % faster than forming
			 * the slot number every time); batch-major: seg_slot */
#if SEG_AOSOA
			const double *__restrict__ sx = sp.x, *__restrict__ sy = sp.y;
#define SEG(i) seg_slot(sp.cap, b, (i))
#else
			const double *__restrict__ sx = sp.x + seg_slot(sp.cap, b, 0);
			const double *__restrict__ sy = sp.y + seg_slot(sp.cap, b, 0);
#define SEG(i) (i)
#endif
			/* own segment: two particles per lane and turn; four such pairs per lane are kept in flight
			 * in four register slots that are refilled right after use (no rotation: a slot's loads have
			 * three turns to arrive) */
			double px[4][2], py[4][2];
#pragma unroll
			for(int q = 0; q < 4; q++)
#pragma unroll
				for(int e = 0; e < 2; e++)
				{
					const int kk = lane + 64 * q + 32 * e;
					px[q][e] = 0.0; py[q][e] = 0.0;
					if(kk < cnt) { px[q][e] = sx[SEG(kk)]; py[q][e] = sy[SEG(kk)]; }
				}
			/* the arrival counters travel while the segment is walked */
			__syncwarp();
			const Arrivals A = find_arrivals(sp.acount, sp.nob, g, nb, b, lane, scratch);
			for(int k0 = 0; k0 < cnt; k0 += 256)
			{
#pragma unroll
				for(int q = 0; q < 4; q++)
				{
					const int k = k0 + 64 * q + lane;
					if(k - lane >= cnt) break;           /* warp-uniform */
					const DepContribution ca = dep_contribution(g, px[q][0], py[q][0], vq, k < cnt, cx0, cy0, NW, ncol, col);
					const DepContribution cb = dep_contribution(g, px[q][1], py[q][1], vq, k + 32 < cnt, cx0, cy0, NW, ncol, col);
#pragma unroll
					for(int e = 0; e < 2; e++)
					{
						const int kk = k + 256 + 32 * e;
						px[q][e] = 0.0; py[q][e] = 0.0;
						if(kk < cnt) { px[q][e] = sx[SEG(kk)]; py[q][e] = sy[SEG(kk)]; }
					}
					DEP_TURN(dep_add(t, ca, ncol, rowstep); dep_add(t, cb, ncol, rowstep););
				}
			}
			/* arrivals: (x, y) are the first 16 bytes of a 48-byte record; two per lane and turn as well,
			 * the next two on their way */
			auto arrival = [&](int f) -> double2
			{
				if(f >= A.total) return make_double2(0.0, 0.0);
				return *(const double2 *) (sp.arec + (size_t) arrival_slot(A, sp.roff, sp.rcap, f) * OREC);
			};
			double2 va = arrival(lane), vb = arrival(lane + 32);
			for(int f = lane; f - lane < A.total; f += 64)
			{
				const double2 na = arrival(f + 64), nb2 = arrival(f + 96);
				const DepContribution ca = dep_contribution(g, va.x, va.y, vq, f < A.total, cx0, cy0, NW, ncol, col);
				const DepContribution cb = dep_contribution(g, vb.x, vb.y, vq, f + 32 < A.total, cx0, cy0, NW, ncol, col);
				DEP_TURN(dep_add(t, ca, ncol, rowstep); dep_add(t, cb, ncol, rowstep););
				va = na; vb = nb2;
			}
		}
#undef DEP_TURN
#undef SEG

		/* the columns of every node, in a fixed order that starts at a different column for
		 * neighbouring nodes (conflict-free reads) */
		double *out = tiles + (size_t) b * NN;
		for(int n = lane; n < NN; n += 32)
		{
			const double *q = t + n * ncol;
			double s = 0.0;
			for(int k = 0; k < ncol; k++) s = ADD(s, q[(k + n) & (ncol - 1)]);
			if(FIRST) out[n] = s;
			else out[n] += s;
		}
		__syncwarp();
	}
}

/* Grid node (x, y) of the slab, y in [0, ny] (row ny: the south ghost row of `_rho`), from the
 * block tiles that hold it: the tile of block (x / BX, y / BY) at local node (x % BX, y % BY),
 * and, on a block boundary, the right / bottom edge of the block before it (X periodic: node nx
 * is node 0, src/interpolate.c:161-276). Order: upper left, upper right, lower left, lower right
 * block. fold: row 0 also takes the ghost row (one rank). */
static __global__ void
k_rho_assemble(const double *__restrict__ tiles, double *__restrict__ rho, Geom g, int fold)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = blockIdx.y;                /* 0 .. ny */
	if(x >= g.nx) return;
	const int NW = g.BX + 1, NN = NW * (g.BY + 1);
	auto node = [&](int yy)
	{
		const int bxr = x >> g.lBX, lx = x & (g.BX - 1);
		const int bxl = lx == 0 ? (bxr == 0 ? g.nbx - 1 : bxr - 1) : -1;
		const int byr = yy < g.ny ? yy >> g.lBY : -1, ly = yy & (g.BY - 1);
		const int byl = (ly == 0 && yy > 0) ? (yy >> g.lBY) - 1 : -1;
		double v = 0.0;
		if(byl >= 0)
		{
			const double *tl = tiles + (size_t) byl * g.nbx * NN + g.BY * NW;
			if(bxl >= 0) v = ADD(v, tl[(size_t) bxl * NN + g.BX]);
			v = ADD(v, tl[(size_t) bxr * NN + lx]);
		}
		if(byr >= 0)
		{
			const double *tr = tiles + (size_t) byr * g.nbx * NN + ly * NW;
			if(bxl >= 0) v = ADD(v, tr[(size_t) bxl * NN + g.BX]);
			v = ADD(v, tr[(size_t) bxr * NN + lx]);
		}
		return v;
	};
	double v = node(y);
	if(fold && y == 0) v = ADD(v, node(g.ny));
	rho[(size_t) y * g.S + x] = v;
}

/* Compact image of a species for host round trips: the live particles of every block, block
 * after block, without the slack of the segments. pack 1: segments -> image; 0: image ->
 * segments, replacing their content; -1: image appended to the segments (streamed
 * initialisation). The image holds n values per array (x y ux uy uz id), `off` the
 * exclusive prefix of the block counts. One warp per block. */
static __global__ void __launch_bounds__(256)
k_image_copy(SpeciesDev sp, int nb, const int *__restrict__ cnt, const long long *__restrict__ off,
		double *__restrict__ img, long long n, int pack)
{
	const int lane = threadIdx.x & 31;
	const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if(b >= nb) return;
	const int c = cnt[b];
	const long long o = off[b];
	const int at = pack < 0 ? sp.count[b] : 0;       /* first slot written */
	double *arr[6] = { sp.x, sp.y, sp.ux, sp.uy, sp.uz, (double *) sp.id };
#pragma unroll
	for(int a = 0; a < 6; a++)
		for(int i = lane; i < c; i += 32)
		{
			if(pack > 0) img[(size_t) a * n + o + i] = arr[a][seg_slot(sp.cap, b, i)];
			else arr[a][seg_slot(sp.cap, b, at + i)] = img[(size_t) a * n + o + i];
		}
	__syncwarp();                /* every lane has read the old count */
	if(pack <= 0 && lane == 0) sp.count[b] = at + c;
}

/* k_image_copy for a band of blocks [b0, b0 + nblk): `cnt` and `off` are band-local (entry 0 = block b0),
 * the band's image holds `stride` slots per array. pack 1: segments -> image (cnt = the device counts of
 * the band); 0: image -> segments, replacing their content. */
static __global__ void __launch_bounds__(256)
k_band_copy(SpeciesDev sp, int b0, int nblk, const int *__restrict__ cnt, const long long *__restrict__ off,
		double *__restrict__ img, long long stride, int pack, int *__restrict__ errflag)
{
	const int lane = threadIdx.x & 31;
	const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if(k >= nblk) return;
	const int b = b0 + k;
	const int c = cnt[k];
	const long long o = off[k];
	if(o + c > stride) { if(lane == 0) atomicOr(errflag, 1); return; }      /* the band outgrew its image */
	double *arr[6] = { sp.x, sp.y, sp.ux, sp.uy, sp.uz, (double *) sp.id };
#pragma unroll
	for(int a = 0; a < 6; a++)
		for(int i = lane; i < c; i += 32)
		{
			if(pack) img[(size_t) a * stride + o + i] = arr[a][seg_slot(sp.cap, b, i)];
			else arr[a][seg_slot(sp.cap, b, i)] = img[(size_t) a * stride + o + i];
		}
	if(!pack && lane == 0) sp.count[b] = c;
}

/* Exclusive prefix of the block counts (the offsets of the compact image), by one CTA: every thread
 * sums a contiguous share, the shares are scanned in shared memory, then written out. */
static __global__ void __launch_bounds__(1024)
k_count_offsets(const int *__restrict__ cnt, int nb, long long *__restrict__ off)
{
	__shared__ long long part[1024];
	const int per = (nb + 1023) / 1024;
	const int b0 = threadIdx.x * per, b1 = min(nb, b0 + per);
	long long sum = 0;
	for(int b = b0; b < b1; b++) sum += cnt[b];
	part[threadIdx.x] = sum;
	__syncthreads();
	for(int o = 1; o < 1024; o <<= 1)
	{
		const long long v = (int) threadIdx.x >= o ? part[threadIdx.x - o] : 0;
		__syncthreads();
		part[threadIdx.x] += v;
		__syncthreads();
	}
	long long run = part[threadIdx.x] - sum;
	for(int b = b0; b < b1; b++) { off[b] = run; run += cnt[b]; }
	if(threadIdx.x == 1023) off[nb] = part[1023];
}

/* Particles in the HOST's order (the drop-in binding keeps the reference's lists, which never
 * reorder): pos[id] is the place of particle `id` in the host's walk over its lists; entry
 * pos[id] of each of the seven output arrays (x y ux uy uz E_x E_y, n values each) receives the
 * particle. The permutation is made here, so that the host writes its lists front to back. */
static __global__ void __launch_bounds__(256)
k_gather_ordered(SpeciesDev sp, int nb, const int *__restrict__ pos, long long npos, double *__restrict__ out,
		long long n, int *__restrict__ errflag)
{
	const int lane = threadIdx.x & 31;
	const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if(b >= nb) return;
	const int c = sp.count[b];
	for(int i = lane; i < c; i += 32)
	{
		const seg_index_t q = seg_slot(sp.cap, b, i);
		const long long id = sp.id[q];
		const long long p = (id >= 0 && id < npos) ? pos[id] : -1;
		if(p < 0 || p >= n) { atomicOr(errflag, 1); continue; }
		out[p] = sp.x[q]; out[n + p] = sp.y[q];
		out[2 * n + p] = sp.ux[q]; out[3 * n + p] = sp.uy[q]; out[4 * n + p] = sp.uz[q];
		const size_t e = (size_t) b * sp.cap + i;
		out[5 * n + p] = sp.pEx ? sp.pEx[e] : 0.0;
		out[6 * n + p] = sp.pEy ? sp.pEy[e] : 0.0;
	}
}

/* Kinetic energy per species: sum(ux^2 + uy^2), reference src/sim.c:366-395 (compiled
 * out there). One warp per block, fixed order; block sums land in out[b]. */
static __global__ void __launch_bounds__(256)
k_kinetic(SpeciesDev sp, int nb, double *__restrict__ out)
{
	const int lane = threadIdx.x & 31;
	const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if(b >= nb) return;
	const int cnt = sp.count[b];
	double acc = 0.0;
	for(int i = lane; i < cnt; i += 32)
	{
		double ux = sp.ux[seg_slot(sp.cap, b, i)], uy = sp.uy[seg_slot(sp.cap, b, i)];
		acc += ux * ux + uy * uy;
	}
	for(int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(FULL, acc, o);
	if(lane == 0) out[b] = acc;
}

/* Potential energy: per-row sum of rho*phi over the slab (reference src/sim.c:356-363) */
static __global__ void __launch_bounds__(256)
k_potential(const double *__restrict__ rho, const double *__restrict__ phi, Geom g,
		double *__restrict__ out)
{
	__shared__ double part[8];
	const int y = blockIdx.x;
	double acc = 0.0;
	for(int x = threadIdx.x; x < g.nx; x += blockDim.x)
		acc += rho[(size_t) y * g.S + x] * phi[(size_t) (y + 1) * g.S + x];
	for(int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(FULL, acc, o);
	if((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
	__syncthreads();
	if(threadIdx.x == 0)
	{
		double s = 0.0;
		for(int w = 0; w < (int) (blockDim.x >> 5); w++) s += part[w];
		out[y] = s;
	}
}

/* Fixed-order sum of n doubles by one CTA (n is a few thousand: block/row partials) */
static __global__ void __launch_bounds__(1024)
k_sum(const double *__restrict__ in, int n, double *__restrict__ out)
{
	__shared__ double part[1024];
	double acc = 0.0;
	for(int i = threadIdx.x; i < n; i += blockDim.x) acc += in[i];
	part[threadIdx.x] = acc;
	__syncthreads();
	for(int o = 512; o > 0; o >>= 1)
	{
		if((int) threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
		__syncthreads();
	}
	if(threadIdx.x == 0) *out = part[0];
}

/* ---- the reference's own initial conditions, drawn on the device ----
 * One run = the particles one (process, chunk, species) triple of the reference initialises in one
 * go (src/plasma.c:292-316, src/particle.c:178-213): ids first, first+step, ...; "random position"
 * draws four rand() values per particle (x, y, ux, uy: src/particle.c:69-73) from the process's glibc
 * stream, "position delta" places particle i at WRAP(r0 + dr*i, L) with the drift velocity
 * (src/particle.c:150-158). Every thread takes REFINIT_K consecutive particles of the run; the host
 * hands it the state of the generator at its first draw (host/glibc_rand.h: the recurrence
 * r[i] = r[i-31] + r[i-3] advanced by matrix powers). Bit for bit the values of the host
 * initialiser (host/front.cpp: generate_particles). */
#define REFINIT_K 256
#define REFINIT_THREADS 128
struct RefInitRun {
	int method;              /* 0 "random position", 1 "position delta" */
	long long first, step;   /* particle ids of the run */
	double v[2], dr[2], r0[2];
	double Lx, Ly;
};

static __global__ void __launch_bounds__(REFINIT_THREADS)
k_refinit_generate(RefInitRun run, long long k0, long long n, const unsigned *__restrict__ states,
		double *__restrict__ x, double *__restrict__ y, double *__restrict__ ux, double *__restrict__ uy,
		long long *__restrict__ id)
{
	__shared__ unsigned buf[31][REFINIT_THREADS];
	const int tid = threadIdx.x;
	const long long t = (long long) blockIdx.x * blockDim.x + tid;
	const long long j0 = t * REFINIT_K;
	if(j0 >= n) return;
	if(run.method == 0)
		for(int i = 0; i < 31; i++) buf[i][tid] = states[t * 31 + i];
	int p = 0;                       /* buf[p]: the oldest of the last 31 values */
	auto draw = [&]() -> double
	{
		const int q = p + 28 >= 31 ? p - 3 : p + 28;
		const unsigned v = buf[p][tid] + buf[q][tid];
		buf[p][tid] = v;
		p = p == 30 ? 0 : p + 1;
		/* rand() / (RAND_MAX + 1.0): exact */
		return (double) (v >> 1) * (1.0 / 2147483648.0);
	};
	const long long j1 = j0 + REFINIT_K < n ? j0 + REFINIT_K : n;
	for(long long j = j0; j < j1; j++)
	{
		const long long i = run.first + (k0 + j) * run.step;
		double px, py, pux, puy;
		if(run.method == 0)
		{
			/* uniform(a, b) = rand() / (RAND_MAX + 1.0) * (b - a) + a, src/particle.c:17-21 */
			px = ADD(MUL(draw(), SUB(run.Lx, 0.0)), 0.0);
			py = ADD(MUL(draw(), SUB(run.Ly, 0.0)), 0.0);
			pux = ADD(MUL(draw(), SUB(run.v[0], -run.v[0])), -run.v[0]);
			puy = ADD(MUL(draw(), SUB(run.v[1], -run.v[1])), -run.v[1]);
		}
		else
		{
			/* the reference build fuses r0 + dr*i (see host/front.cpp) */
			px = fmod(FMA(run.dr[0], (double) i, run.r0[0]), run.Lx);
			if(px < 0.0) px = ADD(px, run.Lx);
			py = fmod(FMA(run.dr[1], (double) i, run.r0[1]), run.Ly);
			if(py < 0.0) py = ADD(py, run.Ly);
			pux = run.v[0]; puy = run.v[1];
		}
		x[j] = px; y[j] = py; ux[j] = pux; uy[j] = puy; id[j] = i;
	}
}

/* Particle block of every particle of a batch: over all slabs (tally: capacity, pass 1) or inside
 * this rank's slab (key, value = position in the batch; a particle of another slab gets key nb) */
static __global__ void
k_refinit_keys(const double *__restrict__ x, const double *__restrict__ y, long long n, Geom g, int nb,
		int *__restrict__ tally, int *__restrict__ key, int *__restrict__ val, int *__restrict__ cnt)
{
	const long long j = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if(j >= n) return;
	/* what the host binning does with a particle on the upper edge (particle_comm_initial wraps it) */
	const double px = x[j] >= g.Lx ? x[j] - g.Lx : x[j], py = y[j] >= g.Ly ? y[j] - g.Ly : y[j];
	const int row = global_row(g, py);
	const int bxg = cell_ix(g, px) >> g.lBX;
	if(tally) atomicAdd(&tally[(size_t) (row >> g.lBY) * g.nbx + bxg], 1);
	if(key)
	{
		const bool mine = row >= g.row0 && row < g.row0 + g.ny;
		const int k = mine ? ((row - g.row0) >> g.lBY) * g.nbx + bxg : nb;
		key[j] = k;
		val[j] = (int) j;
		atomicAdd(&cnt[k], 1);
	}
}

/* The batch in block order (val: the stable sort's permutation) as the compact image k_image_copy appends */
static __global__ void
k_refinit_gather(const int *__restrict__ val, long long n_in, const double *__restrict__ x, const double *__restrict__ y,
		const double *__restrict__ ux, const double *__restrict__ uy, const long long *__restrict__ id,
		double *__restrict__ img, Geom g)
{
	const long long j = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if(j >= n_in) return;
	const int s = val[j];
	img[j] = x[s] >= g.Lx ? x[s] - g.Lx : x[s];
	img[n_in + j] = y[s] >= g.Ly ? y[s] - g.Ly : y[s];
	img[2 * n_in + j] = ux[s];
	img[3 * n_in + j] = uy[s];
	img[4 * n_in + j] = 0.0;
	img[5 * n_in + j] = __longlong_as_double(id[s]);
}

/* Throughput-only initialiser: particle k of block b sits uniformly inside block b, so
 * the plasma is uniform with equal block populations; u ~ U(-v, v) per axis as in the
 * reference's "random position" (src/particle.c:72-73), plus a drift (a beam: the
 * "position delta" initialiser gives every particle the drift velocity, src/particle.c:152-153).
 * Counter-based (splitmix64). */
__device__ __forceinline__ double
u01(uint64_t &s)
{
	s += 0x9e3779b97f4a7c15ULL;
	uint64_t z = s;
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
	z ^= z >> 31;
	return (double) (z >> 11) * (1.0 / 9007199254740992.0);
}

static __global__ void __launch_bounds__(256)
k_init_uniform(SpeciesDev sp, Geom g, int nb, long long n, long long id0, double vx, double vy,
		double dux, double duy, uint64_t seed)
{
	const int lane = threadIdx.x & 31;
	const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if(b >= nb) return;
	const long long cnt = n / nb + (b < n % nb ? 1 : 0);
	const int bx = b % g.nbx, by = b / g.nbx;
	for(long long k = lane; k < cnt; k += 32)
	{
		long long gid = id0 + k * nb + b;
		uint64_t s = seed ^ ((uint64_t) gid * 0xd1342543de82ef95ULL);
		double fx = u01(s), fy = u01(s);
		double x = ((double) (bx * g.BX) + fx * g.BX) * g.dx;
		double y = g.y0 + ((double) (by * g.BY) + fy * g.BY) * g.dy;
		/* keep the particle strictly inside its block whatever the rounding */
		if(block_of(g, x, y) != b)
		{
			x = ((double) (bx * g.BX) + 0.5 * g.BX) * g.dx;
			y = g.y0 + ((double) (by * g.BY) + 0.5 * g.BY) * g.dy;
		}
		const size_t ds = seg_slot(sp.cap, b, (int) k);
		sp.x[ds] = x;
		sp.y[ds] = y;
		sp.ux[ds] = dux + (2.0 * u01(s) - 1.0) * vx;
		sp.uy[ds] = duy + (2.0 * u01(s) - 1.0) * vy;
		sp.uz[ds] = 0.0;
		sp.id[ds] = gid;
	}
	if(lane == 0) sp.count[b] = (int) cnt;
}

#endif
