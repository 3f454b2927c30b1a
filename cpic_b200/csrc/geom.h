/* Geometry and per-particle arithmetic shared by the host binning code and the device
 * kernels, so that both place a particle in exactly the same cell. */
#ifndef CPIC_B200_GEOM_H
#define CPIC_B200_GEOM_H

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#else
#define HD static inline
#endif

/* Every floating point operation of the per-particle arithmetic is pinned (no compiler
 * contraction), so that the separately staged kernels and the fused ones, and the host
 * binning code, produce identical bits. */
#ifdef __CUDA_ARCH__
#define MUL(a, b) __dmul_rn((a), (b))
#define ADD(a, b) __dadd_rn((a), (b))
#define SUB(a, b) __dsub_rn((a), (b))
#define FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define MUL(a, b) ((a) * (b))
#define ADD(a, b) ((a) + (b))
#define SUB(a, b) ((a) - (b))
#define FMA(a, b, c) fma((a), (b), (c))
#endif

/* Destination codes of a particle that leaves its particle block: (ddy+1)*3 + (ddx+1)
 * for a move to an adjacent block, 4 = stays. */
#define DEST_STAY 4
#define DEST_FAR 9

/* Deferred error bits (device flag word) */
#define ERRBIT_VELOCITY 1
#define ERRBIT_CAPACITY 2
#define ERRBIT_FAR 4
#define ERRBIT_TMA 8
#define ERRBIT_REGION 16      /* an exchange region overflowed */
#define ERRBIT_FARLIST 32     /* more far movers in one step than the list holds */
#define ERRBIT_ABSORB 64      /* a block could not take its arrivals */
#define ERRBIT_PEER 128       /* a neighbour rank did not arrive at an exchange */

/* One rank's slab and its decomposition into particle blocks of BX x BY cells */
struct Geom {
	int nx, ny;          /* grid points of the slab: all columns, ny = ny_glob / nranks rows */
	int ny_glob;
	int row0;            /* first global row of the slab = rank * ny */
	int S;               /* row stride of rho and phi: 2*(nx/2+1) (reference src/solver.c:381-433) */
	int SE;              /* row stride of the device E arrays: nx + wrap columns, even */
	int BX, BY;          /* cells per particle block (powers of two) */
	int lBX, lBY;        /* their base-2 logarithms */
	int nbx, nby;        /* blocks per slab row / column */
	int nby_glob;        /* nby * nranks */
	int brow0;           /* first global block row = rank * nby */
	int WPC;             /* particle blocks (warps) per CTA, divides nbx */
	int TW, TH;          /* E tile: TW = WPC*BX + 2 columns (even), TH = BY + 1 rows */
	int tile_dbl;        /* doubles between the E_x and E_y tiles in shared memory (128 B multiple) */
	double Lx, Ly;       /* global lengths */
	double dx, dy;       /* reference src/sim.c:172-176 */
	double idx, idy;     /* 1.0 / dx, as the reference forms them (src/interpolate.c:300-303) */
	double y0;           /* slab origin: field.x0[Y] = rank * ny * dy (src/field.c:149-151) */
};

/* Cell of a position, the reference's way (src/interpolate.c:38-66: floor(x * idx), one
 * round-down conversion here), clamped into the domain so that a particle sitting exactly
 * on the upper edge (x == L after a wrap from -eps, src/comm_plasma.c:738-742) stays
 * addressable. Rows are always formed from the GLOBAL coordinate (floor(y * idy)) and then
 * made slab-relative as integers: every kernel and the host binning agree on the cell of a
 * particle whatever the rank, and the arithmetic is that of the single-rank reference. */
/* clamp into [0, hi]: one min and one max instruction on the device */
HD int
clamp0(int c, int hi)
{
#ifdef __CUDA_ARCH__
	return min(max(c, 0), hi);
#else
	return c < 0 ? 0 : (c > hi ? hi : c);
#endif
}

HD int
cell_ix(const Geom &g, double x)
{
#ifdef __CUDA_ARCH__
	int c = __double2int_rd(MUL(x, g.idx));
#else
	int c = (int) floor(x * g.idx);
#endif
	return clamp0(c, g.nx - 1);
}

/* Global row of a position */
HD int
global_row(const Geom &g, double y)
{
#ifdef __CUDA_ARCH__
	int c = __double2int_rd(MUL(y, g.idy));
#else
	int c = (int) floor(y * g.idy);
#endif
	return clamp0(c, g.ny_glob - 1);
}

/* Row inside this rank's slab */
HD int
cell_iy(const Geom &g, double y)
{
	return clamp0(global_row(g, y) - g.row0, g.ny - 1);
}

/* Bilinear (CIC) weights, restating reference src/interpolate.c:77-100 (weights),
 * :38-66 (relative_position_grid) and :11-32 (linear_interpolation). The Y offset uses
 * dx[X] as the reference does (src/interpolate.c:87-88). i0y is slab-relative. */
HD void
cic_weights(const Geom &g, double x, double y, int &i0x, int &i0y,
		double &w00, double &w01, double &w10, double &w11)
{
	double bd, bs, relx, rely, delx, dely;

	bd = x;
	i0x = cell_ix(g, x);
	bs = (double) i0x;
	relx = MUL(SUB(bd, MUL(bs, g.dx)), g.idx);

	/* global row and global offset: (y - row*dx) is what the single-rank reference forms */
	bd = y;
	i0y = cell_iy(g, y);
	bs = (double) (i0y + g.row0);
	rely = MUL(SUB(bd, MUL(bs, g.dx)), g.idy);

	delx = SUB(1.0, relx);
	dely = SUB(1.0, rely);
	w00 = MUL(delx, dely);
	w01 = MUL(delx, rely);
	w10 = MUL(relx, dely);
	w11 = MUL(relx, rely);
}

/* Local particle block of a position inside the slab */
HD int
block_of(const Geom &g, double x, double y)
{
	int cx = cell_ix(g, x);
	int cy = cell_iy(g, y);
	return (cy >> g.lBY) * g.nbx + (cx >> g.lBX);
}

/* Shortest signed distance between two indices on a ring of n */
HD int
ring_delta(int to, int from, int n)
{
	int d = to - from;
	if(2 * d > n) d -= n;
	else if(2 * d < -n) d += n;
	return d;
}

#endif
