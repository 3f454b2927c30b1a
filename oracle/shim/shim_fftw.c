/* ORACLE TEST INFRASTRUCTURE — not product code.
 *
 * Plain double-precision restatement of the two FFTW-MPI transforms the reference's
 * MFT solver plans (src/solver.c:314-330) and executes (src/solver.c:485,491), for a
 * single rank:
 *   r2c 2D: out[l][k] = sum_{y<n0} sum_{x<n1} in[y][x] exp(-2 pi i (k x/n1 + l y/n0)),
 *           k in [0, n1/2], real rows padded to 2*(n1/2+1) doubles (FFTW manual 4.3.4,
 *           6.5 "Multi-dimensional MPI DFTs of Real Data");
 *   c2r 2D: the unnormalised inverse of the above, imaginary parts of the k=0 and
 *           k=n1/2 columns ignored as FFTW does after the n0 pass.
 * Powers of two use an iterative radix-2 transform with a long-double twiddle table;
 * other lengths fall back to a direct O(n^2) DFT (small test grids only).
 * No value is ever scaled into the denormal range (the reference traps FE_UNDERFLOW,
 * src/sim.c:102-106).
 *
 * -DSHIM_MP (cpic_ref_mp, with shim_mpi_mp.c): the same transforms over P forked ranks, each
 * owning n0/P rows as FFTW-MPI's slab decomposition does (fftw_mpi_local_size_2d): row
 * transforms of the local rows into a work array in the shared mapping, barrier, column
 * transforms of a 1/P share of the columns, barrier, local rows back out.
 */
#define _GNU_SOURCE
#include "fftw3.h"
#include "fftw3-mpi.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef double _Complex cplx;

typedef struct fft1d {
	ptrdiff_t n;
	int pow2;
	cplx *w;          /* w[k] = exp(-2 pi i k / n), k < n (forward sign) */
	ptrdiff_t *rev;   /* bit reversal (pow2 only) */
	cplx *tmp;
} fft1d_t;

#ifdef SHIM_MP
int shim_mp_rank(void);
int shim_mp_size(void);
void shim_mp_barrier(void);
void *shim_mp_shared_alloc(size_t bytes);
#endif

struct shim_fftw_plan {
	int kind;         /* 0 = r2c, 1 = c2r */
	ptrdiff_t n0, n1;
	ptrdiff_t l0, y0; /* local rows and first local row (n0, 0 on one rank) */
	cplx *work;       /* n0 x (n1/2+1): the whole spectrum (shared between the ranks with SHIM_MP) */
	double *real;
	cplx *cpx;
	fft1d_t f0, f1;
	cplx *col;        /* n0 scratch for strided columns */
	cplx *row;        /* n1 scratch */
};

static void
fft1d_init(fft1d_t *f, ptrdiff_t n)
{
	ptrdiff_t k, bits = 0, i;

	f->n = n;
	f->pow2 = (n & (n - 1)) == 0;
	f->w = malloc((size_t) n * sizeof(cplx));
	f->tmp = malloc((size_t) n * sizeof(cplx));
	f->rev = NULL;
	if(!f->w || !f->tmp) abort();

	for(k = 0; k < n; k++)
	{
		long double a = -2.0L * M_PIl * (long double) k / (long double) n;
		double c = (double) cosl(a), s = (double) sinl(a);
		/* Exact zeros and ones at the quadrant points */
		if(4 * k == n) { c = 0.0; s = -1.0; }
		if(2 * k == n) { c = -1.0; s = 0.0; }
		if(4 * k == 3 * n) { c = 0.0; s = 1.0; }
		if(k == 0) { c = 1.0; s = 0.0; }
		f->w[k] = c + s * I;
	}

	if(f->pow2)
	{
		while(((ptrdiff_t) 1 << bits) < n) bits++;
		f->rev = malloc((size_t) n * sizeof(ptrdiff_t));
		if(!f->rev) abort();
		for(i = 0; i < n; i++)
		{
			ptrdiff_t r = 0, b;
			for(b = 0; b < bits; b++)
				if(i & ((ptrdiff_t) 1 << b)) r |= (ptrdiff_t) 1 << (bits - 1 - b);
			f->rev[i] = r;
		}
	}
}

static void
fft1d_free(fft1d_t *f)
{
	free(f->w); free(f->rev); free(f->tmp);
}

/* In-place transform of x[0..n). sign = -1 forward, +1 backward (unnormalised). */
static void
fft1d_exec(fft1d_t *f, cplx *x, int sign)
{
	ptrdiff_t n = f->n, i, j, k, len;

	if(n == 1) return;

	if(!f->pow2)
	{
		for(k = 0; k < n; k++)
		{
			cplx acc = 0.0;
			for(j = 0; j < n; j++)
			{
				cplx w = f->w[(j * k) % n];
				if(sign > 0) w = conj(w);
				acc += x[j] * w;
			}
			f->tmp[k] = acc;
		}
		memcpy(x, f->tmp, (size_t) n * sizeof(cplx));
		return;
	}

	for(i = 0; i < n; i++)
	{
		j = f->rev[i];
		if(i < j) { cplx t = x[i]; x[i] = x[j]; x[j] = t; }
	}

	for(len = 2; len <= n; len <<= 1)
	{
		ptrdiff_t half = len >> 1, step = n / len;
		for(i = 0; i < n; i += len)
		{
			for(k = 0; k < half; k++)
			{
				cplx w = f->w[k * step];
				double wr = creal(w), wi = sign > 0 ? -cimag(w) : cimag(w);
				double ar = creal(x[i + k + half]), ai = cimag(x[i + k + half]);
				double tr = ar * wr - ai * wi;
				double ti = ar * wi + ai * wr;
				double ur = creal(x[i + k]), ui = cimag(x[i + k]);
				x[i + k] = (ur + tr) + (ui + ti) * I;
				x[i + k + half] = (ur - tr) + (ui - ti) * I;
			}
		}
	}
}

static void
exec_r2c(struct shim_fftw_plan *p)
{
	ptrdiff_t n0 = p->n0, n1 = p->n1, nc = n1 / 2 + 1, ld = 2 * nc;
	ptrdiff_t l0 = p->l0;
	ptrdiff_t y, x, k;
	cplx *W = p->work + p->y0 * nc;        /* this rank's rows of the whole spectrum */

	/* Pass 1: rows. Two real rows share one complex transform. */
	for(y = 0; y + 1 < l0; y += 2)
	{
		double *a = p->real + y * ld, *b = a + ld;
		cplx *oa = W + y * nc, *ob = oa + nc;
		for(x = 0; x < n1; x++) p->row[x] = a[x] + b[x] * I;
		fft1d_exec(&p->f1, p->row, -1);
		for(k = 0; k < nc; k++)
		{
			cplx z = p->row[k], zc = conj(p->row[(n1 - k) % n1]);
			oa[k] = 0.5 * (z + zc);
			/* (z - zc) / (2i) */
			cplx d = z - zc;
			ob[k] = 0.5 * (cimag(d) - creal(d) * I);
		}
	}
	if(y < l0)
	{
		double *a = p->real + y * ld;
		cplx *oa = W + y * nc;
		for(x = 0; x < n1; x++) p->row[x] = a[x];
		fft1d_exec(&p->f1, p->row, -1);
		for(k = 0; k < nc; k++) oa[k] = p->row[k];
	}

	/* Pass 2: columns (a share of them per rank) */
	ptrdiff_t k0 = 0, k1 = nc;
#ifdef SHIM_MP
	k0 = nc * shim_mp_rank() / shim_mp_size();
	k1 = nc * (shim_mp_rank() + 1) / shim_mp_size();
	shim_mp_barrier();
#endif
	for(k = k0; k < k1; k++)
	{
		for(y = 0; y < n0; y++) p->col[y] = p->work[y * nc + k];
		fft1d_exec(&p->f0, p->col, -1);
		for(y = 0; y < n0; y++) p->work[y * nc + k] = p->col[y];
	}
#ifdef SHIM_MP
	shim_mp_barrier();
#endif
	if(p->cpx != W) memcpy(p->cpx, W, (size_t) (l0 * nc) * sizeof(cplx));
#ifdef SHIM_MP
	shim_mp_barrier();      /* nobody refills the work array before everybody has copied */
#endif
}

static void
exec_c2r(struct shim_fftw_plan *p)
{
	ptrdiff_t n0 = p->n0, n1 = p->n1, nc = n1 / 2 + 1, ld = 2 * nc;
	ptrdiff_t l0 = p->l0;
	ptrdiff_t y, x, k;
	cplx *W = p->work + p->y0 * nc;

	if(p->cpx != W) memcpy(W, p->cpx, (size_t) (l0 * nc) * sizeof(cplx));

	/* Pass 1: columns, backward (a share of them per rank) */
	ptrdiff_t k0 = 0, k1 = nc;
#ifdef SHIM_MP
	k0 = nc * shim_mp_rank() / shim_mp_size();
	k1 = nc * (shim_mp_rank() + 1) / shim_mp_size();
	shim_mp_barrier();
#endif
	for(k = k0; k < k1; k++)
	{
		for(y = 0; y < n0; y++) p->col[y] = p->work[y * nc + k];
		fft1d_exec(&p->f0, p->col, +1);
		for(y = 0; y < n0; y++) p->work[y * nc + k] = p->col[y];
	}
#ifdef SHIM_MP
	shim_mp_barrier();
#endif

	/* Pass 2: rows, Hermitian-extended, two at a time */
	for(y = 0; y < l0; y += 2)
	{
		int pair = (y + 1 < l0);
		cplx *ia = W + y * nc, *ib = ia + nc;
		double *a = p->real + y * ld, *b = a + ld;

		for(k = 0; k < nc; k++)
		{
			cplx A = ia[k], B = pair ? ib[k] : 0.0;
			if(k == 0 || 2 * k == n1)
			{
				/* A real signal has a real DC and Nyquist term */
				A = creal(A);
				B = creal(B);
			}
			/* Z_k = A_k + i B_k */
			p->row[k] = (creal(A) - cimag(B)) + (cimag(A) + creal(B)) * I;
			if(k != 0 && 2 * k != n1)
			{
				/* Z_{n-k} = conj(A_k) + i conj(B_k) */
				p->row[n1 - k] = (creal(A) + cimag(B)) + (creal(B) - cimag(A)) * I;
			}
		}
		fft1d_exec(&p->f1, p->row, +1);
		for(x = 0; x < n1; x++)
		{
			a[x] = creal(p->row[x]);
			if(pair) b[x] = cimag(p->row[x]);
		}
	}
#ifdef SHIM_MP
	shim_mp_barrier();
#endif
}

void
fftw_execute(const fftw_plan p)
{
	if(p->kind == 0) exec_r2c(p);
	else exec_c2r(p);
}

static fftw_plan
plan_new(int kind, ptrdiff_t n0, ptrdiff_t n1, double *real, cplx *cpx)
{
	struct shim_fftw_plan *p = calloc(1, sizeof(*p));
	if(!p) abort();
	p->kind = kind;
	p->n0 = n0;
	p->n1 = n1;
	p->real = real;
	p->cpx = cpx;
#ifdef SHIM_MP
	p->l0 = n0 / shim_mp_size();
	p->y0 = p->l0 * shim_mp_rank();
	p->work = shim_mp_shared_alloc((size_t) (n0 * (n1 / 2 + 1)) * sizeof(cplx));
#else
	/* one rank: the caller's complex array is the whole spectrum; transformed in place */
	p->l0 = n0;
	p->y0 = 0;
	p->work = cpx;
#endif
	fft1d_init(&p->f0, n0);
	fft1d_init(&p->f1, n1);
	p->col = malloc((size_t) n0 * sizeof(cplx));
	p->row = malloc((size_t) n1 * sizeof(cplx));
	if(!p->col || !p->row) abort();
	return p;
}

fftw_plan
fftw_mpi_plan_dft_r2c_2d(ptrdiff_t n0, ptrdiff_t n1, double *in, fftw_complex *out,
		MPI_Comm comm, unsigned flags)
{
	(void) comm; (void) flags;
	return plan_new(0, n0, n1, in, out);
}

fftw_plan
fftw_mpi_plan_dft_c2r_2d(ptrdiff_t n0, ptrdiff_t n1, fftw_complex *in, double *out,
		MPI_Comm comm, unsigned flags)
{
	(void) comm; (void) flags;
	return plan_new(1, n0, n1, out, in);
}

void
fftw_destroy_plan(fftw_plan p)
{
	if(!p) return;
	fft1d_free(&p->f0);
	fft1d_free(&p->f1);
	free(p->col);
	free(p->row);
	free(p);
}

ptrdiff_t
fftw_mpi_local_size_2d(ptrdiff_t n0, ptrdiff_t n1, MPI_Comm comm,
		ptrdiff_t *local_n0, ptrdiff_t *local_0_start)
{
	(void) comm;
#ifdef SHIM_MP
	*local_n0 = n0 / shim_mp_size();
	*local_0_start = *local_n0 * shim_mp_rank();
	return *local_n0 * n1;
#else
	*local_n0 = n0;
	*local_0_start = 0;
	return n0 * n1;
#endif
}

void fftw_mpi_init(void) {}
int fftw_init_threads(void) { return 1; }
void fftw_plan_with_nthreads(int n) { (void) n; }

void *
fftw_malloc(size_t n)
{
	void *p = NULL;
	if(posix_memalign(&p, 64, n ? n : 64)) abort();
	return p;
}

fftw_complex *fftw_alloc_complex(size_t n) { return fftw_malloc(n * sizeof(fftw_complex)); }
double *fftw_alloc_real(size_t n) { return fftw_malloc(n * sizeof(double)); }
void fftw_free(void *p) { free(p); }
