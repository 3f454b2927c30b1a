/* ORACLE — TEST INFRASTRUCTURE ONLY (see cpic_oracle.h). Plain C, one rank, scalar.
 * Every function cites the reference lines it restates (paths relative to the
 * reference tree). Built with -ffp-contract=off: one rounding per operation. */
#define _GNU_SOURCE
#include "cpic_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define X 0
#define Y 1
#define Z 2

static void *
xcalloc(size_t n, size_t sz)
{
	void *p = calloc(n ? n : 1, sz);
	if(!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
	return p;
}

/* ------------------------------------------------------------------ set-up */

/* src/sim.c:87-206 (sim_prepare), src/field.c:20-160 (field_init),
 * src/solver.c:207-280 (MFT_init: the G table) */
oracle_sim_t *
oracle_create(long long nx, long long ny, double Lx, double Ly, double dt, double e0,
		const double B[3], long long plasma_chunks, int nspecies,
		const double *q, const double *m)
{
	oracle_sim_t *s = xcalloc(1, sizeof(*s));
	long long nc = nx / 2 + 1, ix, iy;
	double cx, cy;
	int i;

	s->nx = nx;
	s->ny = ny;
	s->S = 2 * nc;
	s->L[X] = Lx;
	s->L[Y] = Ly;
	s->dx[X] = Lx / (double) nx;          /* src/sim.c:172-176 */
	s->dx[Y] = Ly / (double) ny;
	s->dt = dt;
	s->e0 = e0;
	memcpy(s->B, B, 3 * sizeof(double));
	/* src/sim.c:198-200: umax = chunksize * dx / dt, chunksize[Z] = 1, dx[Z] = 0 */
	s->umax[X] = (double) (nx / plasma_chunks) * s->dx[X] / dt;
	s->umax[Y] = (double) ny * s->dx[Y] / dt;
	s->umax[Z] = 1.0 * 0.0 / dt;
	s->iter = -1;
	s->nspecies = nspecies;
	s->sp = xcalloc((size_t) nspecies, sizeof(oracle_species_t));
	for(i = 0; i < nspecies; i++) { s->sp[i].q = q[i]; s->sp[i].m = m[i]; }

	s->rho = xcalloc((size_t) ((ny + 1) * s->S), sizeof(double));
	s->phi = xcalloc((size_t) ((ny + 3) * s->S), sizeof(double));
	s->Ex = xcalloc((size_t) ((ny + 1) * nx), sizeof(double));
	s->Ey = xcalloc((size_t) ((ny + 1) * nx), sizeof(double));
	s->G = xcalloc((size_t) (ny * nc), sizeof(double));
	s->gre = xcalloc((size_t) (ny * nc), sizeof(double));
	s->gim = xcalloc((size_t) (ny * nc), sizeof(double));

	/* src/solver.c:257-280 */
	cx = 2.0 * M_PI / (double) nx;
	cy = 2.0 * M_PI / (double) ny;
	for(iy = 0; iy < ny; iy++)
		for(ix = 0; ix < nc; ix++)
		{
			if(ix == 0 && iy == 0) s->G[0] = 0.0;
			else s->G[iy * nc + ix] = 1.0 /
				(2.0 * (cos(cx * (double) ix) + cos(cy * (double) iy)) - 4.0);
		}
	return s;
}

static void
free_species(oracle_species_t *p)
{
	free(p->id); free(p->x); free(p->y); free(p->z);
	free(p->ux); free(p->uy); free(p->uz); free(p->Ex); free(p->Ey);
}

void
oracle_destroy(oracle_sim_t *s)
{
	int i;
	if(!s) return;
	for(i = 0; i < s->nspecies; i++) free_species(&s->sp[i]);
	free(s->sp); free(s->rho); free(s->phi); free(s->Ex); free(s->Ey);
	free(s->G); free(s->gre); free(s->gim);
	free(s);
}

int
oracle_alloc_species(oracle_sim_t *s, int is, long long n)
{
	oracle_species_t *p = &s->sp[is];
	double q = p->q, m = p->m;
	free_species(p);
	memset(p, 0, sizeof(*p));
	p->q = q; p->m = m; p->n = n;
	p->id = xcalloc((size_t) n, sizeof(long long));
	p->x = xcalloc((size_t) n, sizeof(double));
	p->y = xcalloc((size_t) n, sizeof(double));
	p->z = xcalloc((size_t) n, sizeof(double));
	p->ux = xcalloc((size_t) n, sizeof(double));
	p->uy = xcalloc((size_t) n, sizeof(double));
	p->uz = xcalloc((size_t) n, sizeof(double));
	p->Ex = xcalloc((size_t) n, sizeof(double));
	p->Ey = xcalloc((size_t) n, sizeof(double));
	return 0;
}

int
oracle_set_particles(oracle_sim_t *s, int is, long long n, const long long *id,
		const double *x, const double *y, const double *ux, const double *uy,
		const double *uz)
{
	oracle_species_t *p;
	oracle_alloc_species(s, is, n);
	p = &s->sp[is];
	memcpy(p->id, id, (size_t) n * sizeof(long long));
	memcpy(p->x, x, (size_t) n * sizeof(double));
	memcpy(p->y, y, (size_t) n * sizeof(double));
	memcpy(p->ux, ux, (size_t) n * sizeof(double));
	memcpy(p->uy, uy, (size_t) n * sizeof(double));
	if(uz) memcpy(p->uz, uz, (size_t) n * sizeof(double));
	return 0;
}

/* src/sim.c:153: srand(seed + rank) */
void oracle_srand(unsigned int seed) { srand(seed); }

/* src/particle.c:17-21 */
static double
uniform(double a, double b)
{
	return rand() / (RAND_MAX + 1.0) * (b - a) + a;
}

/* src/particle.c:23-89 (init_randpos) for the particles that src/plasma.c:62-128 gives
 * to chunk ic: ids ic, ic+nchunks, ... (one rank). The particle with id i is stored at
 * array slot i, so arrays stay in id order whatever the chunk count. The caller loops
 * chunk-major, species-minor (src/plasma.c:307-314, :275-283) to keep rand() order. */
void
oracle_init_randpos_chunk(oracle_sim_t *s, int is, long long ic, long long nchunks,
		const double v[2])
{
	oracle_species_t *p = &s->sp[is];
	long long i;
	for(i = ic; i < p->n; i += nchunks)
	{
		p->id[i] = i;
		p->x[i] = uniform(0.0, s->L[X]);
		p->y[i] = uniform(0.0, s->L[Y]);
		p->z[i] = 0.0;
		p->ux[i] = uniform(-v[X], v[X]);
		p->uy[i] = uniform(-v[Y], v[Y]);
		p->uz[i] = 0.0;
	}
}

/* src/particle.c:91-168 (init_position_delta); WRAP is src/mat.h:62 */
void
oracle_init_delta(oracle_sim_t *s, int is, const double r0[2], const double dr[2],
		const double v[2])
{
	oracle_species_t *p = &s->sp[is];
	long long i;
	for(i = 0; i < p->n; i++)
	{
		double r;
		p->id[i] = i;
		r = fmod(r0[X] + dr[X] * (double) i, s->L[X]); if(r < 0.0) r += s->L[X];
		p->x[i] = r;
		r = fmod(r0[Y] + dr[Y] * (double) i, s->L[Y]); if(r < 0.0) r += s->L[Y];
		p->y[i] = r;
		p->z[i] = 0.0;
		p->ux[i] = v[X];
		p->uy[i] = v[Y];
		p->uz[i] = 0.0;
	}
}

/* ------------------------------------------------------------ interpolation */

/* src/simd_avx2.h:78-88: double -> i32 -> zero-extended i64 */
static long long
to_index(double f)
{
	int v = (int) f;
	return (long long) (unsigned int) v;
}

/* src/interpolate.c:38-66 (relative_position_grid), :77-100 (weights, including the
 * dx[X]-for-Y quirk at :87-88), :11-32 (linear_interpolation).
 * w = {w00, w01, w10, w11} with wXY: X offset first. Field origin is (0,0) for one rank. */
void
oracle_weights(const oracle_sim_t *s, double x, double y, double w[4],
		long long *i0x, long long *i0y)
{
	double idx = 1.0 / s->dx[X], idy = 1.0 / s->dx[Y];
	double bd, br, bs, gd, relx, rely, delx, dely;

	bd = x - 0.0;
	br = bd * idx;
	bs = floor(br);
	*i0x = to_index(bs);
	gd = bd - bs * s->dx[X];
	relx = gd * idx;

	bd = y - 0.0;
	br = bd * idy;
	bs = floor(br);
	*i0y = to_index(bs);
	gd = bd - bs * s->dx[X];        /* sic: dx[X], src/interpolate.c:87-88 */
	rely = gd * idy;

	delx = 1.0 - relx;
	dely = 1.0 - rely;
	w[0] = delx * dely;
	w[1] = delx * rely;
	w[2] = relx * dely;
	w[3] = relx * rely;
}

/* src/interpolate.c:102-155 (interpolate_f2p): X wraps by remod, Y uses the ghost row */
static double
gather_one(const oracle_sim_t *s, const double *F, const double w[4],
		long long i0x, long long i0y)
{
	long long i1x = i0x + 1, i1y = i0y + 1, nx = s->nx;
	double val;
	if(i1x >= nx) i1x -= nx;              /* src/simd_avx2.h:33-43 */
	val  = w[0] * F[nx * i0y + i0x];
	val += w[1] * F[nx * i1y + i0x];
	val += w[2] * F[nx * i0y + i1x];
	val += w[3] * F[nx * i1y + i1x];
	return val;
}

/* src/particle.c:215-248 (stage_plasma_E) -> src/interpolate.c:349-404 */
void
oracle_stage_plasma_E(oracle_sim_t *s)
{
	int is;
	long long i, i0x, i0y;
	double w[4];
	for(is = 0; is < s->nspecies; is++)
	{
		oracle_species_t *p = &s->sp[is];
		for(i = 0; i < p->n; i++)
		{
			oracle_weights(s, p->x[i], p->y[i], w, &i0x, &i0y);
			p->Ex[i] = gather_one(s, s->Ex, w, i0x, i0y);
			p->Ey[i] = gather_one(s, s->Ey, w, i0x, i0y);
		}
	}
}

/* src/interpolate.c:161-276 (interpolate_p2f), accumulate-correct (no F1 loss) */
static void
deposit_one(oracle_sim_t *s, double x, double y, double vq)
{
	long long i0x, i0y, i1x, i1y, S = s->S;
	double w[4];
	oracle_weights(s, x, y, w, &i0x, &i0y);
	i1x = i0x + 1; i1y = i0y + 1;
	if(i1x >= s->nx) i1x -= s->nx;
	s->rho[S * i0y + i0x] += w[0] * vq;
	s->rho[S * i1y + i0x] += w[1] * vq;
	s->rho[S * i0y + i1x] += w[2] * vq;
	s->rho[S * i1y + i1x] += w[3] * vq;
}

/* src/simd_avx2.h:226-249: gather 4, add, store lane by lane => the highest lane that
 * shares a node wins */
void
oracle_deposit_lossy(oracle_sim_t *s, double q, long long n, const double *x,
		const double *y, double gx, double gy)
{
	double vq0 = -q / s->e0;
	long long S = s->S, ip, iv, c;

	for(ip = 0; ip < (n + 3) / 4; ip++)
	{
		double w[4][4], vq[4];
		long long i0x[4], i0y[4], i1x[4], i1y[4];
		for(iv = 0; iv < 4; iv++)
		{
			long long k = ip * 4 + iv;
			double px = k < n ? x[k] : gx, py = k < n ? y[k] : gy;
			vq[iv] = k < n ? vq0 : 0.0;
			oracle_weights(s, px, py, w[iv], &i0x[iv], &i0y[iv]);
			i1x[iv] = i0x[iv] + 1; i1y[iv] = i0y[iv] + 1;
			if(i1x[iv] >= s->nx) i1x[iv] -= s->nx;
		}
		for(c = 0; c < 4; c++)
		{
			double old[4];
			long long at[4];
			for(iv = 0; iv < 4; iv++)
			{
				long long ix = (c & 2) ? i1x[iv] : i0x[iv];
				long long iy = (c & 1) ? i1y[iv] : i0y[iv];
				at[iv] = S * iy + ix;
				old[iv] = s->rho[at[iv]];
			}
			for(iv = 0; iv < 4; iv++)
				s->rho[at[iv]] = old[iv] + w[iv][c] * vq[iv];
		}
	}
}

/* src/field.c:268-356 (stage_field_rho): rho_reset :163-210, rho_update :215-230 ->
 * src/interpolate.c:282-346 with vq = -q/e0 (:307); then the ghost row is sent to
 * rank+1 and added to row 0 (src/comm_field.c:51-136; one rank: onto itself) */
void
oracle_stage_field_rho(oracle_sim_t *s)
{
	long long ix, iy, i, S = s->S;
	int is;

	for(iy = 0; iy < s->ny + 1; iy++)
		for(ix = 0; ix < s->nx; ix++)
			s->rho[S * iy + ix] = 0.0;

	for(is = 0; is < s->nspecies; is++)
	{
		oracle_species_t *p = &s->sp[is];
		double vq = -p->q / s->e0;
		for(i = 0; i < p->n; i++)
			deposit_one(s, p->x[i], p->y[i], vq);
	}

	for(ix = 0; ix < s->nx; ix++)
		s->rho[ix] += s->rho[S * s->ny + ix];
}

/* ------------------------------------------------------------------ mover */

/* src/mover.c:141-188 (plist_update_r) with :22-70 (boris_rotation), :75-82 (move),
 * :85-93 (update_u), :97-137 (check_velocity); dtqm2 from :191-226 */
static int
update_species(oracle_sim_t *s, oracle_species_t *p, double dt, double dtqm2, int set_r)
{
	long long i;
	int d;
	for(i = 0; i < p->n; i++)
	{
		double u0[3] = { p->ux[i], p->uy[i], p->uz[i] };
		double E[3] = { p->Ex[i], p->Ey[i], 0.0 };
		double t[3], sd[3], sv[3], vm[3], vp[3], vq[3], u[3];

		for(d = 0; d < 3; d++)
		{
			sd[d] = 1.0;
			t[d] = s->B[d] * dtqm2;
			sd[d] += t[d] * t[d];
			vm[d] = u0[d] + dtqm2 * E[d];
			sv[d] = 2.0 * t[d] / sd[d];
		}
		vp[X] = vm[Y] * t[Z] - vm[Z] * t[Y];
		vp[Y] = vm[Z] * t[X] - vm[X] * t[Z];
		vp[Z] = vm[X] * t[Y] - vm[Y] * t[X];
		for(d = 0; d < 3; d++) vp[d] += vm[d];
		vq[X] = vp[Y] * sv[Z] - vp[Z] * sv[Y];
		vq[Y] = vp[Z] * sv[X] - vp[X] * sv[Z];
		vq[Z] = vp[X] * sv[Y] - vp[Y] * sv[X];
		for(d = 0; d < 3; d++)
		{
			vq[d] += vm[d];
			u[d] = vq[d] + dtqm2 * E[d];
		}
		for(d = 0; d < 3; d++)
			if(fabs(u[d]) > s->umax[d]) { s->aborted = 1; return -1; }

		if(set_r)
		{
			p->x[i] += u[X] * dt;
			p->y[i] += u[Y] * dt;
			p->z[i] += u[Z] * dt;
		}
		p->ux[i] = u[X]; p->uy[i] = u[Y]; p->uz[i] = u[Z];
	}
	return 0;
}

/* src/mover.c:331-362 (stage_plasma_r): mover, then comm_plasma (one rank: only the
 * periodic wrap survives, src/comm_plasma.c:725-747, X then Y) */
int
oracle_stage_plasma_r(oracle_sim_t *s)
{
	int is, d;
	long long i;

	for(is = 0; is < s->nspecies; is++)
	{
		oracle_species_t *p = &s->sp[is];
		double dt, dtqm2;
		int set_r;
		if(s->iter == 0) { dt = -s->dt / 2; set_r = 0; }   /* src/mover.c:204-215 */
		else { dt = s->dt; set_r = 1; }
		dtqm2 = 0.5 * dt * p->q / p->m;
		if(update_species(s, p, dt, dtqm2, set_r)) return -1;
	}

	for(is = 0; is < s->nspecies; is++)
	{
		oracle_species_t *p = &s->sp[is];
		for(d = 0; d < 2; d++)
		{
			double *r = d == X ? p->x : p->y;
			double L = s->L[d];
			for(i = 0; i < p->n; i++)
			{
				if(r[i] >= L) r[i] -= L;
				else if(r[i] < 0.0) r[i] += L;
			}
		}
	}
	return 0;
}

/* -------------------------------------------------------------------- FFT */

typedef struct { double re, im; } cpx_t;

/* exp(-2 pi i k/n), k < n/2, from long-double trigonometry; cached per length */
static const cpx_t *
twiddles(long long n)
{
	static struct { long long n; cpx_t *w; } cache[16];
	static int used;
	long long k;
	int i;
	cpx_t *w;
	for(i = 0; i < used; i++) if(cache[i].n == n) return cache[i].w;
	w = xcalloc((size_t) (n / 2 + 1), sizeof(cpx_t));
	for(k = 0; k < n / 2; k++)
	{
		long double ang = -2.0L * M_PIl * (long double) k / (long double) n;
		w[k].re = (double) cosl(ang);
		w[k].im = (double) sinl(ang);
		if(4 * k == n) { w[k].re = 0.0; w[k].im = -1.0; }
	}
	w[0].re = 1.0; w[0].im = 0.0;
	if(used < 16) { cache[used].n = n; cache[used].w = w; used++; }
	return w;
}

/* Unnormalised DFT of length n (power of two: radix-2; otherwise direct sum).
 * sign = -1: exp(-2 pi i jk/n). This stands in for FFTW 3.3.6, a third-party
 * dependency of the reference that is not vendored (src/build.mk:47-48). */
static void
dft(cpx_t *a, long long n, int sign, cpx_t *tmp)
{
	long long i, j, k, len;
	if(n <= 1) return;
	if(n & (n - 1))
	{
		for(k = 0; k < n; k++)
		{
			double sr = 0.0, si = 0.0;
			for(j = 0; j < n; j++)
			{
				long double ang = sign * 2.0L * M_PIl * (long double) ((j * k) % n) / (long double) n;
				double c = (double) cosl(ang), sn = (double) sinl(ang);
				sr += a[j].re * c - a[j].im * sn;
				si += a[j].re * sn + a[j].im * c;
			}
			tmp[k].re = sr; tmp[k].im = si;
		}
		memcpy(a, tmp, (size_t) n * sizeof(cpx_t));
		return;
	}
	for(i = 1, j = 0; i < n; i++)
	{
		long long bit = n >> 1;
		for(; j & bit; bit >>= 1) j ^= bit;
		j ^= bit;
		if(i < j) { cpx_t t = a[i]; a[i] = a[j]; a[j] = t; }
	}
	{
		const cpx_t *w = twiddles(n);
		for(len = 2; len <= n; len <<= 1)
		{
			long long half = len >> 1, step = n / len;
			for(k = 0; k < half; k++)
			{
				double wr = w[k * step].re, wi = sign < 0 ? w[k * step].im : -w[k * step].im;
				for(i = k; i < n; i += len)
				{
					cpx_t u = a[i], v = a[i + half], t;
					t.re = v.re * wr - v.im * wi;
					t.im = v.re * wi + v.im * wr;
					a[i].re = u.re + t.re; a[i].im = u.im + t.im;
					a[i + half].re = u.re - t.re; a[i + half].im = u.im - t.im;
				}
			}
		}
	}
}

/* out[l][k] = sum_y sum_x in[y][x] exp(-2 pi i (kx/nx + ly/ny)), k in [0, nx/2]
 * (FFTW r2c convention, src/solver.c:314-321). Straight separable transform. */
void
oracle_rfft2(long long ny, long long nx, const double *in, long long ld,
		double *ore, double *oim)
{
	long long nc = nx / 2 + 1, n = nx > ny ? nx : ny, x, y, k;
	cpx_t *a = xcalloc((size_t) n, sizeof(cpx_t)), *tmp = xcalloc((size_t) n, sizeof(cpx_t));

	for(y = 0; y < ny; y++)
	{
		for(x = 0; x < nx; x++) { a[x].re = in[y * ld + x]; a[x].im = 0.0; }
		dft(a, nx, -1, tmp);
		for(k = 0; k < nc; k++) { ore[y * nc + k] = a[k].re; oim[y * nc + k] = a[k].im; }
	}
	for(k = 0; k < nc; k++)
	{
		for(y = 0; y < ny; y++) { a[y].re = ore[y * nc + k]; a[y].im = oim[y * nc + k]; }
		dft(a, ny, -1, tmp);
		for(y = 0; y < ny; y++) { ore[y * nc + k] = a[y].re; oim[y * nc + k] = a[y].im; }
	}
	free(a); free(tmp);
}

/* Unnormalised inverse (FFTW c2r, src/solver.c:323-330); the spectrum is taken as
 * Hermitian: only k in [0, nx/2] is read, DC/Nyquist imaginary parts are ignored. */
void
oracle_irfft2(long long ny, long long nx, const double *ire, const double *iim,
		double *out, long long ld)
{
	long long nc = nx / 2 + 1, n = nx > ny ? nx : ny, x, y, k;
	cpx_t *a = xcalloc((size_t) n, sizeof(cpx_t)), *tmp = xcalloc((size_t) n, sizeof(cpx_t));
	double *wre = xcalloc((size_t) (ny * nc), sizeof(double));
	double *wim = xcalloc((size_t) (ny * nc), sizeof(double));

	for(k = 0; k < nc; k++)
	{
		for(y = 0; y < ny; y++) { a[y].re = ire[y * nc + k]; a[y].im = iim[y * nc + k]; }
		dft(a, ny, +1, tmp);
		for(y = 0; y < ny; y++) { wre[y * nc + k] = a[y].re; wim[y * nc + k] = a[y].im; }
	}
	for(y = 0; y < ny; y++)
	{
		for(k = 0; k < nc; k++)
		{
			double re = wre[y * nc + k], im = wim[y * nc + k];
			if(k == 0 || 2 * k == nx) im = 0.0;
			a[k].re = re; a[k].im = im;
			if(k != 0 && 2 * k != nx) { a[nx - k].re = re; a[nx - k].im = -im; }
		}
		dft(a, nx, +1, tmp);
		for(x = 0; x < nx; x++) out[y * ld + x] = a[x].re;
	}
	free(a); free(tmp); free(wre); free(wim);
}

/* ------------------------------------------------------------------ fields */

/* src/solver.c:465-509 (MFT_solve): r2c, MFT_kernel :337-363 (g *= G),
 * c2r, MFT_normalize :365-379 (phi /= nx*ny, N an int) */
void
oracle_solve(oracle_sim_t *s)
{
	long long nc = s->nx / 2 + 1, i, ix, iy;
	double *phi = s->phi + s->S;      /* slab starts at row 1 (PHI_NG_NORTH, src/def.h:11) */
	int N = (int) (s->nx * s->ny);

	oracle_rfft2(s->ny, s->nx, s->rho, s->S, s->gre, s->gim);
	for(i = 0; i < s->ny * nc; i++) { s->gre[i] *= s->G[i]; s->gim[i] *= s->G[i]; }
	oracle_irfft2(s->ny, s->nx, s->gre, s->gim, phi, s->S);
	for(iy = 0; iy < s->ny; iy++)
		for(ix = 0; ix < s->nx; ix++)
			phi[iy * s->S + ix] /= N;
}

/* src/comm_field.c:139-201 on one rank: slab rows 0,1 -> the two south ghost rows,
 * slab row ny-1 -> the north ghost row (padding columns travel too) */
void
oracle_phi_ghosts(oracle_sim_t *s)
{
	long long S = s->S, ny = s->ny;
	double *phi = s->phi + S;
	memcpy(phi + ny * S, phi, (size_t) (2 * S) * sizeof(double));
	memcpy(s->phi, phi + (ny - 1) * S, (size_t) S * sizeof(double));
}

/* src/field.c:358-416 (field_E_compute), rows [0, ny] */
void
oracle_field_E(oracle_sim_t *s)
{
	long long nx = s->nx, ny = s->ny, S = s->S, ix, iy;
	double dx2 = 2 * s->dx[X], dy2 = 2 * s->dx[Y];
	const double *phi = s->phi + S;   /* phi(ix, iy) = phi[iy*S + ix], iy in [-1, ny+1] */

	for(iy = 0; iy < ny + 1; iy++)
		for(ix = 0; ix < nx; ix++)
		{
			long long x0 = (ix + nx - 1) % nx, x1 = (ix + 1) % nx;
			s->Ey[iy * nx + ix] = (phi[(iy - 1) * S + ix] - phi[(iy + 1) * S + ix]) / dy2;
			s->Ex[iy * nx + ix] = (phi[iy * S + x0] - phi[iy * S + x1]) / dx2;
		}
}

/* src/field.c:450-501 (stage_field_E) */
void
oracle_stage_field_E(oracle_sim_t *s)
{
	oracle_solve(s);
	oracle_phi_ghosts(s);
	oracle_field_E(s);
}

/* ------------------------------------------------------------------ driver */

/* src/sim.c:208-236 (sim_pre_step) and :305-317: particles are already in place on one
 * rank (particle_comm_initial only re-bins), so: rho, dummy field solve, iter = 0 */
void
oracle_pre_step(oracle_sim_t *s)
{
	s->iter = -1;
	oracle_stage_field_rho(s);
	oracle_stage_field_E(s);
	s->iter = 0;
}

/* src/sim.c:481-581 (sim_step) */
int
oracle_step(oracle_sim_t *s)
{
	oracle_stage_field_E(s);
	oracle_stage_plasma_E(s);
	if(oracle_stage_plasma_r(s)) return -1;
	oracle_stage_field_rho(s);
	s->iter++;
	return 0;
}

/* src/sim.c:356-399 (conservation_energy, compiled out in the reference) */
double
oracle_kinetic_energy(const oracle_sim_t *s)
{
	double KE = 0.0;
	int is;
	long long i;
	for(is = 0; is < s->nspecies; is++)
	{
		const oracle_species_t *p = &s->sp[is];
		double ke = 0.0;
		for(i = 0; i < p->n; i++) ke += p->ux[i] * p->ux[i] + p->uy[i] * p->uy[i];
		KE += ke * p->m / 2.0;
	}
	return KE;
}

double
oracle_field_energy(const oracle_sim_t *s)
{
	double PE = 0.0;
	long long ix, iy;
	const double *phi = s->phi + s->S;
	for(iy = 0; iy < s->ny; iy++)
		for(ix = 0; ix < s->nx; ix++)
			PE += s->rho[iy * s->S + ix] * phi[iy * s->S + ix];
	return PE;
}
