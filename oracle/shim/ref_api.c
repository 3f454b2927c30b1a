/* ORACLE TEST INFRASTRUCTURE — not product code.
 *
 * Flat C entry points (ctypes-friendly) around the UNMODIFIED reference objects, which
 * are compiled from /root/reference/src by oracle/Makefile. This file only calls the
 * reference's public functions (src/sim.h, src/field.h, src/particle.h, src/mover.h)
 * and reads its public structures (src/def.h); it holds no algorithm of its own apart
 * from the F1 collision census (a read-only scan of the particle lists).
 *
 * The reference enables floating point traps (src/sim.c:102-106). They are armed on
 * entry to every call and disarmed on exit so that the calling Python process is not
 * left with SIGFPE-on-underflow semantics.
 */
#define _GNU_SOURCE
#include <fenv.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <libgen.h>

#include "sim.h"
#include "field.h"
#include "particle.h"
#include "mover.h"
#include "solver.h"
#include "perf.h"

typedef struct ref_handle {
	config_t conf;
	sim_t *sim;
} ref_handle_t;

#define TRAPS (FE_INVALID | FE_DIVBYZERO | FE_OVERFLOW | FE_UNDERFLOW)

static void enter(void) { feclearexcept(FE_ALL_EXCEPT); feenableexcept(TRAPS); }
static void leave(void) { fedisableexcept(FE_ALL_EXCEPT); feclearexcept(FE_ALL_EXCEPT); }

void *
ref_open(const char *conf_path)
{
	ref_handle_t *h = calloc(1, sizeof(*h));
	char *dup, *dir;
	int prov;

	if(!h) return NULL;
	MPI_Init_thread(NULL, NULL, MPI_THREAD_MULTIPLE, &prov);

	config_init(&h->conf);
	dup = strdup(conf_path);
	dir = dirname(dup);
	config_set_include_dir(&h->conf, dir);
	if(config_read_file(&h->conf, conf_path) == CONFIG_FALSE)
	{
		fprintf(stderr, "ref_open: %s:%d - %s\n", conf_path,
				config_error_line(&h->conf), config_error_text(&h->conf));
		free(dup);
		free(h);
		return NULL;
	}
	free(dup);

	enter();
	h->sim = sim_init(&h->conf, 1);
	leave();
	if(!h->sim) { free(h); return NULL; }
	return h;
}

int ref_step(void *vh)
{
	ref_handle_t *h = vh;
	int rc;
	enter();
	rc = sim_step(h->sim);
	leave();
	return rc;
}

/* The four stages, individually (src/sim.c:503,517,525,536) */
int ref_stage_field_E(void *vh) { ref_handle_t *h = vh; int rc; enter(); rc = stage_field_E(h->sim); leave(); return rc; }
void ref_stage_plasma_E(void *vh) { ref_handle_t *h = vh; enter(); stage_plasma_E(h->sim); leave(); }
void ref_stage_plasma_r(void *vh) { ref_handle_t *h = vh; enter(); stage_plasma_r(h->sim); leave(); }
void ref_stage_field_rho(void *vh) { ref_handle_t *h = vh; enter(); stage_field_rho(h->sim); leave(); }

/* What sim_step does to the clock after the stages (src/sim.c:574-575) */
void ref_advance_iter(void *vh)
{
	ref_handle_t *h = vh;
	h->sim->iter += 1;
	h->sim->t = (double) h->sim->iter * h->sim->dt;
}

long long ref_iter(void *vh) { return ((ref_handle_t *) vh)->sim->iter; }
long long ref_cycles(void *vh) { return ((ref_handle_t *) vh)->sim->cycles; }
int ref_nspecies(void *vh) { return (int) ((ref_handle_t *) vh)->sim->nspecies; }
long long ref_nchunks(void *vh) { return ((ref_handle_t *) vh)->sim->plasma.nchunks; }

/* Scalars: 0 dt, 1 e0, 2 Lx, 3 Ly, 4 dx, 5 dy, 6 Bx, 7 By, 8 Bz, 9 umax_x, 10 umax_y, 11 umax_z */
double ref_scalar(void *vh, int which)
{
	sim_t *s = ((ref_handle_t *) vh)->sim;
	switch(which)
	{
		case 0: return s->dt;
		case 1: return s->e0;
		case 2: return s->L[X];
		case 3: return s->L[Y];
		case 4: return s->dx[X];
		case 5: return s->dx[Y];
		case 6: return s->B[X];
		case 7: return s->B[Y];
		case 8: return s->B[Z];
		case 9: return s->umax[X];
		case 10: return s->umax[Y];
		case 11: return s->umax[Z];
	}
	return NAN;
}

void ref_grid(void *vh, long long *nx, long long *ny)
{
	sim_t *s = ((ref_handle_t *) vh)->sim;
	*nx = s->blocksize[X];
	*ny = s->blocksize[Y];
}

void ref_specie(void *vh, int is, double *q, double *m, long long *n)
{
	sim_t *s = ((ref_handle_t *) vh)->sim;
	*q = s->species[is].q;
	*m = s->species[is].m;
	*n = s->species[is].nparticles;
}

/* Copy the particles of species `is`, chunk by chunk in list order (the order the
 * reference's own loops visit them). Any output pointer may be NULL. Returns the
 * number of particles (also when cap is too small, without writing past cap). */
long long
ref_get_particles(void *vh, int is, long long cap, long long *id,
		double *x, double *y, double *z,
		double *ux, double *uy, double *uz,
		double *Ex, double *Ey, long long *chunk_of)
{
	sim_t *s = ((ref_handle_t *) vh)->sim;
	long long n = 0;
	i64 ic, ip, iv;

	for(ic = 0; ic < s->plasma.nchunks; ic++)
	{
		plist_t *l = &s->plasma.chunks[ic].species[is].list;
		pblock_t *b;
		for(b = l->b; b; b = b->next)
		{
			for(ip = 0; ip < b->npacks; ip++)
			{
				ppack_t *p = &b->p[ip];
				for(iv = 0; iv < MAX_VEC; iv++)
				{
					if(ip * MAX_VEC + iv >= b->n) break;
					if(n < cap)
					{
						if(id) id[n] = p->i[iv];
						if(x) x[n] = p->r[X][iv];
						if(y) y[n] = p->r[Y][iv];
						if(z) z[n] = p->r[Z][iv];
						if(ux) ux[n] = p->u[X][iv];
						if(uy) uy[n] = p->u[Y][iv];
						if(uz) uz[n] = p->u[Z][iv];
						if(Ex) Ex[n] = p->E[X][iv];
						if(Ey) Ey[n] = p->E[Y][iv];
						if(chunk_of) chunk_of[n] = ic;
					}
					n++;
				}
			}
		}
	}
	return n;
}

/* F1 census: number of 4-lane packs, in list order, in which two
 * live lanes share a cell, i.e. where vmat_add_xy (src/simd_avx2.h:226-249) drops
 * deposits; `lost` counts the dropped particle contributions. Tail-pack garbage lanes
 * take part with q=0 exactly as in src/interpolate.c:329-344. */
long long
ref_collision_census(void *vh, long long *lost)
{
	sim_t *s = ((ref_handle_t *) vh)->sim;
	long long packs = 0, nlost = 0;
	i64 ic, is, ip, iv, jv;
	double idx0 = 1.0 / s->dx[X], idx1 = 1.0 / s->dx[Y];

	for(ic = 0; ic < s->plasma.nchunks; ic++)
	for(is = 0; is < s->nspecies; is++)
	{
		plist_t *l = &s->plasma.chunks[ic].species[is].list;
		pblock_t *b;
		for(b = l->b; b; b = b->next)
		{
			for(ip = 0; ip < b->npacks; ip++)
			{
				ppack_t *p = &b->p[ip];
				long long cx[MAX_VEC], cy[MAX_VEC];
				int hit = 0;
				/* Only packs that interpolate_p2f_rho visits: full packs of every
				 * block, and the partial pack of the last block */
				if(ip >= b->nfpacks && b->next) continue;
				for(iv = 0; iv < MAX_VEC; iv++)
				{
					cx[iv] = (long long) floor((p->r[X][iv] - s->field.x0[X]) * idx0);
					cy[iv] = (long long) floor((p->r[Y][iv] - s->field.x0[Y]) * idx1);
				}
				for(iv = 0; iv < MAX_VEC; iv++)
				{
					int live = ip * MAX_VEC + iv < b->n;
					/* lane iv is overwritten if a higher lane has the same cell */
					for(jv = iv + 1; jv < MAX_VEC; jv++)
					{
						if(cx[iv] == cx[jv] && cy[iv] == cy[jv])
						{
							if(live) { nlost++; hit = 1; }
							break;
						}
					}
				}
				packs += hit;
			}
		}
	}
	if(lost) *lost = nlost;
	return packs;
}

/* Field access. which: 0 rho (ny x nx view), 1 phi (ny x nx view), 2 E_X, 3 E_Y
 * ((ny+1) x nx, ghost row included), 4 _rho rows [0, ny+1) x nx (ghost row included),
 * 5 _phi (ny+3 rows x nx: north ghost, slab, two south ghosts). Output is dense
 * row-major rows x nx. Returns the number of rows. */
static mat_t *
pick(sim_t *s, int which, i64 *rows, i64 *row0)
{
	field_t *f = &s->field;
	*row0 = 0;
	switch(which)
	{
		case 0: *rows = s->blocksize[Y]; return f->rho;
		case 1: *rows = s->blocksize[Y]; return f->phi;
		case 2: *rows = s->blocksize[Y] + 1; return f->_E[X];
		case 3: *rows = s->blocksize[Y] + 1; return f->_E[Y];
		case 4: *rows = s->blocksize[Y] + 1; return f->_rho;
		case 5: *rows = s->blocksize[Y] + 3; return f->_phi;
	}
	return NULL;
}

long long
ref_get_field(void *vh, int which, double *out)
{
	sim_t *s = ((ref_handle_t *) vh)->sim;
	i64 rows, row0, ix, iy, nx = s->blocksize[X];
	mat_t *m = pick(s, which, &rows, &row0);
	if(!m) return -1;
	if(out)
		for(iy = 0; iy < rows; iy++)
			for(ix = 0; ix < nx; ix++)
				out[iy * nx + ix] = MAT_XY(m, ix, iy + row0);
	return rows;
}

long long
ref_set_field(void *vh, int which, const double *in)
{
	sim_t *s = ((ref_handle_t *) vh)->sim;
	i64 rows, row0, ix, iy, nx = s->blocksize[X];
	mat_t *m = pick(s, which, &rows, &row0);
	if(!m) return -1;
	for(iy = 0; iy < rows; iy++)
		for(ix = 0; ix < nx; ix++)
			MAT_XY(m, ix, iy + row0) = in[iy * nx + ix];
	return rows;
}

/* Raw padded storage exactly as output.c writes it (src/output.c:627-630):
 * which 0 _rho, 1 _phi, 2 _E[X], 3 _E[Y]. Returns the number of doubles. */
long long
ref_raw_field(void *vh, int which, double *out, long long cap, long long *stride, long long *nrows)
{
	sim_t *s = ((ref_handle_t *) vh)->sim;
	field_t *f = &s->field;
	mat_t *m = which == 0 ? f->_rho : which == 1 ? f->_phi : which == 2 ? f->_E[X] : f->_E[Y];
	long long n = m->real_shape[X] * m->real_shape[Y];
	if(stride) *stride = m->real_shape[X];
	if(nrows) *nrows = m->real_shape[Y];
	if(out && cap >= n) memcpy(out, m->real_data, (size_t) n * sizeof(double));
	return n;
}

/* Seconds accumulated by one of the reference's own timers (src/def.h:394-408) */
double ref_timer(void *vh, int which)
{
	sim_t *s = ((ref_handle_t *) vh)->sim;
	if(which < 0 || which >= MAX_TIMERS) return NAN;
	return perf_measure(&s->timers[which]);
}
