/* Reference-side binding of cpic_b200: the four stage functions of cpic's sim_step
 * (src/sim.c:503,517,525,536) with their original signatures, implemented over the C ABI of
 * include/cpic_b200.h. A cpic maintainer adds this one file to src/, drops the bodies of
 * stage_field_E / stage_field_rho (src/field.c:268-356,450-501), stage_plasma_E
 * (src/particle.c:232-248) and stage_plasma_r (src/mover.c:331-362), and links
 * libcpic_b200.so; sim.c, cpic.c, the .conf handling, plasma_init and the test drivers stay
 * as they are. It is compiled against the reference's own headers (never copied here).
 *
 * Host/device coherence: the particles and grids live on the GPU between stages. The host
 * structures the reference's drivers read after a step (`chunks[ic].species[is].list.b->p[..]`
 * in test/cyclotron.c:65-73, test/harmonic.c:40-55; `sim->field._rho/_phi/_E` in
 * src/output.c:627-630) are refreshed at the end of stage_field_rho, the last stage of a
 * step, unless CPIC_B200_SYNC=lazy is set, in which case cpic_b200_dropin_sync(sim) does
 * it on demand.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "sim.h"
#include "field.h"
#include "particle.h"
#include "mover.h"
#include "perf.h"

#include "cpic_b200.h"

typedef struct slot { ppack_t *p; int lane; } slot_t;

static struct {
	sim_t *sim;
	cpic_b200_sim_t *gpu;
	slot_t **slots;          /* per species: host slot of the k-th particle of the walk over the lists */
	i64 *count;              /* per species: particles in the lists */
	double *stage[7];        /* pinned: x y ux uy uz E_x E_y of one species, in walk order */
	int lazy;
	int threads;
} D;

static void
fatal(const char *what)
{
	/* the reference's convention for unrecoverable errors (src/log.h:69-70) */
	fprintf(stderr, "cpic_b200: %s: %s\n", what, cpic_b200_last_error());
	abort();
}

/* plist -> SoA upload (after plasma_init and particle_comm_initial have run on the host) */
static void
attach(sim_t *sim)
{
	cpic_b200_params_t p;
	i64 is, ic;

	if(D.gpu) return;
	if(sim->nprocs > 1)
	{
		/* one process = one GPU slab needs the NCCL bootstrap (cpic_b200_comm_id on rank 0, the id
		 * broadcast with MPI_Bcast, cpic_b200_comm_init on every rank) before the first stage, and a
		 * host id -> slot map that follows the particles between ranks: this binding does neither */
		fprintf(stderr, "cpic_b200 drop-in: %d MPI processes: this binding drives one rank on one GPU; "
				"several GPUs run through cpic_b200_sim_from_conf + cpic_b200_comm_init (INTEGRATION.md)\n", sim->nprocs);
		abort();
	}
	memset(&p, 0, sizeof(p));
	p.nx = sim->ntpoints[X];
	p.ny = sim->ntpoints[Y];
	p.Lx = sim->L[X];
	p.Ly = sim->L[Y];
	p.dt = sim->dt;
	p.e0 = sim->e0;
	p.B[0] = sim->B[X]; p.B[1] = sim->B[Y]; p.B[2] = sim->B[Z];
	p.plasma_chunks = sim->plasma_chunks;
	p.nspecies = (int) sim->nspecies;
	for(is = 0; is < sim->nspecies; is++) { p.q[is] = sim->species[is].q; p.m[is] = sim->species[is].m; }
	p.rank = sim->rank;
	p.nranks = sim->nprocs;
	p.device = -1;
	p.keep_particle_E = 1;   /* ppack.E is part of the host structure the drivers read */
	if(cpic_b200_create(&p, &D.gpu)) fatal("create");
	D.sim = sim;
	D.lazy = getenv("CPIC_B200_SYNC") && strcmp(getenv("CPIC_B200_SYNC"), "lazy") == 0;
	D.slots = calloc((size_t) sim->nspecies, sizeof(slot_t *));
	D.count = calloc((size_t) sim->nspecies, sizeof(i64));
	i64 nstage = 0;
	{
		long nc = sysconf(_SC_NPROCESSORS_ONLN);
		D.threads = nc < 1 ? 1 : (nc > 16 ? 16 : (int) nc);
		if(getenv("CPIC_B200_SYNC_THREADS")) D.threads = atoi(getenv("CPIC_B200_SYNC_THREADS"));
		if(D.threads < 1) D.threads = 1;
	}

	for(is = 0; is < sim->nspecies; is++)
	{
		i64 n = 0, k = 0, ip, iv, nmax = sim->species[is].nparticles;
		i64 *id = malloc((size_t) nmax * sizeof(i64));
		double *x = malloc((size_t) nmax * sizeof(double)), *y = malloc((size_t) nmax * sizeof(double));
		double *ux = malloc((size_t) nmax * sizeof(double)), *uy = malloc((size_t) nmax * sizeof(double));
		double *uz = malloc((size_t) nmax * sizeof(double));
		D.slots[is] = calloc((size_t) nmax + 1, sizeof(slot_t));
		for(ic = 0; ic < sim->plasma.nchunks; ic++)
		{
			pblock_t *b;
			for(b = sim->plasma.chunks[ic].species[is].list.b; b; b = b->next)
				for(ip = 0; ip < b->npacks; ip++)
					for(iv = 0; iv < MAX_VEC && ip * MAX_VEC + iv < b->n; iv++)
					{
						ppack_t *pk = &b->p[ip];
						if(k >= nmax) fatal("more particles in the lists than the species declares");
						id[k] = pk->i[iv];
						x[k] = pk->r[X][iv]; y[k] = pk->r[Y][iv];
						ux[k] = pk->u[X][iv]; uy[k] = pk->u[Y][iv]; uz[k] = pk->u[Z][iv];
						D.slots[is][k].p = pk;
						D.slots[is][k].lane = (int) iv;
						k++;
					}
		}
		n = k;
		D.count[is] = n;
		if(n > nstage) nstage = n;
		if(cpic_b200_set_particles(D.gpu, (int) is, n, (const int64_t *) id, x, y, ux, uy, uz)) fatal("set_particles");
		/* the lists never reorder on the host (comm_plasma runs on the device): downloads come back
		 * in this walk's order */
		if(cpic_b200_set_host_order(D.gpu, (int) is, n, (const int64_t *) id)) fatal("set_host_order");
		free(id); free(x); free(y); free(ux); free(uy); free(uz);
	}
	for(int k = 0; k < 7; k++)
	{
		D.stage[k] = cpic_b200_host_alloc((size_t) (nstage + 1) * sizeof(double));
		if(!D.stage[k]) fatal("pinned staging");
	}
}

/* grids -> mat_t */
static void
sync_fields(sim_t *sim)
{
	field_t *f = &sim->field;
	int64_t rows, stride;
	/* same strides as the reference's padded arrays (src/field.c:20-160) */
	if(cpic_b200_field_shape(D.gpu, CPIC_B200_RHO, &rows, &stride) || stride != f->_rho->real_shape[X])
		fatal("rho stride differs from the reference layout");
	if(cpic_b200_get_field(D.gpu, CPIC_B200_RHO, f->_rho->data)) fatal("get rho");
	if(cpic_b200_get_field(D.gpu, CPIC_B200_PHI, f->_phi->data)) fatal("get phi");
	if(cpic_b200_get_field(D.gpu, CPIC_B200_EX, f->_E[X]->data)) fatal("get E_X");
	if(cpic_b200_get_field(D.gpu, CPIC_B200_EY, f->_E[Y]->data)) fatal("get E_Y");
}

/* One thread's share of the walk: entry k of the staged arrays into the k-th slot of the lists */
typedef struct fill_job { const slot_t *slots; i64 k0, k1; } fill_job_t;

static void *
fill_lists(void *arg)
{
	const fill_job_t *j = arg;
	for(i64 k = j->k0; k < j->k1; k++)
	{
		const slot_t s = j->slots[k];
		s.p->r[X][s.lane] = D.stage[0][k]; s.p->r[Y][s.lane] = D.stage[1][k];
		s.p->u[X][s.lane] = D.stage[2][k]; s.p->u[Y][s.lane] = D.stage[3][k]; s.p->u[Z][s.lane] = D.stage[4][k];
		s.p->E[X][s.lane] = D.stage[5][k]; s.p->E[Y][s.lane] = D.stage[6][k];
	}
	return NULL;
}

/* SoA -> plist and grids -> mat_t. The device permutes the particles into the order of the host's
 * lists and copies them into pinned staging arrays (cpic_b200_get_particles_ordered); the lists are
 * then written front to back by a few threads. */
void
cpic_b200_dropin_sync(sim_t *sim)
{
	i64 is;

	if(!D.gpu) return;
	for(is = 0; is < sim->nspecies; is++)
	{
		const i64 n = D.count[is];
		if(cpic_b200_get_particles_ordered(D.gpu, (int) is, n, D.stage[0], D.stage[1], D.stage[2], D.stage[3],
					D.stage[4], D.stage[5], D.stage[6])) fatal("get_particles_ordered");
		int nt = n < 65536 ? 1 : D.threads;
		pthread_t th[16];
		fill_job_t job[16];
		for(int t = 0; t < nt; t++)
		{
			job[t].slots = D.slots[is];
			job[t].k0 = n * t / nt;
			job[t].k1 = n * (t + 1) / nt;
			if(t > 0 && pthread_create(&th[t], NULL, fill_lists, &job[t])) { fill_lists(&job[t]); th[t] = 0; }
		}
		fill_lists(&job[0]);
		for(int t = 1; t < nt; t++) if(th[t]) pthread_join(th[t], NULL);
	}
	sync_fields(sim);
}

int
stage_field_E(sim_t *sim)
{
	attach(sim);
	perf_start(&sim->timers[TIMER_FIELD_E]);
	cpic_b200_set_iter(D.gpu, sim->iter);
	if(cpic_b200_stage_field_E(D.gpu)) fatal("stage_field_E");
	/* output_fields runs right after this stage (src/sim.c:507) and reads the host grids */
	if(!D.lazy && sim->output && sim->output->enabled) sync_fields(sim);
	perf_stop(&sim->timers[TIMER_FIELD_E]);
	return 0;
}

void
stage_plasma_E(sim_t *sim)
{
	attach(sim);
	perf_start(&sim->timers[TIMER_PARTICLE_E]);
	cpic_b200_set_iter(D.gpu, sim->iter);
	if(cpic_b200_stage_plasma_E(D.gpu)) fatal("stage_plasma_E");
	perf_stop(&sim->timers[TIMER_PARTICLE_E]);
}

void
stage_plasma_r(sim_t *sim)
{
	attach(sim);
	perf_start(&sim->timers[TIMER_PARTICLE_X]);
	cpic_b200_set_iter(D.gpu, sim->iter);
	if(cpic_b200_stage_plasma_r(D.gpu)) fatal("stage_plasma_r");
	perf_stop(&sim->timers[TIMER_PARTICLE_X]);
}

void
stage_field_rho(sim_t *sim)
{
	attach(sim);
	perf_start(&sim->timers[TIMER_FIELD_RHO]);
	cpic_b200_set_iter(D.gpu, sim->iter);
	if(cpic_b200_stage_field_rho(D.gpu)) fatal("stage_field_rho");
	/* velocity limit etc. surface here, where the reference would have aborted */
	if(cpic_b200_sync(D.gpu)) fatal("step");
	if(!D.lazy) cpic_b200_dropin_sync(sim);
	perf_stop(&sim->timers[TIMER_FIELD_RHO]);
}
