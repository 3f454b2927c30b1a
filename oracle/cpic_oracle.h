/* ORACLE — TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product
 * (cpic_b200/), only by tests/, __graft_entry__.smoke() and bench.py's CPU baseline.
 *
 * Plain-C, single-rank restatement of cpic's per-timestep hot path. Parity status:
 * PINNED — checked in tests/test_oracle.py against (1) the unmodified reference built
 * from /root/reference by oracle/Makefile (oracle/_ref/libcpic_ref*.so), step by step,
 * and (2) the reference's own golden vectors (test/harmonic/harm.r0x, harm.E0x, E.csv;
 * the CIC known-answer of test/interpolate.c.disabled; the analytic checks of
 * test/cyclotron.c and test/constant-speed.c), committed under tests/golden/.
 */
#ifndef CPIC_ORACLE_H
#define CPIC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_species {
	double q, m;
	long long n;
	long long *id;
	double *x, *y, *z;
	double *ux, *uy, *uz;
	double *Ex, *Ey;
} oracle_species_t;

typedef struct oracle_sim {
	long long nx, ny;        /* grid points */
	long long S;             /* row stride of rho/phi: 2*(nx/2+1) (src/solver.c:381-433) */
	double L[2], dx[2], dt, e0, B[3], umax[3];
	long long iter;
	int nspecies;
	oracle_species_t *sp;
	double *rho;             /* (ny+1) x S : rows [0,ny) slab, row ny south ghost */
	double *phi;             /* (ny+3) x S : row 0 north ghost, 1..ny slab, ny+1..ny+2 south ghosts */
	double *Ex, *Ey;         /* (ny+1) x nx */
	double *G;               /* ny x (nx/2+1) */
	double *gre, *gim;       /* spectrum scratch ny x (nx/2+1) */
	int aborted;             /* set when check_velocity would abort (src/mover.c:97-137) */
} oracle_sim_t;

oracle_sim_t *oracle_create(long long nx, long long ny, double Lx, double Ly, double dt,
		double e0, const double B[3], long long plasma_chunks,
		int nspecies, const double *q, const double *m);
void oracle_destroy(oracle_sim_t *s);

/* Particle initialisation, restating src/plasma.c:62-128 + src/particle.c:17-168.
 * oracle_srand must be called once before the first oracle_init_randpos (src/sim.c:153). */
void oracle_srand(unsigned int seed);
int oracle_alloc_species(oracle_sim_t *s, int is, long long n);
/* Fills chunk `ic` of `nchunks` of one species, in the reference's rand() order. */
void oracle_init_randpos_chunk(oracle_sim_t *s, int is, long long ic, long long nchunks,
		const double v[2]);
void oracle_init_delta(oracle_sim_t *s, int is, const double r0[2], const double dr[2],
		const double v[2]);
int oracle_set_particles(oracle_sim_t *s, int is, long long n, const long long *id,
		const double *x, const double *y, const double *ux, const double *uy,
		const double *uz);

/* Stages (src/sim.c:481-581) */
void oracle_stage_field_rho(oracle_sim_t *s);
void oracle_stage_field_E(oracle_sim_t *s);
void oracle_stage_plasma_E(oracle_sim_t *s);
int oracle_stage_plasma_r(oracle_sim_t *s);
void oracle_pre_step(oracle_sim_t *s);      /* src/sim.c:208-236, leaves iter = 0 */
int oracle_step(oracle_sim_t *s);           /* one sim_step */

/* Building blocks, exported for unit parity tests */
void oracle_weights(const oracle_sim_t *s, double x, double y, double w[4],
		long long *i0x, long long *i0y);
void oracle_solve(oracle_sim_t *s);         /* rho -> phi slab rows (src/solver.c:465-509) */
void oracle_phi_ghosts(oracle_sim_t *s);    /* src/comm_field.c:139-201, one rank */
void oracle_field_E(oracle_sim_t *s);       /* src/field.c:358-416 */
void oracle_rfft2(long long ny, long long nx, const double *in, long long ld,
		double *ore, double *oim);
void oracle_irfft2(long long ny, long long nx, const double *ire, const double *iim,
		double *out, long long ld);

/* F1 emulation: deposit one species' particles in the given order in packs of four with
 * the reference's "last lane wins" store (src/simd_avx2.h:226-249). `npad` garbage lanes
 * complete the tail pack at (gx, gy) with zero charge (src/interpolate.c:329-344). */
void oracle_deposit_lossy(oracle_sim_t *s, double q, long long n, const double *x,
		const double *y, double gx, double gy);

/* Diagnostics the reference only sketches (src/sim.c:332-405, disabled there) */
double oracle_kinetic_energy(const oracle_sim_t *s);
double oracle_field_energy(const oracle_sim_t *s);

#ifdef __cplusplus
}
#endif

#endif
