/* cpic_b200 — C ABI of the B200-native replacement for cpic's per-timestep hot path.
 *
 * cpic has no plugin interface: its seams are plain C functions on `sim_t *`
 * (reference src/field.h:5-12, src/particle.h:7-25, src/mover.h:3-7, src/solver.h:46-59)
 * called in a fixed order by sim_step (src/sim.c:481-581). This header is the boundary
 * those functions bind to: plain pointers and sizes, no CUDA or torch types. The
 * `sim_t *` flavoured symbols a cpic maintainer links (stage_field_E(sim_t *) ...) are a
 * thin translation onto these entry points; see INTEGRATION.md and dropin/.
 *
 * Conventions (reference src/sim.c:507-511, src/log.h:69-70): functions return 0 on
 * success and non-zero on failure, with a message available from cpic_b200_last_error().
 * The velocity-limit check (src/mover.c:97-137), which aborts in the reference, is
 * reported as CPIC_B200_EVELOCITY by the next call that synchronises.
 *
 * All device work of one simulation is issued on one CUDA stream; calls are
 * asynchronous unless they return data to the host.
 */
#ifndef CPIC_B200_H
#define CPIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPIC_B200_MAX_SPECIES 8

enum cpic_b200_status {
	CPIC_B200_OK = 0,
	CPIC_B200_EINVAL = 1,       /* bad argument / unsupported configuration */
	CPIC_B200_ECUDA = 2,        /* CUDA, cuFFT or NCCL failure */
	CPIC_B200_EVELOCITY = 3,    /* |u| exceeded umax (src/mover.c:123-136) */
	CPIC_B200_ECAPACITY = 4,    /* a particle block or exchange buffer overflowed */
	CPIC_B200_EFAR = 5,         /* a particle crossed more than one particle block in a step */
	CPIC_B200_ENOMEM = 6
};

/* Grid arrays, in the reference's own layouts (src/field.c:20-160):
 *   RHO  `_rho`: rows [0, ny] (row ny = south ghost), stride 2*(nx/2+1)  (src/solver.c:381-433)
 *   PHI  `_phi`: ny+3 rows (north ghost, slab, two south ghosts), same stride (src/def.h:11-12)
 *   EX, EY `_E[X]`, `_E[Y]`: ny+1 rows, stride nx                       (src/def.h:13-14)
 * ny is the number of rows of this rank's slab. */
enum cpic_b200_field { CPIC_B200_RHO = 0, CPIC_B200_PHI = 1, CPIC_B200_EX = 2, CPIC_B200_EY = 3 };

/* What sim_read_config + sim_prepare produce (src/sim.c:38-206), for one rank */
typedef struct cpic_b200_params {
	int64_t nx, ny;             /* global grid points: grid.points */
	double Lx, Ly;              /* simulation.space_length */
	double dt;                  /* simulation.time_step */
	double e0;                  /* constants.vacuum_permittivity */
	double B[3];                /* field.magnetic */
	int64_t plasma_chunks;      /* simulation.plasma_chunks: only sets umax[X] (src/sim.c:198) */
	int32_t nspecies;
	double q[CPIC_B200_MAX_SPECIES];   /* species[i].charge */
	double m[CPIC_B200_MAX_SPECIES];   /* species[i].mass */
	int32_t rank, nranks;       /* Y-slab decomposition, ny % nranks == 0 (src/sim.c:116-122) */
	int32_t device;             /* CUDA device ordinal, -1 = current */
	double capacity_factor;     /* particle-block slack over the fullest block, 0 = default (1.5) */
	int32_t keep_particle_E;    /* 1: stage_plasma_r also keeps the gathered E per particle */
	int32_t block_cells;        /* cells per particle-block side (power of two <= 32), 0 = default (8) */
	double outbox_fraction;     /* exchange buffer per block side as a share of the block capacity,
	                             * 0 = default (0.3); corners get a quarter of it */
} cpic_b200_params_t;

typedef struct cpic_b200_sim cpic_b200_sim_t;

const char *cpic_b200_version(void);
const char *cpic_b200_last_error(void);

/* ---- life cycle: sim_init (src/sim.c:238-320) / sim_end (:322-330) ---- */
int cpic_b200_create(const cpic_b200_params_t *params, cpic_b200_sim_t **out);
void cpic_b200_destroy(cpic_b200_sim_t *sim);

/* Multi-GPU: each rank passes the same 128-byte NCCL unique id (obtained on rank 0 with
 * cpic_b200_comm_id) before the first stage call. Replaces MPI_COMM_WORLD (src/cpic.c:82-96). */
int cpic_b200_comm_id(void *id128);
int cpic_b200_comm_init(cpic_b200_sim_t *sim, const void *id128);

/* ---- particles: plasma_init + particle_comm_initial (src/plasma.c:292-316,
 * src/particle.h:19-20). Host SoA in, any order; particles outside this rank's slab are
 * rejected. Replaces the whole population of one species. ---- */
int cpic_b200_set_particles(cpic_b200_sim_t *sim, int species, int64_t n,
		const int64_t *id, const double *x, const double *y,
		const double *ux, const double *uy, const double *uz);
/* Streamed plasma_init for populations whose host arrays do not fit at once (1e9 particles; the
 * reference's own host lists need 112 B per particle, src/def.h:88-133). The caller generates the
 * population twice, batch by batch: cpic_b200_count_particles tallies every batch per particle
 * block on the host (particles of every rank's slab are accepted and tallied, so that ranks which
 * all see the whole stream arrive at one common capacity); cpic_b200_reserve_counted sizes the
 * species from the tallies; then cpic_b200_add_particles bins every batch of this rank's slab
 * (stable) and appends it to the blocks. A block ends up
 * holding its particles in batch order, then input order (cpic_b200_set_particles: input order). */
int cpic_b200_count_particles(cpic_b200_sim_t *sim, int species, int64_t n, const double *x, const double *y);
int cpic_b200_reserve_counted(cpic_b200_sim_t *sim, int species);
int cpic_b200_add_particles(cpic_b200_sim_t *sim, int species, int64_t n, const int64_t *id,
		const double *x, const double *y, const double *ux, const double *uy, const double *uz);

/* Block capacity (slots per particle block) of a species. With several ranks every rank must
 * use the same capacity (the exchange buffers are sized from it): set_particles on every
 * rank, take the maximum of cpic_b200_capacity over the ranks, and if it differs from the
 * local value call cpic_b200_reserve with it and set_particles again. */
int64_t cpic_b200_capacity(cpic_b200_sim_t *sim, int species);
int cpic_b200_reserve(cpic_b200_sim_t *sim, int species, int64_t capacity);
/* Fill of the fullest particle block and exchange regions of a species, next to their
 * capacities: out[6] = { max block count, block capacity, max side-region count, side
 * capacity, max corner-region count, corner capacity }. */
int cpic_b200_occupancy(cpic_b200_sim_t *sim, int species, int64_t out[6]);
int64_t cpic_b200_num_particles(cpic_b200_sim_t *sim, int species);
/* Host SoA out, in device (particle-block) order; any pointer may be NULL.
 * Ex/Ey are the fields last gathered onto the particles (ppack.E, src/def.h:96).
 * Returns the number of particles, or -1. */
int64_t cpic_b200_get_particles(cpic_b200_sim_t *sim, int species, int64_t cap,
		int64_t *id, double *x, double *y, double *ux, double *uy, double *uz,
		double *Ex, double *Ey);

/* Host lists that keep their own order -- the reference's plist/pblock/ppack lists (src/def.h:88-211),
 * which the drop-in binding refreshes after a step and which never reorder there because comm_plasma
 * runs on the device. cpic_b200_set_host_order takes the particle ids in the order the host walks its
 * lists; cpic_b200_get_particles_ordered then fills arrays whose entry k belongs to ids[k]: the
 * permutation from particle-block order is made on the device and the host writes its lists front to
 * back. Arrays from cpic_b200_host_alloc (pinned) are filled by DMA. Any array may be NULL. */
int cpic_b200_set_host_order(cpic_b200_sim_t *sim, int species, int64_t n, const int64_t *ids);
int cpic_b200_get_particles_ordered(cpic_b200_sim_t *sim, int species, int64_t n, double *x, double *y,
		double *ux, double *uy, double *uz, double *Ex, double *Ey);

/* Throughput-only initialiser on the device (no host arrays): uniform positions,
 * u ~ U(-v, v) per axis like "random position" (src/particle.c:23-89) but from a
 * counter-based generator, not glibc rand(). n is this rank's share. */
int cpic_b200_init_uniform(cpic_b200_sim_t *sim, int species, int64_t n, int64_t id0,
		double vx, double vy, uint64_t seed);
/* Same with a drift: u = (dux, duy) + U(-v, v) per axis -- a warm beam (the reference's
 * "position delta" initialiser gives every particle the drift velocity, src/particle.c:152-153) */
int cpic_b200_init_beam(cpic_b200_sim_t *sim, int species, int64_t n, int64_t id0,
		double dux, double duy, double vx, double vy, uint64_t seed);

/* The reference's own initial conditions drawn on the device. A run is what one (process, chunk,
 * species) triple of the reference initialises in one go (src/plasma.c:292-316 -> src/particle.c:178-213):
 * the ids first, first + step, ... (`count` of them); method 0 = "random position" (four rand() calls
 * per particle from the process's glibc stream: srand(seed), `draw0` calls made before this run),
 * method 1 = "position delta". Positions and velocities are bit-identical to cpic_b200_conf_init_particles;
 * inside a particle block the particles are in the order they were drawn. Replaces the populations of
 * the species the runs name. cpic_b200_sim_from_conf_device builds the runs from a `.conf`. */
typedef struct cpic_b200_init_run {
	int32_t species, method;
	int64_t first, step, count;
	uint32_t seed;
	int64_t draw0;
	double v[2], dr[2], r0[2];
} cpic_b200_init_run_t;
int cpic_b200_init_reference(cpic_b200_sim_t *sim, int nruns, const cpic_b200_init_run_t *runs, int64_t batch);

/* ---- the four stages, in sim_step order (src/sim.c:503,517,525,536) ---- */
int cpic_b200_stage_field_E(cpic_b200_sim_t *sim);     /* src/field.c:450-501 */
int cpic_b200_stage_plasma_E(cpic_b200_sim_t *sim);    /* src/particle.c:232-248 */
int cpic_b200_stage_plasma_r(cpic_b200_sim_t *sim);    /* src/mover.c:331-362 (mover + comm_plasma) */
int cpic_b200_stage_field_rho(cpic_b200_sim_t *sim);   /* src/field.c:268-356 */

/* sim_pre_step (src/sim.c:208-236): rho, dummy field solve; leaves iter = 0 */
int cpic_b200_pre_step(cpic_b200_sim_t *sim);
/* One sim_step (src/sim.c:481-581) with gather and push fused (E is not stored per
 * particle unless keep_particle_E). Advances iter. */
int cpic_b200_step(cpic_b200_sim_t *sim);
int cpic_b200_run(cpic_b200_sim_t *sim, int64_t steps);
/* Same, bracketed by CUDA events on the simulation's stream; *ms = device time of the steps */
int cpic_b200_run_timed(cpic_b200_sim_t *sim, int64_t steps, double *ms);

int64_t cpic_b200_iter(cpic_b200_sim_t *sim);
int cpic_b200_set_iter(cpic_b200_sim_t *sim, int64_t iter);

/* Wait for the device and report deferred errors (velocity limit, capacity) */
int cpic_b200_sync(cpic_b200_sim_t *sim);

/* ---- fields ---- */
/* Geometry of one grid array on this rank: rows and row stride (doubles) */
int cpic_b200_field_shape(cpic_b200_sim_t *sim, int field, int64_t *rows, int64_t *stride);
/* Copy a whole array in the reference's padded layout (what output.c writes,
 * src/output.c:627-630). `host` holds rows*stride doubles. */
int cpic_b200_get_field(cpic_b200_sim_t *sim, int field, double *host);
int cpic_b200_set_field(cpic_b200_sim_t *sim, int field, const double *host);
/* The four grids on their way to host buffers (pinned: cpic_b200_host_alloc) without stalling the step:
 * the copies run on a stream of their own behind the work issued so far; stages that overwrite a grid
 * wait for them on the device. host[field] == NULL skips a grid. cpic_b200_get_fields_end waits. */
int cpic_b200_get_fields_begin(cpic_b200_sim_t *sim, double *const host[4]);
int cpic_b200_get_fields_end(cpic_b200_sim_t *sim);

/* Solver seam (src/solver.h:46-59): phi slab rows <- solve(rho slab rows); no ghosts */
int cpic_b200_solve(cpic_b200_sim_t *sim);

/* ---- diagnostics the reference sketches but compiles out (src/sim.c:332-405) ---- */
int cpic_b200_energy(cpic_b200_sim_t *sim, double *kinetic, double *potential);

/* ---- measurement ---- */
/* Device time (ms, CUDA events on the simulation's stream) spent in each stage since the
 * last reset: [0] field_E without the solver [1] the push kernels (fused gather+push in
 * cpic_b200_step, push alone in stage_plasma_r) [2] particle exchange [3] field_rho
 * [4] solver alone [5] the gather kernels of stage_plasma_E. Enable with
 * cpic_b200_timing(sim, 1). */
int cpic_b200_timing(cpic_b200_sim_t *sim, int enable);
int cpic_b200_get_timing(cpic_b200_sim_t *sim, double ms[6], int64_t launches[1]);

/* Compact particle image for host<->device round trips (bench e2e, checkpoints): the live
 * particles of every block, block after block, with the block counts. The caller owns a
 * (pinned) buffer of cpic_b200_image_bytes(); download fills it, upload restores the exact
 * device state (order included). */
int64_t cpic_b200_image_bytes(cpic_b200_sim_t *sim);
int cpic_b200_image_download(cpic_b200_sim_t *sim, void *host, int64_t bytes);
int cpic_b200_image_upload(cpic_b200_sim_t *sim, const void *host, int64_t bytes);
/* One sim_step with the particle state living in pinned host memory: `host` is an image as written by
 * cpic_b200_image_download; every species is uploaded, pushed and downloaded in turn on three streams
 * (the upload of one species overlaps the push and the download of the previous one: PCIe runs in both
 * directions at once); the image comes back with the new block counts and particles. One rank. */
int cpic_b200_step_host(cpic_b200_sim_t *sim, void *host, int64_t bytes);
/* The same in bands of block rows: the image is laid out band by band (cpic_b200_banded_image_download;
 * bands <= 0: 8), every band is uploaded, unpacked and pushed while the next ones are on their way, and
 * absorbed, packed and downloaded as soon as its neighbour bands are pushed: uploads and downloads overlap
 * all along. A band's image has a quarter of slack for its population to change; a band that outgrows it
 * fails with CPIC_B200_ECAPACITY (make a new image). One rank. */
int64_t cpic_b200_banded_image_bytes(cpic_b200_sim_t *sim, int bands);
int cpic_b200_banded_image_download(cpic_b200_sim_t *sim, void *host, int64_t bytes, int bands);
int cpic_b200_step_host_banded(cpic_b200_sim_t *sim, void *host, int64_t bytes);
void *cpic_b200_host_alloc(size_t bytes);   /* pinned */
void cpic_b200_host_free(void *p);

/* ---- front end: cpic's `.conf` files (reference src/sim.c:38-206, src/specie.c:15-36,
 * src/particle.c:17-213, src/plasma.c:62-128, src/output.c:63-130) ---- */
typedef struct cpic_b200_conf cpic_b200_conf_t;

/* Run control that is not part of the hot path's parameters */
typedef struct cpic_b200_run {
	int64_t cycles;               /* simulation.cycles */
	uint32_t seed;                /* simulation.random_seed */
	double stop_SEM;              /* simulation.stop_SEM */
	int64_t period_energy, period_field, period_particle;
	char solver[16];              /* simulation.solver: "MFT" (others are rejected) */
	int32_t output_enabled;       /* output.path present */
	char output_path[4096];
	int64_t output_slices;        /* output.slices, default 1 */
	int64_t output_alignment;     /* output.alignment, default 512 */
	int64_t nparticles[CPIC_B200_MAX_SPECIES];
} cpic_b200_run_t;

/* config_read_file (reference src/cpic.c:150-164); include dir = the file's directory */
int cpic_b200_conf_load(const char *path, cpic_b200_conf_t **out);
void cpic_b200_conf_free(cpic_b200_conf_t *conf);
/* sim_read_config + sim_prepare + species_init + output_init. rank/nranks/device are
 * copied from the arguments; missing required keys fail like the reference's
 * `Failed to read parameter "..."`. */
int cpic_b200_conf_params(const cpic_b200_conf_t *conf, int rank, int nranks, int device,
		cpic_b200_params_t *params, cpic_b200_run_t *run);
/* plasma_init on the host, bit-identical to the reference run with `ref_nprocs` MPI
 * processes (ids striped over ref_nprocs*plasma_chunks chunks, srand(seed + rank) per
 * process, rand() drawn chunk -> species -> particle -> {x, y, ux, uy}). Fills, for every
 * species s, arrays of run->nparticles[s] entries in id order. uz is 0. */
int cpic_b200_conf_init_particles(const cpic_b200_conf_t *conf, int ref_nprocs,
		int64_t *const *id, double *const *x, double *const *y,
		double *const *ux, double *const *uy);
/* The stream of cpic_b200_conf_init_particles delivered in batches of at most `batch` particles of
 * one species, in the order the reference draws them (process -> chunk -> species -> particle);
 * `sink` returns non-zero to stop. */
typedef int (*cpic_b200_particle_sink_t)(void *ctx, int species, int64_t n, const int64_t *id,
		const double *x, const double *y, const double *ux, const double *uy);
int cpic_b200_conf_stream_particles(const cpic_b200_conf_t *conf, int ref_nprocs, int64_t batch,
		cpic_b200_particle_sink_t sink, void *ctx);
/* cpic_b200_sim_from_conf with the streamed initialisation above: host memory stays at a few
 * batches whatever the population (the particles are the same ones; inside a particle block they
 * are ordered by batch instead of by id). */
int cpic_b200_sim_from_conf_streamed(const char *path, int rank, int nranks, int device, int ref_nprocs,
		int64_t batch, cpic_b200_sim_t **sim, cpic_b200_run_t *run);
/* The same with the particles drawn on the device (cpic_b200_init_reference): no host arrays, no serial
 * rand() calls -- 1e9 particles in seconds. */
int cpic_b200_sim_from_conf_device(const char *path, int rank, int nranks, int device, int ref_nprocs,
		int64_t batch, cpic_b200_sim_t **sim, cpic_b200_run_t *run);
/* sim_init (reference src/sim.c:238-320): params, create, host init of all species,
 * upload of this rank's slab, and (single rank only) the pre-step. With several ranks
 * call cpic_b200_comm_init and cpic_b200_pre_step afterwards. */
int cpic_b200_sim_from_conf(const char *path, int rank, int nranks, int device, int ref_nprocs,
		cpic_b200_sim_t **sim, cpic_b200_run_t *run);
/* output_fields (reference src/output.c:594-635): <path>/bin/<iter>/{rho,phi,E_X,E_Y}.bin -- the
 * padded arrays as they are, rounded up to `alignment` bytes -- and
 * <path>/xdmf/fields-iter<iter>.xdmf with the reference's hyperslab descriptors. */
int cpic_b200_write_fields(cpic_b200_sim_t *sim, const char *path, int64_t iter, int64_t alignment,
		int64_t nx, int64_t ny, double dx, double dy);
/* The same without waiting: the grids are copied to pinned staging buffers on a copy stream and written by
 * a background thread -- aligned slices with O_DIRECT as the reference does (src/output.c:482-590; `slices`
 * = output.slices) -- while the step goes on. One output in flight per simulation: the next call, or
 * cpic_b200_output_wait, waits for the previous one and returns its status. */
int cpic_b200_write_fields_async(cpic_b200_sim_t *sim, const char *path, int64_t iter, int64_t alignment,
		int64_t slices, int64_t nx, int64_t ny, double dx, double dy);
int cpic_b200_output_wait(cpic_b200_sim_t *sim);
/* The reference's command line, `cpic [-q] <conf>` (src/cpic.c:50-191). `mpirun -n P cpic <conf>` is P
 * processes of it, one per GPU: rank, number of ranks and device from CPIC_B200_RANK / CPIC_B200_NRANKS /
 * CPIC_B200_DEVICE (or torchrun's RANK / WORLD_SIZE / LOCAL_RANK, or Open MPI's OMPI_COMM_WORLD_*), the
 * communicator id through the file CPIC_B200_ID_FILE */
int cpic_b200_main(int argc, char **argv);

#ifdef __cplusplus
}
#endif

#endif
