/* Host front end of cpic_b200: reads cpic's `.conf` files and reproduces the reference's
 * host-side initialisation bit for bit, so that a run starts from the reference's own
 * initial conditions. Nothing here is on the per-timestep path. */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "cpic_b200.h"
#include "front.h"

/* shares the error slot of the C ABI through a setter in sim.cu */
extern "C" void cpic_b200_set_error_(const char *msg);

void
front_set_error(const char *fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	cpic_b200_set_error_(buf);
}

extern "C" int
cpic_b200_conf_load(const char *path, cpic_b200_conf_t **out)
{
	char err[256];
	if(!path || !out) { front_set_error("null argument"); return CPIC_B200_EINVAL; }
	conf_node_t *root = conf_parse_file(path, NULL, err, sizeof(err));
	if(!root)
	{
		/* reference src/cpic.c:158-164 */
		front_set_error("Configuration read failed: %s - %s", path, err);
		return CPIC_B200_EINVAL;
	}
	*out = new cpic_b200_conf{ root };
	return 0;
}

extern "C" void
cpic_b200_conf_free(cpic_b200_conf_t *c)
{
	if(!c) return;
	conf_free(c->root);
	delete c;
}

/* config_array_float, reference src/config.c:9-61 */
static int
array_float(conf_node_t *cs, double *v, int size, const char *what)
{
	if(!cs) { front_set_error("setting \"%s\" is missing", what); return 1; }
	if(cs->type == CONF_ARRAY)
	{
		if(conf_length(cs) != size)
		{
			front_set_error("Line %d: The setting %s should have %d dimensions, but %d found.",
					cs->line, what, size, conf_length(cs));
			return 1;
		}
		for(int i = 0; i < size; i++)
		{
			/* config_setting_get_float_elem: non-float elements read as 0 */
			v[i] = 0.0;
			conf_get_float(conf_elem(cs, i), &v[i]);
		}
		return 0;
	}
	if(cs->type == CONF_FLOAT && size == 1) { v[0] = cs->fval; return 0; }
	front_set_error("Line %d: The setting %s is expected to be and array of %d dimensions.", cs->line, what, size);
	return 1;
}

#define READ(getter, path, var) do { if(!getter(conf_lookup(root, path), var)) { \
	front_set_error("Failed to read parameter \"%s\"", path); return CPIC_B200_EINVAL; } } while(0)

extern "C" int
cpic_b200_conf_params(const cpic_b200_conf_t *c, int rank, int nranks, int device,
		cpic_b200_params_t *p, cpic_b200_run_t *run)
{
	if(!c || !p) { front_set_error("null argument"); return CPIC_B200_EINVAL; }
	conf_node_t *root = c->root;
	cpic_b200_run_t local;
	if(!run) run = &local;
	memset(p, 0, sizeof(*p));
	memset(run, 0, sizeof(*run));

	/* sim_read_config, reference src/sim.c:46-82 (same keys, same order) */
	long long dim, ll;
	int iv;
	double dv;
	const char *sv;
	READ(conf_get_int64, "simulation.dimensions", &dim);
	READ(conf_get_int64, "simulation.cycles", &ll); run->cycles = ll;
	READ(conf_get_float, "simulation.time_step", &p->dt);
	READ(conf_get_int, "simulation.random_seed", &iv); run->seed = (uint32_t) iv;
	READ(conf_get_float, "constants.light_speed", &dv);
	READ(conf_get_float, "constants.vacuum_permittivity", &p->e0);
	READ(conf_get_int64, "simulation.sampling_period.energy", &ll); run->period_energy = ll;
	READ(conf_get_int64, "simulation.sampling_period.field", &ll); run->period_field = ll;
	READ(conf_get_int64, "simulation.sampling_period.particle", &ll); run->period_particle = ll;
	READ(conf_get_float, "simulation.stop_SEM", &run->stop_SEM);
	READ(conf_get_int, "simulation.realtime_plot", &iv);
	READ(conf_get_string, "simulation.solver", &sv);
	snprintf(run->solver, sizeof(run->solver), "%s", sv);
	READ(conf_get_int64, "simulation.enable_fftw_threads", &ll);
	READ(conf_get_int64, "simulation.plasma_chunks", &ll); p->plasma_chunks = ll;
	READ(conf_get_int64, "simulation.pblock_nmax", &ll);

	/* reference src/sim.c:110-114 */
	if(dim != 2) { front_set_error("Only 2 dimensions supported by now..."); return CPIC_B200_EINVAL; }
	/* reference src/solver.c:564-577: the only production solver is MFT */
	if(strcmp(run->solver, "MFT") != 0)
	{
		front_set_error("Unknown solver method \"%s\" (cpic_b200 implements MFT)", run->solver);
		return CPIC_B200_EINVAL;
	}

	double L[2], B[3];
	if(array_float(conf_lookup(root, "simulation.space_length"), L, 2, "space_length")) return CPIC_B200_EINVAL;
	if(array_float(conf_lookup(root, "field.magnetic"), B, 3, "magnetic")) return CPIC_B200_EINVAL;
	p->Lx = L[0]; p->Ly = L[1];
	memcpy(p->B, B, sizeof(B));
	conf_node_t *pts = conf_lookup(root, "grid.points");
	long long nx = 0, ny = 0;
	if(!conf_get_int64(conf_elem(pts, 0), &nx) || !conf_get_int64(conf_elem(pts, 1), &ny))
	{
		front_set_error("Failed to read parameter \"grid.points\"");
		return CPIC_B200_EINVAL;
	}
	p->nx = nx; p->ny = ny;

	/* species_init, reference src/specie.c:15-68 */
	conf_node_t *sp = conf_lookup(root, "species");
	int ns = conf_length(sp);
	if(ns <= 0) { front_set_error("The number of species must be at least 1."); return CPIC_B200_EINVAL; }
	if(ns > CPIC_B200_MAX_SPECIES) { front_set_error("at most %d species are supported", CPIC_B200_MAX_SPECIES); return CPIC_B200_EINVAL; }
	p->nspecies = ns;
	for(int i = 0; i < ns; i++)
	{
		conf_node_t *s = conf_elem(sp, i);
		long long n = 0;
		conf_get_float(conf_member(s, "charge"), &p->q[i]);
		conf_get_float(conf_member(s, "mass"), &p->m[i]);
		conf_get_int64(conf_member(s, "particles"), &n);
		if(n <= 0) { front_set_error("The number of particles must be greater than 0"); return CPIC_B200_EINVAL; }
		run->nparticles[i] = n;
	}

	/* output_init, reference src/output.c:63-130 */
	if(conf_get_string(conf_lookup(root, "output.path"), &sv))
	{
		run->output_enabled = 1;
		snprintf(run->output_path, sizeof(run->output_path), "%s", sv);
	}
	run->output_slices = 1;
	run->output_alignment = 512;
	if(conf_get_int64(conf_lookup(root, "output.slices"), &ll)) run->output_slices = ll;
	if(conf_get_int64(conf_lookup(root, "output.alignment"), &ll)) run->output_alignment = ll;

	p->rank = rank;
	p->nranks = nranks;
	p->device = device;
	return 0;
}

/* particles_init, reference src/particle.c:178-213, for every chunk of every reference
 * process, in the reference's loop order (src/plasma.c:292-316 -> :192-283 -> :17-128).
 * `put(species, i, x, y, ux, uy)` receives every particle as it is drawn; a non-zero return stops. */
template <typename Put>
static int
generate_particles(const cpic_b200_conf_t *c, int ref_nprocs, Put put)
{
	cpic_b200_params_t p;
	cpic_b200_run_t run;
	int rc = cpic_b200_conf_params(c, 0, 1, -1, &p, &run);
	if(rc) return rc;
	if(ref_nprocs < 1) ref_nprocs = 1;
	conf_node_t *species = conf_lookup(c->root, "species");
	const long long nchunks = p.plasma_chunks;
	const long long step = (long long) ref_nprocs * nchunks;      /* src/plasma.c:62 */

	struct Init { int method; double v[2], dr[2], r0[2]; };
	std::vector<Init> init((size_t) p.nspecies);
	for(int is = 0; is < p.nspecies; is++)
	{
		conf_node_t *s = conf_elem(species, is);
		const char *method = NULL;
		if(!conf_get_string(conf_member(s, "init_method"), &method))
		{
			front_set_error("Particle init method for specie %d not specified.", is);
			return CPIC_B200_EINVAL;
		}
		Init &in = init[(size_t) is];
		memset(&in, 0, sizeof(in));
		if(array_float(conf_member(s, "drift_velocity"), in.v, 2, "drift_velocity")) return CPIC_B200_EINVAL;
		if(strcmp(method, "random position") == 0) in.method = 0;
		else if(strcmp(method, "position delta") == 0)
		{
			in.method = 1;
			if(array_float(conf_member(s, "position_delta"), in.dr, 2, "position_delta")) return CPIC_B200_EINVAL;
			if(array_float(conf_member(s, "position_init"), in.r0, 2, "position_init")) return CPIC_B200_EINVAL;
		}
		else
		{
			front_set_error("Unknown init method \"%s\", aborting.", method);
			return CPIC_B200_EINVAL;
		}
	}

	const double L[2] = { p.Lx, p.Ly };
	for(int proc = 0; proc < ref_nprocs; proc++)
	{
		srand(run.seed + (unsigned int) proc);                   /* src/sim.c:153 */
		for(long long ic = 0; ic < nchunks; ic++)
		{
			const long long first = ic * ref_nprocs + proc;      /* src/plasma.c:63 */
			for(int is = 0; is < p.nspecies; is++)
			{
				const Init &in = init[(size_t) is];
				for(long long i = first; i < run.nparticles[is]; i += step)
				{
					double x, y, ux, uy;
					if(in.method == 0)
					{
						/* src/particle.c:17-21, :69-73: rand() / (RAND_MAX + 1.0) * (b - a) + a */
						x = rand() / (RAND_MAX + 1.0) * (L[0] - 0.0) + 0.0;
						y = rand() / (RAND_MAX + 1.0) * (L[1] - 0.0) + 0.0;
						ux = rand() / (RAND_MAX + 1.0) * (in.v[0] - -in.v[0]) + -in.v[0];
						uy = rand() / (RAND_MAX + 1.0) * (in.v[1] - -in.v[1]) + -in.v[1];
					}
					else
					{
						/* src/particle.c:150-158, WRAP = src/mat.h:62 */
						double r[2];
						for(int d = 0; d < 2; d++)
						{
							/* the reference build (-O3 -march=core-avx2, GNU fp-contract) fuses
							 * r0 + dr*i into one FMA; mirrored for bit-identical positions */
							r[d] = fmod(fma(in.dr[d], (double) i, in.r0[d]), L[d]);
							if(r[d] < 0.0) r[d] += L[d];
						}
						x = r[0]; y = r[1];
						ux = in.v[0]; uy = in.v[1];
					}
					if((rc = put(is, i, x, y, ux, uy))) return rc;
				}
			}
		}
	}
	return 0;
}

extern "C" int
cpic_b200_conf_init_particles(const cpic_b200_conf_t *c, int ref_nprocs,
		int64_t *const *id, double *const *x, double *const *y,
		double *const *ux, double *const *uy)
{
	return generate_particles(c, ref_nprocs, [&](int is, long long i, double px, double py, double pux, double puy) {
		id[is][i] = i;
		x[is][i] = px; y[is][i] = py;
		ux[is][i] = pux; uy[is][i] = puy;
		return 0;
	});
}

extern "C" int
cpic_b200_conf_stream_particles(const cpic_b200_conf_t *c, int ref_nprocs, int64_t batch,
		cpic_b200_particle_sink_t sink, void *ctx)
{
	if(!sink || batch < 1) { front_set_error("stream_particles: no sink or empty batch"); return CPIC_B200_EINVAL; }
	std::vector<int64_t> id; std::vector<double> x, y, ux, uy;
	id.reserve((size_t) batch); x.reserve((size_t) batch); y.reserve((size_t) batch);
	ux.reserve((size_t) batch); uy.reserve((size_t) batch);
	int cur = -1;
	auto flush = [&]() {
		int rc = 0;
		if(!id.empty()) rc = sink(ctx, cur, (int64_t) id.size(), id.data(), x.data(), y.data(), ux.data(), uy.data());
		id.clear(); x.clear(); y.clear(); ux.clear(); uy.clear();
		return rc;
	};
	int rc = generate_particles(c, ref_nprocs, [&](int is, long long i, double px, double py, double pux, double puy) {
		if(is != cur || (int64_t) id.size() >= batch)
		{
			int r = flush();
			if(r) return r;
			cur = is;
		}
		id.push_back(i); x.push_back(px); y.push_back(py); ux.push_back(pux); uy.push_back(puy);
		return 0;
	});
	if(!rc) rc = flush();
	return rc;
}

/* The slab filter of particle_comm_initial (reference src/particle.h:19-20) over a batch */
struct SlabSink {
	cpic_b200_sim_t *sim;
	double dy;
	long long ny, rows, rank;
	int pass;                /* 0: count, 1: add */
	std::vector<int64_t> id; std::vector<double> x, y, ux, uy;
};

static int
slab_sink(void *ctx, int is, int64_t n, const int64_t *id, const double *x, const double *y,
		const double *ux, const double *uy)
{
	SlabSink &k = *(SlabSink *) ctx;
	if(k.pass == 0) return cpic_b200_count_particles(k.sim, is, n, x, y);
	k.id.clear(); k.x.clear(); k.y.clear(); k.ux.clear(); k.uy.clear();
	for(int64_t i = 0; i < n; i++)
	{
		long long row = (long long) floor(y[i] * (1.0 / k.dy));
		if(row < 0) row = 0;
		if(row > k.ny - 1) row = k.ny - 1;
		if(row / k.rows != k.rank) continue;
		k.id.push_back(id[i]); k.x.push_back(x[i]); k.y.push_back(y[i]); k.ux.push_back(ux[i]); k.uy.push_back(uy[i]);
	}
	return cpic_b200_add_particles(k.sim, is, (int64_t) k.id.size(), k.id.data(), k.x.data(), k.y.data(),
			k.ux.data(), k.uy.data(), NULL);
}

extern "C" int
cpic_b200_sim_from_conf_streamed(const char *path, int rank, int nranks, int device, int ref_nprocs,
		int64_t batch, cpic_b200_sim_t **out, cpic_b200_run_t *run)
{
	cpic_b200_conf_t *c = NULL;
	cpic_b200_params_t p;
	cpic_b200_run_t local;
	if(!run) run = &local;
	if(batch < 1) batch = 1 << 22;
	int rc = cpic_b200_conf_load(path, &c);
	if(rc) return rc;
	rc = cpic_b200_conf_params(c, rank, nranks, device, &p, run);
	cpic_b200_sim_t *sim = NULL;
	if(!rc) rc = cpic_b200_create(&p, &sim);
	if(rc) { cpic_b200_conf_free(c); return rc; }
	SlabSink k;
	k.sim = sim;
	k.dy = p.Ly / (double) p.ny;
	k.ny = p.ny; k.rows = p.ny / nranks; k.rank = rank;
	/* the population is drawn twice from the same seed: once to size the blocks, once to fill them */
	k.pass = 0;
	rc = cpic_b200_conf_stream_particles(c, ref_nprocs, batch, slab_sink, &k);
	for(int is = 0; is < p.nspecies && !rc; is++) rc = cpic_b200_reserve_counted(sim, is);
	k.pass = 1;
	if(!rc) rc = cpic_b200_conf_stream_particles(c, ref_nprocs, batch, slab_sink, &k);
	cpic_b200_conf_free(c);
	if(!rc && nranks == 1) rc = cpic_b200_pre_step(sim);
	if(rc) { cpic_b200_destroy(sim); return rc; }
	*out = sim;
	return 0;
}

/* sim_init with the reference's initial conditions drawn on the device: the runs of
 * generate_particles (process -> chunk -> species) handed to cpic_b200_init_reference */
extern "C" int
cpic_b200_sim_from_conf_device(const char *path, int rank, int nranks, int device, int ref_nprocs,
		int64_t batch, cpic_b200_sim_t **out, cpic_b200_run_t *run)
{
	cpic_b200_conf_t *c = NULL;
	cpic_b200_params_t p;
	cpic_b200_run_t local;
	if(!run) run = &local;
	if(ref_nprocs < 1) ref_nprocs = 1;
	int rc = cpic_b200_conf_load(path, &c);
	if(rc) return rc;
	rc = cpic_b200_conf_params(c, rank, nranks, device, &p, run);
	if(rc) { cpic_b200_conf_free(c); return rc; }
	conf_node_t *species = conf_lookup(c->root, "species");
	const long long nchunks = p.plasma_chunks, step = (long long) ref_nprocs * nchunks;
	std::vector<cpic_b200_init_run_t> runs;
	for(int proc = 0; proc < ref_nprocs && !rc; proc++)
	{
		long long draws = 0;              /* rand() calls of this process so far */
		for(long long ic = 0; ic < nchunks && !rc; ic++)
			for(int is = 0; is < p.nspecies && !rc; is++)
			{
				conf_node_t *sn = conf_elem(species, is);
				const char *method = NULL;
				cpic_b200_init_run_t r;
				memset(&r, 0, sizeof(r));
				if(!conf_get_string(conf_member(sn, "init_method"), &method))
				{ front_set_error("Particle init method for specie %d not specified.", is); rc = CPIC_B200_EINVAL; break; }
				if(array_float(conf_member(sn, "drift_velocity"), r.v, 2, "drift_velocity")) { rc = CPIC_B200_EINVAL; break; }
				if(strcmp(method, "random position") == 0) r.method = 0;
				else if(strcmp(method, "position delta") == 0)
				{
					r.method = 1;
					if(array_float(conf_member(sn, "position_delta"), r.dr, 2, "position_delta")
							|| array_float(conf_member(sn, "position_init"), r.r0, 2, "position_init")) { rc = CPIC_B200_EINVAL; break; }
				}
				else { front_set_error("Unknown init method \"%s\", aborting.", method); rc = CPIC_B200_EINVAL; break; }
				r.species = is;
				r.first = ic * ref_nprocs + proc;              /* src/plasma.c:63 */
				r.step = step;
				r.count = r.first < run->nparticles[is] ? (run->nparticles[is] - r.first + step - 1) / step : 0;
				r.seed = run->seed + (unsigned int) proc;       /* src/sim.c:153 */
				r.draw0 = draws;
				if(r.method == 0) draws += 4 * r.count;
				if(r.count > 0) runs.push_back(r);
			}
	}
	cpic_b200_conf_free(c);
	cpic_b200_sim_t *sim = NULL;
	if(!rc) rc = cpic_b200_create(&p, &sim);
	if(rc) return rc;
	rc = cpic_b200_init_reference(sim, (int) runs.size(), runs.data(), batch);
	if(!rc && nranks == 1) rc = cpic_b200_pre_step(sim);
	if(rc) { cpic_b200_destroy(sim); return rc; }
	*out = sim;
	return 0;
}

extern "C" int
cpic_b200_sim_from_conf(const char *path, int rank, int nranks, int device, int ref_nprocs,
		cpic_b200_sim_t **out, cpic_b200_run_t *run)
{
	cpic_b200_conf_t *c = NULL;
	cpic_b200_params_t p;
	cpic_b200_run_t local;
	if(!run) run = &local;
	int rc = cpic_b200_conf_load(path, &c);
	if(rc) return rc;
	rc = cpic_b200_conf_params(c, rank, nranks, device, &p, run);
	if(rc) { cpic_b200_conf_free(c); return rc; }

	std::vector<std::vector<int64_t>> id((size_t) p.nspecies);
	std::vector<std::vector<double>> x((size_t) p.nspecies), y((size_t) p.nspecies), ux((size_t) p.nspecies), uy((size_t) p.nspecies);
	int64_t *pid[CPIC_B200_MAX_SPECIES];
	double *px[CPIC_B200_MAX_SPECIES], *py[CPIC_B200_MAX_SPECIES], *pux[CPIC_B200_MAX_SPECIES], *puy[CPIC_B200_MAX_SPECIES];
	for(int is = 0; is < p.nspecies; is++)
	{
		size_t n = (size_t) run->nparticles[is];
		id[(size_t) is].resize(n); x[(size_t) is].resize(n); y[(size_t) is].resize(n);
		ux[(size_t) is].resize(n); uy[(size_t) is].resize(n);
		pid[is] = id[(size_t) is].data(); px[is] = x[(size_t) is].data(); py[is] = y[(size_t) is].data();
		pux[is] = ux[(size_t) is].data(); puy[is] = uy[(size_t) is].data();
	}
	rc = cpic_b200_conf_init_particles(c, ref_nprocs, pid, px, py, pux, puy);
	cpic_b200_conf_free(c);
	if(rc) return rc;

	cpic_b200_sim_t *sim = NULL;
	rc = cpic_b200_create(&p, &sim);
	if(rc) return rc;

	/* particle_comm_initial (reference src/particle.h:19-20): keep this rank's slab */
	const double dy = p.Ly / (double) p.ny;
	const long long rows = p.ny / nranks;
	for(int is = 0; is < p.nspecies && !rc; is++)
	{
		const size_t n = (size_t) run->nparticles[is];
		if(nranks == 1)
		{
			rc = cpic_b200_set_particles(sim, is, (int64_t) n, pid[is], px[is], py[is], pux[is], puy[is], NULL);
			continue;
		}
		std::vector<int64_t> sid; std::vector<double> sx, sy, sux, suy;
		std::vector<size_t> per_rank((size_t) nranks, 0);
		for(size_t i = 0; i < n; i++)
		{
			long long row = (long long) floor(py[is][i] * (1.0 / dy));
			if(row < 0) row = 0;
			if(row > p.ny - 1) row = p.ny - 1;
			per_rank[(size_t) (row / rows)]++;
			if(row / rows != rank) continue;
			sid.push_back(pid[is][i]); sx.push_back(px[is][i]); sy.push_back(py[is][i]);
			sux.push_back(pux[is][i]); suy.push_back(puy[is][i]);
		}
		rc = cpic_b200_set_particles(sim, is, (int64_t) sid.size(), sid.data(), sx.data(), sy.data(), sux.data(), suy.data(), NULL);
		/* every rank sees the whole population here, but only bins its own slab: the common
		 * block capacity is agreed by the caller (cpic_b200_capacity / cpic_b200_reserve) */
		(void) per_rank;
	}
	if(!rc && nranks == 1) rc = cpic_b200_pre_step(sim);
	if(rc) { cpic_b200_destroy(sim); return rc; }
	*out = sim;
	return 0;
}
