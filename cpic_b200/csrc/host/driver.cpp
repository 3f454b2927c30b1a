/* Stand-alone host driver of cpic_b200: cpic's command line (`cpic [-q] <conf>`, reference
 * src/cpic.c:50-191), run loop with the sampling statistics (src/sim.c:440-479, 621-654) and
 * field output layout (src/output.c:401-635), over the C ABI. Nothing here is on the
 * per-timestep path except the call to cpic_b200_step. */
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <map>
#include <string>
#include <vector>

#include "cpic_b200.h"
#include "front.h"

static int
mkdir_ok(const std::string &p)
{
	if(mkdir(p.c_str(), 0700) && errno != EEXIST)
	{
		front_set_error("mkdir %s: %s", p.c_str(), strerror(errno));
		return -1;
	}
	return 0;
}

/* One output in flight per simulation: pinned staging buffers for the four grids (whole `alignment`
 * blocks, padded with 0xca as the reference's allocation is, src/mat.c:116) and the thread that writes
 * them */
struct OutCtx {
	cpic_b200_sim_t *sim;
	double *buf[4];
	size_t total[4];
	int64_t rows[4], stride[4];
	pthread_t th;
	bool running;
	int status;
	char err[512];
	/* the job of the thread */
	std::string root;
	int64_t iter, alignment, slices, nx, ny;
	double dx, dy;
};

static std::map<cpic_b200_sim_t *, OutCtx *> g_out;
static pthread_mutex_t g_out_lock = PTHREAD_MUTEX_INITIALIZER;

/* write_field, reference src/output.c:482-590: the padded array as it is, in `slices` runs of whole
 * `alignment` blocks, O_DIRECT | O_SYNC like the reference (a file system that refuses O_DIRECT gets a
 * plain write) */
static int
write_slices(const std::string &file, const unsigned char *data, size_t total, int64_t alignment, int64_t slices, char *err, size_t errlen)
{
	int fd = open(file.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_DIRECT, S_IRUSR | S_IWUSR);
	if(fd < 0) fd = open(file.c_str(), O_WRONLY | O_CREAT | O_TRUNC, S_IRUSR | S_IWUSR);
	if(fd < 0) { snprintf(err, errlen, "cannot write %s: %s", file.c_str(), strerror(errno)); return -1; }
	if(slices < 1) slices = 1;
	const size_t blocks = total / (size_t) alignment;
	const size_t per = blocks / (size_t) slices;
	size_t left = blocks % (size_t) slices, offset = 0;
	for(int64_t i = 0; i < slices; i++)
	{
		size_t n = per * (size_t) alignment;
		if(left > 0) { n += (size_t) alignment; left--; }
		size_t done = 0;
		while(done < n)
		{
			const ssize_t w = pwrite(fd, data + offset + done, n - done, (off_t) (offset + done));
			if(w < 0 && errno == EINVAL)
			{
				/* O_DIRECT refused for this buffer or file system: write the rest the plain way */
				const int fl = fcntl(fd, F_GETFL);
				if(fl >= 0 && (fl & O_DIRECT) && fcntl(fd, F_SETFL, fl & ~O_DIRECT) == 0) continue;
			}
			if(w <= 0) { snprintf(err, errlen, "cannot write %s: %s", file.c_str(), strerror(errno)); close(fd); return -1; }
			done += (size_t) w;
		}
		offset += n;
	}
	close(fd);
	return 0;
}

static int write_xdmf(const OutCtx &o);

static void *
out_thread(void *arg)
{
	OutCtx *o = (OutCtx *) arg;
	o->status = 0;
	if(cpic_b200_get_fields_end(o->sim)) { snprintf(o->err, sizeof(o->err), "%s", cpic_b200_last_error()); o->status = CPIC_B200_ECUDA; return NULL; }
	const std::string dir = o->root + "/bin/" + std::to_string(o->iter);
	const char *names[4] = { "rho", "phi", "E_X", "E_Y" };
	for(int k = 0; k < 4 && !o->status; k++)
		if(write_slices(dir + "/" + names[k] + ".bin", (const unsigned char *) o->buf[k], o->total[k], o->alignment, o->slices, o->err, sizeof(o->err)))
			o->status = CPIC_B200_EINVAL;
	if(!o->status && write_xdmf(*o)) o->status = CPIC_B200_EINVAL;
	return NULL;
}

/* write_field_attribute, reference src/output.c:401-421 */
static void
xdmf_attribute(FILE *f, long iter, const char *name, long nx, long ny, long dy, long rows, long stride)
{
	fprintf(f, "      <Attribute Center=\"Node\" Name=\"%s\" DataType=\"Scalar\">\n", name);
	fprintf(f, "        <DataItem ItemType=\"HyperSlab\" Dimensions=\"1 %ld %ld\" Type=\"HyperSlab\">\n", ny, nx);
	fprintf(f, "          <DataItem Dimensions=\"3 3\" Format=\"XML\">\n");
	fprintf(f, "            0 %ld %ld\n", dy, 0L);
	fprintf(f, "            1 1 1\n");
	fprintf(f, "            1 %ld %ld\n", dy + ny, nx);
	fprintf(f, "          </DataItem>\n");
	fprintf(f, "          <DataItem Dimensions=\"1 %ld %ld\" DataType=\"Float\" Precision=\"8\" Format=\"Binary\">\n", rows, stride);
	fprintf(f, "            ../bin/%ld/%s.bin\n", iter, name);
	fprintf(f, "          </DataItem>\n");
	fprintf(f, "        </DataItem>\n");
	fprintf(f, "      </Attribute>\n");
}

/* write_xdmf_fields, reference src/output.c:423-460 */
static int
write_xdmf(const OutCtx &o)
{
	const std::string xf = o.root + "/xdmf/fields-iter" + std::to_string(o.iter) + ".xdmf";
	FILE *f = fopen(xf.c_str(), "w");
	if(!f) return -1;
	const int64_t *rows = o.rows, *stride = o.stride;
	const long nx = (long) o.nx, ny = (long) o.ny;
	fprintf(f, "<?xml version=\"1.0\" encoding=\"utf-8\"?>\n");
	fprintf(f, "<Xdmf xmlns:xi=\"http://www.w3.org/2001/XInclude\" Version=\"3.0\">\n");
	fprintf(f, "  <Domain>\n");
	fprintf(f, "    <Grid Name=\"fields\">\n");
	fprintf(f, "      <Topology TopologyType=\"3DCoRectMesh\" NumberOfElements=\"%ld %ld %ld\"/>\n", 1L, ny, nx);
	fprintf(f, "      <Geometry Origin=\"\" Type=\"ORIGIN_DXDYDZ\">\n");
	fprintf(f, "        <DataItem Format=\"XML\" Dimensions=\"3\">\n");
	fprintf(f, "            0.0 0.0 0.0\n");
	fprintf(f, "        </DataItem>\n");
	fprintf(f, "        <DataItem Format=\"XML\" Dimensions=\"3\">\n");
	fprintf(f, "            %f %f %f\n", 0.0, o.dy, o.dx);
	fprintf(f, "        </DataItem>\n");
	fprintf(f, "      </Geometry>\n");
	/* phi is a view one row into `_phi` (PHI_NG_NORTH, reference src/def.h:11, src/field.c:96) */
	xdmf_attribute(f, (long) o.iter, "phi", nx, ny, 1, (long) rows[1], (long) stride[1]);
	xdmf_attribute(f, (long) o.iter, "rho", nx, ny, 0, (long) rows[0], (long) stride[0]);
	xdmf_attribute(f, (long) o.iter, "E_X", nx, ny, 0, (long) rows[2], (long) stride[2]);
	xdmf_attribute(f, (long) o.iter, "E_Y", nx, ny, 0, (long) rows[3], (long) stride[3]);
	fprintf(f, "    </Grid>\n");
	fprintf(f, "  </Domain>\n");
	fprintf(f, "</Xdmf>\n");
	fclose(f);
	return 0;
}

static OutCtx *
out_ctx(cpic_b200_sim_t *sim)
{
	pthread_mutex_lock(&g_out_lock);
	OutCtx *&o = g_out[sim];
	if(!o) { o = new OutCtx(); o->sim = sim; o->running = false; o->status = 0; for(int k = 0; k < 4; k++) { o->buf[k] = NULL; o->total[k] = 0; } }
	OutCtx *r = o;
	pthread_mutex_unlock(&g_out_lock);
	return r;
}

extern "C" int
cpic_b200_output_wait(cpic_b200_sim_t *sim)
{
	if(!sim) { front_set_error("null argument"); return CPIC_B200_EINVAL; }
	OutCtx *o = out_ctx(sim);
	if(!o->running) return 0;
	pthread_join(o->th, NULL);
	o->running = false;
	if(o->status) front_set_error("%s", o->err);
	return o->status;
}

/* output_fields, reference src/output.c:594-635: the grids start their way to pinned staging on a copy
 * stream, a thread writes them (aligned slices, write_field src/output.c:482-590) and the XDMF descriptor */
extern "C" int
cpic_b200_write_fields_async(cpic_b200_sim_t *sim, const char *path, int64_t iter, int64_t alignment,
		int64_t slices, int64_t nx, int64_t ny, double dx, double dy)
{
	if(!sim || !path) { front_set_error("null argument"); return CPIC_B200_EINVAL; }
	if(alignment <= 0) alignment = 512;
	int rc = cpic_b200_output_wait(sim);        /* the staging buffers are free again */
	if(rc) return rc;
	OutCtx *o = out_ctx(sim);
	o->root = path;
	if(mkdir_ok(o->root) || mkdir_ok(o->root + "/xdmf") || mkdir_ok(o->root + "/bin")
			|| mkdir_ok(o->root + "/bin/" + std::to_string(iter))) return CPIC_B200_EINVAL;
	const int fields[4] = { CPIC_B200_RHO, CPIC_B200_PHI, CPIC_B200_EX, CPIC_B200_EY };
	for(int k = 0; k < 4; k++)
	{
		if(cpic_b200_field_shape(sim, fields[k], &o->rows[k], &o->stride[k])) return CPIC_B200_EINVAL;
		const size_t bytes = (size_t) o->rows[k] * o->stride[k] * sizeof(double);
		const size_t total = (bytes + alignment - 1) / alignment * alignment;
		if(total != o->total[k])
		{
			cpic_b200_host_free(o->buf[k]);
			o->buf[k] = (double *) cpic_b200_host_alloc(total);
			if(!o->buf[k]) { o->total[k] = 0; front_set_error("pinned staging of %zu bytes failed", total); return CPIC_B200_ENOMEM; }
			o->total[k] = total;
		}
		memset((unsigned char *) o->buf[k] + bytes, 0xca, total - bytes);
	}
	if(cpic_b200_get_fields_begin(sim, o->buf)) return CPIC_B200_ECUDA;
	o->iter = iter; o->alignment = alignment; o->slices = slices; o->nx = nx; o->ny = ny; o->dx = dx; o->dy = dy;
	if(pthread_create(&o->th, NULL, out_thread, o))
	{
		out_thread(o);          /* no thread: write here */
		if(o->status) front_set_error("%s", o->err);
		return o->status;
	}
	o->running = true;
	return 0;
}

extern "C" int
cpic_b200_write_fields(cpic_b200_sim_t *sim, const char *path, int64_t iter, int64_t alignment,
		int64_t nx, int64_t ny, double dx, double dy)
{
	int rc = cpic_b200_write_fields_async(sim, path, iter, alignment, 1, nx, ny, dx, dy);
	if(rc) return rc;
	return cpic_b200_output_wait(sim);
}

static double
now(void)
{
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return (double) t.tv_sec + 1e-9 * (double) t.tv_nsec;
}

static int
env_int(const char *name, int fallback)
{
	const char *e = getenv(name);
	return (e && *e) ? atoi(e) : fallback;
}

/* MPI_Bcast of the NCCL id without MPI: rank 0 writes it to a file (written under another name, then
 * renamed: readers never see a partial file), the others wait for the file */
static int
share_id(int rank, unsigned char id[128])
{
	char path[4096], tmp[4200];
	const char *e = getenv("CPIC_B200_ID_FILE");
	if(e && *e) snprintf(path, sizeof(path), "%s", e);
	else snprintf(path, sizeof(path), "/tmp/cpic_b200_id_%s_%ld", getenv("MASTER_PORT") ? getenv("MASTER_PORT") : "0", (long) getppid());
	if(rank == 0)
	{
		if(cpic_b200_comm_id(id)) return -1;
		snprintf(tmp, sizeof(tmp), "%s.tmp", path);
		FILE *f = fopen(tmp, "wb");
		if(!f || fwrite(id, 1, 128, f) != 128) { front_set_error("cannot write %s", tmp); if(f) fclose(f); return -1; }
		fclose(f);
		if(rename(tmp, path)) { front_set_error("cannot rename %s", tmp); return -1; }
		return 0;
	}
	for(int tries = 0; tries < 6000; tries++)
	{
		FILE *f = fopen(path, "rb");
		if(f)
		{
			const size_t n = fread(id, 1, 128, f);
			fclose(f);
			if(n == 128) return 0;
		}
		usleep(20000);
	}
	front_set_error("rank %d: no communicator id in %s after two minutes", rank, path);
	return -1;
}

/* sim_init + sim_run of one rank among several (Y slabs, src/sim.c:116-130, 177-187): the reference's initial
 * conditions drawn on the device for this rank's slab, the exchanges inside the library */
static int
run_ranks(const char *fn, int rank, int nranks, int device, double t0)
{
	cpic_b200_sim_t *sim = NULL;
	cpic_b200_run_t run;
	unsigned char id[128];
	if(cpic_b200_sim_from_conf_device(fn, rank, nranks, device, 1, 1 << 22, &sim, &run))
	{
		fprintf(stderr, "sim_init failed\n%s\n", cpic_b200_last_error());
		return 1;
	}
	if(share_id(rank, id) || cpic_b200_comm_init(sim, id) || cpic_b200_pre_step(sim) || cpic_b200_sync(sim))
	{
		fprintf(stderr, "rank %d: communicator or pre-step failed\n%s\n", rank, cpic_b200_last_error());
		return 1;
	}
	if(rank == 0)
	{
		if(run.output_enabled) fprintf(stderr, "output of several ranks is not written (the reference's ranks overwrite one file, src/output.c:517)\n");
		fprintf(stderr, "Using %d GPU ranks\n", nranks);
		printf("%e init-time\n", now() - t0);
		printf("Simulation runs now\n");
	}
	double mean = 0.0, m2 = 0.0;
	long n = 0;
	/* the sampling decision is rank 0's (MPI_Bcast(&sim->running), src/sim.c:578); without a broadcast every
	 * rank runs all the cycles when sampling is on */
	while(cpic_b200_iter(sim) < run.cycles)
	{
		const double t1 = now();
		if(cpic_b200_step(sim) || cpic_b200_sync(sim))
		{
			fprintf(stderr, "rank %d: sim_step failed\n%s\n", rank, cpic_b200_last_error());
			return 1;
		}
		const double t = now() - t1;
		if(rank == 0 && run.stop_SEM > 0.0)
		{
			const double mean0 = mean;
			const double std = n >= 2 ? sqrt(m2 / (double) (n - 1)) : 0.0;
			const double sem = n >= 2 ? std / sqrt((double) n) : 0.0;
			const double rsem = mean0 != 0.0 ? sem / mean0 : sem;
			n++;
			mean = mean0 + (t - mean0) / (double) n;
			m2 += (t - mean0) * (t - mean);
			printf("stats iter=%ld last=%e mean=%e std=%e sem=%e rsem=%e mem=%ld solver=%e\n",
					(long) cpic_b200_iter(sim) - 1, t, mean0, std, sem, rsem, 0L, 0.0);
		}
	}
	double ke = 0.0, pe = 0.0;
	if(!cpic_b200_energy(sim, &ke, &pe)) printf("rank %d: kinetic %.17g potential %.17g\n", rank, ke, pe);
	if(rank == 0) printf("Simulation ends\n");
	cpic_b200_destroy(sim);
	return 0;
}

/* main of cpic (reference src/cpic.c:50-191) + sim_run (src/sim.c:621-654) */
extern "C" int
cpic_b200_main(int argc, char **argv)
{
	int opt;
	optind = 1;
	while((opt = getopt(argc, argv, "qd")) != -1)
	{
		if(opt != 'q' && opt != 'd')
		{
			fprintf(stderr, "Simulation of plasma using particle in cell method (cpic_b200).\n\nUsage: %s [-q] <config file>\n", argv[0]);
			return 1;
		}
	}
	if(optind != argc - 1)
	{
		fprintf(stderr, "Simulation of plasma using particle in cell method (cpic_b200).\n\nUsage: %s [-q] <config file>\n", argv[0]);
		return 1;
	}
	const char *fn = argv[optind];
	const double t0 = now();
	cpic_b200_sim_t *sim = NULL;
	cpic_b200_run_t run;
	cpic_b200_conf_t *conf = NULL;
	cpic_b200_params_t p;
	/* `mpirun -n P cpic <conf>` (src/cpic.c:82-96: one MPI process per Y slab) is P processes of this driver, one
	 * per GPU, told their place by the environment a launcher sets (CPIC_B200_RANK / CPIC_B200_NRANKS, else
	 * torchrun's or Open MPI's variables); the 128-byte communicator id travels through a file */
	const int rank = env_int("CPIC_B200_RANK", env_int("RANK", env_int("OMPI_COMM_WORLD_RANK", 0)));
	const int nranks = env_int("CPIC_B200_NRANKS", env_int("WORLD_SIZE", env_int("OMPI_COMM_WORLD_SIZE", 1)));
	const int device = nranks > 1 ? env_int("CPIC_B200_DEVICE", env_int("LOCAL_RANK", env_int("OMPI_COMM_WORLD_LOCAL_RANK", rank))) : -1;
	if(nranks > 1) return run_ranks(fn, rank, nranks, device, t0);
	if(cpic_b200_conf_load(fn, &conf) || cpic_b200_conf_params(conf, 0, 1, -1, &p, &run))
	{
		fprintf(stderr, "Configuration read failed:\n%s\n", cpic_b200_last_error());
		return 1;
	}
	cpic_b200_conf_free(conf);
	if(!run.output_enabled) fprintf(stderr, "No output path specified, output will not be saved\n");
	if(run.stop_SEM > 0.0) fprintf(stderr, "Sampling enabled with relative error limit %e\n", run.stop_SEM);
	/* populations beyond 2^27 particles are drawn and uploaded in batches: the host arrays of the
	 * one-shot path would not fit next to the binning copies (1e9 particles = 48 GB each) */
	long long total = 0;
	for(int is = 0; is < p.nspecies; is++) total += run.nparticles[is];
	const int rc_init = total > (1LL << 27) ? cpic_b200_sim_from_conf_streamed(fn, 0, 1, -1, 1, 1 << 24, &sim, NULL)
			: cpic_b200_sim_from_conf(fn, 0, 1, -1, 1, &sim, NULL);
	if(rc_init)
	{
		fprintf(stderr, "sim_init failed\n%s\n", cpic_b200_last_error());
		return 1;
	}
	printf("%e init-time\n", now() - t0);
	printf("Simulation runs now\n");

	/* Welford statistics of the iteration time, as perf_record/perf_stats (src/perf.c:85-133) */
	double mean = 0.0, m2 = 0.0;
	long n = 0;
	const double dx = p.Lx / (double) p.nx, dy = p.Ly / (double) p.ny;
	int running = 1;
	while(running && cpic_b200_iter(sim) < run.cycles)
	{
		const double t1 = now();
		/* sim_step writes the fields right after stage_field_E (src/sim.c:503-511) */
		if(run.output_enabled)
		{
			if(cpic_b200_stage_field_E(sim)
					|| cpic_b200_write_fields_async(sim, run.output_path, cpic_b200_iter(sim), run.output_alignment, run.output_slices, p.nx, p.ny, dx, dy)
					|| cpic_b200_stage_plasma_E(sim) || cpic_b200_stage_plasma_r(sim) || cpic_b200_stage_field_rho(sim)
					|| cpic_b200_set_iter(sim, cpic_b200_iter(sim) + 1))
			{
				fprintf(stderr, "sim_step failed\n%s\n", cpic_b200_last_error());
				return 1;
			}
		}
		else if(cpic_b200_step(sim))
		{
			fprintf(stderr, "sim_step failed\n%s\n", cpic_b200_last_error());
			return 1;
		}
		if(cpic_b200_sync(sim))
		{
			fprintf(stderr, "sim_step failed\n%s\n", cpic_b200_last_error());
			return 1;
		}
		const double t = now() - t1;
		if(run.stop_SEM > 0.0)
		{
			/* sampling_complete, src/sim.c:440-479: perf_stats BEFORE perf_record -- the mean, std and
			 * sem that are printed and tested are those of the samples before this one */
			const double mean0 = mean;
			const double std = n >= 2 ? sqrt(m2 / (double) (n - 1)) : 0.0;
			const double sem = n >= 2 ? std / sqrt((double) n) : 0.0;
			const double rsem = mean0 != 0.0 ? sem / mean0 : sem;
			/* perf_record, src/perf.c:85-108 (Welford) */
			n++;
			mean = mean0 + (t - mean0) / (double) n;
			m2 += (t - mean0) * (t - mean);
			printf("stats iter=%ld last=%e mean=%e std=%e sem=%e rsem=%e mem=%ld solver=%e\n",
					(long) cpic_b200_iter(sim) - 1, t, mean0, std, sem, rsem, 0L, 0.0);
			if(cpic_b200_iter(sim) - 1 >= 30 && 1.96 * sem < run.stop_SEM * mean0)
			{
				printf("sampling complete\n");
				running = 0;
			}
		}
	}
	if(cpic_b200_output_wait(sim))
	{
		fprintf(stderr, "output failed\n%s\n", cpic_b200_last_error());
		return 1;
	}
	printf("Simulation ends\n");
	cpic_b200_destroy(sim);
	return 0;
}
