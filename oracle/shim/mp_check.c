/* ORACLE TEST INFRASTRUCTURE — not product code.
 *
 * ref_mp_check <conf> <steps> <outdir>: the unmodified reference (accumulate-correct deposit
 * variant) over CPIC_SHIM_NPROCS forked ranks (shim_mpi_mp.c); after `steps` sim_steps every
 * rank writes its slab of rho, phi, E_X, E_Y and its particles to <outdir>/rank<r>.bin.
 * tests/test_oracle.py compares the assembled slabs with the single-rank run: that is what
 * pins the multi-process shims (sockets, barrier, shared-mapping FFT) the CPU baseline uses. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mpi.h>

void *ref_open(const char *conf_path);
int ref_step(void *vh);
int ref_nspecies(void *vh);
void ref_grid(void *vh, long long *nx, long long *ny);
long long ref_get_field(void *vh, int which, double *out);
long long ref_get_particles(void *vh, int is, long long cap, long long *id, double *x, double *y, double *z,
		double *ux, double *uy, double *uz, double *Ex, double *Ey, long long *chunk_of);

int
main(int argc, char **argv)
{
	if(argc < 4) { fprintf(stderr, "usage: ref_mp_check conf steps outdir\n"); return 2; }
	void *h = ref_open(argv[1]);          /* MPI_Init_thread inside: the ranks are forked here */
	if(!h) return 1;
	int rank;
	MPI_Comm_rank(MPI_COMM_WORLD, &rank);
	for(int i = 0; i < atoi(argv[2]); i++)
		if(ref_step(h)) return 1;

	char path[4096];
	snprintf(path, sizeof(path), "%s/rank%d.bin", argv[3], rank);
	FILE *f = fopen(path, "wb");
	if(!f) { perror(path); return 1; }
	long long nx, ny, ns = ref_nspecies(h);
	ref_grid(h, &nx, &ny);
	long long hdr[3] = { nx, ny, ns };
	fwrite(hdr, sizeof(hdr), 1, f);
	double *buf = malloc((size_t) ((ny + 3) * nx) * sizeof(double));
	for(int which = 0; which < 4; which++)
	{
		long long rows = ref_get_field(h, which, buf);
		fwrite(&rows, sizeof(rows), 1, f);
		fwrite(buf, sizeof(double), (size_t) (rows * nx), f);
	}
	free(buf);
	for(int is = 0; is < ns; is++)
	{
		long long n = ref_get_particles(h, is, 0, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL);
		long long *id = malloc((size_t) (n + 1) * sizeof(long long));
		double *a[4];
		for(int k = 0; k < 4; k++) a[k] = malloc((size_t) (n + 1) * sizeof(double));
		ref_get_particles(h, is, n, id, a[0], a[1], NULL, a[2], a[3], NULL, NULL, NULL, NULL);
		fwrite(&n, sizeof(n), 1, f);
		fwrite(id, sizeof(long long), (size_t) n, f);
		for(int k = 0; k < 4; k++) { fwrite(a[k], sizeof(double), (size_t) n, f); free(a[k]); }
		free(id);
	}
	fclose(f);
	MPI_Finalize();
	return 0;
}
