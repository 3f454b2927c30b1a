/* Multi-GPU layer of cpic_b200: one Y slab per rank, NCCL point-to-point over NVLink.
 *
 * What travels (reference -> here):
 *   rho ghost row      src/comm_field.c:51-136   ncclSend/Recv of nx doubles on the rank ring
 *   phi ghost rows     src/comm_field.c:139-201  2 rows north, 1 row south, padding included
 *   particles (Y pass) src/comm_plasma.c:887-1120  the outbox regions of the edge block rows
 *                      that point across the slab face, gathered into one message per face
 *                      for all species, into the neighbour's ghost outbox rows
 *   FFT transposes     FFTW-MPI inside src/solver.c:485,491: row FFTs, all-to-all, column
 *                      FFTs with the Green's function applied in the transposed layout,
 *                      all-to-all back, row FFTs (2 exchanges per solve; FFTW does 4)
 *
 * NCCL is resolved at run time (dlopen "libnccl.so.2", the copy PyTorch already loaded
 * when the ranks were started by torchrun), so the library also loads where NCCL is absent.
 */
#include "comm.h"
#include "kernels.cuh"

#include <cufft.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

/* ---- the part of nccl.h that is used (stable ABI since NCCL 2.7) ---- */
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclInt32 = 2, ncclInt64 = 4, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };

struct Nccl {
	void *lib;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *);
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*GroupStart)(void);
	ncclResult_t (*GroupEnd)(void);
	const char *(*GetErrorString)(ncclResult_t);
};

static Nccl g_nccl;

static int
load_nccl(char *err, size_t errlen)
{
	if(g_nccl.lib) return 0;
	const char *names[] = { getenv("CPIC_B200_NCCL"), "libnccl.so.2", "libnccl.so" };
	void *lib = NULL;
	for(const char *n : names)
	{
		if(!n || !*n) continue;
		lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
		if(lib) break;
	}
	if(!lib)
	{
		snprintf(err, errlen, "cannot load NCCL (libnccl.so.2): %s; set CPIC_B200_NCCL to its path", dlerror());
		return 2;
	}
#define SYM(field, name) do { *(void **) &g_nccl.field = dlsym(lib, name); \
	if(!g_nccl.field) { snprintf(err, errlen, "NCCL symbol %s missing", name); return 2; } } while(0)
	SYM(GetUniqueId, "ncclGetUniqueId");
	SYM(CommInitRank, "ncclCommInitRank");
	SYM(CommDestroy, "ncclCommDestroy");
	SYM(Send, "ncclSend");
	SYM(Recv, "ncclRecv");
	SYM(AllReduce, "ncclAllReduce");
	SYM(GroupStart, "ncclGroupStart");
	SYM(GroupEnd, "ncclGroupEnd");
	SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
	g_nccl.lib = lib;
	return 0;
}

struct Comm {
	ncclComm_t nc;
	int rank, n;
	Geom g;
	char *err;
	size_t errlen;

	double *rho_recv;            /* nx doubles */
	double *face[4];             /* particle face buffers: send north/south, receive from south/north */
	size_t face_cap;             /* doubles each */

	/* distributed FFT */
	int nc_, cw;                 /* complex columns nx/2+1; columns per rank (ceil) */
	cufftHandle rows_fwd, rows_inv, cols;
	cufftDoubleComplex *a;       /* ny_loc x nc   row spectra */
	cufftDoubleComplex *sb;      /* n blocks of ny_loc x cw   (send / receive staging) */
	cufftDoubleComplex *tb;      /* ny_glob x cw  this rank's columns, all rows */
	double *GT;                  /* ny_glob x cw  Green's function in that layout */
};

#define NCK(call) do { ncclResult_t r_ = (call); if(r_ != 0) { \
	snprintf(c->err, c->errlen, "%s:%d: %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); return 2; } } while(0)
#define CCK(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) { \
	snprintf(c->err, c->errlen, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return 2; } } while(0)
#define FCK(call) do { cufftResult r_ = (call); if(r_ != CUFFT_SUCCESS) { \
	snprintf(c->err, c->errlen, "%s:%d: %s: cufft error %d", __FILE__, __LINE__, #call, (int) r_); return 2; } } while(0)

int
comm_unique_id(void *id128, char *err, size_t errlen)
{
	int rc = load_nccl(err, errlen);
	if(rc) return rc;
	ncclUniqueId id;
	ncclResult_t r = g_nccl.GetUniqueId(&id);
	if(r != 0)
	{
		snprintf(err, errlen, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
		return 2;
	}
	memcpy(id128, &id, sizeof(id));
	return 0;
}

static int
comm_setup(Comm *c, const void *id128, cudaStream_t stream)
{
	const Geom &g = c->g;
	ncclUniqueId id;
	memcpy(&id, id128, sizeof(id));
	NCK(g_nccl.CommInitRank(&c->nc, c->n, id, c->rank));

	CCK(cudaMalloc(&c->rho_recv, (size_t) g.nx * sizeof(double)));

	c->nc_ = g.nx / 2 + 1;
	c->cw = (c->nc_ + c->n - 1) / c->n;
	const size_t blk = (size_t) g.ny * c->cw;
	CCK(cudaMalloc(&c->a, (size_t) g.ny * c->nc_ * sizeof(cufftDoubleComplex)));
	CCK(cudaMalloc(&c->sb, blk * c->n * sizeof(cufftDoubleComplex)));
	CCK(cudaMalloc(&c->tb, blk * c->n * sizeof(cufftDoubleComplex)));

	/* MFT_init, reference src/solver.c:257-280: G[l][k] for this rank's columns, all rows */
	std::vector<double> GT((size_t) g.ny_glob * c->cw, 0.0);
	const double cx = 2.0 * M_PI / (double) g.nx, cy = 2.0 * M_PI / (double) g.ny_glob;
	for(int iy = 0; iy < g.ny_glob; iy++)
		for(int kl = 0; kl < c->cw; kl++)
		{
			const int ix = c->rank * c->cw + kl;
			if(ix >= c->nc_) continue;
			GT[(size_t) iy * c->cw + kl] = (ix == 0 && iy == 0) ? 0.0 :
				1.0 / (2.0 * (cos(cx * (double) ix) + cos(cy * (double) iy)) - 4.0);
		}
	CCK(cudaMalloc(&c->GT, GT.size() * sizeof(double)));
	CCK(cudaMemcpy(c->GT, GT.data(), GT.size() * sizeof(double), cudaMemcpyHostToDevice));

	int nrow[1] = { g.nx };
	int rembed[1] = { g.S }, cembed[1] = { c->nc_ };
	FCK(cufftPlanMany(&c->rows_fwd, 1, nrow, rembed, 1, g.S, cembed, 1, c->nc_, CUFFT_D2Z, g.ny));
	FCK(cufftPlanMany(&c->rows_inv, 1, nrow, cembed, 1, c->nc_, rembed, 1, g.S, CUFFT_Z2D, g.ny));
	int ncol[1] = { g.ny_glob };
	int embed[1] = { g.ny_glob };
	/* column k of the ny_glob x cw array: stride cw between rows, 1 between columns */
	FCK(cufftPlanMany(&c->cols, 1, ncol, embed, c->cw, 1, embed, c->cw, 1, CUFFT_Z2Z, c->cw));
	FCK(cufftSetStream(c->rows_fwd, stream));
	FCK(cufftSetStream(c->rows_inv, stream));
	FCK(cufftSetStream(c->cols, stream));
	return 0;
}

Comm *
comm_create(const void *id128, int rank, int nranks, const Geom &g, cudaStream_t stream,
		char *err, size_t errlen)
{
	if(load_nccl(err, errlen)) return NULL;
	Comm *c = new Comm();
	memset(c, 0, sizeof(*c));
	c->rank = rank;
	c->n = nranks;
	c->g = g;
	c->err = err;
	c->errlen = errlen;
	if(comm_setup(c, id128, stream))
	{
		comm_destroy(c);
		return NULL;
	}
	return c;
}

void
comm_destroy(Comm *c)
{
	if(!c) return;
	if(c->rows_fwd) cufftDestroy(c->rows_fwd);
	if(c->rows_inv) cufftDestroy(c->rows_inv);
	if(c->cols) cufftDestroy(c->cols);
	for(int k = 0; k < 4; k++) cudaFree(c->face[k]);
	cudaFree(c->rho_recv); cudaFree(c->a); cudaFree(c->sb); cudaFree(c->tb); cudaFree(c->GT);
	if(c->nc) g_nccl.CommDestroy(c->nc);
	delete c;
}

/* Element-wise maximum of a few ints over the ranks, in place (error words, capacity
 * requests): what MPI_Bcast(&sim->running) is to the reference (src/sim.c:578) */
int
comm_allreduce_max(Comm *c, int *dev, int n, cudaStream_t stream)
{
	NCK(g_nccl.AllReduce(dev, dev, (size_t) n, ncclInt32, ncclMax, c->nc, stream));
	return 0;
}

/* comm_send_ghost_rho + comm_recv_ghost_rho, reference src/comm_field.c:51-136 */
int
comm_rho_halo(Comm *c, double *rho, cudaStream_t stream, long long *launches)
{
	const Geom &g = c->g;
	const int south = (c->rank + 1) % c->n, north = (c->rank + c->n - 1) % c->n;
	NCK(g_nccl.GroupStart());
	NCK(g_nccl.Send(rho + (size_t) g.ny * g.S, (size_t) g.nx, ncclFloat64, south, c->nc, stream));
	NCK(g_nccl.Recv(c->rho_recv, (size_t) g.nx, ncclFloat64, north, c->nc, stream));
	NCK(g_nccl.GroupEnd());
	k_rho_fold<<<(g.nx + 127) / 128, 128, 0, stream>>>(rho, c->rho_recv, g);
	CCK(cudaGetLastError());
	if(launches) (*launches)++;
	return 0;
}

/* comm_phi_send + comm_phi_recv, reference src/comm_field.c:139-201 */
int
comm_phi_halo(Comm *c, double *phi, cudaStream_t stream)
{
	const Geom &g = c->g;
	const int south = (c->rank + 1) % c->n, north = (c->rank + c->n - 1) % c->n;
	const size_t S = (size_t) g.S;
	NCK(g_nccl.GroupStart());
	/* slab rows 0,1 (array rows 1,2) -> north rank's two south ghost rows */
	NCK(g_nccl.Send(phi + 1 * S, 2 * S, ncclFloat64, north, c->nc, stream));
	/* slab row ny-1 (array row ny) -> south rank's north ghost row */
	NCK(g_nccl.Send(phi + (size_t) g.ny * S, S, ncclFloat64, south, c->nc, stream));
	NCK(g_nccl.Recv(phi + (size_t) (g.ny + 1) * S, 2 * S, ncclFloat64, south, c->nc, stream));
	NCK(g_nccl.Recv(phi, S, ncclFloat64, north, c->nc, stream));
	NCK(g_nccl.GroupEnd());
	return 0;
}

#define FAR_SECTION (1 + 8 * FAR_FACE)      /* doubles: count, then eight arrays */

/* Pack / unpack of the regions that cross a slab face. A face buffer holds, for the three
 * codes k of that direction, the records of the nbx edge blocks' regions (nbx * rcap[code]
 * records of OREC doubles, one contiguous chunk of the outbox), then their (E_x, E_y) pairs
 * when the per-particle field is kept, followed by the 3 * nbx counts. */
struct FaceLayout {
	unsigned off[3];         /* first value of code k's chunk */
	unsigned total;          /* doubles before the counts */
	int na;                  /* doubles per slot: OREC, + 2 with the field */
};

/* What one launch moves: every species, both faces. Grid: (blocks, 2 * species); the last CTA
 * of a grid row handles the far-mover section of its (species, face). */
struct FaceJob {
	SpeciesSet set;
	FaceLayout L[SET_MAX_SPECIES][2];     /* [species][0: codes 0,1,2 | 1: codes 6,7,8] */
	unsigned long long off[SET_MAX_SPECIES + 1];   /* first double of a species' section in a face buffer */
	int nbx;
	int first_block[2];      /* of the block row whose regions are copied, per face */
	double *buf[2];          /* face buffers: pack [to north, to south]; unpack [from south, from north] */
	int pack;
};

static_assert(sizeof(FaceJob) <= 4096, "kernel parameters beyond the classic 4 KB limit");

/* pack != 0: the regions of row 0 with codes 0,1,2 (to the north rank) and of the last row with
 * codes 6,7,8 (south) -> buffers; else buffers -> ghost rows (what the south rank sent north
 * fills the south ghost row and vice versa). */
static __global__ void __launch_bounds__(256)
k_faces(const __grid_constant__ FaceJob job, int *__restrict__ errflag)
{
	const int is = blockIdx.y >> 1, dir = blockIdx.y & 1;
	const SpeciesDev &sp = job.set.sp[is];
	const FaceLayout &L = job.L[is][dir];
	const int nbx = job.nbx, first_block = job.first_block[dir], code0 = dir ? 6 : 0;
	double *__restrict__ buf = job.buf[dir] + job.off[is];
	if(blockIdx.x == gridDim.x - 1)
	{
		/* unpack: the list received on face `dir` came from the opposite direction's sender; it is
		 * appended to the local far-mover list whatever its origin */
		far_face(sp, dir, job.buf[dir] + job.off[is + 1] - FAR_SECTION, job.pack, errflag);
		return;
	}
	const Outbox &ob = sp.ob[job.set.arr[is]];
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i < L.total)
	{
		int k = i >= L.off[2] ? 2 : i >= L.off[1] ? 1 : 0;
		const int c = code0 + k;
		const size_t slots = (size_t) nbx * (size_t) sp.rcap[c];          /* block-major inside the row */
		const size_t slot0 = (size_t) sp.roff[c] + (size_t) first_block * (size_t) sp.rcap[c];
		const size_t r = i - L.off[k];
		double *p = r < slots * OREC ? ob.rec + slot0 * OREC + r : ob.recE + slot0 * 2 + (r - slots * OREC);
		if(job.pack) buf[i] = *p;
		else *p = buf[i];
	}
	else if(i < L.total + 3u * nbx)
	{
		const unsigned r = i - L.total;
		const int k = r / nbx, bx = r % nbx;
		int *cnt = ob.count + (size_t) (code0 + k) * sp.nob + first_block + bx;
		int *bc = (int *) (buf + L.total) + r;
		if(job.pack) *bc = *cnt;
		else *cnt = *bc;
	}
}

static FaceLayout
face_layout(const SpeciesDev *sp, int nbx, int code0)
{
	FaceLayout L;
	L.na = sp->ob[0].recE ? OREC + 2 : OREC;
	unsigned off = 0;
	for(int k = 0; k < 3; k++)
	{
		L.off[k] = off;
		off += (unsigned) L.na * (unsigned) nbx * (unsigned) sp->rcap[code0 + k];
	}
	L.total = off;
	return L;
}

/* The Y pass of comm_plasma between ranks (reference src/comm_plasma.c:1039-1120), all
 * species at once. Row 0's regions with codes 0,1,2 (moving north) land in the north rank's
 * south ghost row; the last row's regions with codes 6,7,8 in the south rank's north ghost
 * row. One message per face: one kernel gathers every species' regions, counts and far movers
 * into the two face buffers, one NCCL group moves both faces, and the same kernel scatters
 * what arrived into the ghost rows. */
int
comm_particles(Comm *c, SpeciesDev *const *sps, const int *arrs, int nsp, const Geom &g, int nb,
		cudaStream_t stream, int *errflag, long long *launches)
{
	const int south = (c->rank + 1) % c->n, north = (c->rank + c->n - 1) % c->n;
	const int nbx = g.nbx;
	FaceLayout Ln[8], Ls[8];
	size_t off[9];
	off[0] = 0;
	for(int i = 0; i < nsp; i++)
	{
		Ln[i] = face_layout(sps[i], nbx, 0);
		Ls[i] = face_layout(sps[i], nbx, 6);
		/* both layouts have the same size (corner, side, corner); then the far-mover section */
		off[i + 1] = off[i] + (size_t) Ln[i].total + (3 * (size_t) nbx + 1) / 2 + FAR_SECTION;
	}
	const size_t doubles = off[nsp];
	if(doubles > c->face_cap)
	{
		for(int k = 0; k < 4; k++) cudaFree(c->face[k]);
		c->face_cap = doubles + doubles / 4;
		for(int k = 0; k < 4; k++) CCK(cudaMalloc(&c->face[k], c->face_cap * sizeof(double)));
	}
	double *send_n = c->face[0], *send_s = c->face[1], *recv_s = c->face[2], *recv_n = c->face[3];
	FaceJob job;
	unsigned most = 0;
	job.set.n = nsp;
	for(int i = 0; i < nsp; i++)
	{
		job.set.sp[i] = *sps[i];
		job.set.arr[i] = arrs[i];
		job.L[i][0] = Ln[i];
		job.L[i][1] = Ls[i];
		job.off[i] = off[i];
		most = std::max(most, Ln[i].total + 3u * nbx);
	}
	job.off[nsp] = off[nsp];
	job.nbx = nbx;
	const dim3 grid((most + 255) / 256 + 1, 2 * nsp);
	/* one launch gathers every species' regions, counts and far movers of both faces */
	job.first_block[0] = 0; job.first_block[1] = nb - nbx;
	job.buf[0] = send_n; job.buf[1] = send_s;
	job.pack = 1;
	k_faces<<<grid, 256, 0, stream>>>(job, errflag);
	NCK(g_nccl.GroupStart());
	NCK(g_nccl.Send(send_n, doubles, ncclFloat64, north, c->nc, stream));
	NCK(g_nccl.Send(send_s, doubles, ncclFloat64, south, c->nc, stream));
	NCK(g_nccl.Recv(recv_s, doubles, ncclFloat64, south, c->nc, stream));
	NCK(g_nccl.Recv(recv_n, doubles, ncclFloat64, north, c->nc, stream));
	NCK(g_nccl.GroupEnd());
	/* what the south rank sent north (codes 0,1,2) fills our south ghost row, and vice versa */
	job.first_block[0] = nb + nbx; job.first_block[1] = nb;
	job.buf[0] = recv_s; job.buf[1] = recv_n;
	job.pack = 0;
	k_faces<<<grid, 256, 0, stream>>>(job, errflag);
	CCK(cudaGetLastError());
	if(launches) *launches += 2;
	return 0;
}

/* ---- distributed MFT solve ---- */

/* a[iy][k] (ny x nc) -> sb[r][iy][kl] with k = r*cw + kl (zero beyond nc) */
__global__ void
k_fft_pack(const cufftDoubleComplex *__restrict__ a, cufftDoubleComplex *__restrict__ sb,
		int ny, int nc, int cw, int n)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;     /* padded column */
	const int iy = blockIdx.y;
	if(k >= cw * n) return;
	const int r = k / cw, kl = k % cw;
	cufftDoubleComplex v = { 0.0, 0.0 };
	if(k < nc) v = a[(size_t) iy * nc + k];
	sb[((size_t) r * ny + iy) * cw + kl] = v;
}

__global__ void
k_fft_unpack(const cufftDoubleComplex *__restrict__ sb, cufftDoubleComplex *__restrict__ a,
		int ny, int nc, int cw)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	const int iy = blockIdx.y;
	if(k >= nc) return;
	const int r = k / cw, kl = k % cw;
	a[(size_t) iy * nc + k] = sb[((size_t) r * ny + iy) * cw + kl];
}

static int
all_to_all(Comm *c, cufftDoubleComplex *send, cufftDoubleComplex *recv, cudaStream_t stream)
{
	const size_t blk = (size_t) c->g.ny * c->cw;             /* complex elements per pair */
	NCK(g_nccl.GroupStart());
	for(int r = 0; r < c->n; r++)
	{
		if(r == c->rank) continue;
		NCK(g_nccl.Send(send + r * blk, 2 * blk, ncclFloat64, r, c->nc, stream));
		NCK(g_nccl.Recv(recv + r * blk, 2 * blk, ncclFloat64, r, c->nc, stream));
	}
	NCK(g_nccl.GroupEnd());
	CCK(cudaMemcpyAsync(recv + c->rank * blk, send + c->rank * blk, blk * sizeof(cufftDoubleComplex),
				cudaMemcpyDeviceToDevice, stream));
	return 0;
}

/* MFT_solve, reference src/solver.c:465-509, over the ranks: rho slab rows -> unnormalised
 * phi slab rows (MFT_normalize is applied by k_phi_finish) */
int
comm_solve(Comm *c, const double *rho, double *phi_raw, cudaStream_t stream, long long *launches)
{
	const Geom &g = c->g;
	const int ncp = c->cw * c->n;
	FCK(cufftExecD2Z(c->rows_fwd, (double *) rho, c->a));
	k_fft_pack<<<dim3((ncp + 127) / 128, g.ny), 128, 0, stream>>>(c->a, c->sb, g.ny, c->nc_, c->cw, c->n);
	if(all_to_all(c, c->sb, c->tb, stream)) return 2;
	FCK(cufftExecZ2Z(c->cols, c->tb, c->tb, CUFFT_FORWARD));
	const size_t n = (size_t) g.ny_glob * c->cw;
	int blocks = (int) ((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
	k_green<<<blocks, 256, 0, stream>>>(c->tb, c->GT, n);
	FCK(cufftExecZ2Z(c->cols, c->tb, c->tb, CUFFT_INVERSE));
	if(all_to_all(c, c->tb, c->sb, stream)) return 2;
	k_fft_unpack<<<dim3((c->nc_ + 127) / 128, g.ny), 128, 0, stream>>>(c->sb, c->a, g.ny, c->nc_, c->cw);
	FCK(cufftExecZ2D(c->rows_inv, c->a, phi_raw));
	CCK(cudaGetLastError());
	if(launches) *launches += 3;
	return 0;
}
