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
#include <unistd.h>
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

#define FAR_SECTION (1 + 8 * FAR_FACE)      /* doubles: count, then eight arrays */

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

/* ---- peer memory over NVLink (CUDA IPC between the ranks of one box) ----
 * What a rank lets its peers write: its exchange flags, the row that receives the rho ghost row,
 * phi (ghost rows), the far-mover inbox, the transposed spectra of the distributed FFT and the
 * outboxes of every species (ghost outbox rows). Every such buffer is one cudaMalloc allocation,
 * exported with cudaIpcGetMemHandle; the handles travel once over NCCL and are opened by the
 * ranks that write there. */
enum { CH_PART = 0, CH_RHO, CH_PHI, CH_FFT_FWD, CH_FFT_INV, CH_COUNT };
enum { X_FLAGS = 0, X_RHO_RECV, X_PHI, X_FAR_IN, X_FFT_TB, X_FFT_A, X_SPECIES, X_SPECIES_E = X_SPECIES + SET_MAX_SPECIES,
	X_COUNT = X_SPECIES_E + SET_MAX_SPECIES };

struct ExportEntry {
	cudaIpcMemHandle_t h;
	unsigned long long bytes;
	int valid;
	int pad;
};

struct Comm {
	ncclComm_t nc;
	int rank, n;
	Geom g;
	char *err;
	size_t errlen;

	bool p2p;                    /* every rank could export its buffers: exchanges go through peer memory */
	void *local[X_COUNT];        /* this rank's exported allocations */
	ExportEntry mine[X_COUNT];
	ExportEntry *table;          /* host: n x X_COUNT, what every rank exports */
	ExportEntry *dtable;         /* device staging of the same */
	void *remote[X_COUNT][PEER_MAX];       /* peers' allocations mapped here (NULL: not mapped) */
	ExportEntry opened[X_COUNT][PEER_MAX]; /* the handle behind each mapping */
	int *flags;                  /* CH_COUNT x PEER_MAX words, written by the peers */
	int seq[CH_COUNT];
	double *far_in;              /* [species][face] sections written by the neighbours' k_far_ship */

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

static int p2p_setup(Comm *c, cudaStream_t stream);

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
	if(p2p_setup(c, stream)) return 2;
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
	/* this rank's mappings of its peers' memory first, then its own buffers */
	for(int what = 0; what < X_COUNT; what++)
		for(int r = 0; r < PEER_MAX; r++)
			if(c->remote[what][r]) cudaIpcCloseMemHandle(c->remote[what][r]);
	free(c->table);
	cudaFree(c->dtable); cudaFree(c->flags); cudaFree(c->far_in);
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

/* ------------------------------------------------------------ peer memory */

static bool
same_handle(const ExportEntry &a, const ExportEntry &b)
{
	return a.valid == b.valid && a.bytes == b.bytes && memcmp(&a.h, &b.h, sizeof(a.h)) == 0;
}

/* Registers (or replaces, or with ptr == NULL withdraws) one exported allocation of this rank. The
 * peers see it after the next comm_p2p_refresh. */
int
comm_p2p_export(Comm *c, int what, void *ptr, size_t bytes)
{
	if(!c->p2p) return 0;
	ExportEntry e;
	memset(&e, 0, sizeof(e));
	c->local[what] = ptr;
	if(ptr)
	{
		CCK(cudaIpcGetMemHandle(&e.h, ptr));
		e.bytes = bytes;
		e.valid = 1;
	}
	c->mine[what] = e;
	return 0;
}

/* Closes this rank's mappings of what the peers export under `what` (before the owners free it) */
int
comm_p2p_unmap(Comm *c, int what)
{
	if(!c->p2p) return 0;
	for(int r = 0; r < c->n; r++)
		if(c->remote[what][r])
		{
			CCK(cudaIpcCloseMemHandle(c->remote[what][r]));
			c->remote[what][r] = NULL;
			memset(&c->opened[what][r], 0, sizeof(ExportEntry));
		}
	return 0;
}

/* Does rank r's export `what` have to be mapped here? Flags and the FFT buffers: every rank;
 * everything else: the two neighbours on the ring. */
static bool
wanted(const Comm *c, int what, int r)
{
	if(r == c->rank) return false;
	if(what == X_FLAGS || what == X_FFT_TB || what == X_FFT_A) return true;
	return r == (c->rank + 1) % c->n || r == (c->rank + c->n - 1) % c->n;
}

/* Collective: every rank's table of exports goes to every other rank (grouped NCCL send/recv),
 * and the mappings that changed are (re)opened. */
int
comm_p2p_refresh(Comm *c, cudaStream_t stream)
{
	if(!c->p2p) return 0;
	const size_t one = sizeof(ExportEntry) * X_COUNT;
	memcpy(c->table + (size_t) c->rank * X_COUNT, c->mine, one);
	CCK(cudaMemcpyAsync((char *) c->dtable + c->rank * one, c->mine, one, cudaMemcpyHostToDevice, stream));
	NCK(g_nccl.GroupStart());
	for(int r = 0; r < c->n; r++)
	{
		if(r == c->rank) continue;
		NCK(g_nccl.Send((char *) c->dtable + c->rank * one, one, ncclInt8, r, c->nc, stream));
		NCK(g_nccl.Recv((char *) c->dtable + r * one, one, ncclInt8, r, c->nc, stream));
	}
	NCK(g_nccl.GroupEnd());
	CCK(cudaMemcpyAsync(c->table, c->dtable, one * c->n, cudaMemcpyDeviceToHost, stream));
	CCK(cudaStreamSynchronize(stream));
	for(int what = 0; what < X_COUNT; what++)
		for(int r = 0; r < c->n; r++)
		{
			if(!wanted(c, what, r)) continue;
			const ExportEntry &e = c->table[(size_t) r * X_COUNT + what];
			if(same_handle(e, c->opened[what][r])) continue;
			if(c->remote[what][r])
			{
				CCK(cudaIpcCloseMemHandle(c->remote[what][r]));
				c->remote[what][r] = NULL;
			}
			if(e.valid) CCK(cudaIpcOpenMemHandle(&c->remote[what][r], e.h, cudaIpcMemLazyEnablePeerAccess));
			c->opened[what][r] = e;
		}
	return 0;
}

bool comm_p2p(const Comm *c) { return c && c->p2p; }
int comm_rank_north(const Comm *c) { return (c->rank + c->n - 1) % c->n; }
int comm_rank_south(const Comm *c) { return (c->rank + 1) % c->n; }
static_assert((int) COMM_EXPORT_PHI == (int) X_PHI, "comm.h names X_PHI");

void *
comm_p2p_remote(const Comm *c, int what, int rank)
{
	return c->p2p ? c->remote[what][rank] : NULL;
}

int comm_export_species(int is, int with_E) { return (with_E ? X_SPECIES_E : X_SPECIES) + is; }

/* The ranks listed meet: every one of them signals the others and waits for them (channel = what
 * the exchange carries; the sequence number of a channel grows with every use, the same on all
 * ranks because all of them make the same calls). */
static int
peer_meet(Comm *c, int channel, const int *ranks, int nr, cudaStream_t stream, int *errflag, long long *launches)
{
	PeerJob sig, wt;
	const int value = ++c->seq[channel];
	sig.n = wt.n = 0;
	sig.value = wt.value = value;
	for(int k = 0; k < nr; k++)
	{
		const int r = ranks[k];
		if(r == c->rank) continue;
		bool seen = false;
		for(int j = 0; j < sig.n; j++) seen = seen || wt.slot[j] == channel * PEER_MAX + r;
		if(seen) continue;          /* two ranks: north and south are the same peer */
		sig.flag[sig.n] = (int *) c->remote[X_FLAGS][r];
		sig.slot[sig.n++] = channel * PEER_MAX + c->rank;
		wt.flag[wt.n] = c->flags;
		wt.slot[wt.n++] = channel * PEER_MAX + r;
	}
	k_peer_signal<<<1, 32, 0, stream>>>(sig);
	k_peer_wait<<<1, 32, 0, stream>>>(wt, errflag);
	CCK(cudaGetLastError());
	if(launches) *launches += 2;
	return 0;
}

static int
meet_neighbours(Comm *c, int channel, cudaStream_t stream, int *errflag, long long *launches)
{
	const int r[2] = { (c->rank + c->n - 1) % c->n, (c->rank + 1) % c->n };
	return peer_meet(c, channel, r, 2, stream, errflag, launches);
}

static int
meet_all(Comm *c, int channel, cudaStream_t stream, int *errflag, long long *launches)
{
	int r[PEER_MAX];
	for(int k = 0; k < c->n; k++) r[k] = k;
	return peer_meet(c, channel, r, c->n, stream, errflag, launches);
}

/* Sets up the peer-memory path: decides collectively whether every rank can export, allocates the
 * flags and the far-mover inbox and exports the static buffers. */
static int
p2p_setup(Comm *c, cudaStream_t stream)
{
	const char *e = getenv("CPIC_B200_P2P");
	int fail = (e && atoi(e) == 0) || c->n > PEER_MAX;
	int *probe = NULL;
	if(cudaMalloc(&probe, 256) != cudaSuccess) return 2;
	if(!fail)
	{
		cudaIpcMemHandle_t h;
		if(cudaIpcGetMemHandle(&h, probe) != cudaSuccess) { fail = 1; cudaGetLastError(); }
	}
	/* one rank that cannot export (no IPC: a container without it, the CPU test interpreter), or ranks on
	 * different boxes (the host names differ: max(h) != -max(-h)), send everybody down the NCCL path */
	char host[256] = "";
	gethostname(host, sizeof(host) - 1);
	unsigned hh = 2166136261u;
	for(const char *q = host; *q; q++) hh = (hh ^ (unsigned char) *q) * 16777619u;
	int v[3] = { fail, (int) (hh & 0x3fffffff), -(int) (hh & 0x3fffffff) };
	CCK(cudaMemcpyAsync(probe, v, sizeof(v), cudaMemcpyHostToDevice, stream));
	if(comm_allreduce_max(c, probe, 3, stream)) return 2;
	CCK(cudaMemcpyAsync(v, probe, sizeof(v), cudaMemcpyDeviceToHost, stream));
	CCK(cudaStreamSynchronize(stream));
	cudaFree(probe);
	fail = v[0] || v[1] != -v[2];
	c->p2p = !fail;
	if(!c->p2p) return 0;
	c->table = (ExportEntry *) calloc((size_t) c->n * X_COUNT, sizeof(ExportEntry));
	CCK(cudaMalloc(&c->dtable, (size_t) c->n * X_COUNT * sizeof(ExportEntry)));
	CCK(cudaMalloc(&c->flags, CH_COUNT * PEER_MAX * sizeof(int)));
	CCK(cudaMemset(c->flags, 0, CH_COUNT * PEER_MAX * sizeof(int)));
	const size_t fin = (size_t) SET_MAX_SPECIES * 2 * FAR_SECTION * sizeof(double);
	CCK(cudaMalloc(&c->far_in, fin));
	CCK(cudaMemset(c->far_in, 0, fin));
	if(comm_p2p_export(c, X_FLAGS, c->flags, CH_COUNT * PEER_MAX * sizeof(int))) return 2;
	if(comm_p2p_export(c, X_FAR_IN, c->far_in, fin)) return 2;
	if(comm_p2p_export(c, X_RHO_RECV, c->rho_recv, (size_t) c->g.nx * sizeof(double))) return 2;
	if(comm_p2p_export(c, X_FFT_TB, c->tb, (size_t) c->g.ny * c->cw * c->n * sizeof(cufftDoubleComplex))) return 2;
	if(comm_p2p_export(c, X_FFT_A, c->a, (size_t) c->g.ny * c->nc_ * sizeof(cufftDoubleComplex))) return 2;
	return 0;
}

/* comm_send_ghost_rho + comm_recv_ghost_rho, reference src/comm_field.c:51-136 */
int
comm_rho_halo(Comm *c, double *rho, cudaStream_t stream, int *errflag, long long *launches)
{
	const Geom &g = c->g;
	const int south = (c->rank + 1) % c->n, north = (c->rank + c->n - 1) % c->n;
	if(c->p2p)
	{
		/* the ghost row goes straight into the south rank's receive row (peer memory) */
		CCK(cudaMemcpyAsync(c->remote[X_RHO_RECV][south], rho + (size_t) g.ny * g.S, (size_t) g.nx * sizeof(double),
					cudaMemcpyDefault, stream));
		if(meet_neighbours(c, CH_RHO, stream, errflag, launches)) return 2;
	}
	else
	{
		NCK(g_nccl.GroupStart());
		NCK(g_nccl.Send(rho + (size_t) g.ny * g.S, (size_t) g.nx, ncclFloat64, south, c->nc, stream));
		NCK(g_nccl.Recv(c->rho_recv, (size_t) g.nx, ncclFloat64, north, c->nc, stream));
		NCK(g_nccl.GroupEnd());
	}
	k_rho_fold<<<(g.nx + 127) / 128, 128, 0, stream>>>(rho, c->rho_recv, g);
	CCK(cudaGetLastError());
	if(launches) (*launches)++;
	return 0;
}

/* comm_phi_send + comm_phi_recv, reference src/comm_field.c:139-201 */
int
comm_phi_halo(Comm *c, double *phi, cudaStream_t stream, int *errflag, long long *launches)
{
	const Geom &g = c->g;
	const int south = (c->rank + 1) % c->n, north = (c->rank + c->n - 1) % c->n;
	const size_t S = (size_t) g.S;
	if(c->p2p)
	{
		/* phi is exported by sim.cu (X_PHI): the rows go straight into the neighbours' ghost rows */
		double *pn = (double *) c->remote[X_PHI][north], *ps = (double *) c->remote[X_PHI][south];
		if(!pn || !ps) { snprintf(c->err, c->errlen, "phi is not mapped on the neighbour ranks"); return 2; }
		CCK(cudaMemcpyAsync(pn + (size_t) (g.ny + 1) * S, phi + 1 * S, 2 * S * sizeof(double), cudaMemcpyDefault, stream));
		CCK(cudaMemcpyAsync(ps, phi + (size_t) g.ny * S, S * sizeof(double), cudaMemcpyDefault, stream));
		return meet_neighbours(c, CH_PHI, stream, errflag, launches);
	}
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

/* Peer-memory path: the leavers that cross a face were written into the neighbours' ghost outbox
 * rows by the push itself; what is left are the (rare) far movers. pack: the lists k_far_insert
 * left for the north / south rank go into that rank's inbox (peer memory); else: what the
 * neighbours left in this rank's inbox joins the local far-mover list. Grid (1, 2 * species). */
struct FarShipJob {
	SpeciesSet set;
	double *buf[2];          /* pack: [north rank's inbox, south rank's inbox]; unpack: twice the own inbox */
	int pack;
};

static __global__ void __launch_bounds__(256)
k_far_ship(const __grid_constant__ FarShipJob job, int *__restrict__ errflag)
{
	const int is = blockIdx.y >> 1, dir = blockIdx.y & 1;
	/* section [species][face]: what travels north arrives "from the south" in the north rank's
	 * section 0, what travels south in the south rank's section 1 */
	far_face(job.set.sp[is], dir, job.buf[dir] + (size_t) (is * 2 + dir) * FAR_SECTION, job.pack, errflag);
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
	if(c->p2p)
	{
		FarShipJob job;
		job.set.n = nsp;
		for(int i = 0; i < nsp; i++) { job.set.sp[i] = *sps[i]; job.set.arr[i] = arrs[i]; }
		job.buf[0] = (double *) c->remote[X_FAR_IN][north];
		job.buf[1] = (double *) c->remote[X_FAR_IN][south];
		job.pack = 1;
		k_far_ship<<<dim3(1, 2 * nsp), 256, 0, stream>>>(job, errflag);
		/* the push's stores into the neighbours' ghost rows and the far movers are complete before
		 * the neighbours read them, and theirs before this rank does */
		if(meet_neighbours(c, CH_PART, stream, errflag, launches)) return 2;
		job.buf[0] = job.buf[1] = c->far_in;
		job.pack = 0;
		k_far_ship<<<dim3(1, 2 * nsp), 256, 0, stream>>>(job, errflag);
		CCK(cudaGetLastError());
		if(launches) *launches += 2;
		return 0;
	}
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
		/* this rank's mappings of its peers' memory first, then its own buffers */
	for(int what = 0; what < X_COUNT; what++)
		for(int r = 0; r < PEER_MAX; r++)
			if(c->remote[what][r]) cudaIpcCloseMemHandle(c->remote[what][r]);
	free(c->table);
	cudaFree(c->dtable); cudaFree(c->flags); cudaFree(c->far_in);
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

/* The transposes of the distributed transform over peer memory: the pack kernel IS the exchange.
 * Forward: element (iy, k) of this rank's row spectra goes straight to row row0+iy, column k % cw of
 * rank k / cw's column array; inverse: this rank's columns go back into the owners' row spectra. */
struct FftPeers { cufftDoubleComplex *p[PEER_MAX]; };

static __global__ void
k_fft_scatter_fwd(const cufftDoubleComplex *__restrict__ a, const __grid_constant__ FftPeers tb,
		int ny, int nc, int cw, int n, int row0)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;     /* padded column */
	const int iy = blockIdx.y;
	if(k >= cw * n) return;
	const int r = k / cw, kl = k % cw;
	cufftDoubleComplex v = { 0.0, 0.0 };
	if(k < nc) v = a[(size_t) iy * nc + k];
	tb.p[r][(size_t) (row0 + iy) * cw + kl] = v;
}

static __global__ void
k_fft_scatter_inv(const cufftDoubleComplex *__restrict__ tb, const __grid_constant__ FftPeers a,
		int ny, int nc, int cw, int rank)
{
	const int kl = blockIdx.x * blockDim.x + threadIdx.x;
	const int row = blockIdx.y;              /* global row: rank row / ny owns it */
	const int k = rank * cw + kl;
	if(kl >= cw || k >= nc) return;
	a.p[row / ny][(size_t) (row % ny) * nc + k] = tb[(size_t) row * cw + kl];
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
comm_solve(Comm *c, const double *rho, double *phi_raw, cudaStream_t stream, int *errflag, long long *launches)
{
	const Geom &g = c->g;
	const int ncp = c->cw * c->n;
	FCK(cufftExecD2Z(c->rows_fwd, (double *) rho, c->a));
	if(c->p2p)
	{
		FftPeers tbs, as;
		for(int r = 0; r < c->n; r++)
		{
			tbs.p[r] = r == c->rank ? c->tb : (cufftDoubleComplex *) c->remote[X_FFT_TB][r];
			as.p[r] = r == c->rank ? c->a : (cufftDoubleComplex *) c->remote[X_FFT_A][r];
		}
		k_fft_scatter_fwd<<<dim3((ncp + 127) / 128, g.ny), 128, 0, stream>>>(c->a, tbs, g.ny, c->nc_, c->cw, c->n, g.row0);
		if(meet_all(c, CH_FFT_FWD, stream, errflag, launches)) return 2;
		FCK(cufftExecZ2Z(c->cols, c->tb, c->tb, CUFFT_FORWARD));
		const size_t n = (size_t) g.ny_glob * c->cw;
		int blocks = (int) ((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
		k_green<<<blocks, 256, 0, stream>>>(c->tb, c->GT, n);
		FCK(cufftExecZ2Z(c->cols, c->tb, c->tb, CUFFT_INVERSE));
		k_fft_scatter_inv<<<dim3((c->cw + 127) / 128, g.ny_glob), 128, 0, stream>>>(c->tb, as, g.ny, c->nc_, c->cw, c->rank);
		if(meet_all(c, CH_FFT_INV, stream, errflag, launches)) return 2;
		FCK(cufftExecZ2D(c->rows_inv, c->a, phi_raw));
		CCK(cudaGetLastError());
		if(launches) *launches += 3;
		return 0;
	}
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
