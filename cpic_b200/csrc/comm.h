/* Multi-GPU layer of cpic_b200: one Y slab per rank, NCCL over NVLink.
 * Replaces the reference's MPI traffic: rho ghost row (src/comm_field.c:51-136), phi
 * ghost rows (:139-201), particles crossing a slab face (src/comm_plasma.c:887-1120)
 * and FFTW-MPI's transposes inside the MFT solver (src/solver.c:314-330, :485-491). */
#ifndef CPIC_B200_COMM_H
#define CPIC_B200_COMM_H

#include <cuda_runtime.h>
#include <stddef.h>

#include "geom.h"

struct SpeciesDev;
struct Comm;

int comm_unique_id(void *id128, char *err, size_t errlen);
Comm *comm_create(const void *id128, int rank, int nranks, const Geom &g, cudaStream_t stream,
		char *err, size_t errlen);
void comm_destroy(Comm *c);

/* element-wise maximum of n ints over the ranks, in place */
int comm_allreduce_max(Comm *c, int *dev, int n, cudaStream_t stream);

/* rho: ghost row ny -> rank+1, received row added to row 0 */
int comm_rho_halo(Comm *c, double *rho, cudaStream_t stream, int *errflag, long long *launches);
/* phi: slab rows 0,1 -> rank-1's south ghosts; slab row ny-1 -> rank+1's north ghost */
int comm_phi_halo(Comm *c, double *phi, cudaStream_t stream, int *errflag, long long *launches);
/* particles: outbox entries of the edge block rows that leave the slab are sent to the
 * neighbour ranks and land in the ghost outbox rows nb .. nb+2*nbx */
int comm_particles(Comm *c, SpeciesDev *const *sps, const int *arrs, int nsp, const Geom &g, int nb,
		cudaStream_t stream, int *errflag, long long *launches);
/* distributed MFT solve: rho slab rows -> unnormalised phi slab rows (ny x S) */
int comm_solve(Comm *c, const double *rho, double *phi_raw, cudaStream_t stream, int *errflag, long long *launches);

/* ---- peer memory (CUDA IPC over NVLink between the ranks of one box). When every rank can
 * export its buffers (decided collectively when the communicator is created; CPIC_B200_P2P=0
 * forces the NCCL path), the exchanges above are plain stores into the neighbour's memory -- the
 * push writes a leaver that crosses a slab face straight into the neighbour's ghost outbox row --
 * followed by a flag handshake; NCCL then only carries the handles and the error words. ---- */
bool comm_p2p(const Comm *c);
enum { COMM_EXPORT_PHI = 2 };          /* X_PHI of comm.cu */
int comm_export_species(int is, int with_E);
/* registers / replaces / withdraws (ptr NULL) one exported allocation; seen by the peers after the
 * next comm_p2p_refresh (collective) */
int comm_p2p_export(Comm *c, int what, void *ptr, size_t bytes);
int comm_p2p_unmap(Comm *c, int what);
int comm_p2p_refresh(Comm *c, cudaStream_t stream);
/* rank's allocation `what` mapped into this process (NULL: not mapped) */
void *comm_p2p_remote(const Comm *c, int what, int rank);
int comm_rank_north(const Comm *c);
int comm_rank_south(const Comm *c);

#endif
