/* placeholder, replaced below */
#include "comm.h"
#include <stdio.h>
struct Comm { int dummy; };
int comm_unique_id(void *, char *err, size_t n) { snprintf(err, n, "multi-rank support not built"); return 2; }
Comm *comm_create(const void *, int, int, const Geom &, cudaStream_t, char *err, size_t n) { snprintf(err, n, "multi-rank support not built"); return NULL; }
void comm_destroy(Comm *) {}
int comm_rho_halo(Comm *, double *, cudaStream_t, long long *) { return 2; }
int comm_phi_halo(Comm *, double *, cudaStream_t) { return 2; }
int comm_particles(Comm *, SpeciesDev *, int, const Geom &, int, cudaStream_t, int *, long long *) { return 2; }
int comm_solve(Comm *, const double *, double *, cudaStream_t, long long *) { return 2; }
