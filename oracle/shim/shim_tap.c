/* ORACLE TEST INFRASTRUCTURE. The MFT_TAP solver (src/mft_tap.c, src/tap.c) spawns
 * FFTW worker processes over MPI-3 shared windows; it is out of scope
 * and only reachable with simulation.solver = "MFT_TAP". */
#include <stdio.h>
#include <stdlib.h>
struct sim; struct solver;
int MFT_TAP_init(struct sim *sim, struct solver *s) { (void) sim; (void) s; fprintf(stderr, "MFT_TAP is not available in the oracle build\n"); abort(); }
int MFT_TAP_solve(struct solver *s) { (void) s; abort(); }
int MFT_TAP_end(struct solver *s) { (void) s; return 0; }
