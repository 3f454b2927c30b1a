/* ORACLE TEST INFRASTRUCTURE — not product code.
 * Stand-in for <fftw3.h>/<fftw3-mpi.h>: the reference's MFT solver calls FFTW 3.3.6's
 * MPI r2c/c2r 2D plans (src/solver.c:246,254,290-330,485,491). FFTW is a third-party
 * dependency that is not vendored; shim_fftw.c restates the published transform
 * (unnormalised DFT, padded in-row real layout 2*(n1/2+1)) in plain double precision. */
#ifndef ORACLE_SHIM_FFTW3_H
#define ORACLE_SHIM_FFTW3_H

#include <stddef.h>
#include <complex.h>

typedef double _Complex fftw_complex;
typedef struct shim_fftw_plan *fftw_plan;

#define FFTW_MEASURE 0U
#define FFTW_ESTIMATE (1U << 6)

void *fftw_malloc(size_t n);
fftw_complex *fftw_alloc_complex(size_t n);
double *fftw_alloc_real(size_t n);
void fftw_free(void *p);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
int fftw_init_threads(void);
void fftw_plan_with_nthreads(int nthreads);

#endif
