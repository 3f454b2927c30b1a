/* ORACLE TEST INFRASTRUCTURE. See fftw3.h in this directory. */
#ifndef ORACLE_SHIM_FFTW3_MPI_H
#define ORACLE_SHIM_FFTW3_MPI_H

#include "fftw3.h"
#include <mpi.h>

void fftw_mpi_init(void);
ptrdiff_t fftw_mpi_local_size_2d(ptrdiff_t n0, ptrdiff_t n1, MPI_Comm comm,
		ptrdiff_t *local_n0, ptrdiff_t *local_0_start);
fftw_plan fftw_mpi_plan_dft_r2c_2d(ptrdiff_t n0, ptrdiff_t n1, double *in,
		fftw_complex *out, MPI_Comm comm, unsigned flags);
fftw_plan fftw_mpi_plan_dft_c2r_2d(ptrdiff_t n0, ptrdiff_t n1, fftw_complex *in,
		double *out, MPI_Comm comm, unsigned flags);

#endif
