/* ORACLE shim: see fftw3.h */
#ifndef ORACLE_SHIM_FFTW3_MPI_H
#define ORACLE_SHIM_FFTW3_MPI_H
#include <fftw3.h>
#include <mpi.h>
#define FFTW_MPI_TRANSPOSED_IN (1u << 29)
#define FFTW_MPI_TRANSPOSED_OUT (1u << 30)
ptrdiff_t fftw_mpi_local_size(int rnk, const ptrdiff_t *n, MPI_Comm comm, ptrdiff_t *local_n0, ptrdiff_t *local_0_start);
ptrdiff_t fftw_mpi_local_size_transposed(int rnk, const ptrdiff_t *n, MPI_Comm comm,
        ptrdiff_t *local_n0, ptrdiff_t *local_0_start, ptrdiff_t *local_n1, ptrdiff_t *local_1_start);
fftwf_plan fftwf_mpi_plan_dft_r2c(int rnk, const ptrdiff_t *n, float *in, fftwf_complex *out, MPI_Comm comm, unsigned flags);
fftwf_plan fftwf_mpi_plan_dft_c2r(int rnk, const ptrdiff_t *n, fftwf_complex *in, float *out, MPI_Comm comm, unsigned flags);
void fftwf_mpi_execute_dft_r2c(fftwf_plan p, float *in, fftwf_complex *out);
void fftwf_mpi_execute_dft_c2r(fftwf_plan p, fftwf_complex *in, float *out);
#endif
