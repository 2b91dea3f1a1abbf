/* ORACLE shim: the reference's optional FFTW-MPI slab path (`-f`, pmpfft.c:67-106,
 * 266-279) is never taken by the oracle; these declarations only let
 * pmpfft.c compile, the definitions abort(). */
#ifndef ORACLE_SHIM_FFTW3_H
#define ORACLE_SHIM_FFTW3_H
#include <stddef.h>
typedef float fftwf_complex[2];
typedef double fftw_complex[2];
typedef struct oracle_fft3_plan *fftwf_plan;
typedef struct oracle_fft3_plan *fftw_plan;
#define FFTW_ESTIMATE (1u << 6)
#define FFTW_DESTROY_INPUT (1u << 0)
void fftwf_destroy_plan(fftwf_plan p);
void fftw_destroy_plan(fftw_plan p);
#endif
