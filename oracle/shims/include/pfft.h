/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * PFFT-API stand-in (PFFT 1.0.8-alpha3-fftw3, depends/Makefile.pfft:3 of the
 * reference, is downloaded at build time and is not in the tree).  Only the
 * single-precision entry points pmpfft.c binds (pmpfft.c:15-42,160-175,
 * 281-303,376-397) are provided, for ONE rank: local sizes are the global
 * sizes; r2c is an unnormalised forward DFT of the padded real array
 * [n0][n1][2*(n2/2+1)], written as [n0][n1][n2/2+1] complex, or as
 * [n1][n2/2+1][n0] when PFFT_TRANSPOSED_OUT is given; c2r is the
 * unnormalised inverse.  The arithmetic is ../src/cpufft.c.
 */
#ifndef ORACLE_SHIM_PFFT_H
#define ORACLE_SHIM_PFFT_H
#include <stddef.h>
#include <mpi.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef float pfftf_complex[2];
typedef double pfft_complex[2];
typedef struct oracle_fft3_plan *pfftf_plan;
typedef struct oracle_fft3_plan *pfft_plan;

#define PFFT_FORWARD (-1)
#define PFFT_BACKWARD (+1)
#define PFFT_TRANSPOSED_NONE 0u
#define PFFT_TRANSPOSED_IN  (1u << 0)
#define PFFT_TRANSPOSED_OUT (1u << 1)
#define PFFT_PADDED_R2C (1u << 2)
#define PFFT_PADDED_C2R (1u << 3)
#define PFFT_ESTIMATE (1u << 4)
#define PFFT_MEASURE (1u << 5)
#define PFFT_TUNE (1u << 6)
#define PFFT_DESTROY_INPUT (1u << 7)
#define PFFT_PRESERVE_INPUT (1u << 8)

void pfftf_init(void);
void pfftf_cleanup(void);
void pfft_init(void);
void pfft_cleanup(void);
int pfft_create_procmesh(int rnk, MPI_Comm comm, const int *np, MPI_Comm *comm_cart);
ptrdiff_t pfft_local_size_dft_r2c(int rnk_n, const ptrdiff_t *n, MPI_Comm comm_cart, unsigned pfft_flags,
        ptrdiff_t *local_ni, ptrdiff_t *local_i_start, ptrdiff_t *local_no, ptrdiff_t *local_o_start);
pfftf_plan pfftf_plan_dft_r2c(int rnk_n, const ptrdiff_t *n, float *in, pfftf_complex *out,
        MPI_Comm comm_cart, int sign, unsigned pfft_flags);
pfftf_plan pfftf_plan_dft_c2r(int rnk_n, const ptrdiff_t *n, pfftf_complex *in, float *out,
        MPI_Comm comm_cart, int sign, unsigned pfft_flags);
void pfftf_execute_dft_r2c(const pfftf_plan plan, float *in, pfftf_complex *out);
void pfftf_execute_dft_c2r(const pfftf_plan plan, pfftf_complex *in, float *out);
void pfftf_destroy_plan(pfftf_plan plan);
#ifdef __cplusplus
}
#endif
#endif
