/* ORACLE / TEST INFRASTRUCTURE ONLY.
 * abort() definitions for symbols that the compiled reference files reference from
 * out-of-scope reference files (ncdm, lightcone, healpix, mpsort: SURVEY.md section 8c);
 * none is reached in the configurations the oracle runs (ncdm off, pgdc off). */
#include <stdio.h>
#include <stdlib.h>
/* -------- symbols referenced from out-of-scope reference files (never reached) */
#define UNREACHED(name) do { fprintf(stderr, "oracle: out-of-scope reference function %s reached\n", name); abort(); } while (0)
void delta_nu_from_power(void) { UNREACHED("delta_nu_from_power"); }
void fastpm_fd_interp_init(void *p) { (void) p; UNREACHED("fastpm_fd_interp_init"); }
void fastpm_fd_interp_destroy(void *p) { (void) p; UNREACHED("fastpm_fd_interp_destroy"); }
double fastpm_do_fd_interp(void *p, int id, double y) { (void) p; (void) id; (void) y; UNREACHED("fastpm_do_fd_interp"); return 0; }
int fastpm_lc_inside(void) { UNREACHED("fastpm_lc_inside"); return 0; }
void mpsort_mpi_newarray(void) { UNREACHED("mpsort_mpi_newarray"); }
long nside2npix(long n) { (void) n; UNREACHED("nside2npix"); return 0; }
void pix2ang_ring(void) { UNREACHED("pix2ang_ring"); }
void pix2vec_ring(void) { UNREACHED("pix2vec_ring"); }
