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
/* -------- libfastpmio/io.c (snapshot writer, row N2): linear-response neutrino tables are off, the mesh/healpix writers unused */
void mpsort_mpi(void) { UNREACHED("mpsort_mpi"); }
void ncdm_lr_save_neutrinos(void *bf, int task) { (void) bf; (void) task; }        /* neutrinos_lra.c:329-332: returns at once when the table was never initialised */
int ncdm_lr_read_neutrinos(void *bf, int task) { (void) bf; (void) task; return 1; }  /* neutrinos_lra.c:405-411: no "Neutrino" block */
long nside2npix64(long n) { (void) n; UNREACHED("nside2npix64"); return 0; }
void vec2pix_nest64(void) { UNREACHED("vec2pix_nest64"); }
const char *LIBFASTPM_VERSION = "1.0.oracle";     /* libfastpm/Makefile:66 generates "1.0.<git describe>" */
long nside2npix(long n) { (void) n; UNREACHED("nside2npix"); return 0; }
void pix2ang_ring(void) { UNREACHED("pix2ang_ring"); }
void pix2vec_ring(void) { UNREACHED("pix2vec_ring"); }
