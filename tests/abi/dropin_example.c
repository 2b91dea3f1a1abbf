/* A libfastpm user program for the PM force step, written against the reference's API names only (cf. the reference's
 * tests/testpm.c:21-109 and src/fastpm.c:186-397) and linked with libfastpm_b200.so instead of libfastpm.a:
 *
 *   dropin_example delta_k.f32 steps.f64 out_x.f64 out_pk.f64
 *
 * reads a linear density field (host, reference layout of pm_alloc(solver->lptpm)), sets up 2LPT initial conditions at the
 * first time step, evolves through the given steps with a FORCE/after handler that measures P(k), and writes the final
 * positions and the last P(k).  The only lines that differ from a libfastpm program are the two host <-> device mirror calls
 * (fastpm_b200_mesh_set_complex, fastpm_b200_store_get_column): mesh buffers and store columns are device memory.
 * tests/test_abi_layout.py compiles and links it (CPU); tests/test_zz_gpu_ic.py runs it on the GPU against the fixture. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "fastpm_b200_api.h"

static double last_pk[64], last_k[64];
static int last_n = 0;

static int measure_pk(FastPMSolver *solver, FastPMForceEvent *event, void *userdata)
{
    (void) solver; (void) userdata;
    FastPMPowerSpectrum ps;
    fastpm_powerspectrum_init_from_delta(&ps, event->pm, event->delta_k, event->delta_k);
    last_n = ps.base.size < 64 ? (int) ps.base.size : 64;
    memcpy(last_k, ps.base.k, sizeof(double) * last_n);
    memcpy(last_pk, ps.base.f, sizeof(double) * last_n);
    fastpm_powerspectrum_destroy(&ps);
    return 0;
}

static void *slurp(const char *fn, size_t *bytes)
{
    FILE *f = fopen(fn, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", fn); exit(2); }
    fseek(f, 0, SEEK_END); *bytes = (size_t) ftell(f); fseek(f, 0, SEEK_SET);
    void *p = malloc(*bytes);
    if (fread(p, 1, *bytes, f) != *bytes) exit(2);
    fclose(f);
    return p;
}

int main(int argc, char **argv)
{
    if (argc < 5) { fprintf(stderr, "usage: %s delta_k.f32 steps.f64 out_x.f64 out_pk.f64\n", argv[0]); return 2; }
    libfastpm_init();
    FastPMCosmology cosmology;
    memset(&cosmology, 0, sizeof(cosmology));
    cosmology.h = 0.6774; cosmology.Omega_m = 0.307494; cosmology.T_cmb = 0; cosmology.Omega_k = 0; cosmology.w0 = -1; cosmology.wa = 0;
    cosmology.N_eff = 3.046; cosmology.N_nu = 0; cosmology.N_ncdm = 0; cosmology.ncdm_matterlike = 1; cosmology.ncdm_freestreaming = 1;
    cosmology.growth_mode = FASTPM_GROWTH_MODE_LCDM;
    VPMInit vpminit[] = { { .a_start = 0, .pm_nc_factor = 2 }, { .a_start = 1, .pm_nc_factor = 0 } };
    FastPMConfig config;
    memset(&config, 0, sizeof(config));
    config.nc = 16; config.boxsize = 32.; config.alloc_factor = 2.0; config.lpt_nc_factor = 1; config.cosmology = &cosmology;
    config.vpminit = vpminit; config.nLPT = -2.5; config.PAINTER_TYPE = FASTPM_PAINTER_CIC; config.painter_support = 2;
    config.FORCE_TYPE = FASTPM_FORCE_FASTPM; config.KERNEL_TYPE = FASTPM_KERNEL_1_4; config.SOFTENING_TYPE = FASTPM_SOFTENING_NONE;

    FastPMSolver solver[1];
    fastpm_solver_init(solver, &config, MPI_COMM_WORLD);

    size_t nbytes, sbytes;
    float *delta_k_host = slurp(argv[1], &nbytes);
    double *time_step = slurp(argv[2], &sbytes);
    const int nstep = (int) (sbytes / sizeof(double));
    if (nbytes != sizeof(float) * fastpm_b200_mesh_host_size(solver->lptpm)) { fprintf(stderr, "delta_k has the wrong size\n"); return 2; }

    FastPMFloat *delta_k = pm_alloc(solver->lptpm);
    fastpm_b200_mesh_set_complex(solver->lptpm, delta_k, delta_k_host);
    fastpm_solver_setup_lpt(solver, FASTPM_SPECIES_CDM, delta_k, NULL, time_step[0]);
    pm_free(solver->lptpm, delta_k);

    fastpm_add_event_handler(&solver->event_handlers, FASTPM_EVENT_FORCE, FASTPM_EVENT_STAGE_AFTER,
                             (FastPMEventHandlerFunction) measure_pk, NULL);
    fastpm_solver_evolve(solver, time_step, nstep);

    FastPMStore *cdm = fastpm_solver_get_species(solver, FASTPM_SPECIES_CDM);
    const size_t np = cdm->np;
    double *x = malloc(sizeof(double) * 3 * np);
    if (fastpm_b200_store_get_column(cdm, COLUMN_POS, x, 0, np) != 0) return 3;
    FILE *f = fopen(argv[3], "wb"); fwrite(x, sizeof(double), 3 * np, f); fclose(f);
    f = fopen(argv[4], "wb"); fwrite(last_k, sizeof(double), last_n, f); fwrite(last_pk, sizeof(double), last_n, f); fclose(f);
    printf("dropin_example: %zu particles, a_x = %g, a_v = %g, P(k) bins %d\n", np, cdm->meta.a_x, cdm->meta.a_v, last_n);

    fastpm_solver_destroy(solver);
    libfastpm_cleanup();
    free(x); free(delta_k_host); free(time_step);
    return 0;
}
