/* The run loop of the reference's command-line program (src/fastpm.c: prepare_deltak :415-545, prepare_cdm, check_snapshots
 * :1130-1208, take_a_snapshot :1473-1486, write_powerspectrum) without the Lua front end, written with the reference's API names
 * only and linked with libfastpm_b200.so instead of libfastpm.a / libfastpmio.a -- this file contains NO fastpm_b200_* call:
 *
 *   cli_like_example powerspec.txt outbase nc boxsize seed aout1,aout2,...
 *
 * Gaussian initial conditions from the seed (Gadget scheme) coloured by the P(k) table, 2LPT at a = 0.1, five COLA steps to a = 1
 * on a mesh twice as fine as the particle grid, a snapshot "<outbase>_<aout>" for every requested scale factor, a power spectrum
 * file "<outbase>_powerspec_<a>.txt" after every force evaluation.
 * tests/test_abi_layout.py compiles and links it (CPU); tests/first_gpu_run_cases.py runs it on the GPU against the reference. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "fastpm_b200_api.h"

typedef struct { const char *base; double aout[32]; int nout, iout; } RunData;

static int check_snapshots(FastPMSolver *fastpm, FastPMInterpolationEvent *event, RunData *prr)
{
    for (int iout = prr->iout; iout < prr->nout; iout++) {
        if (event->a1 == event->a2) {
            if (event->a1 != prr->aout[iout]) continue;
        } else {
            if (event->a1 >= prr->aout[iout]) continue;
            if (event->a2 < prr->aout[iout]) continue;
        }
        FastPMSolver snapshot[1];
        FastPMStore cdm[1];
        memcpy(snapshot, fastpm, sizeof(FastPMSolver));
        fastpm_solver_add_species(snapshot, FASTPM_SPECIES_CDM, cdm);
        fastpm_set_snapshot(fastpm, snapshot, event->drift, event->kick, prr->aout[iout]);

        char *filebase = fastpm_strdup_printf("%s_%0.04f", prr->base, prr->aout[iout]);
        write_snapshot_header(snapshot, filebase, fastpm->comm);
        fastpm_store_write(cdm, filebase, "w", 0, fastpm->comm);
        fastpm_info("snapshot %s written at a = %6.4f\n", filebase, prr->aout[iout]);
        free(filebase);

        fastpm_unset_snapshot(fastpm, snapshot, event->drift, event->kick, prr->aout[iout]);
        prr->iout = iout + 1;
    }
    return 0;
}

static int write_powerspectrum(FastPMSolver *fastpm, FastPMForceEvent *event, RunData *prr)
{
    FastPMPowerSpectrum ps;
    fastpm_powerspectrum_init_from_delta(&ps, event->pm, event->delta_k, event->delta_k);
    char *fn = fastpm_strdup_printf("%s_powerspec_%0.04f.txt", prr->base, event->a_f);
    fastpm_path_ensure_dirname(fn);
    if (fastpm->ThisTask == 0) fastpm_powerspectrum_write(&ps, fn, event->N);
    free(fn);
    fastpm_powerspectrum_destroy(&ps);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 7) { fprintf(stderr, "usage: %s powerspec.txt outbase nc boxsize seed aout1,aout2,...\n", argv[0]); return 2; }
    libfastpm_init();
    fastpm_set_msg_handler(fastpm_void_msg_handler, MPI_COMM_WORLD, NULL);
    RunData prr[1];
    memset(prr, 0, sizeof(prr));
    prr->base = argv[2];
    char **parts = fastpm_strsplit(argv[6], ",");
    for (int i = 0; parts[i] && prr->nout < 32; i++) prr->aout[prr->nout++] = atof(parts[i]);
    free(parts);

    FastPMCosmology cosmology;
    memset(&cosmology, 0, sizeof(cosmology));
    cosmology.h = 0.6774; cosmology.Omega_m = 0.307494; cosmology.T_cmb = 0; cosmology.Omega_k = 0; cosmology.w0 = -1; cosmology.wa = 0;
    cosmology.N_eff = 3.046; cosmology.N_nu = 0; cosmology.N_ncdm = 0; cosmology.ncdm_matterlike = 1; cosmology.ncdm_freestreaming = 1;
    cosmology.growth_mode = FASTPM_GROWTH_MODE_LCDM;
    VPMInit vpminit[] = { { .a_start = 0, .pm_nc_factor = 2 }, { .a_start = 1, .pm_nc_factor = 0 } };
    FastPMConfig config;
    memset(&config, 0, sizeof(config));
    config.nc = atoi(argv[3]); config.boxsize = atof(argv[4]); config.alloc_factor = 2.0; config.lpt_nc_factor = 1; config.cosmology = &cosmology;
    config.vpminit = vpminit; config.nLPT = -2.5; config.PAINTER_TYPE = FASTPM_PAINTER_CIC; config.painter_support = 2;
    config.FORCE_TYPE = FASTPM_FORCE_COLA; config.KERNEL_TYPE = FASTPM_KERNEL_1_4; config.SOFTENING_TYPE = FASTPM_SOFTENING_NONE;

    FastPMSolver fastpm[1];
    fastpm_solver_init(fastpm, &config, MPI_COMM_WORLD);

    /* prepare_deltak, src/fastpm.c:415-545 (seed + linear power spectrum table) */
    char *content = fastpm_file_get_content(argv[1]);
    if (!content) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    FastPMPowerSpectrum linear_powerspectrum;
    if (0 != fastpm_powerspectrum_init_from_string(&linear_powerspectrum, content)) { fprintf(stderr, "cannot parse %s\n", argv[1]); return 2; }
    free(content);
    FastPMFloat *delta_k = pm_alloc(fastpm->lptpm);
    fastpm_ic_fill_gaussiank(fastpm->lptpm, delta_k, atoi(argv[5]), FASTPM_DELTAK_GADGET);
    fastpm_ic_induce_correlation(fastpm->lptpm, delta_k, (fastpm_fkfunc) fastpm_powerspectrum_eval2, &linear_powerspectrum);
    ptrdiff_t mode[4] = { 0, 0, 0, 0 };
    fastpm_apply_modify_mode_transfer(fastpm->lptpm, delta_k, delta_k, mode, 1.0);

    double time_step[] = { 0.1, 0.325, 0.55, 0.775, 1.0 };
    fastpm_solver_setup_lpt(fastpm, FASTPM_SPECIES_CDM, delta_k, NULL, time_step[0]);
    pm_free(fastpm->lptpm, delta_k);
    fastpm_powerspectrum_destroy(&linear_powerspectrum);

    fastpm_add_event_handler(&fastpm->event_handlers, FASTPM_EVENT_FORCE, FASTPM_EVENT_STAGE_AFTER,
                             (FastPMEventHandlerFunction) write_powerspectrum, prr);
    fastpm_add_event_handler(&fastpm->event_handlers, FASTPM_EVENT_INTERPOLATION, FASTPM_EVENT_STAGE_BEFORE,
                             (FastPMEventHandlerFunction) check_snapshots, prr);
    fastpm_solver_evolve(fastpm, time_step, 5);

    FastPMStore *cdm = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM);
    printf("cli_like_example: %td particles at a_x = %g, a_v = %g, %d snapshots\n", (ptrdiff_t) cdm->np, cdm->meta.a_x, cdm->meta.a_v, prr->iout);
    fastpm_solver_destroy(fastpm);
    libfastpm_cleanup();
    return 0;
}
