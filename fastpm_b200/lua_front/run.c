/* fastpm_b200_run -- the FastPM command-line run for the force step / integrator path, driven by the reference's own Lua
 * parameter surface.
 *
 *     fastpm_b200_run [-n NGPUS] [-r restart_snapshot] [-W Nwriters] [-m MemoryMB] [--dump-config] paramfile.lua [args ...]
 *
 * -n N (2, 4 or 8; where the reference says `mpirun -n N fastpm`): the process forks N ranks, one per GPU, x-slabs of the mesh, BEFORE
 * anything touches a device; they find each other through a shared-memory segment the parent made (fastpm_b200_comm_init_local).
 *
 * The parameter file is evaluated by the reference's Lua runtime (vendored lua/ + src/lua-runtime-{dump,config,fastpm}.lua,
 * compiled in place by this directory's Makefile; schema and validation are the reference's, lua-runtime-fastpm.lua:14-346) and
 * read through the accessors that runtime generates (lua-config.h, CONF / HAS of src/param.h).  What follows restates the run
 * loop of src/fastpm.c for the rows of SURVEY.md section 8 on top of libfastpm_b200.so:
 *
 *     main            src/fastpm.c:121-228   CONF() -> FastPMConfig, VPMInit from pm_nc_factor, cosmology (src/prepare.c:19-41)
 *     run             :265-397               solver init, handlers, initial conditions, fastpm_solver_evolve
 *     prepare_deltak  :415-586               seed / white-noise file / linear-density file, remove_cosmic_variance, set_mode,
 *                                            inverted_ic, induce_correlation, rescale to a = 1, DC mode
 *     prepare_cdm     :612-717               restart (-r), write_lineark / write_powerspectrum "_linear.txt", setup_lpt
 *     handlers        :1144-1208 check_snapshots, :1403-1486 take_a_snapshot (write_snapshot, write_nonlineark, sort_snapshot),
 *                     :1576-1604 print_transition, :1650-1668 report_lpt, :1671-1708 report_domain, :1711-1776 write_powerspectrum
 * with the same log lines (the reference's tests grep them: tests/run-test-*.sh).  Options outside the path -- light cones,
 * FOF / RFOF, neutrinos, primordial non-Gaussianity, constrained ICs, GrafIC / RunPB files, particle subsampling -- raise, as
 * DESIGN.md section 8 lists.  --dump-config prints the parsed configuration (the string the Lua runtime hands to C) and exits
 * without touching a GPU.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <unistd.h>
#include <signal.h>
#include <errno.h>
#include <sys/wait.h>
#include <fastpm/libfastpm.h>          /* shim/fastpm/libfastpm.h: fastpm_b200_api.h + the option enums of out-of-scope features */
#include "lua-config.h"
#include "param.h"

typedef struct {
    CLIParameters *cli;
    LUAParameters *lua;
    int iout;
} RunData;

static int cmp_double(const void *a, const void *b)
{
    const double x = *(const double *) a, y = *(const double *) b;
    return (x > y) - (x < y);
}

static void refuse(RunData *prr, int nranks)
{
#define OUT_OF_SCOPE(cond, what) do { if (cond) fastpm_raise(-1, "fastpm_b200_run: %s is outside the force-step path of this build (DESIGN.md section 8)\n", what); } while (0)
    OUT_OF_SCOPE(CONF(prr->lua, lc_write_usmesh) != NULL, "lc_write_usmesh (light cone)");
    OUT_OF_SCOPE(CONF(prr->lua, write_fof) != NULL, "write_fof");
    OUT_OF_SCOPE(CONF(prr->lua, write_rfof) != NULL, "write_rfof");
    OUT_OF_SCOPE(CONF(prr->lua, n_m_ncdm) != 0, "m_ncdm (massive neutrinos)");
    OUT_OF_SCOPE(CONF(prr->lua, f_nl_type) != FASTPM_FNL_NONE, "f_nl_type");
    OUT_OF_SCOPE(CONF(prr->lua, constraints) != NULL, "constraints");
    OUT_OF_SCOPE(CONF(prr->lua, read_grafic) != NULL, "read_grafic");
    OUT_OF_SCOPE(CONF(prr->lua, read_runpbic) != NULL, "read_runpbic");
    OUT_OF_SCOPE(CONF(prr->lua, write_runpb_snapshot) != NULL, "write_runpb_snapshot");
    OUT_OF_SCOPE(CONF(prr->lua, read_linear_growth_rate) != NULL, "read_linear_growth_rate");
#undef OUT_OF_SCOPE
    /* the distributed sort by id places every row at the file position its id names, which needs the dense ids of a full store: a
     * sub-sampled snapshot written by several ranks comes out sorted inside each rank's part only */
    if (nranks > 1 && CONF(prr->lua, particle_fraction) < 1 && CONF(prr->lua, sort_snapshot) && CONF(prr->lua, write_snapshot))
        fastpm_raise(-1, "fastpm_b200_run: particle_fraction < 1 on several GPUs needs sort_snapshot = false in this build\n");
}

/* ------------------------------------------------------------------ initial conditions: src/fastpm.c:79,399-586,1779-1812 */
static void read_powerspectrum(FastPMPowerSpectrum *ps, const char *filename, double sigma8)
{
    fastpm_info("Powerspecectrum file: %s\n", filename);
    char *content = fastpm_file_get_content(filename);
    if (content == NULL) fastpm_raise(-1, "Failed to read powerspectrum from file %s\n", filename);
    if (0 != fastpm_powerspectrum_init_from_string(ps, content)) fastpm_raise(-1, "Failed to parse the powerspectrum\n");
    free(content);
    fastpm_info("Found %d pairs of values in input spectrum table\n", (int) ps->base.size);
    const double sigma8_input = fastpm_powerspectrum_sigma(ps, 8);
    fastpm_info("Input power spectrum sigma8 %f\n", sigma8_input);
    if (sigma8 > 0) {
        fastpm_info("Expected power spectrum sigma8 %g; correction applied. \n", sigma8);
        fastpm_powerspectrum_scale(ps, pow(sigma8 / sigma8_input, 2));
    }
}

static void rescale_deltak(FastPMSolver *fastpm, PM *pm, FastPMFloat *delta_k, double aout, double linear_density_redshift)
{
    FastPMGrowthInfo gi_out, gi_in;
    fastpm_growth_info_init(&gi_out, aout, fastpm->cosmology);
    fastpm_growth_info_init(&gi_in, 1. / (linear_density_redshift + 1), fastpm->cosmology);
    const double linear_evolve = gi_out.D1 / gi_in.D1;
    fastpm_info("Reference linear density is calibrated at redshift %g; multiply by %g to extract to redshift %g.\n",
                linear_density_redshift, linear_evolve, 1. / aout - 1);
    fastpm_apply_multiply_transfer(pm, delta_k, delta_k, linear_evolve);
}

static void prepare_deltak(FastPMSolver *fastpm, PM *pm, FastPMFloat *delta_k, RunData *prr, double aout)
{
    const double zlin = CONF(prr->lua, linear_density_redshift);
    const char *lineark = CONF(prr->lua, read_lineark), *pkfile = CONF(prr->lua, read_powerspectrum);
    if (lineark) {
        fastpm_info("Reading Fourier space linear overdensity from %s\n", lineark);
        read_complex(pm, delta_k, lineark, "LinearDensityK", prr->cli->Nwriters);
        if (CONF(prr->lua, inverted_ic)) fastpm_apply_multiply_transfer(pm, delta_k, delta_k, -1);
        rescale_deltak(fastpm, pm, delta_k, aout, zlin);
        return;
    }
    if (!pkfile) fastpm_raise(-1, "Need a power spectrum to start the simulation.\n");
    FastPMPowerSpectrum linear_powerspectrum;
    read_powerspectrum(&linear_powerspectrum, pkfile, CONF(prr->lua, sigma8));

    if (CONF(prr->lua, read_whitenoisek)) {
        fastpm_info("Reading Fourier white noise file from '%s'.\n", CONF(prr->lua, read_whitenoisek));
        read_complex(pm, delta_k, CONF(prr->lua, read_whitenoisek), "WhiteNoiseK", prr->cli->Nwriters);
    } else {
        fastpm_ic_fill_gaussiank(pm, delta_k, CONF(prr->lua, random_seed), FASTPM_DELTAK_GADGET);
    }
    if (CONF(prr->lua, remove_cosmic_variance)) {
        fastpm_info("Remove Cosmic variance from initial condition.\n");
        fastpm_ic_remove_variance(pm, delta_k);
    }
    if (CONF(prr->lua, set_mode)) {
        int method = 0;
        if (0 == strcmp(CONF(prr->lua, set_mode_method), "add")) { method = 1; fastpm_info("SetMode is add\n"); }
        else fastpm_info("SetMode is override\n");
        double *c = CONF(prr->lua, set_mode);
        for (int i = 0; i < CONF(prr->lua, n_set_mode); i++) {
            ptrdiff_t mode[4] = { (ptrdiff_t) c[i * 5 + 0], (ptrdiff_t) c[i * 5 + 1], (ptrdiff_t) c[i * 5 + 2], (ptrdiff_t) c[i * 5 + 3] };
            const double value = c[i * 5 + 4];
            fastpm_apply_set_mode_transfer(pm, delta_k, delta_k, mode, value, method);
            const double result = fastpm_apply_get_mode_transfer(pm, delta_k, mode);
            fastpm_info("SetMode %d : %td %td %td %td value = %g, to = %g\n", i, mode[0], mode[1], mode[2], mode[3], value, result);
        }
    }
    if (CONF(prr->lua, inverted_ic)) fastpm_apply_multiply_transfer(pm, delta_k, delta_k, -1);

    const double variance = pm_compute_variance(pm, delta_k);
    fastpm_info("Variance of input white noise is %0.8f, expectation is %0.8f\n", variance, 1.0 - 1.0 / pm_norm(pm));
    if (CONF(prr->lua, write_whitenoisek)) {
        fastpm_info("Writing Fourier white noise to file '%s'.\n", CONF(prr->lua, write_whitenoisek));
        write_complex(pm, delta_k, CONF(prr->lua, write_whitenoisek), "WhiteNoiseK", prr->cli->Nwriters);
    }
    fastpm_info("Inducing correlation to the white noise.\n");
    fastpm_ic_induce_correlation(pm, delta_k, (fastpm_fkfunc) fastpm_powerspectrum_eval2, &linear_powerspectrum);
    rescale_deltak(fastpm, pm, delta_k, aout, zlin);
    ptrdiff_t mode[4] = { 0, 0, 0, 0 };
    fastpm_apply_modify_mode_transfer(pm, delta_k, delta_k, mode, 1.0);
    fastpm_powerspectrum_destroy(&linear_powerspectrum);
}

static double *prepare_time_step(RunData *prr, double a0, size_t *n_time_step)
{
    const int n = CONF(prr->lua, n_time_step);
    double *time_step = malloc((n + 1) * sizeof(double)), *all = CONF(prr->lua, time_step);
    int i;
    for (i = -1; i < n - 1; i++)
        if (all[i + 1] > a0 + 1e-7) break;                 /* some slack to cover an inexact equality */
    time_step[0] = a0;
    int j;
    for (j = 1; j + i < n; j++) time_step[j] = all[j + i];
    *n_time_step = j;
    return time_step;
}

static void prepare_cdm(FastPMSolver *fastpm, RunData *prr, double a0)
{
    MPI_Comm comm = fastpm->comm;
    if (prr->cli->RestartSnapshotPath) {
        if (CONF(prr->lua, particle_fraction) != 1) fastpm_raise(-1, "Cannot restart because subsampling of particles is enabled.\n");
        fastpm_info("Restarting from snapshot at `%s`.\n", prr->cli->RestartSnapshotPath);
        FastPMStore *p = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM);
        FastPMStore po[1];
        fastpm_set_species_snapshot(fastpm, p, NULL, NULL, po, 1.0);
        fastpm_store_read(po, prr->cli->RestartSnapshotPath, prr->cli->Nwriters, comm);
        if (po->meta.a_x != po->meta.a_v)
            fastpm_raise(-1, "Snapshot velocity and position are out of sync. a_x =% g, a_v = %g.\n", p->meta.a_x, p->meta.a_v);
        if (po->meta.a_x != a0)
            fastpm_raise(-1, "Snapshot velocity and position are out of sync. a_x =% g, first step = %g.\n", p->meta.a_x, a0);
        fastpm_unset_species_snapshot(fastpm, p, NULL, NULL, po, po->meta.a_x);
        return;
    }
    FastPMFloat *delta_k = pm_alloc(fastpm->lptpm);
    prepare_deltak(fastpm, fastpm->lptpm, delta_k, prr, 1.0);
    fastpm_info("No cdm linear growth rate file input.\n");
    if (CONF(prr->lua, write_lineark)) {
        fastpm_info("Writing fourier space linear field to %s\n", CONF(prr->lua, write_lineark));
        write_complex(fastpm->lptpm, delta_k, CONF(prr->lua, write_lineark), "LinearDensityK", prr->cli->Nwriters);
    }
    if (CONF(prr->lua, write_linearr)) {
        fastpm_info("Writing real space linear field to %s\n", CONF(prr->lua, write_linearr));
        pm_c2r(fastpm->lptpm, delta_k);
        write_complex(fastpm->lptpm, delta_k, CONF(prr->lua, write_linearr), "LinearDensityR", prr->cli->Nwriters);
        pm_r2c(fastpm->lptpm, delta_k, delta_k);
    }
    if (CONF(prr->lua, write_powerspectrum)) {
        FastPMPowerSpectrum ps;
        fastpm_powerspectrum_init_from_delta(&ps, fastpm->lptpm, delta_k, delta_k);
        char *buf = fastpm_strdup_printf("%s_linear.txt", CONF(prr->lua, write_powerspectrum));
        fastpm_info("writing linear power spectrum to %s\n", buf);
        if (fastpm->ThisTask == 0) {
            fastpm_path_ensure_dirname(CONF(prr->lua, write_powerspectrum));
            fastpm_powerspectrum_write(&ps, buf, pow(fastpm->config->nc, 3.0));
        }
        free(buf);
        fastpm_powerspectrum_destroy(&ps);
    }
    fastpm_solver_setup_lpt(fastpm, FASTPM_SPECIES_CDM, delta_k, NULL, CONF(prr->lua, time_step)[0]);
    pm_free(fastpm->lptpm, delta_k);
}

/* ------------------------------------------------------------------ handlers */
static int take_a_snapshot(FastPMSolver *fastpm, RunData *prr)
{
    FastPMStore *cdm = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM);
    const double aout = cdm->meta.a_x, z_out = 1.0 / aout - 1.0;
    if (CONF(prr->lua, write_nonlineark)) {
        /* before write_snapshot: that one disturbs the domain decomposition (src/fastpm.c:1421) */
        char *filename = fastpm_strdup_printf("%s_%0.04f", CONF(prr->lua, write_nonlineark), aout);
        FastPMPainter painter[1];
        FastPMFloat *rho_x = pm_alloc(fastpm->basepm), *rho_k = pm_alloc(fastpm->basepm);
        fastpm_painter_init(painter, fastpm->basepm, fastpm->config->PAINTER_TYPE, fastpm->config->painter_support);
        FastPMFieldDescr none = { 0, 0 };
        fastpm_paint(painter, rho_x, cdm, none);
        pm_r2c(fastpm->basepm, rho_x, rho_k);
        write_complex(fastpm->basepm, rho_k, filename, "DensityK", prr->cli->Nwriters);
        pm_free(fastpm->basepm, rho_k);
        pm_free(fastpm->basepm, rho_x);
        free(filename);
    }
    /* particle_fraction < 1 (src/fastpm.c:1449-1461): the snapshot holds the particles whose `rand` deviate is below the fraction,
     * compacted into a store of their own (every column but ACC) */
    double particle_fraction = CONF(prr->lua, particle_fraction);
    FastPMStore subsample[1], *all = cdm;
    if (particle_fraction < 1) {
        FastPMParticleMaskType *mask = fastpm_memory_alloc(cdm->mem, "SubsampleMask", sizeof(mask[0]) * cdm->np_upper, FASTPM_MEMORY_FLOATING);
        fastpm_store_fill_subsample_mask(cdm, particle_fraction, mask);
        size_t room = fastpm_store_subsample(cdm, mask, NULL);
        if (fastpm->NTask > 1) {
            /* several GPUs: device blocks come from the symmetric arena, where every rank must ask for the same sizes in the same
             * order -- a bound that is the same everywhere instead of this rank's own count */
            size_t bound = (size_t) (cdm->np_upper * particle_fraction * 1.5) + 4096;
            if (bound > cdm->np_upper) bound = cdm->np_upper;
            if (room > bound) fastpm_raise(-1, "fastpm_b200_run: %td particles of this rank pass particle_fraction = %g, room for %td\n", (ptrdiff_t) room, particle_fraction, (ptrdiff_t) bound);
            room = bound;
        }
        fastpm_store_init(subsample, cdm->name, room, cdm->attributes & (~COLUMN_ACC) & (~COLUMN_MASK), FASTPM_MEMORY_FLOATING);
        fastpm_store_subsample(cdm, mask, subsample);
        fastpm_memory_free(cdm->mem, mask);
        cdm = subsample;
    }
    if (CONF(prr->lua, sort_snapshot)) fastpm_sort_snapshot(cdm, fastpm->comm, FastPMSnapshotSortByID, 0);
    if (CONF(prr->lua, write_snapshot)) {
        char *filebase = fastpm_strdup_printf("%s_%0.04f", CONF(prr->lua, write_snapshot), aout);
        write_snapshot_header(fastpm, filebase, fastpm->comm);
        /* write_parameters (src/fastpm.c:1210-1225): the parameter file as the reference stores it, attribute "ParamFile" of Header */
        write_snapshot_attr(filebase, "Header", "ParamFile", prr->lua->string, "S1", strlen(prr->lua->string) + 1, fastpm->comm);
        write_snapshot_attr(filebase, "Header", "ParticleFraction", &particle_fraction, "f8", 1, fastpm->comm);
        fastpm_store_write(cdm, filebase, "w", prr->cli->Nwriters, fastpm->comm);
        fastpm_info("snapshot %s [%s] written at z = %6.4f a = %6.4f \n", filebase, "1", z_out, aout);
        free(filebase);
    }
    if (cdm != all) fastpm_store_destroy(subsample);
    return 0;
}

static int check_snapshots(FastPMSolver *fastpm, FastPMInterpolationEvent *event, RunData *prr)
{
    fastpm_info("Checking Snapshots (%0.4f %0.4f) with K(%0.4f->%0.4f|%0.4f) D(%0.4f->%0.4f|%0.4f)\n", event->a1, event->a2,
                event->kick->ai, event->kick->af, event->kick->ac, event->drift->ai, event->drift->af, event->drift->ac);
    const int nout = CONF(prr->lua, n_aout);
    double *aout = malloc(sizeof(double) * (nout + 1));
    memcpy(aout, CONF(prr->lua, aout), sizeof(double) * nout);
    qsort(aout, nout, sizeof(double), cmp_double);
    for (int iout = prr->iout; iout < nout; iout++) {
        if (event->a1 == event->a2) {
            if (event->a1 != aout[iout]) continue;                  /* initial condition: only when asked for */
            if (prr->cli->RestartSnapshotPath) continue;            /* restarting from this snapshot: no need to write it again */
        } else {
            if (event->a1 >= aout[iout]) continue;
            if (event->a2 < aout[iout]) continue;
        }
        FastPMSolver snapshot[1];
        FastPMStore cdm[1];
        memcpy(snapshot, fastpm, sizeof(FastPMSolver));
        fastpm_solver_add_species(snapshot, FASTPM_SPECIES_CDM, cdm);
        fastpm_set_snapshot(fastpm, snapshot, event->drift, event->kick, aout[iout]);
        FastPMGrowthInfo gi;
        fastpm_growth_info_init(&gi, aout[iout], fastpm->cosmology);
        fastpm_info("Snapshot a_x = %6.4f, a_v = %6.4f \n", cdm->meta.a_x, cdm->meta.a_v);
        fastpm_info("Growth factor of snapshot %6.4f (a=%0.4f)\n", gi.D1, aout[iout]);
        fastpm_info("Growth rate of snapshot %6.4f (a=%0.4f)\n", gi.f1, aout[iout]);
        take_a_snapshot(snapshot, prr);
        fastpm_unset_snapshot(fastpm, snapshot, event->drift, event->kick, aout[iout]);
        prr->iout = iout + 1;
    }
    free(aout);
    return 0;
}

static int print_transition(FastPMSolver *fastpm, FastPMTransitionEvent *event, RunData *prr)
{
    (void) fastpm; (void) prr;
    FastPMTransition *trans = event->transition;
    const char *action = trans->action == FASTPM_ACTION_FORCE ? "FORCE" : (trans->action == FASTPM_ACTION_KICK ? "KICK" :
                         (trans->action == FASTPM_ACTION_DRIFT ? "DRIFT" : "Unknown"));
    fastpm_info("==== -> %03d [%03d %03d %03d] a_i = %6.4f a_f = %6.4f a_r = %6.4f Action = %s(%d) ====\n",
                trans->iend, trans->end->x, trans->end->v, trans->end->force, trans->a.i, trans->a.f, trans->a.r, action, trans->action);
    return 0;
}

static int report_lpt(FastPMSolver *fastpm, FastPMLPTEvent *event, RunData *prr)
{
    (void) prr;
    double dx1_std[3], dx2_std[3];
    fastpm_store_summary(event->p, COLUMN_DX1, fastpm->comm, "s", dx1_std);
    fastpm_store_summary(event->p, COLUMN_DX2, fastpm->comm, "s", dx2_std);
    fastpm_info("dx1  : %g %g %g %g\n", dx1_std[0], dx1_std[1], dx1_std[2], (dx1_std[0] + dx1_std[1] + dx1_std[2]) / 3.0);
    fastpm_info("dx2  : %g %g %g %g\n", dx2_std[0], dx2_std[1], dx2_std[2], (dx2_std[0] + dx2_std[1] + dx2_std[2]) / 3.0);
    return 0;
}

static int report_domain(FastPMSolver *fastpm, FastPMForceEvent *event, RunData *prr)
{
    (void) prr;
    fastpm_info("Force Calculation Nmesh = %d ====\n", (int) pm_nmesh(event->pm)[0]);
    FastPMStore *p = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM);
    double min[3], max[3], std[3];
    fastpm_store_summary(p, COLUMN_POS, fastpm->comm, "<>", min, max);
    fastpm_store_summary(p, COLUMN_VEL, fastpm->comm, "s", std);
    fastpm_info("Position range (a = %06.4f): min = %g %g %g max = %g %g %g \n", p->meta.a_x, min[0], min[1], min[2], max[0], max[1], max[2]);
    fastpm_info("Velocity dispersion (a = %06.4f): std = %g %g %g\n", p->meta.a_v, std[0], std[1], std[2]);
    return 0;
}

static int write_powerspectrum(FastPMSolver *fastpm, FastPMForceEvent *event, RunData *prr)
{
    const int K_LINEAR = CONF(prr->lua, enforce_broadband_kmax);
    FastPMStore *p = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM);
    double fstd[3];
    fastpm_store_summary(p, COLUMN_ACC, fastpm->comm, "s", fstd);
    fastpm_info("Force dispersion: std = %g %g %g\n", fstd[0], fstd[1], fstd[2]);

    FastPMPowerSpectrum ps;
    fastpm_powerspectrum_init_from_delta(&ps, event->pm, event->delta_k, event->delta_k);
    double Plin = fastpm_powerspectrum_large_scale(&ps, K_LINEAR);
    double Sigma8 = fastpm_powerspectrum_sigma(&ps, 8);
    FastPMGrowthInfo gi;
    fastpm_growth_info_init(&gi, event->a_f, fastpm->cosmology);
    Plin /= pow(gi.D1, 2.0);
    Sigma8 /= pow(gi.D1, 2.0);
    fastpm_info("D^2(%g, 1.0) P(k<%g) = %g Sigma8 = %g\n", event->a_f, K_LINEAR * 6.28 / pm_boxsize(event->pm)[0], Plin, Sigma8);
    if (CONF(prr->lua, write_powerspectrum)) {
        char *buf = fastpm_strdup_printf("%s_%0.04f.txt", CONF(prr->lua, write_powerspectrum), event->a_f);
        fastpm_info("writing power spectrum to %s\n", buf);
        if (fastpm->ThisTask == 0) {
            fastpm_path_ensure_dirname(CONF(prr->lua, write_powerspectrum));
            fastpm_powerspectrum_write(&ps, buf, event->N);
        }
        free(buf);
    }
    fastpm_powerspectrum_destroy(&ps);
    return 0;
}

/* ------------------------------------------------------------------ run_fastpm, src/fastpm.c:265-397 */
static int run_fastpm(FastPMConfig *config, RunData *prr, MPI_Comm comm)
{
    FastPMSolver fastpm[1];
    CLOCK(init);
    CLOCK(cdmic);
    LEAVE(cdmic);
    CLOCK(evolve);
    LEAVE(evolve);
    fastpm_solver_init(fastpm, config, comm);
    fastpm_info("BaseProcMesh : %d x %d\n", pm_nproc(fastpm->basepm)[0], pm_nproc(fastpm->basepm)[1]);
    LEAVE(init);

    fastpm_add_event_handler(&fastpm->event_handlers, FASTPM_EVENT_LPT, FASTPM_EVENT_STAGE_AFTER, (FastPMEventHandlerFunction) report_lpt, prr);
    fastpm_add_event_handler(&fastpm->event_handlers, FASTPM_EVENT_FORCE, FASTPM_EVENT_STAGE_BEFORE, (FastPMEventHandlerFunction) report_domain, prr);
    fastpm_add_event_handler(&fastpm->event_handlers, FASTPM_EVENT_FORCE, FASTPM_EVENT_STAGE_AFTER, (FastPMEventHandlerFunction) write_powerspectrum, prr);
    fastpm_add_event_handler(&fastpm->event_handlers, FASTPM_EVENT_INTERPOLATION, FASTPM_EVENT_STAGE_BEFORE, (FastPMEventHandlerFunction) check_snapshots, prr);
    fastpm_add_event_handler(&fastpm->event_handlers, FASTPM_EVENT_TRANSITION, FASTPM_EVENT_STAGE_BEFORE, (FastPMEventHandlerFunction) print_transition, prr);
    /* print_transition only prints: no need to apply the queued kicks and drifts for it (keeps the fused particle update) */
    fastpm_b200_mark_handler_passive((FastPMEventHandlerFunction) print_transition);

    double a_restart = 0.0;
    if (prr->cli->RestartSnapshotPath) {
        read_snapshot_header(fastpm, prr->cli->RestartSnapshotPath, &a_restart, comm);
        fastpm_info("Restarting from %s at a = %06.4f", prr->cli->RestartSnapshotPath, a_restart);
    } else {
        a_restart = CONF(prr->lua, time_step)[0];
    }
    size_t n_time_step;
    double *time_step = prepare_time_step(prr, a_restart, &n_time_step);

    ENTER(cdmic);
    prepare_cdm(fastpm, prr, time_step[0]);
    LEAVE(cdmic);

    ENTER(evolve);
    fastpm_solver_evolve(fastpm, time_step, (int) n_time_step);
    LEAVE(evolve);
    free(time_step);
    fastpm_solver_destroy(fastpm);
    fastpm_clock_stat(comm);
    return 0;
}

/* -n N: the parent passes an interrupt on to the ranks it forked (exact pids), the wait loop below then cleans up */
static pid_t rank_pids[8];
static int n_rank_pids = 0;
static void forward_signal(int sig)
{
    for (int i = 0; i < n_rank_pids; i++) if (rank_pids[i] > 0) kill(rank_pids[i], sig);
}

int main(int argc, char **argv)
{
    /* --dump-config: handled before anything touches a device */
    int dump = 0, nranks = 1;
    char **av = malloc(sizeof(char *) * (argc + 1));
    int ac = 0;
    for (int i = 0; i < argc; i++) {
        if (i > 0 && !strcmp(argv[i], "--dump-config")) { dump = 1; continue; }
        if (i > 0 && !strcmp(argv[i], "-n") && i + 1 < argc && ac == 1) { nranks = atoi(argv[++i]); continue; }      /* first option only */
        av[ac++] = argv[i];
    }
    if (nranks != 1 && nranks != 2 && nranks != 4 && nranks != 8) { fprintf(stderr, "-n %d: 1, 2, 4 or 8 GPUs of one node\n", nranks); return 1; }
    av[ac] = NULL;
    RunData prr[1];
    memset(prr, 0, sizeof(prr));
    prr->cli = parse_cli_args(ac, av);
    if (!prr->cli) return 1;
    char *error = NULL;
    prr->lua = parse_config(prr->cli->argv[0], prr->cli->argc, prr->cli->argv, &error);
    if (!prr->lua) {
        fprintf(stderr, "Parsing configuration failed with error: %s\n", error ? error : "(none)");
        return 1;
    }
    if (dump) {
        printf("%s\n", prr->lua->string);
        return 0;
    }

    refuse(prr, nranks);                  /* before any device is touched */
    if (nranks > 1) {
        /* one process per GPU, forked before the first CUDA call; the parent only waits (and takes the others down if one fails) */
        char segment[64];
        if (fastpm_b200_local_segment_create(segment, sizeof(segment)) != 0) { fprintf(stderr, "cannot create the shared segment of the ranks\n"); return 1; }
        pid_t pids[8];
        int rank = -1;
        fflush(NULL);
        for (int r = 0; r < nranks; r++) {
            pids[r] = fork();
            if (pids[r] < 0) { perror("fork"); for (int q = 0; q < r; q++) kill(pids[q], SIGTERM); fastpm_b200_local_segment_unlink(segment); return 1; }
            if (pids[r] == 0) { rank = r; break; }
        }
        if (rank < 0) {
            for (int r = 0; r < nranks; r++) rank_pids[r] = pids[r];
            n_rank_pids = nranks;
            signal(SIGINT, forward_signal);
            signal(SIGTERM, forward_signal);
            int failed = 0;
            for (int left = nranks; left > 0; left--) {
                int status = 0;
                pid_t done;
                while ((done = wait(&status)) < 0 && errno == EINTR) { }       /* a forwarded signal interrupted the wait */
                if (done < 0) break;
                if (!(WIFEXITED(status) && WEXITSTATUS(status) == 0) && !failed) {
                    failed = WIFEXITED(status) ? WEXITSTATUS(status) : 128 + WTERMSIG(status);
                    for (int q = 0; q < nranks; q++) if (pids[q] != done) kill(pids[q], SIGTERM);      /* the others would wait for it forever */
                }
            }
            fastpm_b200_local_segment_unlink(segment);
            return failed;
        }
        char dev[16];
        snprintf(dev, sizeof(dev), "%d", rank);
        setenv("FASTPM_B200_DEVICE", dev, 1);                                              /* rank r drives GPU r */
        fastpm_b200_comm_init_local(rank, nranks, segment);
    }
    libfastpm_init();
    MPI_Comm comm = MPI_COMM_WORLD;
    fastpm_set_msg_handler(fastpm_default_msg_handler, comm, NULL);      /* the command line logs like the reference's (src/fastpm.c:136) */
    libfastpm_set_memory_bound(prr->cli->MemoryPerRank * 1024 * 1024);

    /* pm_nc_factor -> VPMInit, src/fastpm.c:158-181 */
    VPMInit *vpminit = NULL;
    if (CONF(prr->lua, ndim_pm_nc_factor) == 0) {
        vpminit = malloc(sizeof(VPMInit) * 2);
        vpminit[0].a_start = 0; vpminit[0].pm_nc_factor = CONF(prr->lua, pm_nc_factor)[0];
        vpminit[1].a_start = 1; vpminit[1].pm_nc_factor = 0;
    } else if (CONF(prr->lua, ndim_pm_nc_factor) == 2) {
        const int n = CONF(prr->lua, shape_pm_nc_factor)[0];
        vpminit = malloc(sizeof(VPMInit) * (n + 1));
        int i;
        for (i = 0; i < n; i++) {
            vpminit[i].a_start = CONF(prr->lua, pm_nc_factor)[2 * i];
            vpminit[i].pm_nc_factor = CONF(prr->lua, pm_nc_factor)[2 * i + 1];
        }
        vpminit[i].a_start = 1; vpminit[i].pm_nc_factor = 0;
    } else {
        fastpm_raise(-1, "Unknown format of pm_nc_factor, either a scalar or a 2d array. ");
    }
    fastpm_info("np_alloc_factor = %g\n", CONF(prr->lua, np_alloc_factor));

    /* prepare_cosmology, src/prepare.c:19-41 */
    FastPMCosmology cosmology[1];
    memset(cosmology, 0, sizeof(cosmology));
    cosmology->h = CONF(prr->lua, h);
    cosmology->Omega_m = CONF(prr->lua, Omega_m);
    cosmology->T_cmb = CONF(prr->lua, T_cmb);
    cosmology->Omega_k = CONF(prr->lua, Omega_k);
    cosmology->w0 = CONF(prr->lua, w0);
    cosmology->wa = CONF(prr->lua, wa);
    cosmology->N_eff = CONF(prr->lua, N_eff);
    cosmology->N_nu = CONF(prr->lua, N_nu);
    cosmology->N_ncdm = CONF(prr->lua, n_m_ncdm);
    cosmology->ncdm_matterlike = CONF(prr->lua, ncdm_matterlike);
    cosmology->ncdm_freestreaming = CONF(prr->lua, ncdm_freestreaming);
    cosmology->ncdm_linearresponse = CONF(prr->lua, ncdm_linearresponse);
    cosmology->growth_mode = CONF(prr->lua, growth_mode);

    FastPMConfig config[1];
    memset(config, 0, sizeof(config));
    config->nc = CONF(prr->lua, nc);
    config->alloc_factor = CONF(prr->lua, np_alloc_factor);
    config->lpt_nc_factor = CONF(prr->lua, lpt_nc_factor);
    config->vpminit = vpminit;
    config->boxsize = CONF(prr->lua, boxsize);
    config->cosmology = cosmology;
    config->USE_DX1_ONLY = CONF(prr->lua, za);
    config->nLPT = -2.5f;
    config->USE_SHIFT = CONF(prr->lua, shift);
    config->FORCE_TYPE = CONF(prr->lua, force_mode);
    config->KERNEL_TYPE = CONF(prr->lua, kernel_type);
    config->SOFTENING_TYPE = CONF(prr->lua, force_softening_type);
    config->PAINTER_TYPE = CONF(prr->lua, painter_type);
    config->painter_support = CONF(prr->lua, painter_support);
    config->NprocY = prr->cli->NprocY;
    config->UseFFTW = prr->cli->UseFFTW;
    config->ExtraAttributes = 0;
    config->pgdc = CONF(prr->lua, pgdc);
    config->pgdc_alpha0 = CONF(prr->lua, pgdc_alpha0);
    config->pgdc_A = CONF(prr->lua, pgdc_A);
    config->pgdc_B = CONF(prr->lua, pgdc_B);
    config->pgdc_kl = CONF(prr->lua, pgdc_kl);
    config->pgdc_ks = CONF(prr->lua, pgdc_ks);
    if (CONF(prr->lua, compute_potential)) config->ExtraAttributes |= COLUMN_POTENTIAL;
    if (CONF(prr->lua, pgdc)) config->ExtraAttributes |= COLUMN_PGDC;
    /* the reference's CDM store always carries MASK and RAND (solver.c:93-97); here RAND is there when sub-sampling needs it */
    if (CONF(prr->lua, particle_fraction) < 1) config->ExtraAttributes |= COLUMN_RAND;

    run_fastpm(config, prr, comm);

    free(vpminit);
    free_lua_parameters(prr->lua);
    free_cli_parameters(prr->cli);
    if (nranks > 1) fastpm_b200_comm_finalize();
    libfastpm_cleanup();
    return 0;
}
