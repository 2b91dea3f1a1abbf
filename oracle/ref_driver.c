/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * Array-level C entry points over the UNMODIFIED reference library
 * (/root/reference/libfastpm/*.c, compiled in place by oracle/Makefile into
 * oracle/_ref/libfastpm_ref.so against the 1-rank MPI / mini-GSL / PFFT shims
 * under oracle/shims).  This file contains no arithmetic of its own: every
 * ref_* function builds the reference's structs and calls the reference's
 * functions, so that tests and the CPU baseline can drive them from numpy
 * arrays through ctypes (oracle/ref.py).
 *
 * Mirrors the call sequence of the reference CLI: src/fastpm.c:186-210
 * (config), :415-587 (prepare_deltak), :616-713 (prepare_cdm), :369 (evolve),
 * :1671-1708 (report_domain), :1711-1776 (write_powerspectrum handler).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <alloca.h>
#include <mpi.h>
#include <fastpm/libfastpm.h>
#include <fastpm/logging.h>
#include <fastpm/prof.h>
#include "pmpfft.h"
#include <fastpm/io.h>

typedef struct {
    int64_t nc;
    double boxsize;
    double pm_nc_factor[8];      /* pairs {a_start, B}; terminated by B = 0 */
    double np_alloc_factor;
    double lpt_nc_factor;
    int force_mode;              /* FastPMForceType */
    int kernel_type;             /* FastPMKernelType */
    int growth_mode;             /* FastPMGrowthMode */
    int compute_potential;
    int use_dx1_only;
    int verbose;
    double nLPT;
    double Omega_m, h, T_cmb, Omega_k, w0, wa, N_eff;
    int N_nu;
    int enforce_broadband_kmax;
    double pgdc[6];              /* {enabled, alpha0, A, B, kl, ks}: the PGD correction of src/fastpm.c:204-217 */
    int softening_type;          /* FastPMSofteningType (gravity.c:244-270); 0 = none */
    int painter_type;            /* FastPMPainterType of the force step (config->PAINTER_TYPE); 0 = CIC */
    int painter_support;
    int use_shift;               /* config->USE_SHIFT: ICs at cell centres (solver.c:142-150,201-209) */
} RefConfig;

#define MAX_FORCE_RECORDS 256
typedef struct {
    double a_f;
    double vel_std[3], pos_min[3], pos_max[3], acc_std[3];
    double Plin, a_x, a_v;
    int nbins;
    double *k, *p, *nmodes;
} ForceRecord;

typedef struct {
    FastPMSolver solver[1];
    FastPMConfig config[1];
    FastPMCosmology cosmology[1];
    VPMInit vpminit[5];
    RefConfig cfg;
    double dx1_std[3], dx2_std[3];
    int nrecords;
    ForceRecord rec[MAX_FORCE_RECORDS];
    double t_evolve;
} RefSession;

static int lib_inited = 0;

static int on_lpt(FastPMSolver *fastpm, FastPMLPTEvent *event, RefSession *s)
{
    fastpm_store_summary(event->p, COLUMN_DX1, fastpm->comm, "s", s->dx1_std);
    fastpm_store_summary(event->p, COLUMN_DX2, fastpm->comm, "s", s->dx2_std);
    return 0;
}

static int on_force_before(FastPMSolver *fastpm, FastPMForceEvent *event, RefSession *s)
{
    if (s->nrecords >= MAX_FORCE_RECORDS) return 0;
    ForceRecord *r = &s->rec[s->nrecords];
    FastPMStore *p = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM);
    memset(r, 0, sizeof(*r));
    r->a_f = event->a_f;
    r->a_x = p->meta.a_x; r->a_v = p->meta.a_v;
    fastpm_store_summary(p, COLUMN_POS, fastpm->comm, "<>", r->pos_min, r->pos_max);
    fastpm_store_summary(p, COLUMN_VEL, fastpm->comm, "s", r->vel_std);
    return 0;
}

static int on_force_after(FastPMSolver *fastpm, FastPMForceEvent *event, RefSession *s)
{
    if (s->nrecords >= MAX_FORCE_RECORDS) return 0;
    ForceRecord *r = &s->rec[s->nrecords];
    FastPMStore *p = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM);
    fastpm_store_summary(p, COLUMN_ACC, fastpm->comm, "s", r->acc_std);

    FastPMPowerSpectrum ps;
    fastpm_powerspectrum_init_from_delta(&ps, event->pm, event->delta_k, event->delta_k);
    double Plin = fastpm_powerspectrum_large_scale(&ps, s->cfg.enforce_broadband_kmax);
    FastPMGrowthInfo gi;
    fastpm_growth_info_init(&gi, event->a_f, fastpm->cosmology);
    r->Plin = Plin / pow(gi.D1, 2.0);
    r->nbins = (int) ps.base.size;
    r->k = malloc(sizeof(double) * r->nbins);
    r->p = malloc(sizeof(double) * r->nbins);
    r->nmodes = malloc(sizeof(double) * r->nbins);
    memcpy(r->k, ps.base.k, sizeof(double) * r->nbins);
    memcpy(r->p, ps.base.f, sizeof(double) * r->nbins);
    memcpy(r->nmodes, ps.Nmodes, sizeof(double) * r->nbins);
    fastpm_powerspectrum_destroy(&ps);
    s->nrecords++;
    return 0;
}

RefSession *ref_session_new(const RefConfig *cfg)
{
    if (!lib_inited) { libfastpm_init(); lib_inited = 1; }
    fastpm_set_msg_handler(cfg->verbose ? fastpm_default_msg_handler : fastpm_void_msg_handler, MPI_COMM_WORLD, NULL);

    RefSession *s = calloc(1, sizeof(*s));
    s->cfg = *cfg;
    int i;
    for (i = 0; i < 4 && cfg->pm_nc_factor[2 * i + 1] > 0; i++) {
        s->vpminit[i].a_start = cfg->pm_nc_factor[2 * i];
        s->vpminit[i].pm_nc_factor = cfg->pm_nc_factor[2 * i + 1];
    }
    s->vpminit[i].a_start = 1; s->vpminit[i].pm_nc_factor = 0;

    /* src/prepare.c:20-39 */
    FastPMCosmology *c = s->cosmology;
    memset(c, 0, sizeof(*c));
    c->h = cfg->h; c->Omega_m = cfg->Omega_m; c->T_cmb = cfg->T_cmb; c->Omega_k = cfg->Omega_k;
    c->w0 = cfg->w0; c->wa = cfg->wa; c->N_eff = cfg->N_eff; c->N_nu = cfg->N_nu; c->N_ncdm = 0;
    c->ncdm_matterlike = 1; c->ncdm_freestreaming = 1; c->ncdm_linearresponse = 0;
    c->growth_mode = cfg->growth_mode;

    /* src/fastpm.c:186-217 */
    FastPMConfig *config = s->config;
    memset(config, 0, sizeof(*config));
    config->nc = cfg->nc;
    config->alloc_factor = cfg->np_alloc_factor;
    config->lpt_nc_factor = cfg->lpt_nc_factor;
    config->vpminit = s->vpminit;
    config->boxsize = cfg->boxsize;
    config->cosmology = c;
    config->USE_DX1_ONLY = cfg->use_dx1_only;
    config->nLPT = cfg->nLPT;
    config->USE_SHIFT = cfg->use_shift;
    config->FORCE_TYPE = cfg->force_mode;
    config->KERNEL_TYPE = cfg->kernel_type;
    config->SOFTENING_TYPE = (FastPMSofteningType) cfg->softening_type;
    config->PAINTER_TYPE = (FastPMPainterType) cfg->painter_type;
    config->painter_support = cfg->painter_type == 0 ? 2 : cfg->painter_support;
    config->NprocY = 0; config->UseFFTW = 0;
    config->ExtraAttributes = 0;
    if (cfg->compute_potential) config->ExtraAttributes |= COLUMN_POTENTIAL;
    if (cfg->pgdc[0] != 0) {
        config->pgdc = 1;
        config->pgdc_alpha0 = cfg->pgdc[1]; config->pgdc_A = cfg->pgdc[2]; config->pgdc_B = cfg->pgdc[3];
        config->pgdc_kl = cfg->pgdc[4]; config->pgdc_ks = cfg->pgdc[5];
        config->ExtraAttributes |= COLUMN_PGDC;
    }

    fastpm_solver_init(s->solver, config, MPI_COMM_WORLD);

    fastpm_add_event_handler(&s->solver->event_handlers, FASTPM_EVENT_LPT, FASTPM_EVENT_STAGE_AFTER,
            (FastPMEventHandlerFunction) on_lpt, s);
    fastpm_add_event_handler(&s->solver->event_handlers, FASTPM_EVENT_FORCE, FASTPM_EVENT_STAGE_BEFORE,
            (FastPMEventHandlerFunction) on_force_before, s);
    fastpm_add_event_handler(&s->solver->event_handlers, FASTPM_EVENT_FORCE, FASTPM_EVENT_STAGE_AFTER,
            (FastPMEventHandlerFunction) on_force_after, s);
    return s;
}

static void clear_records(RefSession *s)
{
    for (int i = 0; i < s->nrecords; i++) { free(s->rec[i].k); free(s->rec[i].p); free(s->rec[i].nmodes); }
    s->nrecords = 0;
}

void ref_session_free(RefSession *s)
{
    clear_records(s);
    fastpm_solver_destroy(s->solver);
    free(s);
}

/* ------------------------------------------------------------------ meshes */
static PM *pick_pm(RefSession *s, int which, double a)
{
    if (which == 0) return fastpm_find_pm(s->solver, a);     /* force mesh in use at time a */
    if (which == 1) return s->solver->lptpm;
    return s->solver->basepm;
}

/* info[0]=Nmesh, [1]=allocsize, [2..4]=IRegion.strides, [5..7]=ORegion.strides (complex units), [8..10] = ORegion.size */
void ref_pm_info(RefSession *s, int which, double a, int64_t *info)
{
    PM *pm = pick_pm(s, which, a);
    info[0] = pm->Nmesh[0]; info[1] = pm->allocsize;
    for (int d = 0; d < 3; d++) {
        info[2 + d] = pm->IRegion.strides[d];
        info[5 + d] = pm->ORegion.strides[d];
        info[8 + d] = pm->ORegion.size[d];
    }
}

/* ---------------------------------------------------------------------- IC */
/* src/fastpm.c:415-587 (prepare_deltak) for the seed + power-spectrum-table path */
double ref_ic_deltak(RefSession *s, int seed, int remove_variance, const char *pk_table_text,
                     double linear_density_redshift, float *delta_k_out, double *sigma8_out)
{
    PM *pm = s->solver->lptpm;
    FastPMFloat *delta_k = pm_alloc(pm);
    FastPMPowerSpectrum lin;
    if (0 != fastpm_powerspectrum_init_from_string(&lin, pk_table_text)) {
        fastpm_raise(-1, "Failed to parse the powerspectrum\n");
    }
    if (sigma8_out) *sigma8_out = fastpm_powerspectrum_sigma(&lin, 8);

    fastpm_ic_fill_gaussiank(pm, delta_k, seed, FASTPM_DELTAK_GADGET);
    if (remove_variance) fastpm_ic_remove_variance(pm, delta_k);
    double variance = pm_compute_variance(pm, delta_k);
    fastpm_ic_induce_correlation(pm, delta_k, (fastpm_fkfunc) fastpm_powerspectrum_eval2, &lin);

    /* rescale_deltak, src/fastpm.c:400-412, aout = 1.0 */
    FastPMGrowthInfo gi_out, gi_in;
    fastpm_growth_info_init(&gi_out, 1.0, s->solver->cosmology);
    fastpm_growth_info_init(&gi_in, 1. / (linear_density_redshift + 1), s->solver->cosmology);
    fastpm_apply_multiply_transfer(pm, delta_k, delta_k, gi_out.D1 / gi_in.D1);

    ptrdiff_t mode[4] = { 0, 0, 0, 0 };
    fastpm_apply_modify_mode_transfer(pm, delta_k, delta_k, mode, 1.0);

    memcpy(delta_k_out, delta_k, sizeof(FastPMFloat) * pm->allocsize);
    fastpm_powerspectrum_destroy(&lin);
    pm_free(pm, delta_k);
    return variance;
}

/* the raw Gaussian field alone: fastpm_ic_fill_gaussiank with the Gadget scheme (initialcondition.c:145-273) on the LPT mesh */
void ref_fill_gaussian(RefSession *s, int seed, float *delta_k_out)
{
    PM *pm = s->solver->lptpm;
    FastPMFloat *delta_k = pm_alloc(pm);
    fastpm_ic_fill_gaussiank(pm, delta_k, seed, FASTPM_DELTAK_GADGET);
    memcpy(delta_k_out, delta_k, sizeof(FastPMFloat) * pm->allocsize);
    pm_free(pm, delta_k);
}

/* fastpm_ic_remove_variance (initialcondition.c:66-99) on the LPT mesh */
void ref_remove_variance(RefSession *s, const float *in, float *out)
{
    PM *pm = s->solver->lptpm;
    FastPMFloat *dk = pm_alloc(pm);
    memcpy(dk, in, sizeof(FastPMFloat) * pm->allocsize);
    fastpm_ic_remove_variance(pm, dk);
    memcpy(out, dk, sizeof(FastPMFloat) * pm->allocsize);
    pm_free(pm, dk);
}

void ref_setup_lpt(RefSession *s, const float *delta_k_in, double a0)
{
    PM *pm = s->solver->lptpm;
    FastPMFloat *delta_k = pm_alloc(pm);
    memcpy(delta_k, delta_k_in, sizeof(FastPMFloat) * pm->allocsize);
    fastpm_solver_setup_lpt(s->solver, FASTPM_SPECIES_CDM, delta_k, NULL, a0);
    pm_free(pm, delta_k);
}

void ref_lpt_std(RefSession *s, double *dx1_std, double *dx2_std)
{
    memcpy(dx1_std, s->dx1_std, sizeof(double) * 3);
    memcpy(dx2_std, s->dx2_std, sizeof(double) * 3);
}

/* pm_2lpt_solve alone (pm2lpt.c:14): returns dx1, dx2 on the unperturbed grid */
void ref_2lpt_solve(RefSession *s, const float *delta_k_in, float *dx1, float *dx2)
{
    extern void pm_2lpt_solve(PM *pm, FastPMFloat *delta_k, FastPMFuncK *growth_rate_func_k, FastPMStore *p, double shift[3], FastPMKernelType type);
    PM *pm = s->solver->lptpm;
    FastPMStore *p = fastpm_solver_get_species(s->solver, FASTPM_SPECIES_CDM);
    FastPMFloat *delta_k = pm_alloc(pm);
    memcpy(delta_k, delta_k_in, sizeof(FastPMFloat) * pm->allocsize);
    int t1 = 0, t2 = 0;
    if (!p->dx1) { p->dx1 = fastpm_memory_alloc(p->mem, "DX1", sizeof(p->dx1[0]) * p->np_upper, FASTPM_MEMORY_STACK); t1 = 1; }
    if (!p->dx2) { p->dx2 = fastpm_memory_alloc(p->mem, "DX2", sizeof(p->dx2[0]) * p->np_upper, FASTPM_MEMORY_STACK); t2 = 1; }
    double shift[3] = { 0, 0, 0 };
    pm_2lpt_solve(pm, delta_k, NULL, p, shift, s->solver->config->KERNEL_TYPE);
    memcpy(dx1, p->dx1, sizeof(p->dx1[0]) * p->np);
    memcpy(dx2, p->dx2, sizeof(p->dx2[0]) * p->np);
    if (t2) { fastpm_memory_free(p->mem, p->dx2); p->dx2 = NULL; }
    if (t1) { fastpm_memory_free(p->mem, p->dx1); p->dx1 = NULL; }
    pm_free(pm, delta_k);
}

/* --------------------------------------------------------------- particles */
int64_t ref_np(RefSession *s) { return (int64_t) fastpm_solver_get_species(s->solver, FASTPM_SPECIES_CDM)->np; }

/* any pointer may be NULL; dx1/dx2 are only present in COLA mode */
void ref_get_particles(RefSession *s, double *x, float *v, float *acc, uint64_t *id, float *dx1, float *dx2, double *meta)
{
    FastPMStore *p = fastpm_solver_get_species(s->solver, FASTPM_SPECIES_CDM);
    if (x) memcpy(x, p->x, sizeof(p->x[0]) * p->np);
    if (v) memcpy(v, p->v, sizeof(p->v[0]) * p->np);
    if (acc) memcpy(acc, p->acc, sizeof(p->acc[0]) * p->np);
    if (id) memcpy(id, p->id, sizeof(p->id[0]) * p->np);
    if (dx1 && p->dx1) memcpy(dx1, p->dx1, sizeof(p->dx1[0]) * p->np);
    if (dx2 && p->dx2) memcpy(dx2, p->dx2, sizeof(p->dx2[0]) * p->np);
    if (meta) { meta[0] = p->meta.a_x; meta[1] = p->meta.a_v; meta[2] = p->meta.M0; }
}

void ref_set_particles(RefSession *s, int64_t np, const double *x, const float *v, const uint64_t *id,
                       const float *dx1, const float *dx2, const double *meta)
{
    FastPMStore *p = fastpm_solver_get_species(s->solver, FASTPM_SPECIES_CDM);
    if ((size_t) np > p->np_upper) fastpm_raise(-1, "ref_set_particles: too many particles\n");
    p->np = np;
    if (x) memcpy(p->x, x, sizeof(p->x[0]) * np);
    if (v) memcpy(p->v, v, sizeof(p->v[0]) * np);
    if (id) memcpy(p->id, id, sizeof(p->id[0]) * np);
    if (dx1 && p->dx1) memcpy(p->dx1, dx1, sizeof(p->dx1[0]) * np);
    if (dx2 && p->dx2) memcpy(p->dx2, dx2, sizeof(p->dx2[0]) * np);
    if (meta) { p->meta.a_x = meta[0]; p->meta.a_v = meta[1]; p->meta.M0 = meta[2]; }
}

/* fastpm_store_wrap (store.c:447-475) on the session's particles, and fastpm_store_summary (store.c:808-909) of one column:
 * out[0..2] = min, out[3..5] = max, out[6..8] = mean, out[9..11] = std ('<', '>', '-', 's') */
void ref_wrap(RefSession *s)
{
    FastPMStore *p = fastpm_solver_get_species(s->solver, FASTPM_SPECIES_CDM);
    fastpm_store_wrap(p, pm_boxsize(s->solver->basepm));
}

void ref_summary(RefSession *s, int column_tag, double *out)
{
    FastPMStore *p = fastpm_solver_get_species(s->solver, FASTPM_SPECIES_CDM);
    fastpm_store_summary(p, (FastPMColumnTags) column_tag, MPI_COMM_WORLD, "<>-s", out, out + 3, out + 6, out + 9);
}

/* ------------------------------------------------------------------ evolve */
double ref_evolve(RefSession *s, const double *time_step, int nstep)
{
    clear_records(s);
    double *ts = malloc(sizeof(double) * nstep);
    memcpy(ts, time_step, sizeof(double) * nstep);
    double t0 = MPI_Wtime();
    fastpm_solver_evolve(s->solver, ts, nstep);
    s->t_evolve = MPI_Wtime() - t0;
    free(ts);
    return s->t_evolve;
}

int ref_nrecords(RefSession *s) { return s->nrecords; }
int ref_record_nbins(RefSession *s, int i) { return s->rec[i].nbins; }
/* scalars[0]=a_f, [1..3]=vel_std, [4..6]=pos_min, [7..9]=pos_max, [10..12]=acc_std, [13]=Plin/D^2, [14]=a_x, [15]=a_v */
void ref_record(RefSession *s, int i, double *scalars, double *k, double *p, double *nmodes)
{
    ForceRecord *r = &s->rec[i];
    scalars[0] = r->a_f;
    for (int d = 0; d < 3; d++) {
        scalars[1 + d] = r->vel_std[d]; scalars[4 + d] = r->pos_min[d];
        scalars[7 + d] = r->pos_max[d]; scalars[10 + d] = r->acc_std[d];
    }
    scalars[13] = r->Plin; scalars[14] = r->a_x; scalars[15] = r->a_v;
    if (k) memcpy(k, r->k, sizeof(double) * r->nbins);
    if (p) memcpy(p, r->p, sizeof(double) * r->nbins);
    if (nmodes) memcpy(nmodes, r->nmodes, sizeof(double) * r->nbins);
}

/* ------------------------------------------------------ per-kernel entries */
static void tmp_store(FastPMStore *p, const double *x, int64_t np)
{
    fastpm_store_init(p, "tmp", np > 0 ? np : 1, COLUMN_POS | COLUMN_ACC, FASTPM_MEMORY_HEAP);
    p->np = np;
    memcpy(p->x, x, sizeof(p->x[0]) * np);
    p->meta.M0 = 1.0;
}

/* fastpm_store_fill (store.c:723-806) of a scratch store with q and rand columns on the particle grid of the LPT mesh */
void ref_set_fake_rank(int r);               /* shims/src/mpi_stub.c */
int64_t ref_fill_probe_as_rank(RefSession *s, int rank, int64_t np_upper, float *q_out, float *rand_out);
int64_t ref_fill_probe(RefSession *s, int64_t np_upper, float *q_out, float *rand_out) { return ref_fill_probe_as_rank(s, 0, np_upper, q_out, rand_out); }
/* the same with MPI_Comm_rank answering `rank` while the store is filled: the seed chain of _fastpm_store_fill_rand for that rank */
int64_t ref_fill_probe_as_rank(RefSession *s, int rank, int64_t np_upper, float *q_out, float *rand_out)
{
    ref_set_fake_rank(rank);
    PM *pm = s->solver->lptpm;
    FastPMStore p[1];
    fastpm_store_init(p, "probe", np_upper, COLUMN_POS | COLUMN_ID | COLUMN_Q | COLUMN_RAND | COLUMN_MASK, FASTPM_MEMORY_HEAP);
    fastpm_store_fill(p, pm, NULL, NULL);
    const int64_t np = p->np;
    memcpy(q_out, p->q, sizeof(p->q[0]) * p->np);
    memcpy(rand_out, p->rand, sizeof(p->rand[0]) * p->np_upper);
    fastpm_store_destroy(p);
    ref_set_fake_rank(0);
    return np;
}

/* the command line's particle_fraction (src/fastpm.c:1449-1461) on a scratch store filled on the LPT mesh's particle grid:
 * fastpm_store_fill_subsample_mask + fastpm_store_subsample into a second store (or in place), fastpm_store_get_mask_sum; then
 * fastpm_store_sort(FastPMLocalSortByID) of the kept particles after they were reversed by fastpm_store_permute (sort_back != 0) */
int64_t ref_subsample_probe(RefSession *s, int64_t np_upper, double fraction, int in_place, int sort_back, uint64_t *id_out, double *x_out, int64_t *mask_sum)
{
    PM *pm = s->solver->lptpm;
    FastPMStore p[1], po[1];
    const FastPMColumnTags attrs = COLUMN_POS | COLUMN_ID | COLUMN_Q | COLUMN_RAND | COLUMN_MASK;
    fastpm_store_init(p, "probe", np_upper, attrs, FASTPM_MEMORY_HEAP);
    fastpm_store_fill(p, pm, NULL, NULL);
    fastpm_store_fill_subsample_mask(p, fraction, p->mask);
    *mask_sum = (int64_t) fastpm_store_get_mask_sum(p, MPI_COMM_WORLD);
    FastPMStore *out = p;
    if (!in_place) {
        fastpm_store_init(po, "kept", fastpm_store_subsample(p, p->mask, NULL) + 1, attrs & ~COLUMN_MASK, FASTPM_MEMORY_HEAP);
        out = po;
    }
    fastpm_store_subsample(p, p->mask, out);
    if (sort_back && out->np) {
        int *ind = malloc(sizeof(int) * out->np);
        for (size_t i = 0; i < out->np; i++) ind[i] = (int) (out->np - 1 - i);
        fastpm_store_permute(out, ind);
        free(ind);
        if (out->id[0] < out->id[out->np - 1]) fastpm_raise(-1, "ref_subsample_probe: the permutation did not reverse the store\n");
        fastpm_store_sort(out, FastPMLocalSortByID);
    }
    const int64_t kept = out->np;
    memcpy(id_out, out->id, sizeof(out->id[0]) * out->np);
    memcpy(x_out, out->x, sizeof(out->x[0]) * out->np);
    if (!in_place) fastpm_store_destroy(po);
    fastpm_store_destroy(p);
    return kept;
}

/* fastpm_paint_local (painter.c:320) of unit-mass particles onto a cleared canvas */
void ref_paint(RefSession *s, int which, double a, const double *x, int64_t np, float *canvas_out)
{
    PM *pm = pick_pm(s, which, a);
    FastPMPainter painter[1];
    fastpm_painter_init(painter, pm, FASTPM_PAINTER_CIC, 2);
    FastPMStore p[1];
    tmp_store(p, x, np);
    FastPMFloat *canvas = pm_alloc(pm);
    FastPMFieldDescr none = { 0, 0 };
    fastpm_paint_local(painter, canvas, p, p->np, none);
    memcpy(canvas_out, canvas, sizeof(FastPMFloat) * pm->allocsize);
    pm_free(pm, canvas);
    fastpm_store_destroy(p);
}

/* fastpm_readout_local (painter.c:358) into acc[:, memb] */
void ref_readout(RefSession *s, int which, double a, const float *canvas_in, const double *x, int64_t np, float *out)
{
    PM *pm = pick_pm(s, which, a);
    FastPMPainter painter[1];
    fastpm_painter_init(painter, pm, FASTPM_PAINTER_CIC, 2);
    FastPMStore p[1];
    tmp_store(p, x, np);
    FastPMFloat *canvas = pm_alloc(pm);
    memcpy(canvas, canvas_in, sizeof(FastPMFloat) * pm->allocsize);
    FastPMFieldDescr f = { COLUMN_ACC, 0 };
    fastpm_readout_local(painter, canvas, p, p->np, f);
    for (int64_t i = 0; i < np; i++) out[i] = p->acc[i][0];
    pm_free(pm, canvas);
    fastpm_store_destroy(p);
}

/* the same two with the generic windows (fastpm_painter_init, painter.c:128-174): type = FastPMPainterType */
void ref_paint_window(RefSession *s, int which, double a, int type, int support, const double *x, int64_t np, float *canvas_out)
{
    PM *pm = pick_pm(s, which, a);
    FastPMPainter painter[1];
    fastpm_painter_init(painter, pm, (FastPMPainterType) type, support);
    FastPMStore p[1];
    tmp_store(p, x, np);
    FastPMFloat *canvas = pm_alloc(pm);
    FastPMFieldDescr none = { 0, 0 };
    fastpm_paint_local(painter, canvas, p, p->np, none);
    memcpy(canvas_out, canvas, sizeof(FastPMFloat) * pm->allocsize);
    pm_free(pm, canvas);
    fastpm_store_destroy(p);
}

void ref_readout_window(RefSession *s, int which, double a, int type, int support, const float *canvas_in, const double *x, int64_t np, float *out)
{
    PM *pm = pick_pm(s, which, a);
    FastPMPainter painter[1];
    fastpm_painter_init(painter, pm, (FastPMPainterType) type, support);
    FastPMStore p[1];
    tmp_store(p, x, np);
    FastPMFloat *canvas = pm_alloc(pm);
    memcpy(canvas, canvas_in, sizeof(FastPMFloat) * pm->allocsize);
    FastPMFieldDescr f = { COLUMN_ACC, 0 };
    fastpm_readout_local(painter, canvas, p, p->np, f);
    for (int64_t i = 0; i < np; i++) out[i] = p->acc[i][0];
    pm_free(pm, canvas);
    fastpm_store_destroy(p);
}

/* the same two through a derivative painter (fastpm_painter_init_diff, painter.c:178-205; type 0 = CIC: painter-cic.c:57-60) */
void ref_paint_window_diff(RefSession *s, int which, double a, int type, int support, int diffdir, const double *x, int64_t np, float *canvas_out)
{
    PM *pm = pick_pm(s, which, a);
    FastPMPainter base[1], painter[1];
    fastpm_painter_init(base, pm, (FastPMPainterType) type, support);
    fastpm_painter_init_diff(painter, base, diffdir);
    FastPMStore p[1];
    tmp_store(p, x, np);
    FastPMFloat *canvas = pm_alloc(pm);
    FastPMFieldDescr none = { 0, 0 };
    fastpm_paint_local(painter, canvas, p, p->np, none);
    memcpy(canvas_out, canvas, sizeof(FastPMFloat) * pm->allocsize);
    pm_free(pm, canvas);
    fastpm_store_destroy(p);
}

void ref_readout_window_diff(RefSession *s, int which, double a, int type, int support, int diffdir, const float *canvas_in, const double *x, int64_t np, float *out)
{
    PM *pm = pick_pm(s, which, a);
    FastPMPainter base[1], painter[1];
    fastpm_painter_init(base, pm, (FastPMPainterType) type, support);
    fastpm_painter_init_diff(painter, base, diffdir);
    FastPMStore p[1];
    tmp_store(p, x, np);
    FastPMFloat *canvas = pm_alloc(pm);
    memcpy(canvas, canvas_in, sizeof(FastPMFloat) * pm->allocsize);
    FastPMFieldDescr f = { COLUMN_ACC, 0 };
    fastpm_readout_local(painter, canvas, p, p->np, f);
    for (int64_t i = 0; i < np; i++) out[i] = p->acc[i][0];
    pm_free(pm, canvas);
    fastpm_store_destroy(p);
}

void ref_r2c(RefSession *s, int which, double a, const float *in, float *out)
{
    PM *pm = pick_pm(s, which, a);
    FastPMFloat *from = pm_alloc(pm), *to = pm_alloc(pm);
    memcpy(from, in, sizeof(FastPMFloat) * pm->allocsize);
    pm_r2c(pm, from, to);
    memcpy(out, to, sizeof(FastPMFloat) * pm->allocsize);
    pm_free(pm, to); pm_free(pm, from);
}

void ref_c2r(RefSession *s, int which, double a, const float *in, float *out)
{
    PM *pm = pick_pm(s, which, a);
    FastPMFloat *buf = pm_alloc(pm);
    memcpy(buf, in, sizeof(FastPMFloat) * pm->allocsize);
    pm_c2r(pm, buf);
    memcpy(out, buf, sizeof(FastPMFloat) * pm->allocsize);
    pm_free(pm, buf);
}

/* gravity_apply_kernel_transfer (gravity.c:174): attr 0 = ACC[memb], 1 = POTENTIAL */
void ref_kernel_transfer(RefSession *s, int which, double a, const float *delta_k_in, int attr, int memb, float *out)
{
    PM *pm = pick_pm(s, which, a);
    FastPMFloat *dk = pm_alloc(pm), *canvas = pm_alloc(pm);
    memcpy(dk, delta_k_in, sizeof(FastPMFloat) * pm->allocsize);
    /* attr: 0 acceleration, 1 potential, 2 density, 3 tidal (memb 0..5) */
    FastPMFieldDescr f = { attr == 0 ? COLUMN_ACC : (attr == 1 ? COLUMN_POTENTIAL : (attr == 2 ? COLUMN_DENSITY : COLUMN_TIDAL)), memb };
    gravity_apply_kernel_transfer(s->solver->config->KERNEL_TYPE, pm, dk, canvas, f);
    memcpy(out, canvas, sizeof(FastPMFloat) * pm->allocsize);
    pm_free(pm, canvas); pm_free(pm, dk);
}

/* fastpm_pgdc_calculate (pgdcorrection.c:61-137) for np positions: par = {alpha0, A, B, kl, ks}; out[np][3] */
void ref_pgdc(RefSession *s, int which, double a, const float *delta_k_in, const double *x, int64_t np, const double *par, float *out)
{
    PM *pm = pick_pm(s, which, a);
    FastPMPGDCorrection pgdc[1] = {{ FASTPM_PAINTER_CIC, 2, par[0], par[1], par[2], par[3], par[4] }};
    FastPMStore p[1];
    fastpm_store_init(p, "tmp", np > 0 ? np : 1, COLUMN_POS | COLUMN_PGDC, FASTPM_MEMORY_HEAP);
    p->np = np;
    memcpy(p->x, x, sizeof(p->x[0]) * np);
    FastPMFloat *dk = pm_alloc(pm);
    memcpy(dk, delta_k_in, sizeof(FastPMFloat) * pm->allocsize);
    fastpm_pgdc_calculate(pgdc, pm, p, dk, a, 1.0);
    memcpy(out, p->pgdc, sizeof(p->pgdc[0]) * np);
    pm_free(pm, dk);
    fastpm_store_destroy(p);
}

/* the pgdc column of the session's particles (zeros when the correction is off) */
void ref_get_pgdc(RefSession *s, float *out)
{
    FastPMStore *p = fastpm_solver_get_species(s->solver, FASTPM_SPECIES_CDM);
    if (p->pgdc) memcpy(out, p->pgdc, sizeof(p->pgdc[0]) * p->np);
    else memset(out, 0, sizeof(float) * 3 * p->np);
}

void ref_set_pgdc(RefSession *s, const float *in)
{
    FastPMStore *p = fastpm_solver_get_species(s->solver, FASTPM_SPECIES_CDM);
    if (p->pgdc) memcpy(p->pgdc, in, sizeof(p->pgdc[0]) * p->np);
}

/* One snapshot the way src/fastpm.c:1190-1200,1473-1486 writes it, without the Lua-dependent "ParamFile" attribute and without the
 * distributed sort: fastpm_set_species_snapshot (unit conversion + wrap, solver.c:647-702; no drift / kick: the particles already
 * sit at aout), write_snapshot_header, fastpm_store_write, fastpm_unset_species_snapshot. */
void ref_write_snapshot(RefSession *s, const char *filebase)
{
    FastPMSolver *fastpm = s->solver;
    FastPMStore *p = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM);
    FastPMStore po[1];
    double aout = p->meta.a_x;
    fastpm_set_species_snapshot(fastpm, p, NULL, NULL, po, aout);
    FastPMSolver snapshot[1];
    memcpy(snapshot, fastpm, sizeof(FastPMSolver));
    fastpm_solver_add_species(snapshot, FASTPM_SPECIES_CDM, po);
    write_snapshot_header(snapshot, filebase, MPI_COMM_WORLD);
    fastpm_store_write(po, filebase, "w", 0, MPI_COMM_WORLD);
    fastpm_unset_species_snapshot(fastpm, p, NULL, NULL, po, aout);
}

/* the same store appended to the catalog once more (mode "a", io.c:522-537: every block grows) */
void ref_append_snapshot(RefSession *s, const char *filebase)
{
    FastPMSolver *fastpm = s->solver;
    FastPMStore *p = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM);
    FastPMStore po[1];
    double aout = p->meta.a_x;
    fastpm_set_species_snapshot(fastpm, p, NULL, NULL, po, aout);
    fastpm_store_write(po, filebase, "a", 0, MPI_COMM_WORLD);
    fastpm_unset_species_snapshot(fastpm, p, NULL, NULL, po, aout);
}

/* fastpm_solver_evolve with snapshots at aout[] the way the CLI takes them: check_snapshots (src/fastpm.c:1130-1208) restated as
 * an INTERPOLATION handler (it is static in the CLI), then write_snapshot_header + fastpm_store_write to "<base>_%0.04f". */
typedef struct { const char *base; const double *aout; int nout, iout; } RefSnapPlan;
static int ref_check_snapshots(FastPMSolver *fastpm, FastPMInterpolationEvent *event, RefSnapPlan *plan)
{
    for (int iout = plan->iout; iout < plan->nout; iout++) {
        double aout = plan->aout[iout];
        if (event->a1 == event->a2) { if (event->a1 != aout) continue; }
        else { if (event->a1 >= aout) continue; if (event->a2 < aout) continue; }
        FastPMSolver snapshot[1];
        FastPMStore cdm[1];
        memcpy(snapshot, fastpm, sizeof(FastPMSolver));
        fastpm_solver_add_species(snapshot, FASTPM_SPECIES_CDM, cdm);
        fastpm_set_snapshot(fastpm, snapshot, event->drift, event->kick, aout);
        char filebase[1024];
        sprintf(filebase, "%s_%0.04f", plan->base, aout);
        write_snapshot_header(snapshot, filebase, fastpm->comm);
        fastpm_store_write(cdm, filebase, "w", 0, fastpm->comm);
        fastpm_unset_snapshot(fastpm, snapshot, event->drift, event->kick, aout);
        plan->iout = iout + 1;
    }
    return 0;
}

void ref_evolve_snapshots(RefSession *s, const double *time_step, int nstep, const char *base, const double *aout_sorted, int nout)
{
    RefSnapPlan plan = { base, aout_sorted, nout, 0 };
    fastpm_add_event_handler(&s->solver->event_handlers, FASTPM_EVENT_INTERPOLATION, FASTPM_EVENT_STAGE_BEFORE,
            (FastPMEventHandlerFunction) ref_check_snapshots, &plan);
    clear_records(s);
    fastpm_solver_evolve(s->solver, (double *) time_step, nstep);
    fastpm_remove_event_handler(&s->solver->event_handlers, FASTPM_EVENT_INTERPOLATION, FASTPM_EVENT_STAGE_BEFORE,
            (FastPMEventHandlerFunction) ref_check_snapshots, &plan);
}

/* the unit-converted, wrapped particles that ref_write_snapshot hands to fastpm_store_write */
void ref_snapshot_particles(RefSession *s, double *x, float *v)
{
    FastPMSolver *fastpm = s->solver;
    FastPMStore *p = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM);
    FastPMStore po[1];
    double aout = p->meta.a_x;
    fastpm_set_species_snapshot(fastpm, p, NULL, NULL, po, aout);
    memcpy(x, po->x, sizeof(po->x[0]) * po->np);
    memcpy(v, po->v, sizeof(po->v[0]) * po->np);
    fastpm_unset_species_snapshot(fastpm, p, NULL, NULL, po, aout);
}

/* reads the particle columns of a snapshot back into the session's store (fastpm_store_read, io.c:590-597) and returns the
 * ScalingFactor of the header (read_snapshot_header, io.c:163-226) */
double ref_read_snapshot(RefSession *s, const char *filebase, int restart)
{
    FastPMStore *p = fastpm_solver_get_species(s->solver, FASTPM_SPECIES_CDM);
    double aout = 0;
    read_snapshot_header(s->solver, filebase, &aout, MPI_COMM_WORLD);
    if (!restart) {                       /* the columns as they are in the file (v in km/s) */
        p->np = 0;
        fastpm_store_read(p, filebase, 0, MPI_COMM_WORLD);
        return aout;
    }
    /* prepare_cdm with a restart path, src/fastpm.c:618-635: back to the integrator's units */
    FastPMStore po[1];
    fastpm_set_species_snapshot(s->solver, p, NULL, NULL, po, 1.0);
    fastpm_store_read(po, filebase, 0, MPI_COMM_WORLD);
    fastpm_unset_species_snapshot(s->solver, p, NULL, NULL, po, po->meta.a_x);
    return aout;
}

/* write_complex (io.c:641-719) needs the distributed sort of depends/mpsort, not built here: the oracle writes the c8 block the same
 * way through the bigfile library, in the [x][y][N/2+1] order that sort produces */
void ref_write_complex(RefSession *s, int which, double a, const float *dk_in, const char *filename, const char *blockname)
{
    PM *pm = pick_pm(s, which, a);
    int Nmesh = pm_nmesh(pm)[0];
    double BoxSize = pm_boxsize(pm)[0];
    int64_t strides[3] = { (int64_t) Nmesh * (Nmesh / 2 + 1), Nmesh / 2 + 1, 1 };
    int64_t shape[3] = { Nmesh, Nmesh, Nmesh / 2 + 1 };
    size_t size = (size_t) Nmesh * Nmesh * (Nmesh / 2 + 1);
    float *buf = malloc(sizeof(float) * 2 * size);
    PMKIter kiter;
    for (pm_kiter_init(pm, &kiter); !pm_kiter_stop(&kiter); pm_kiter_next(&kiter)) {
        size_t iabs = kiter.iabs[0] * strides[0] + kiter.iabs[1] * strides[1] + kiter.iabs[2] * strides[2];
        buf[2 * iabs] = dk_in[kiter.ind]; buf[2 * iabs + 1] = dk_in[kiter.ind + 1];
    }
    BigFile bf; BigBlock bb; BigArray array; BigBlockPtr ptr;
    fastpm_path_ensure_dirname(filename);
    if (0 != big_file_mpi_create(&bf, filename, MPI_COMM_WORLD)) { fprintf(stderr, "%s\n", big_file_get_error_message()); abort(); }
    if (0 != big_file_mpi_create_block(&bf, &bb, blockname, "c8", 1, 1, size, MPI_COMM_WORLD)) { fprintf(stderr, "%s\n", big_file_get_error_message()); abort(); }
    big_array_init(&array, buf, "c8", 1, (size_t[]) { size, 1 }, NULL);
    big_block_seek(&bb, &ptr, 0);
    big_block_mpi_write(&bb, &ptr, &array, 1, MPI_COMM_WORLD);
    big_block_set_attr(&bb, "ndarray.ndim", (int[]) { 3, }, "i4", 1);
    big_block_set_attr(&bb, "ndarray.strides", strides, "i8", 3);
    big_block_set_attr(&bb, "ndarray.shape", shape, "i8", 3);
    big_block_set_attr(&bb, "Nmesh", &Nmesh, "i4", 1);
    big_block_set_attr(&bb, "BoxSize", &BoxSize, "f8", 1);
    big_block_mpi_close(&bb, MPI_COMM_WORLD);
    big_file_mpi_close(&bf, MPI_COMM_WORLD);
    free(buf);
}

/* the single-mode transfers of transfer.c on the LPT mesh: op 0 = set_mode (mode[4], value, method), 1 = normalize, 2 = c2r_weight;
 * returns get_mode(out, mode) */
double ref_mode_op(RefSession *s, int op, const float *in, float *out, const int64_t *mode4, double value, int method)
{
    PM *pm = s->solver->lptpm;
    ptrdiff_t mode[4] = { mode4[0], mode4[1], mode4[2], mode4[3] };
    FastPMFloat *a = pm_alloc(pm), *b = pm_alloc(pm);
    memcpy(a, in, sizeof(FastPMFloat) * pm->allocsize);
    if (op == 0) fastpm_apply_set_mode_transfer(pm, a, b, mode, value, method);
    else if (op == 1) fastpm_apply_normalize_transfer(pm, a, b);
    else fastpm_apply_c2r_weight_transfer(pm, a, b);
    double r = fastpm_apply_get_mode_transfer(pm, b, mode);
    memcpy(out, b, sizeof(FastPMFloat) * pm->allocsize);
    pm_free(pm, b); pm_free(pm, a);
    return r;
}

void ref_decic(RefSession *s, int which, double a, const float *in, float *out)
{
    PM *pm = pick_pm(s, which, a);
    FastPMFloat *dk = pm_alloc(pm);
    memcpy(dk, in, sizeof(FastPMFloat) * pm->allocsize);
    fastpm_apply_decic_transfer(pm, dk, dk);
    memcpy(out, dk, sizeof(FastPMFloat) * pm->allocsize);
    pm_free(pm, dk);
}

/* laplace + diff (the IC-side gradient, transfer.c:116,154), used by pm_2lpt_solve */
void ref_laplace_diff(RefSession *s, int which, double a, const float *in, int d1, int d2, float *out)
{
    PM *pm = pick_pm(s, which, a);
    int potorder, gradorder, difforder, deconvolveorder;
    fastpm_kernel_type_get_orders(s->solver->config->KERNEL_TYPE, &potorder, &gradorder, &difforder, &deconvolveorder);
    FastPMFloat *dk = pm_alloc(pm), *w = pm_alloc(pm);
    memcpy(dk, in, sizeof(FastPMFloat) * pm->allocsize);
    fastpm_apply_laplace_transfer(pm, dk, w, potorder);
    if (d1 >= 0) fastpm_apply_diff_transfer(pm, w, w, d1, difforder);
    if (d2 >= 0) fastpm_apply_diff_transfer(pm, w, w, d2, difforder);
    memcpy(out, w, sizeof(FastPMFloat) * pm->allocsize);
    pm_free(pm, w); pm_free(pm, dk);
}

int ref_powerspectrum(RefSession *s, int which, double a, const float *delta_k_in, double *k, double *p, double *nmodes)
{
    PM *pm = pick_pm(s, which, a);
    FastPMFloat *dk = pm_alloc(pm);
    memcpy(dk, delta_k_in, sizeof(FastPMFloat) * pm->allocsize);
    FastPMPowerSpectrum ps;
    fastpm_powerspectrum_init_from_delta(&ps, pm, dk, dk);
    int n = (int) ps.base.size;
    memcpy(k, ps.base.k, sizeof(double) * n);
    memcpy(p, ps.base.f, sizeof(double) * n);
    memcpy(nmodes, ps.Nmodes, sizeof(double) * n);
    fastpm_powerspectrum_destroy(&ps);
    pm_free(pm, dk);
    return n;
}

/* the cross spectrum of two fields (delta1_k != delta2_k, powerspectrum.c:87-91) */
int ref_cross_powerspectrum(RefSession *s, int which, double a, const float *delta1_k_in, const float *delta2_k_in, double *k, double *p, double *nmodes)
{
    PM *pm = pick_pm(s, which, a);
    FastPMFloat *dk1 = pm_alloc(pm), *dk2 = pm_alloc(pm);
    memcpy(dk1, delta1_k_in, sizeof(FastPMFloat) * pm->allocsize);
    memcpy(dk2, delta2_k_in, sizeof(FastPMFloat) * pm->allocsize);
    FastPMPowerSpectrum ps;
    fastpm_powerspectrum_init_from_delta(&ps, pm, dk1, dk2);
    int n = (int) ps.base.size;
    memcpy(k, ps.base.k, sizeof(double) * n);
    memcpy(p, ps.base.f, sizeof(double) * n);
    memcpy(nmodes, ps.Nmodes, sizeof(double) * n);
    fastpm_powerspectrum_destroy(&ps);
    pm_free(pm, dk2);
    pm_free(pm, dk1);
    return n;
}

/* One force evaluation on the session's particles at time a (solver.c:404 without events):
 * wrap + decompose + fastpm_solver_compute_force; delta_k (before decic) optionally returned. */
void ref_compute_force(RefSession *s, double a, float *delta_k_out)
{
    FastPMSolver *fastpm = s->solver;
    PM *pm = fastpm_find_pm(fastpm, a);
    FastPMStore *p = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM);
    FastPMPainter painter[1];
    FastPMFloat *delta_k = pm_alloc(pm);
    fastpm_painter_init(painter, pm, fastpm->config->PAINTER_TYPE, fastpm->config->painter_support);
    fastpm_store_wrap(p, pm->BoxSize);
    fastpm_store_decompose(p, (fastpm_store_target_func) FastPMTargetPM, pm, fastpm->comm);
    fastpm_solver_compute_force(fastpm, pm, painter, fastpm->config->SOFTENING_TYPE, fastpm->config->KERNEL_TYPE, delta_k, a);
    if (delta_k_out) memcpy(delta_k_out, delta_k, sizeof(FastPMFloat) * pm->allocsize);
    pm_free(pm, delta_k);
}

/* out = {ai, ac, af, q1, q2, dda[32], Dv1[32], Dv2[32]} */
void ref_kick_factor(RefSession *s, double ai, double ac, double af, double *out)
{
    FastPMKickFactor k;
    fastpm_kick_init(&k, s->solver, ai, ac, af);
    out[0] = k.ai; out[1] = k.ac; out[2] = k.af; out[3] = k.q1; out[4] = k.q2;
    memcpy(out + 5, k.dda, sizeof(double) * 32);
    memcpy(out + 37, k.Dv1, sizeof(double) * 32);
    memcpy(out + 69, k.Dv2, sizeof(double) * 32);
}

/* out = {ai, ac, af, Dv1, Dv2, dyyy[32], da1[32], da2[32]} */
void ref_drift_factor(RefSession *s, double ai, double ac, double af, double *out)
{
    FastPMDriftFactor d;
    fastpm_drift_init(&d, s->solver, ai, ac, af);
    out[0] = d.ai; out[1] = d.ac; out[2] = d.af; out[3] = d.Dv1; out[4] = d.Dv2;
    memcpy(out + 5, d.dyyy, sizeof(double) * 32);
    memcpy(out + 37, d.da1, sizeof(double) * 32);
    memcpy(out + 69, d.da2, sizeof(double) * 32);
}

void ref_kick(RefSession *s, double ai, double ac, double af)
{
    FastPMStore *p = fastpm_solver_get_species(s->solver, FASTPM_SPECIES_CDM);
    FastPMKickFactor k;
    fastpm_kick_init(&k, s->solver, ai, ac, af);
    fastpm_kick_store(&k, p, p, af);
}

void ref_drift(RefSession *s, double ai, double ac, double af)
{
    FastPMStore *p = fastpm_solver_get_species(s->solver, FASTPM_SPECIES_CDM);
    FastPMDriftFactor d;
    fastpm_drift_init(&d, s->solver, ai, ac, af);
    fastpm_drift_store(&d, p, p, af);
}

/* out = {D1, D2, f1, f2, E, dE/da, d2E/da2, dD1/da, d2D1/da2, Omega_source(a), Omega_Lambda, Omega_cdm} */
void ref_growth(RefSession *s, double a, double *out)
{
    FastPMCosmology *c = s->solver->cosmology;
    FastPMGrowthInfo gi;
    fastpm_growth_info_init(&gi, a, c);
    out[0] = gi.D1; out[1] = gi.D2; out[2] = gi.f1; out[3] = gi.f2;
    out[4] = HubbleEa(a, c); out[5] = DHubbleEaDa(a, c); out[6] = D2HubbleEaDa2(a, c);
    out[7] = DGrowthFactorDa(&gi); out[8] = D2GrowthFactorDa2(&gi);
    out[9] = Omega_source(a, c); out[10] = c->Omega_Lambda; out[11] = c->Omega_cdm;
}

/* the KDK schedule (timemachine.c:23-140) for a time table: rows of {action, a_i, a_f, a_r, x, v, force} */
int ref_schedule(const double *time_step, int nstep, double *rows, int maxrows)
{
    FastPMStates states[1];
    FastPMState templ[] = { {0, 0, 1}, {0, 1, 1}, {0, 2, 1}, {2, 2, 1}, {2, 2, 2}, {-1, -1, -1} };
    double *ts = malloc(sizeof(double) * nstep);
    memcpy(ts, time_step, sizeof(double) * nstep);
    fastpm_tevo_generate_states(states, nstep - 1, templ, ts);
    int n = 0;
    for (int i = 1; states->table[i].force != -1 && n < maxrows; i++, n++) {
        FastPMTransition tr[1];
        fastpm_tevo_transition_init(tr, states, i - 1, i);
        double *r = rows + 7 * n;
        r[0] = tr->action; r[1] = tr->a.i; r[2] = tr->a.f; r[3] = tr->a.r;
        r[4] = tr->end->x; r[5] = tr->end->v; r[6] = tr->end->force;
    }
    fastpm_tevo_destroy_states(states);
    free(ts);
    return n;
}

void ref_clock_stat(void) { fastpm_clock_stat(MPI_COMM_WORLD); }

/* the table fastpm_clock_stat prints (prof.c:144-178: "min max mean name : func : file" per clock, cumulative seconds
 * since the library was loaded), captured as text: the clock list itself is private to prof.c */
static char *clk_buf; static size_t clk_cap, clk_len;
static void clk_capture(const enum FastPMLogLevel level, const enum FastPMLogType type, const int errcode,
                        const char *message, MPI_Comm comm, void *userdata)
{
    (void) level; (void) type; (void) errcode; (void) comm; (void) userdata;
    size_t n = strlen(message);
    if (clk_len + n + 1 < clk_cap) { memcpy(clk_buf + clk_len, message, n); clk_len += n; clk_buf[clk_len] = 0; }
}
int ref_clock_table(char *out, int cap)
{
    clk_buf = out; clk_cap = (size_t) cap; clk_len = 0;
    if (cap > 0) out[0] = 0;
    fastpm_push_msg_handler(clk_capture, MPI_COMM_WORLD, NULL);
    fastpm_clock_stat(MPI_COMM_WORLD);
    fastpm_pop_msg_handler();
    return (int) clk_len;
}

