/* fastpm_b200 host layer -- the solver object and the kick-drift-kick main loop
 * (reference: libfastpm/solver.c, pm2lpt.c).  Host code only sequences device work: every particle or
 * mesh operation below is a call into the CUDA library. */
#include "internal.h"

static void fastpm_decompose(FastPMSolver *fastpm, PM *pm);
static void do_interpolation(FastPMSolver *fastpm, FastPMDriftFactor *drift, FastPMKickFactor *kick, double a1, double a2, int whence);

void fastpm_solver_init(FastPMSolver *fastpm, FastPMConfig *config, MPI_Comm comm)
{
    libfastpm_init();
    fastpm->config[0] = *config;
    if (!config->cosmology) {
        /* the reference's fiducial cosmology, solver.c:30-47 (radiation on, LCDM growth) */
        FastPMCosmology c;
        memset(&c, 0, sizeof(c));
        c.h = 0.6772; c.Omega_m = 0.323839; c.Omega_cdm = 0.3; c.Omega_Lambda = 0.67616; c.T_cmb = 2.725;
        c.w0 = -1; c.wa = 0; c.N_eff = 3.046; c.m_ncdm[0] = 1.; c.N_nu = 3; c.growth_mode = FASTPM_GROWTH_MODE_LCDM;
        fastpm->cosmology[0] = c;
    } else {
        fastpm->cosmology[0] = *config->cosmology;
    }
    fastpm_cosmology_init(fastpm->cosmology);
    memset(fastpm->pgdc, 0, sizeof(fastpm->pgdc));
    if (config->pgdc) {                                  /* solver.c:55-66 */
        fastpm->pgdc[0].PainterType = config->PAINTER_TYPE; fastpm->pgdc[0].PainterSupport = config->painter_support;
        fastpm->pgdc[0].alpha0 = config->pgdc_alpha0; fastpm->pgdc[0].A = config->pgdc_A; fastpm->pgdc[0].B = config->pgdc_B;
        fastpm->pgdc[0].kl = config->pgdc_kl; fastpm->pgdc[0].ks = config->pgdc_ks;
    }
    fastpm->event_handlers = NULL;
    fastpm->comm = comm;
    fastpm->ThisTask = fpm_comm_rank(comm);
    fastpm->NTask = fpm_comm_size(comm);
    if (config->FORCE_TYPE == FASTPM_FORCE_COLA) {
        config->ExtraAttributes |= COLUMN_DX1;          /* solver.c:84-88 */
        config->ExtraAttributes |= COLUMN_DX2;
    }
    memset(fastpm->has_species, 0, FASTPM_SOLVER_NSPECIES);
    /* MASK and RAND are allocated by the reference too (solver.c:93-97) but feed only sub-sampling and
     * snapshots; they are left out of the device store unless asked for through ExtraAttributes */
    fastpm_store_init_evenly(fastpm->cdm, fastpm_species_get_name(FASTPM_SPECIES_CDM), pow(1.0 * config->nc, 3),
            COLUMN_POS | COLUMN_VEL | COLUMN_ID | COLUMN_ACC | config->ExtraAttributes, config->alloc_factor, comm);
    fastpm_solver_add_species(fastpm, FASTPM_SPECIES_CDM, fastpm->cdm);

    fastpm->vpm_list = vpm_create(config->vpminit, (int) config->nc, config->boxsize, comm);
    fastpm->basepm = pm_new((int) config->nc, config->boxsize, comm);
    if (pm_unbalanced(fastpm->basepm)) fastpm_raise(-1, "Base PM mesh is not divided by the process mesh.\n");
    fastpm->lptpm = pm_new((int) (config->nc * (config->lpt_nc_factor ? config->lpt_nc_factor : 1)), config->boxsize, comm);
    if (pm_unbalanced(fastpm->lptpm)) fastpm_raise(-1, "LPT PM mesh is not divided by the process mesh.\n");
    double shift0 = config->USE_SHIFT ? config->boxsize / config->nc * 0.5 : 0;
    double shift[3] = { shift0, shift0, shift0 };
    fastpm_store_fill(fastpm->cdm, fastpm->basepm, shift, NULL);
}

void fastpm_solver_destroy(FastPMSolver *fastpm)
{
    fpm_comm_release_migration();
    pm_delete(fastpm->lptpm);
    pm_delete(fastpm->basepm);
    fastpm_store_destroy(fastpm->cdm);
    vpm_free(fastpm->vpm_list);
    fastpm_cosmology_destroy(fastpm->cosmology);
    fastpm_destroy_event_handlers(&fastpm->event_handlers);
}

FastPMStore *fastpm_solver_get_species(FastPMSolver *fastpm, enum FastPMSpecies species)
{ return fastpm->has_species[species] ? fastpm->species[species] : NULL; }
void fastpm_solver_add_species(FastPMSolver *fastpm, enum FastPMSpecies species, FastPMStore *store)
{ fastpm->species[species] = store; fastpm->has_species[species] = 1; }
PM *fastpm_find_pm(FastPMSolver *fastpm, double a) { return vpm_find(fastpm->vpm_list, a)->pm; }

/* ------------------------------------------------------------------ 2LPT (pm2lpt.c) */
static fpm_transfer lpt_kernel(int potorder, int difforder, int d1, int d2)
{
    fpm_transfer t;
    memset(&t, 0, sizeof(t));
    t.active = 1; t.potorder = potorder; t.negate = 0; t.gradorder = difforder; t.zero_selfconj = 1; t.scale = 1.0;
    t.ngrad = 0;
    if (d1 >= 0) t.graddir[t.ngrad++] = d1;
    if (d2 >= 0) t.graddir[t.ngrad++] = d2;
    return t;
}

void pm_2lpt_solve(PM *pm, FastPMFloat *delta_k, FastPMFuncK *growth_rate_func_k, FastPMStore *p, double shift[3], FastPMKernelType type)
{
    if (growth_rate_func_k || p->dv1) fastpm_raise(-1, "fastpm_b200: scale-dependent growth (dv1) is out of scope of this build\n");
    /* pm2lpt.c:30-34: the displacements are read out at the de-shifted (grid) positions; the shift is put back at the end */
    const int shifted = shift[0] != 0 || shift[1] != 0 || shift[2] != 0;
    if (shifted) FPM_MUST(fpm_shift_positions((double *) p->x, (int64_t) p->np, -shift[0], -shift[1], -shift[2]));
    int potorder, gradorder, difforder, deconvolveorder;
    fastpm_kernel_type_get_orders(type, &potorder, &gradorder, &difforder, &deconvolveorder);
    const size_t nf = pm->allocsize;
    FastPMFloat *source = pm_alloc(pm);
    FastPMFloat *workspace = pm_alloc_noclear(pm, __FILE__, __LINE__);
    FastPMFloat *field[3];
    for (int d = 0; d < 3; d++) field[d] = pm_alloc_noclear(pm, __FILE__, __LINE__);
    const int D1[3] = { 1, 2, 0 }, D2[3] = { 2, 0, 1 };
    fpm_transfer t;

    /* dx1_d = c2r( i k_d / k^2 delta ), pm2lpt.c:62-74 */
    for (int d = 0; d < 3; d++) {
        t = lpt_kernel(potorder, difforder, d, -1);
        fpm_mesh_c2r(pm, delta_k, workspace, &t);
        fpm_mesh_readout(pm, workspace, (const double *) p->x, (int64_t) p->np, (float *) p->dx1 + d, 3, 1.0);
    }
    /* phi,dd for the three axes, pm2lpt.c:90-96 */
    for (int d = 0; d < 3; d++) {
        t = lpt_kernel(potorder, difforder, d, d);
        fpm_mesh_c2r(pm, delta_k, field[d], &t);
    }
    for (int d = 0; d < 3; d++) FPM_MUST(fpm_muladd(source, field[D1[d]], field[D2[d]], nf, +1));
    /* cross terms phi,d1d2, pm2lpt.c:108-122 */
    for (int d = 0; d < 3; d++) {
        t = lpt_kernel(potorder, difforder, D1[d], D2[d]);
        fpm_mesh_c2r(pm, delta_k, workspace, &t);
        FPM_MUST(fpm_muladd(source, workspace, workspace, nf, -1));
    }
    /* delta2_k = r2c(source); pm2lpt.c:123-124 copies it back, here the two buffers just swap roles */
    fpm_mesh_r2c(pm, source, workspace, 1.0 / pm->Norm);
    FastPMFloat *delta2_k = workspace, *w2 = source;
    for (int d = 0; d < 3; d++) {
        t = lpt_kernel(potorder, difforder, d, -1);
        fpm_mesh_c2r(pm, delta2_k, w2, &t);
        /* the 3/7 of pm2lpt.c:133 is applied to each mesh value (rounded to float) inside the gather */
        fpm_mesh_readout(pm, w2, (const double *) p->x, (int64_t) p->np, (float *) p->dx2 + d, 3, 3.0 / 7);
    }
    if (shifted) FPM_MUST(fpm_shift_positions((double *) p->x, (int64_t) p->np, shift[0], shift[1], shift[2]));      /* pm2lpt.c:141-145 */
    for (int d = 0; d < 3; d++) pm_free(pm, field[2 - d]);
    pm_free(pm, workspace);
    pm_free(pm, source);
}

void pm_2lpt_evolve(double aout, FastPMStore *p, FastPMCosmology *c, int zaonly)
{
    FastPMGrowthInfo gi;
    fastpm_growth_info_init(&gi, aout, c);
    double D1 = gi.D1, D2 = gi.D2, E = HubbleEa(aout, c);
    double Dv1 = D1 * aout * aout * E * gi.f1, Dv2 = D2 * aout * aout * E * gi.f2;
    fastpm_info("2LPT ICs set at z=%g: E=%g D1=%g, D2=%g, f1=%g, f2=%g\n", 1. / aout - 1, E, D1, D2, gi.f1, gi.f2);
    if (zaonly) { D2 = 0; Dv2 = 0; }
    FPM_MUST(fpm_lpt_evolve((double *) p->x, (float *) p->v, (const float *) p->dx1, (const float *) p->dx2, (int64_t) p->np, D1, D2, Dv1, Dv2));
    p->meta.a_x = p->meta.a_v = aout;
}

void fastpm_solver_setup_lpt(FastPMSolver *fastpm, enum FastPMSpecies species, FastPMFloat *delta_k_ic,
                             FastPMFuncK *growth_rate_func_k_ic, double a0)
{
    FastPMStore *p = fastpm_solver_get_species(fastpm, species);
    if (!p) fastpm_raise(-1, "Species requested (%d) does not exist", species);
    fpm_store_flush(NULL);
    PM *pm = fastpm->lptpm;
    FastPMConfig *config = fastpm->config;
    if (species == FASTPM_SPECIES_CDM) {
        const double M0 = fastpm->cosmology->Omega_cdm * FASTPM_CRITICAL_DENSITY * pow(config->boxsize / config->nc, 3.0);
        fastpm_info("mass of a CDM particle is %g 1e10 Msun/h\n", M0);
        p->meta.M0 = M0;
    }
    int temp_dx1 = 0, temp_dx2 = 0;
    if (p->dx1 == NULL) { p->dx1 = fastpm_memory_alloc(p->mem, "DX1", sizeof(p->dx1[0]) * p->np_upper, FASTPM_MEMORY_STACK); temp_dx1 = 1; }
    if (p->dx2 == NULL) { p->dx2 = fastpm_memory_alloc(p->mem, "DX2", sizeof(p->dx2[0]) * p->np_upper, FASTPM_MEMORY_STACK); temp_dx2 = 1; }

    FastPMLPTEvent event[1];
    event->pm = pm; event->delta_k = delta_k_ic; event->p = p;
    fastpm_emit_event(fastpm->event_handlers, FASTPM_EVENT_LPT, FASTPM_EVENT_STAGE_BEFORE, (FastPMEvent *) event, fastpm);
    if (delta_k_ic) {
        double shift0 = config->USE_SHIFT ? config->boxsize / config->nc * 0.5 : 0;
        double shift[3] = { shift0, shift0, shift0 };
        pm_2lpt_solve(pm, delta_k_ic, growth_rate_func_k_ic, p, shift, config->KERNEL_TYPE);
    }
    if (config->USE_DX1_ONLY == 1) FPM_MUST(fpm_memset(p->dx2, 0, sizeof(p->dx2[0]) * p->np));
    pm_2lpt_evolve(a0, p, fastpm->cosmology, config->USE_DX1_ONLY);
    fastpm_emit_event(fastpm->event_handlers, FASTPM_EVENT_LPT, FASTPM_EVENT_STAGE_AFTER, (FastPMEvent *) event, fastpm);
    if (temp_dx2) { fastpm_memory_free(p->mem, p->dx2); p->dx2 = NULL; }
    if (temp_dx1) { fastpm_memory_free(p->mem, p->dx1); p->dx1 = NULL; }
}

/* ------------------------------------------------------------------ main loop (solver.c:283-555) */
static void do_force(FastPMSolver *fastpm, FastPMTransition *trans)
{
    CLOCK(decompose);
    LEAVE(decompose);
    CLOCK(force);
    LEAVE(force);
    CLOCK(event);
    LEAVE(event);
    PM *pm = fastpm_find_pm(fastpm, trans->a.f);
    FastPMPainter painter[1];
    FastPMFloat *delta_k = pm_alloc_noclear(pm, __FILE__, __LINE__);
    FastPMForceEvent event[1];
    FastPMStore *p = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM);
    int64_t N = p->np;
    fastpm_painter_init(painter, pm, fastpm->config->PAINTER_TYPE, fastpm->config->painter_support);
    fpm_comm_allreduce_i64(fastpm->comm, &N, 1, 0);
    event->delta_k = delta_k; event->a_f = trans->a.f; event->pm = pm; event->N = N;
    event->painter = painter; event->kernel = fastpm->config->KERNEL_TYPE;
    FastPMTransition next[1];
    if (!fastpm_tevo_transition_find_next(trans, next)) event->a_n = -1;
    else {
        if (next->a.i != trans->a.f) fastpm_raise(-1, "Failed to find next Force calculation\n");
        event->a_n = next->a.f;
    }
    ENTER(decompose);
    fastpm_decompose(fastpm, pm);
    LEAVE(decompose);
    fastpm_emit_event(fastpm->event_handlers, FASTPM_EVENT_FORCE, FASTPM_EVENT_STAGE_BEFORE, (FastPMEvent *) event, fastpm);
    ENTER(force);
    fastpm_solver_compute_force(fastpm, pm, painter, fastpm->config->SOFTENING_TYPE, fastpm->config->KERNEL_TYPE, delta_k, trans->a.f);
    if (fpm_pending_wrap) { FastPMStore *pw = fpm_pending_wrap; fpm_pending_wrap = NULL; fastpm_store_wrap(pw, pm->BoxSize); }
    LEAVE(force);
    if (p->pgdc) {
        /* solver.c:458-464; delta_k is input, unchanged (and not yet deconvolved) */
        CLOCK(pgdc);
        fastpm_pgdc_calculate(fastpm->pgdc, pm, p, delta_k, trans->a.f, 1.0);
        LEAVE(pgdc);
    }
    ENTER(event);
    /* solver.c:471: the event sees the CIC-compensated density.  Nobody else reads delta_k afterwards, so the
     * sweep is skipped when no FORCE/after handler is installed. */
    int has_after = 0;
    for (FastPMEventHandler *h = fastpm->event_handlers; h; h = h->next)
        if (h->stage == FASTPM_EVENT_STAGE_AFTER && !strcmp(h->type, FASTPM_EVENT_FORCE)) has_after = 1;
    if (has_after) {
        /* deferred: fastpm_powerspectrum_init_from_delta folds the factor into its read; any other use of the buffer through
         * this library applies the sweep first (fpm_decic_defer, include/fastpm_b200.h) */
        if (getenv("FASTPM_B200_NO_LAZY_DECIC")) fastpm_apply_decic_transfer(pm, delta_k, delta_k);
        else FPM_MUST(fpm_decic_defer(pm->mesh, delta_k));
        fastpm_emit_event(fastpm->event_handlers, FASTPM_EVENT_FORCE, FASTPM_EVENT_STAGE_AFTER, (FastPMEvent *) event, fastpm);
        FPM_MUST(fpm_decic_cancel(delta_k));
    }
    LEAVE(event);
    pm_free(pm, delta_k);
}

static void do_kick(FastPMSolver *fastpm, FastPMTransition *trans)
{
    CLOCK(kick);
    LEAVE(kick);
    FastPMKickFactor kick;
    fastpm_kick_init(&kick, fastpm, trans->a.i, trans->a.r, trans->a.f);
    if (trans->end->v == trans->end->x) {
        FastPMDriftFactor drift;
        FastPMTransition dual[1];
        if (!fastpm_tevo_transition_find_dual(trans, dual))
            fastpm_raise(-1, "Dual transition not found. The state table is likely wrong. Look at states->table.\n");
        fastpm_drift_init(&drift, fastpm, dual->a.i, dual->a.r, dual->a.f);
        do_interpolation(fastpm, &drift, &kick, trans->a.i, trans->a.f, TIMESTEP_CUR);
    }
    ENTER(kick);
    for (int si = 0; si < FASTPM_SOLVER_NSPECIES; si++) {
        FastPMStore *p = fastpm_solver_get_species(fastpm, si);
        if (!p) continue;
        if (kick.ai != p->meta.a_v) fastpm_raise(-1, "kick is inconsitant with state.\n");
        if (kick.ac != p->meta.a_x) fastpm_raise(-1, "kick is inconsitant with state.\n");
        fastpm_kick_store(&kick, p, p, trans->a.f);
    }
    LEAVE(kick);
}

static void do_drift(FastPMSolver *fastpm, FastPMTransition *trans)
{
    CLOCK(drift);
    LEAVE(drift);
    FastPMDriftFactor drift;
    fastpm_drift_init(&drift, fastpm, trans->a.i, trans->a.r, trans->a.f);
    if (trans->end->v == trans->end->x) {
        FastPMKickFactor kick;
        FastPMTransition dual[1];
        if (!fastpm_tevo_transition_find_dual(trans, dual))
            fastpm_raise(-1, "Dual transition not found. The state table is likely wrong. Look at states->table.\n");
        fastpm_kick_init(&kick, fastpm, dual->a.i, dual->a.r, dual->a.f);
        do_interpolation(fastpm, &drift, &kick, trans->a.i, trans->a.f, TIMESTEP_CUR);
    }
    ENTER(drift);
    for (int si = 0; si < FASTPM_SOLVER_NSPECIES; si++) {
        FastPMStore *p = fastpm_solver_get_species(fastpm, si);
        if (!p) continue;
        if (drift.ai != p->meta.a_x) fastpm_raise(-1, "drift is inconsitant with state.\n");
        if (drift.ac != p->meta.a_v) fastpm_raise(-1, "drift is inconsitant with state.\n");
        fastpm_drift_store(&drift, p, p, trans->a.f);
    }
    LEAVE(drift);
}

static void do_interpolation(FastPMSolver *fastpm, FastPMDriftFactor *drift, FastPMKickFactor *kick, double a1, double a2, int whence)
{
    FastPMInterpolationEvent event[1];
    event->drift = drift; event->kick = kick; event->a1 = a1; event->a2 = a2; event->whence = whence;
    fastpm_emit_event(fastpm->event_handlers, FASTPM_EVENT_INTERPOLATION, FASTPM_EVENT_STAGE_BEFORE, (FastPMEvent *) event, fastpm);
}

void fastpm_solver_evolve(FastPMSolver *fastpm, double *time_step, int nstep)
{
    /* warm-up: clear acc (solver.c:378-391) */
    fpm_store_flush(NULL);
    for (int si = 0; si < FASTPM_SOLVER_NSPECIES; si++) {
        FastPMStore *p = fastpm_solver_get_species(fastpm, si);
        if (p) FPM_MUST(fpm_memset(p->acc, 0, sizeof(p->acc[0]) * p->np));
    }
    FastPMStates states[1];
    FastPMState templ[] = { {0, 0, 1}, {0, 1, 1}, {0, 2, 1}, {2, 2, 1}, {2, 2, 2}, {-1, -1, -1} };   /* K D D F K */
    fastpm_tevo_generate_states(states, nstep - 1, templ, time_step);
    FastPMTransition transition[1];
    for (int i = 1; states->table[i].force != -1; i++) {
        fastpm_tevo_transition_init(transition, states, i - 1, i);
        FastPMTransitionEvent event[1];
        event->transition = transition;
        fastpm_emit_event(fastpm->event_handlers, FASTPM_EVENT_TRANSITION, FASTPM_EVENT_STAGE_BEFORE, (FastPMEvent *) event, fastpm);
        switch (transition->action) {
            case FASTPM_ACTION_KICK: do_kick(fastpm, transition); break;
            case FASTPM_ACTION_DRIFT: do_drift(fastpm, transition); break;
            case FASTPM_ACTION_FORCE: do_force(fastpm, transition); break;
        }
        fastpm_emit_event(fastpm->event_handlers, FASTPM_EVENT_TRANSITION, FASTPM_EVENT_STAGE_AFTER, (FastPMEvent *) event, fastpm);
        if (i == 1) {
            /* the interpolation ranges are (,]: the initial step needs its own event (solver.c:333-343) */
            double a0 = time_step[0];
            FastPMKickFactor kick; FastPMDriftFactor drift;
            fastpm_kick_init(&kick, fastpm, a0, a0, a0);
            fastpm_drift_init(&drift, fastpm, a0, a0, a0);
            do_interpolation(fastpm, &drift, &kick, a0, a0, TIMESTEP_START);
        }
    }
    double a1 = time_step[nstep - 1];
    FastPMKickFactor kick; FastPMDriftFactor drift;
    fastpm_kick_init(&kick, fastpm, a1, a1, a1);
    fastpm_drift_init(&drift, fastpm, a1, a1, a1);
    do_interpolation(fastpm, &drift, &kick, a1, a1, TIMESTEP_END);
    fastpm_tevo_destroy_states(states);
    fpm_store_flush(NULL);                          /* the state is complete when evolve returns */
    if (fpm_wrap_check() != 0) fastpm_raise(-1, "%s\n", fpm_last_error());
}

/* The store whose periodic wrap has been postponed into the mass deposit that follows (one GPU, nobody looks at the
 * positions in between): fastpm_paint_local wraps while it reads x, see cic_paint_kernel<.., WRAP> in csrc/paint.cu. */
FastPMStore *fpm_pending_wrap = NULL;

/* A FORCE / TRANSITION / INTERPOLATION handler (or any caller between library calls) that reads p->x, p->v, event->delta_k ... with
 * its OWN device code must call this first: it applies the queued in-place kicks and drifts (factors.c) and the deferred CIC
 * deconvolution of event->delta_k (do_force), then waits for the device.  Handlers that only use this library's entry points need
 * not: every entry point does it for the buffers it is handed. */
void fastpm_b200_sync_state(void)
{
    fpm_store_flush(NULL);
    FPM_MUST(fpm_sync_deferred());
}

extern int fpm_decompose_skip_acc;       /* host/comm.c */
static void fastpm_decompose_inner(FastPMSolver *fastpm, PM *pm);
static void fastpm_decompose(FastPMSolver *fastpm, PM *pm)
{
    fpm_decompose_skip_acc = 1;          /* a force evaluation follows: ACC is recomputed for every particle */
    fastpm_decompose_inner(fastpm, pm);
    fpm_decompose_skip_acc = 0;
}
static void fastpm_decompose_inner(FastPMSolver *fastpm, PM *pm)
{
    if (fastpm->NTask > 1) fpm_store_flush(NULL);
    int before_handlers = 0, nspecies = 0;
    for (FastPMEventHandler *h = fastpm->event_handlers; h; h = h->next)
        if (h->stage == FASTPM_EVENT_STAGE_BEFORE && !strcmp(h->type, FASTPM_EVENT_FORCE)) before_handlers = 1;
    for (int si = 0; si < FASTPM_SOLVER_NSPECIES; si++) if (fastpm_solver_get_species(fastpm, si)) nspecies++;
    for (int si = 0; si < FASTPM_SOLVER_NSPECIES; si++) {
        FastPMStore *p = fastpm_solver_get_species(fastpm, si);
        if (!p) continue;
        if (nspecies == 1 && !getenv("FASTPM_B200_NO_FUSED_WRAP")) {
            if (fastpm->NTask == 1 && !before_handlers) {
                fpm_pending_wrap = p;                   /* one slab owns every particle: nothing to migrate; the deposit wraps */
                continue;
            }
            if (fastpm->NTask > 1) {
                fpm_pending_wrap = p;                   /* the classification pass of the migration wraps */
                if (0 != fastpm_store_decompose(p, (fastpm_store_target_func) FastPMTargetPM, pm, fastpm->comm))
                    fastpm_raise(-1, "Out of particle storage space\n");
                continue;
            }
        }
        fastpm_store_wrap(p, pm->BoxSize);
        if (0 != fastpm_store_decompose(p, (fastpm_store_target_func) FastPMTargetPM, pm, fastpm->comm))
            fastpm_raise(-1, "Out of particle storage space\n");
    }
}

/* ------------------------------------------------------------------ snapshots (solver.c:647-761)
 * po aliases every column of p (fastpm_store_steal, store.c:911-921): the drift/kick to aout and the unit
 * conversion happen in place and are reverted by fastpm_unset_species_snapshot. */
void fastpm_set_species_snapshot(FastPMSolver *fastpm, FastPMStore *p, FastPMDriftFactor *drift, FastPMKickFactor *kick,
                                 FastPMStore *po, double aout)
{
    fpm_store_flush(NULL);
    memcpy(po, p, sizeof(FastPMStore));
    if (drift) fastpm_drift_store(drift, p, po, aout);
    if (kick) fastpm_kick_store(kick, p, po, aout);
    /* a^2 dx/dt / H0 [Mpc/h]  ->  a dx/dt [km/s] */
    FPM_MUST(fpm_scale((const float *) po->v, (float *) po->v, 3 * po->np, HubbleConstant / aout));
    if (po->potential) {
        double potfactor = 1.5 * Omega_source(1, fastpm->cosmology) / (HubbleDistance * HubbleDistance);
        FPM_MUST(fpm_scale(po->potential, po->potential, po->np, potfactor / aout));
    }
    fastpm_store_wrap(po, fastpm->basepm->BoxSize);
}

void fastpm_unset_species_snapshot(FastPMSolver *fastpm, FastPMStore *p, FastPMDriftFactor *drift, FastPMKickFactor *kick,
                                   FastPMStore *po, double aout)
{
    fpm_store_flush(NULL);
    FPM_MUST(fpm_divide((const float *) po->v, (float *) po->v, 3 * po->np, HubbleConstant / aout));
    if (po->potential) {
        double potfactor = 1.5 * Omega_source(1, fastpm->cosmology) / (HubbleDistance * HubbleDistance);
        FPM_MUST(fpm_divide(po->potential, po->potential, po->np, potfactor / aout));
    }
    if (kick) fastpm_kick_store(kick, po, po, p->meta.a_v);
    if (drift) fastpm_drift_store(drift, po, po, p->meta.a_x);
    /* the two in-place updates above are queued under `po`, which is the caller's temporary: apply them now, before the wrap
     * below (reference order) and before `po` goes away */
    fpm_store_flush(NULL);
    /* columns are shared; count and meta come back from po like fastpm_store_steal does (store.c:911-921): the reverting kick /
     * drift above have put po's time stamps back to p's, and without them (restart, src/fastpm.c:625-633) po's are what was read */
    p->np = po->np;
    p->meta = po->meta;
    fastpm_store_wrap(p, fastpm->basepm->BoxSize);
    po->attributes = 0;
}

/* ------------------------------------------------------------------ construction from scalars (bindings) */
typedef struct { FastPMSolver solver; FastPMConfig config; FastPMCosmology cosmology; VPMInit vpminit[9]; } SolverBox;

FastPMSolver *fastpm_b200_solver_new(int64_t nc, double boxsize, const double *pm_nc_factor_pairs, int npairs,
                                     double alloc_factor, double lpt_nc_factor, int force_mode, int kernel_type,
                                     int growth_mode, int compute_potential, double nLPT,
                                     double Omega_m_, double h, double T_cmb, double N_eff, int N_nu)
{
    return fastpm_b200_solver_new_ex(nc, boxsize, pm_nc_factor_pairs, npairs, alloc_factor, lpt_nc_factor, force_mode, kernel_type,
                                     growth_mode, compute_potential, nLPT, Omega_m_, h, T_cmb, N_eff, N_nu, NULL, FASTPM_SOFTENING_NONE, FASTPM_PAINTER_CIC, 2);
}

/* options of FastPMConfig that the scalar constructors below have no argument for; they apply to the NEXT solver made, then reset */
static int g_next_use_shift = 0, g_next_use_dx1_only = 0;
void fastpm_b200_solver_next_options(int use_shift, int use_dx1_only) { g_next_use_shift = use_shift; g_next_use_dx1_only = use_dx1_only; }

FastPMSolver *fastpm_b200_solver_new_ex(int64_t nc, double boxsize, const double *pm_nc_factor_pairs, int npairs,
                                        double alloc_factor, double lpt_nc_factor, int force_mode, int kernel_type,
                                        int growth_mode, int compute_potential, double nLPT,
                                        double Omega_m_, double h, double T_cmb, double N_eff, int N_nu, const double *pgdc, int softening_type,
                                        int painter_type, int painter_support)
{
    libfastpm_init();
    SolverBox *b = calloc(1, sizeof(*b));
    if (npairs > 8) npairs = 8;
    for (int i = 0; i < npairs; i++) { b->vpminit[i].a_start = pm_nc_factor_pairs[2 * i]; b->vpminit[i].pm_nc_factor = pm_nc_factor_pairs[2 * i + 1]; }
    b->vpminit[npairs].a_start = 1; b->vpminit[npairs].pm_nc_factor = 0;
    FastPMCosmology *c = &b->cosmology;            /* src/prepare.c:20-39 with the Lua defaults */
    c->h = h; c->Omega_m = Omega_m_; c->T_cmb = T_cmb; c->Omega_k = 0; c->w0 = -1; c->wa = 0; c->N_eff = N_eff; c->N_nu = N_nu;
    c->N_ncdm = 0; c->ncdm_matterlike = 1; c->ncdm_freestreaming = 1; c->ncdm_linearresponse = 0; c->growth_mode = growth_mode;
    FastPMConfig *cfg = &b->config;                /* src/fastpm.c:186-217 */
    cfg->nc = nc; cfg->boxsize = boxsize; cfg->alloc_factor = alloc_factor; cfg->lpt_nc_factor = lpt_nc_factor;
    cfg->cosmology = c; cfg->vpminit = b->vpminit; cfg->USE_DX1_ONLY = g_next_use_dx1_only; cfg->USE_SHIFT = g_next_use_shift;
    g_next_use_dx1_only = g_next_use_shift = 0;
    cfg->ExtraAttributes = compute_potential ? COLUMN_POTENTIAL : 0;
    if (pgdc) {
        cfg->pgdc = 1; cfg->pgdc_alpha0 = pgdc[0]; cfg->pgdc_A = pgdc[1]; cfg->pgdc_B = pgdc[2]; cfg->pgdc_kl = pgdc[3]; cfg->pgdc_ks = pgdc[4];
        cfg->ExtraAttributes |= COLUMN_PGDC;
    }
    cfg->nLPT = nLPT; cfg->PAINTER_TYPE = (FastPMPainterType) painter_type; cfg->painter_support = painter_type == FASTPM_PAINTER_CIC ? 2 : painter_support;
    cfg->FORCE_TYPE = force_mode; cfg->KERNEL_TYPE = kernel_type; cfg->SOFTENING_TYPE = (FastPMSofteningType) softening_type;
    fastpm_solver_init(&b->solver, cfg, MPI_COMM_WORLD);
    return &b->solver;
}

void fastpm_b200_solver_free(FastPMSolver *solver)
{
    fastpm_solver_destroy(solver);
    free(solver);                                  /* the solver is the first member of its SolverBox */
}

/* small accessors for bindings that do not lay out the structs */
FastPMStore *fastpm_b200_solver_cdm(FastPMSolver *s) { return fastpm_solver_get_species(s, FASTPM_SPECIES_CDM); }
int64_t fastpm_b200_store_np(FastPMStore *p) { return (int64_t) p->np; }
/* resets the particle count (a host buffer with np particles is about to be copied in with fastpm_b200_store_set_column) */
int fastpm_b200_store_set_np(FastPMStore *p, int64_t np)
{
    fpm_store_flush(p);
    if (np < 0 || (size_t) np > p->np_upper) return -1;
    p->np = (size_t) np;
    return 0;
}
void fastpm_b200_store_meta(FastPMStore *p, double *out) { out[0] = p->meta.a_x; out[1] = p->meta.a_v; out[2] = p->meta.M0; }
void fastpm_b200_store_set_meta(FastPMStore *p, const double *in) { p->meta.a_x = in[0]; p->meta.a_v = in[1]; p->meta.M0 = in[2]; }
void *fastpm_b200_store_column_ptr(FastPMStore *p, FastPMColumnTags attribute)
{ fpm_store_flush(p); int ci = fastpm_store_find_column_id(p, attribute); return ci < 0 ? NULL : p->columns[ci]; }
PM *fastpm_b200_solver_lptpm(FastPMSolver *s) { return s->lptpm; }
void fastpm_b200_add_handler(FastPMSolver *s, const char *type, int stage, FastPMEventHandlerFunction fn, void *userdata)
{ fastpm_add_event_handler(&s->event_handlers, type, stage, fn, userdata); }
void fastpm_b200_kick_factor(FastPMSolver *s, double ai, double ac, double af, double *out)
{
    FastPMKickFactor k; fastpm_kick_init(&k, s, ai, ac, af);
    out[0] = k.ai; out[1] = k.ac; out[2] = k.af; out[3] = k.q1; out[4] = k.q2;
    memcpy(out + 5, k.dda, sizeof(double) * 32); memcpy(out + 37, k.Dv1, sizeof(double) * 32); memcpy(out + 69, k.Dv2, sizeof(double) * 32);
}
void fastpm_b200_drift_factor(FastPMSolver *s, double ai, double ac, double af, double *out)
{
    FastPMDriftFactor d; fastpm_drift_init(&d, s, ai, ac, af);
    out[0] = d.ai; out[1] = d.ac; out[2] = d.af; out[3] = d.Dv1; out[4] = d.Dv2;
    memcpy(out + 5, d.dyyy, sizeof(double) * 32); memcpy(out + 37, d.da1, sizeof(double) * 32); memcpy(out + 69, d.da2, sizeof(double) * 32);
}
void fastpm_b200_growth(FastPMSolver *s, double a, double *out)
{
    FastPMCosmology *c = s->cosmology;
    FastPMGrowthInfo gi; fastpm_growth_info_init(&gi, a, c);
    out[0] = gi.D1; out[1] = gi.D2; out[2] = gi.f1; out[3] = gi.f2;
    out[4] = HubbleEa(a, c); out[5] = DHubbleEaDa(a, c); out[6] = D2HubbleEaDa2(a, c);
    out[7] = DGrowthFactorDa(&gi); out[8] = D2GrowthFactorDa2(&gi);
    out[9] = Omega_source(a, c); out[10] = c->Omega_Lambda; out[11] = c->Omega_cdm;
}
int fastpm_b200_schedule(const double *time_step, int nstep, double *rows, int maxrows)
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
        r[0] = tr->action; r[1] = tr->a.i; r[2] = tr->a.f; r[3] = tr->a.r; r[4] = tr->end->x; r[5] = tr->end->v; r[6] = tr->end->force;
    }
    fastpm_tevo_destroy_states(states);
    free(ts);
    return n;
}

/* ------------------------------------------------------------------ host-only scalar entry points
 * (no device needed: used by the CPU test-suite to compare the factor tables with the oracle) */
static void host_only_solver(FastPMSolver *s, const double *cosmo, int growth_mode, int force_mode, double nLPT)
{
    memset(s, 0, sizeof(*s));
    FastPMCosmology *c = s->cosmology;
    c->Omega_m = cosmo[0]; c->h = cosmo[1]; c->T_cmb = cosmo[2]; c->N_eff = cosmo[3]; c->N_nu = (int) cosmo[4];
    c->Omega_k = 0; c->w0 = cosmo[5]; c->wa = cosmo[6];
    c->ncdm_matterlike = 1; c->ncdm_freestreaming = 1; c->growth_mode = growth_mode;
    fastpm_cosmology_init(c);
    s->config->FORCE_TYPE = force_mode;
    s->config->nLPT = nLPT;
    fastpm_set_msg_handler(fastpm_void_msg_handler, MPI_COMM_WORLD, NULL);
}
void fastpm_b200_host_kick_factor(const double *cosmo, int growth_mode, int force_mode, double nLPT, double ai, double ac, double af, double *out)
{ FastPMSolver s; host_only_solver(&s, cosmo, growth_mode, force_mode, nLPT); fastpm_b200_kick_factor(&s, ai, ac, af, out); }
void fastpm_b200_host_drift_factor(const double *cosmo, int growth_mode, int force_mode, double nLPT, double ai, double ac, double af, double *out)
{ FastPMSolver s; host_only_solver(&s, cosmo, growth_mode, force_mode, nLPT); fastpm_b200_drift_factor(&s, ai, ac, af, out); }
void fastpm_b200_host_growth(const double *cosmo, int growth_mode, double a, double *out)
{ FastPMSolver s; host_only_solver(&s, cosmo, growth_mode, 0, 0); fastpm_b200_growth(&s, a, out); }
/* the snapshot "Header" numbers (write_snapshot_header, io.c:250-290) without a device: host/io.c */
void fastpm_b200_host_io_header(const double *cosmo, int growth_mode, int64_t nc, double boxsize, double aout, double M0, uint64_t np_total, FpmIoHeader *h)
{
    FastPMSolver s;
    host_only_solver(&s, cosmo, growth_mode, 0, 0);
    s.config->nc = nc; s.config->boxsize = boxsize;
    fastpm_b200_io_header_values(&s, aout, M0, np_total, h);
}

/* ------------------------------------------------------------------ synthetic ICs for benchmark sizes
 * White noise from the counter-based device generator (instead of the serial RANLUX stream of
 * pmic_fill_gaussian_gadget, initialcondition.c:145-273), coloured by the tabulated linear P(k) exactly as
 * src/fastpm.c:415-545 does for a white-noise input (read_grafic path :448-462: real field * sqrt(Norm), r2c,
 * induce_correlation, DC mode = 1), then 2LPT at a0.  Everything stays on the device. */
void fastpm_b200_setup_synthetic_ic(FastPMSolver *fastpm, uint64_t seed, const double *k, const double *p, int size, double a0)
{
    PM *pm = fastpm->lptpm;
    FastPMFloat *delta_k = pm_alloc_noclear(pm, __FILE__, __LINE__);
    FastPMFloat *g_x = pm_alloc_noclear(pm, __FILE__, __LINE__);
    FPM_MUST(fpm_fill_whitenoise(pm->mesh, g_x, seed));
    fpm_mesh_r2c(pm, g_x, delta_k, sqrt(pm->Norm) / pm->Norm);
    pm_free(pm, g_x);
    FPM_MUST(fpm_induce_correlation(pm->mesh, delta_k, k, p, size));
    ptrdiff_t mode[4] = { 0, 0, 0, 0 };
    fastpm_apply_modify_mode_transfer(pm, delta_k, delta_k, mode, 1.0);
    fastpm_solver_setup_lpt(fastpm, FASTPM_SPECIES_CDM, delta_k, NULL, a0);
    pm_free(pm, delta_k);
}

/* The reference's own initial conditions, src/fastpm.c:415-545 for the seed + P(k)-table path (what oracle/ref_driver.c:ref_ic_deltak
 * restates): Gadget-scheme RANLUX white noise, optional fastpm_ic_remove_variance, colouring by the tabulated linear P(k), DC
 * mode = 1, then 2LPT at a0.  Same seed -> same particles as the reference (up to the last bit of the device's libm). */
void fastpm_b200_setup_gadget_ic(FastPMSolver *fastpm, int seed, int remove_variance, const double *k, const double *p, int size, double a0)
{
    PM *pm = fastpm->lptpm;
    FastPMFloat *delta_k = pm_alloc_noclear(pm, __FILE__, __LINE__);
    fastpm_ic_fill_gaussiank(pm, delta_k, seed, FASTPM_DELTAK_GADGET);
    if (remove_variance) fastpm_ic_remove_variance(pm, delta_k);
    FPM_MUST(fpm_induce_correlation(pm->mesh, delta_k, k, p, size));
    ptrdiff_t mode[4] = { 0, 0, 0, 0 };
    fastpm_apply_modify_mode_transfer(pm, delta_k, delta_k, mode, 1.0);
    fastpm_solver_setup_lpt(fastpm, FASTPM_SPECIES_CDM, delta_k, NULL, a0);
    pm_free(pm, delta_k);
}
