/* fastpm_b200 host layer -- kick and drift factors (reference: libfastpm/factors.c).
 *
 * Host part: 32-sample tables between a_i and a_f of
 *   dda  = -1.5 Omega(a_c) a_c E(a_c) (G_f(a_e) - G_f(a_i)) / g_f(a_c)                 (FastPM, factors.c:292-294)
 *        = -1.5 Omega_m0 * Sphi(a_i, a_e, a_c)                                           (PM / COLA, :296-298)
 *   dyyy = (G_p(a_e) - G_p(a_i)) / (a_c^3 E(a_c) g_p(a_c))                               (FastPM, :352-354)
 *        = Sq(a_i, a_e, a_c)                                                             (PM / COLA, :356)
 * plus the LPT terms Dv1, Dv2, da1, da2, q1, q2.  Device part: one streaming pass per kick or drift
 * (fpm_kick / fpm_drift) with the interpolated factor differences passed as scalars.
 */
#include "internal.h"

/* G_p = D1, g_p = dD1/da, G_f = a^3 E g_p, g_f = dG_f/da   (factors.c:199-231) */
static double Gf(FastPMGrowthInfo *gi)
{
    double a = gi->a;
    return a * a * a * HubbleEa(a, gi->c) * DGrowthFactorDa(gi);
}
static double gf(FastPMGrowthInfo *gi)
{
    double a = gi->a;
    FastPMCosmology *c = gi->c;
    double E = HubbleEa(a, c), dEda = DHubbleEaDa(a, c), dDda = DGrowthFactorDa(gi), d2Dda2 = D2GrowthFactorDa2(gi);
    return 3 * a * a * E * dDda + a * a * a * dEda * dDda + a * a * a * E * d2Dda2;
}

/* modified (COLA) and standard time-stepping integrals, factors.c:395-506 */
typedef struct { FastPMCosmology *c; double nLPT; int kind; } StepInt;
static double step_integrand(double a, void *params)
{
    StepInt *s = params;
    double E = HubbleEa(a, s->c);
    switch (s->kind) {
        case 0: return 1 / (pow(a, 3) * E);                          /* standard drift */
        case 1: return pow(a, s->nLPT) / (pow(a, 3) * E);            /* non-standard drift */
        default: return 1 / (pow(a, 2) * E);                         /* standard kick */
    }
}
static double step_integral(double ai, double af, FastPMCosmology *c, double nLPT, int kind)
{
    StepInt s = { c, nLPT, kind };
    return fpm_integrate(step_integrand, &s, ai, af, 0, 1e-8, 30);
}
static double Sq(double ai, double af, double aRef, double nLPT, FastPMCosmology *c, int nonstd)
{
    if (nonstd) return step_integral(ai, af, c, nLPT, 1) / pow(aRef, nLPT);
    return step_integral(ai, af, c, nLPT, 0);
}
static double Sphi(double ai, double af, double aRef, double nLPT, FastPMCosmology *c, int nonstd)
{
    if (nonstd)
        return (pow(af, nLPT) - pow(ai, nLPT)) * aRef / (pow(aRef, 3) * HubbleEa(aRef, c) * (nLPT * pow(aRef, nLPT - 1)));
    return step_integral(ai, af, c, nLPT, 2);
}

void fastpm_kick_init(FastPMKickFactor *kick, FastPMSolver *fastpm, double ai, double ac, double af)
{
    FastPMCosmology *c = fastpm->cosmology;
    kick->forcemode = fastpm->config->FORCE_TYPE;
    FastPMGrowthInfo gi_i, gi_c, gi_e;
    fastpm_growth_info_init(&gi_i, ai, c);
    fastpm_growth_info_init(&gi_c, ac, c);
    const double E_i = HubbleEa(ai, c), E_c = HubbleEa(ac, c);
    const double Omega_m0 = Omega_source(1, c), Omega_mc = Omega_source(ac, c);

    kick->q1 = gi_c.D1;
    if (c->growth_mode == FASTPM_GROWTH_MODE_LCDM) kick->q2 = gi_c.D1 * gi_c.D1 * (1.0 + 7.0 / 3.0 * pow(Omega_mc, 1.0 / 143.0));
    else if (c->growth_mode == FASTPM_GROWTH_MODE_ODE) kick->q2 = gi_c.D1 * gi_c.D1 * (1 - gi_c.D1 * gi_c.D1 / gi_c.D2);
    else fastpm_raise(-1, "Please enter a valid growth mode.\n");

    kick->nsamples = 32;
    const double Dv1i = gi_i.D1 * ai * ai * E_i * gi_i.f1, Dv2i = gi_i.D2 * ai * ai * E_i * gi_i.f2;
    const double Gf_i = Gf(&gi_i), gf_c = gf(&gi_c);
    for (int i = 0; i < kick->nsamples; i++) {
        const double ae = ai * (1.0 * (kick->nsamples - 1 - i) / (kick->nsamples - 1)) + af * (1.0 * i / (kick->nsamples - 1));
        fastpm_growth_info_init(&gi_e, ae, c);
        const double E_e = HubbleEa(ae, c);
        if (kick->forcemode == FASTPM_FORCE_FASTPM)
            kick->dda[i] = -1.5 * Omega_mc * ac * E_c * (Gf(&gi_e) - Gf_i) / gf_c;
        else
            kick->dda[i] = -1.5 * Omega_m0 * Sphi(ai, ae, ac, fastpm->config->nLPT, c, kick->forcemode == FASTPM_FORCE_COLA);
        kick->Dv1[i] = gi_e.D1 * ae * ae * E_e * gi_e.f1 - Dv1i;
        kick->Dv2[i] = gi_e.D2 * ae * ae * E_e * gi_e.f2 - Dv2i;
    }
    kick->ai = ai; kick->ac = ac; kick->af = af;
    fastpm_info("Growth/FastPM factors at a = %6.4f: D1=%g, D2=%g, f1=%g, f2=%g, G_p=%g, G_f=%g, g_p=%g, g_f=%g\n",
                ai, gi_i.D1, gi_i.D2, gi_i.f1, gi_i.f2, gi_i.D1, Gf_i, DGrowthFactorDa(&gi_i), gf(&gi_i));
}

void fastpm_drift_init(FastPMDriftFactor *drift, FastPMSolver *fastpm, double ai, double ac, double af)
{
    FastPMCosmology *c = fastpm->cosmology;
    drift->forcemode = fastpm->config->FORCE_TYPE;
    FastPMGrowthInfo gi_i, gi_c, gi_e;
    fastpm_growth_info_init(&gi_i, ai, c);
    fastpm_growth_info_init(&gi_c, ac, c);
    const double E_c = HubbleEa(ac, c), gp_c = DGrowthFactorDa(&gi_c);
    drift->nsamples = 32;
    for (int i = 0; i < drift->nsamples; i++) {
        const double ae = ai * (1.0 * (drift->nsamples - 1 - i) / (drift->nsamples - 1)) + af * (1.0 * i / (drift->nsamples - 1));
        fastpm_growth_info_init(&gi_e, ae, c);
        if (drift->forcemode == FASTPM_FORCE_FASTPM)
            drift->dyyy[i] = 1 / (ac * ac * ac * E_c) * (gi_e.D1 - gi_i.D1) / gp_c;
        else
            drift->dyyy[i] = Sq(ai, ae, ac, fastpm->config->nLPT, c, drift->forcemode == FASTPM_FORCE_COLA);
        drift->da1[i] = gi_e.D1 - gi_i.D1;
        drift->da2[i] = gi_e.D2 - gi_i.D2;
    }
    drift->af = af; drift->ai = ai; drift->ac = ac;
    drift->Dv1 = gi_c.D1 * ac * ac * E_c * gi_c.f1;
    drift->Dv2 = gi_c.D2 * ac * ac * E_c * gi_c.f2;
}

/* table look-up with the exact end points special-cased (factors.c:40-71, 117-146) */
static void lookup3(const double *t0, const double *t1, const double *t2, int n, double ai, double af_tab, double a,
                    double *o0, double *o1, double *o2, const char *what)
{
    if (a == af_tab) { *o0 = t0[n - 1]; *o1 = t1[n - 1]; *o2 = t2[n - 1]; return; }
    if (a == ai) { *o0 = t0[0]; *o1 = t1[0]; *o2 = t2[0]; return; }
    const double ind = (a - ai) / (af_tab - ai) * (n - 1);
    const int l = (int) floor(ind);
    const double u = l + 1 - ind, v = ind - l;
    if (l + 1 >= n) fastpm_raise(-1, "%s beyond factor's available range. ", what);
    *o0 = t0[l] * u + t0[l + 1] * v;
    *o1 = t1[l] * u + t1[l + 1] * v;
    *o2 = t2[l] * u + t2[l + 1] * v;
}

void fpm_kick_factors_at(FastPMKickFactor *kick, double a_v, double af, double *dda, double *Dv1, double *Dv2)
{
    double f[3], i[3];
    lookup3(kick->dda, kick->Dv1, kick->Dv2, kick->nsamples, kick->ai, kick->af, af, &f[0], &f[1], &f[2], "kick");
    lookup3(kick->dda, kick->Dv1, kick->Dv2, kick->nsamples, kick->ai, kick->af, a_v, &i[0], &i[1], &i[2], "kick");
    *dda = f[0] - i[0]; *Dv1 = f[1] - i[1]; *Dv2 = f[2] - i[2];
}

void fpm_drift_factors_at(FastPMDriftFactor *drift, double a_x, double af, double *dyyy, double *da1, double *da2)
{
    double f[3], i[3];
    lookup3(drift->dyyy, drift->da1, drift->da2, drift->nsamples, drift->ai, drift->af, af, &f[0], &f[1], &f[2], "drift");
    lookup3(drift->dyyy, drift->da1, drift->da2, drift->nsamples, drift->ai, drift->af, a_x, &i[0], &i[1], &i[2], "drift");
    *dyyy = f[0] - i[0]; *da1 = f[1] - i[1]; *da2 = f[2] - i[2];
}

/* ------------------------------------------------------------------ deferred in-place updates
 * fastpm_solver_evolve issues K K D D between two force evaluations (solver.c:283-356).  When a kick or drift is in place
 * (pi == po) it is queued here instead of launched; the queue is applied by ONE kernel (fpm_update_fused: same operations,
 * order and roundings, v and x in registers in between) as soon as anything else touches the store -- every function of this
 * layer that reads or writes store columns calls fpm_store_flush first, and fastpm_emit_event flushes before it calls a
 * handler.  The time stamps meta.a_x / meta.a_v are updated at once, like the reference's. */
static struct { FastPMStore *p; int n; double ops[8][7]; } pending_updates;

void fpm_store_flush(FastPMStore *p)
{
    if (pending_updates.n == 0 || (p != NULL && p != pending_updates.p)) return;
    FastPMStore *q = pending_updates.p;
    const int n = pending_updates.n;
    pending_updates.n = 0; pending_updates.p = NULL;
    FPM_MUST(fpm_update_fused((double *) q->x, (float *) q->v, (const float *) q->acc, (const float *) q->dx1, (const float *) q->dx2,
                              (int64_t) q->np, n, &pending_updates.ops[0][0]));
}

static int defer_update(FastPMStore *p, int kind, int mode, double f0, double f1, double f2, double f3, double f4)
{
    static int enabled = -1;
    if (enabled < 0) enabled = getenv("FASTPM_B200_NO_FUSED_UPDATE") ? 0 : 1;
    if (!enabled) return 0;
    if (pending_updates.n && pending_updates.p != p) fpm_store_flush(NULL);
    if (pending_updates.n == 8) fpm_store_flush(NULL);
    double *o = pending_updates.ops[pending_updates.n++];
    pending_updates.p = p;
    o[0] = kind; o[1] = mode; o[2] = f0; o[3] = f1; o[4] = f2; o[5] = f3; o[6] = f4;
    return 1;
}

/* fastpm_kick_store / fastpm_drift_store (factors.c:176-197, 374-392): pi and po may be the same store or two
 * stores with separate v / x columns (snapshots, solver.c:647-702) */
void fastpm_kick_store(FastPMKickFactor *kick, FastPMStore *pi, FastPMStore *po, double af)
{
    double dda, Dv1, Dv2;
    fpm_kick_factors_at(kick, pi->meta.a_v, af, &dda, &Dv1, &Dv2);
    const int cola = kick->forcemode == FASTPM_FORCE_COLA;
    if (cola && (!pi->dx1 || !pi->dx2)) fastpm_raise(-1, "COLA kick needs the dx1 and dx2 columns (solver.c:84-88)\n");
    if (pi == po && pi->v == po->v && defer_update(pi, 0, cola, dda, kick->q1, kick->q2, Dv1, Dv2)) { po->meta.a_v = af; return; }
    fpm_store_flush(NULL);
    FPM_MUST(fpm_kick((float *) po->v, (const float *) pi->v, (const float *) pi->acc, (const float *) pi->dx1, (const float *) pi->dx2,
                      (int64_t) pi->np, (int) kick->forcemode, dda, kick->q1, kick->q2, Dv1, Dv2));
    po->meta.a_v = af;
}

void fastpm_drift_store(FastPMDriftFactor *drift, FastPMStore *pi, FastPMStore *po, double af)
{
    double dyyy, da1, da2;
    fpm_drift_factors_at(drift, pi->meta.a_x, af, &dyyy, &da1, &da2);
    const int mode = (int) drift->forcemode;
    if (mode >= 2 && (!pi->dx1 || (mode != 4 && !pi->dx2))) fastpm_raise(-1, "drift mode %d needs the dx1/dx2 columns\n", mode);
    /* with a PGD column the drift is followed at once by the PGD displacement (not queued) */
    if (!pi->pgdc && pi == po && pi->x == po->x && defer_update(pi, 1, mode, dyyy, da1, da2, drift->Dv1, drift->Dv2)) { po->meta.a_x = af; return; }
    fpm_store_flush(NULL);
    FPM_MUST(fpm_drift((double *) po->x, (const double *) pi->x, (const float *) pi->v, (const float *) pi->dx1, (const float *) pi->dx2,
                       (int64_t) pi->np, (int) drift->forcemode, dyyy, da1, da2, drift->Dv1, drift->Dv2));
    /* factors.c:108-113: xo += 0.5 * pgdc * dyyy / dyyy[last], except for an empty interval ("no drift; to protect the pgdc line") */
    if (pi->pgdc && drift->ai != drift->af)
        FPM_MUST(fpm_pgd_shift((double *) po->x, (const float *) pi->pgdc, (int64_t) pi->np, dyyy, drift->dyyy[drift->nsamples - 1]));
    po->meta.a_x = af;
}
