/* fastpm_b200 host layer -- background cosmology and growth factors (scalar inputs of the kick/drift
 * factors).  Same model as the reference's libfastpm/cosmology.c: E(a)^2 = O_r a^-4 + O_cdm a^-3 +
 * O_k a^-2 + O_DE(a) [+ matter-like ncdm], CPL dark energy, radiation from T_cmb and N_eff; growth
 * either from the LCDM integral with fitting formulae for D2, f1, f2 (cosmology.c:374-391) or from the
 * 4-variable growth ODE started in matter domination at a = 0.00625 (cosmology.c:301-372).
 * Massive-neutrino (ncdm) species with Fermi-Dirac tables are out of scope. */
#include "internal.h"

double HubbleDistance = 2997.92458;   /* Mpc/h,  cosmology.c:18 */
double HubbleConstant = 100.0;        /* km/s/(Mpc/h), cosmology.c:19 */

#define STEF_BOLT 2.85087e-48         /* h (1e10 Msun/h) s^-3 K^-4 */
#define RHO_CRIT 27.7455
#define LIGHT 9.715614e-15            /* h (Mpc/h) s^-1 */

void fastpm_cosmology_init(FastPMCosmology *c)
{
    if (c->N_ncdm > 0 && !c->ncdm_matterlike)
        fastpm_raise(-1, "fastpm_b200: Fermi-Dirac (non matter-like) ncdm backgrounds are out of scope of this build.\n");
    double O_ncdm = 0;
    if (c->ncdm_matterlike) {
        for (int i = 0; i < c->N_ncdm; i++) O_ncdm += c->m_ncdm[i];
        O_ncdm *= 1. / 93.14 / c->h / c->h;
    }
    c->Omega_ncdm = O_ncdm;
    c->Omega_cdm = c->Omega_m - O_ncdm;
    c->Omega_Lambda = 1 - c->Omega_m - Omega_r(c) - c->Omega_k;      /* close the universe, cosmology.c:49 */
    c->FDinterp = NULL;
}
void fastpm_cosmology_destroy(FastPMCosmology *c) { (void) c; }

double Omega_g(FastPMCosmology *c) { return 4 * STEF_BOLT * pow(c->T_cmb, 4) / pow(LIGHT, 3) / RHO_CRIT / pow(c->h, 2); }
double Gamma_nu(FastPMCosmology *c)
{
    if (c->N_nu == 0) return 0;
    return pow(4. / 11., 1. / 3.) * pow(c->N_eff / c->N_nu, 1. / 4.);
}
double Omega_ur(FastPMCosmology *c)
{
    int N_ur = c->N_nu - c->N_ncdm;
    return 7. / 8. * N_ur * pow(Gamma_nu(c), 4) * Omega_g(c);
}
double Omega_r(FastPMCosmology *c) { return Omega_g(c) + Omega_ur(c); }

double Omega_DE_TimesHubbleEaSq(double a, FastPMCosmology *c)
{
    double exponent = (a - 1) * c->wa - (1 + c->w0 + c->wa) * log(a);
    return c->Omega_Lambda * exp(3 * exponent);
}
double DOmega_DE_TimesHubbleEaSqDa(double a, FastPMCosmology *c)
{ return 3 * (c->wa - (1 + c->w0 + c->wa) / a) * Omega_DE_TimesHubbleEaSq(a, c); }
double D2Omega_DE_TimesHubbleEaSqDa2(double a, FastPMCosmology *c)
{
    double ode = Omega_DE_TimesHubbleEaSq(a, c), d = DOmega_DE_TimesHubbleEaSqDa(a, c);
    return d * d / c->Omega_Lambda + 3 * (1 + c->w0 + c->wa) / (a * a) * ode;
}

double HubbleEa(double a, FastPMCosmology *c)
{
    double ncdm = c->Omega_ncdm / (a * a * a);
    return sqrt(Omega_r(c) / (a * a * a * a) + c->Omega_cdm / (a * a * a) + c->Omega_k / (a * a)
                + Omega_DE_TimesHubbleEaSq(a, c) + ncdm);
}
double Omega_cdm_a(double a, FastPMCosmology *c) { double E = HubbleEa(a, c); return c->Omega_cdm / (a * a * a) / (E * E); }
double Omega_m(double a, FastPMCosmology *c) { double E = HubbleEa(a, c); return c->Omega_m / (a * a * a) / (E * E); }
double Omega_source(double a, FastPMCosmology *c) { return c->ncdm_freestreaming ? Omega_cdm_a(a, c) : Omega_m(a, c); }

double DHubbleEaDa(double a, FastPMCosmology *c)
{
    double E = HubbleEa(a, c);
    return 0.5 / E * (-4 * Omega_r(c) / pow(a, 5) - 3 * c->Omega_cdm / pow(a, 4) - 2 * c->Omega_k / pow(a, 3)
                      + DOmega_DE_TimesHubbleEaSqDa(a, c) - 3 * c->Omega_ncdm / pow(a, 4));
}
double D2HubbleEaDa2(double a, FastPMCosmology *c)
{
    double E = HubbleEa(a, c), dEda = DHubbleEaDa(a, c);
    return 0.5 / E * (20 * Omega_r(c) / pow(a, 6) + 12 * c->Omega_cdm / pow(a, 5) + 6 * c->Omega_k / pow(a, 4)
                      + D2Omega_DE_TimesHubbleEaSqDa2(a, c) + 12 * c->Omega_ncdm / pow(a, 5) - 2 * pow(dEda, 2));
}

/* LCDM growth integral, cosmology.c:268-298 */
static double growth_integrand(double a, void *param)
{
    double *p = param;
    return pow(a / (p[0] + (1 - p[0] - p[1]) * a + p[1] * a * a * a), 1.5);
}
static double growth_lcdm(double a, FastPMCosmology *c)
{
    double p[2] = { c->Omega_m, c->Omega_Lambda };
    static FastPMCosmology memo_c; static double memo_v; static int memo_ok = 0;
    if (a == 1.0 && memo_ok && !memcmp(&memo_c, c, sizeof(*c))) return memo_v;
    double v = HubbleEa(a, c) * fpm_integrate(growth_integrand, p, 0, a, 0, 1e-9, 20);
    if (a == 1.0) { memo_c = *c; memo_v = v; memo_ok = 1; }
    return v;
}

/* growth ODE in ln a, cosmology.c:301-319 */
static void growth_rhs(double a, const double *y, double *dyda, void *params)
{
    FastPMCosmology *c = params;
    const double E = HubbleEa(a, c), dEda = DHubbleEaDa(a, c), src = 1.5 * Omega_source(a, c);
    double d[4];
    d[0] = y[1];
    d[1] = -(2. + a / E * dEda) * y[1] + src * y[0];
    d[2] = y[3];
    d[3] = -(2. + a / E * dEda) * y[3] + src * (y[2] - y[0] * y[0]);
    for (int i = 0; i < 4; i++) dyda[i] = d[i] / a;
}
static int growth_ode_raw(double a, FastPMCosmology *c, double y[4]);
/* the a = 1 normalisation is needed by every sample of every factor table: remember the last solve */
static int growth_ode(double a, FastPMCosmology *c, double y[4])
{
    static FastPMCosmology memo_c; static double memo_y[4]; static int memo_ok = 0;
    if (a == 1.0) {
        if (memo_ok && !memcmp(&memo_c, c, sizeof(*c))) { memcpy(y, memo_y, sizeof(memo_y)); return 0; }
        int rc = growth_ode_raw(a, c, y);
        if (rc == 0) { memo_c = *c; memcpy(memo_y, y, sizeof(memo_y)); memo_ok = 1; }
        return rc;
    }
    return growth_ode_raw(a, c, y);
}
static int growth_ode_raw(double a, FastPMCosmology *c, double y[4])
{
    double t = 0.00625;                       /* matter-dominated start, z = 159 */
    y[0] = t; y[1] = t; y[2] = -3. / 7. * t * t; y[3] = 2 * y[2];
    if (fpm_ode_rkf45(growth_rhs, c, 4, &t, a, y, 1e-6, 1e-8, 1e-8) != 0) {
        if (a >= 0.00625) fastpm_raise(-1, "Growth ODE unsuccesful at a=%g.", a);
        y[0] = y[1] = y[2] = y[3] = 0;
        return -1;
    }
    return 0;
}

void fastpm_growth_info_init(FastPMGrowthInfo *gi, double a, FastPMCosmology *c)
{
    gi->a = a; gi->c = c;
    if (c->growth_mode == FASTPM_GROWTH_MODE_LCDM) {
        double d1 = growth_lcdm(a, c), d1_1 = growth_lcdm(1, c), Om = Omega_m(a, c);
        gi->D1 = d1 / d1_1;
        gi->f1 = pow(Om, 5. / 9.);
        gi->D2 = gi->D1 * gi->D1 * pow(Om / Omega_m(1, c), -1. / 143.);
        gi->f2 = 2 * pow(Om, 6. / 11.);
    } else if (c->growth_mode == FASTPM_GROWTH_MODE_ODE) {
        double y[4], y1[4];
        growth_ode(a, c, y);
        growth_ode(1, c, y1);
        gi->D1 = y[0] / y1[0];
        gi->f1 = y[1] / y[0];
        gi->D2 = y[2] / y1[2];
        gi->f2 = y[3] / y[2];
    } else {
        fastpm_raise(-1, "Please enter a valid growth mode.\n");
    }
}

double DGrowthFactorDa(FastPMGrowthInfo *gi)
{
    double a = gi->a;
    FastPMCosmology *c = gi->c;
    if (c->growth_mode == FASTPM_GROWTH_MODE_LCDM) {
        double E = HubbleEa(a, c), EI = growth_lcdm(1.0, c);
        return DHubbleEaDa(a, c) * gi->D1 / E + E * pow(a * E, -3) / EI;
    }
    return gi->f1 * gi->D1 / a;
}

double D2GrowthFactorDa2(FastPMGrowthInfo *gi)
{
    double a = gi->a;
    FastPMCosmology *c = gi->c;
    double E = HubbleEa(a, c), dEda = DHubbleEaDa(a, c);
    if (c->growth_mode == FASTPM_GROWTH_MODE_LCDM) {
        double EI = growth_lcdm(1., c);
        return D2HubbleEaDa2(a, c) * gi->D1 / E - (dEda + 3 / a * E) * pow(a * E, -3) / EI;
    }
    double ans = -(3. + a / E * dEda) * gi->f1 + 1.5 * Omega_source(a, c);
    return ans * gi->D1 / (a * a);
}

static double comoving_integrand(double a, void *params) { FastPMCosmology *c = params; return 1. / (a * a * HubbleEa(a, c)); }
double ComovingDistance(double a, FastPMCosmology *c) { return fpm_integrate(comoving_integrand, c, a, 1., 0, 1e-8, 20); }
