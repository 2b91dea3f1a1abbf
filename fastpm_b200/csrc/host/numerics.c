/* fastpm_b200 host layer -- the two numerical tools the integrator factors need, in place of GSL
 * (the reference calls gsl_integration_qag with GK41 at 1e-9 / GK61 at 1e-8 and gsl_odeiv2 rkf45 at
 * 1e-8: cosmology.c:276-298,321-372, factors.c:425-447).  Both are standard algorithms:
 *   fpm_integrate  globally adaptive bisection; on each panel a Gauss-Legendre rule of `order` points
 *                  is compared with the rule of 2*order+1 points; the panel with the largest estimated
 *                  error is split until the summed estimate meets max(epsabs, epsrel*|I|).
 *   fpm_ode_rkf45  Runge-Kutta-Fehlberg 4(5) with the usual step controller
 *                  tol_i = epsabs + epsrel*(|y_i| + |h y'_i|), shrink by 0.9 r^-1/5, grow by 0.9 r^-1/6.
 * Results agree with the reference's to the tolerances it requests (checked against the oracle). */
#include "internal.h"
#include <float.h>

#define MAXN 96
typedef struct { int n; double x[MAXN], w[MAXN]; } Rule;

static void legendre(int n, double x, double *p, double *dp)
{
    double p0 = 1, p1 = x;
    for (int k = 2; k <= n; k++) { double p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k; p0 = p1; p1 = p2; }
    *p = p1;
    *dp = n * (x * p1 - p0) / (x * x - 1);
}

static void rule_make(Rule *r, int n)
{
    r->n = n;
    for (int i = 0; i < n; i++) {
        double x = cos(M_PI * (i + 0.75) / (n + 0.5)), p, dp;
        for (int it = 0; it < 64; it++) {
            legendre(n, x, &p, &dp);
            double dx = p / dp;
            x -= dx;
            if (fabs(dx) < 4e-16) break;
        }
        legendre(n, x, &p, &dp);
        r->x[i] = x;
        r->w[i] = 2 / ((1 - x * x) * dp * dp);
    }
}

static double rule_apply(const Rule *r, fpm_func1 f, void *params, double a, double b)
{
    const double mid = 0.5 * (a + b), half = 0.5 * (b - a);
    double s = 0;
    for (int i = 0; i < r->n; i++) s += r->w[i] * f(mid + half * r->x[i], params);
    return s * half;
}

typedef struct { double a, b, val, err; } Panel;

double fpm_integrate(fpm_func1 f, void *params, double a, double b, double epsabs, double epsrel, int order)
{
    static Rule lo[4], hi[4];
    static int have[4];
    int slot = order <= 10 ? 0 : (order <= 15 ? 1 : (order <= 20 ? 2 : 3));
    const int orders[4] = { 10, 15, 20, 30 };
    if (!have[slot]) { rule_make(&lo[slot], orders[slot]); rule_make(&hi[slot], 2 * orders[slot] + 1); have[slot] = 1; }
    int cap = 32, n = 1;
    Panel *pn = malloc(sizeof(Panel) * cap);
    pn[0].a = a; pn[0].b = b;
    pn[0].val = rule_apply(&hi[slot], f, params, a, b);
    pn[0].err = fabs(pn[0].val - rule_apply(&lo[slot], f, params, a, b));
    double total = pn[0].val, err = pn[0].err;
    while (n < 100000) {
        double tol = fmax(epsabs, epsrel * fabs(total));
        if (err <= tol || err <= 64 * DBL_EPSILON * fabs(total)) break;
        int worst = 0;
        for (int i = 1; i < n; i++) if (pn[i].err > pn[worst].err) worst = i;
        if (n == cap) { cap *= 2; pn = realloc(pn, sizeof(Panel) * cap); }
        double mid = 0.5 * (pn[worst].a + pn[worst].b);
        Panel l = { pn[worst].a, mid, 0, 0 }, r = { mid, pn[worst].b, 0, 0 };
        l.val = rule_apply(&hi[slot], f, params, l.a, l.b); l.err = fabs(l.val - rule_apply(&lo[slot], f, params, l.a, l.b));
        r.val = rule_apply(&hi[slot], f, params, r.a, r.b); r.err = fabs(r.val - rule_apply(&lo[slot], f, params, r.a, r.b));
        pn[worst] = l; pn[n++] = r;
        total = 0; err = 0;
        for (int i = 0; i < n; i++) { total += pn[i].val; err += pn[i].err; }
    }
    free(pn);
    return total;
}

#define ODE_MAXDIM 8
int fpm_ode_rkf45(fpm_odefunc f, void *params, int dim, double *t, double t1, double *y, double h0, double epsabs, double epsrel)
{
    /* Fehlberg tableau */
    static const double c2 = 1. / 4, c3 = 3. / 8, c4 = 12. / 13, c6 = 1. / 2;
    static const double a21 = 1. / 4;
    static const double a31 = 3. / 32, a32 = 9. / 32;
    static const double a41 = 1932. / 2197, a42 = -7200. / 2197, a43 = 7296. / 2197;
    static const double a51 = 439. / 216, a52 = -8., a53 = 3680. / 513, a54 = -845. / 4104;
    static const double a61 = -8. / 27, a62 = 2., a63 = -3544. / 2565, a64 = 1859. / 4104, a65 = -11. / 40;
    static const double b1 = 16. / 135, b3 = 6656. / 12825, b4 = 28561. / 56430, b5 = -9. / 50, b6 = 2. / 55;
    static const double e1 = 1. / 360, e3 = -128. / 4275, e4 = -2197. / 75240, e5 = 1. / 50, e6 = 2. / 55;
    if (dim > ODE_MAXDIM) return -1;
    if (t1 < *t) return -1;                 /* the reference's driver refuses to run against the sign of h0 */
    double k1[ODE_MAXDIM], k2[ODE_MAXDIM], k3[ODE_MAXDIM], k4[ODE_MAXDIM], k5[ODE_MAXDIM], k6[ODE_MAXDIM];
    double yt[ODE_MAXDIM], yn[ODE_MAXDIM], ye[ODE_MAXDIM], dn[ODE_MAXDIM];
    double h = h0;
    for (long iter = 0; *t < t1; iter++) {
        if (iter > 5000000) return -1;
        double hs = h;
        int last = 0;
        if (*t + hs >= t1) { hs = t1 - *t; last = 1; }
        f(*t, y, k1, params);
        for (int i = 0; i < dim; i++) yt[i] = y[i] + hs * a21 * k1[i];
        f(*t + c2 * hs, yt, k2, params);
        for (int i = 0; i < dim; i++) yt[i] = y[i] + hs * (a31 * k1[i] + a32 * k2[i]);
        f(*t + c3 * hs, yt, k3, params);
        for (int i = 0; i < dim; i++) yt[i] = y[i] + hs * (a41 * k1[i] + a42 * k2[i] + a43 * k3[i]);
        f(*t + c4 * hs, yt, k4, params);
        for (int i = 0; i < dim; i++) yt[i] = y[i] + hs * (a51 * k1[i] + a52 * k2[i] + a53 * k3[i] + a54 * k4[i]);
        f(*t + hs, yt, k5, params);
        for (int i = 0; i < dim; i++) yt[i] = y[i] + hs * (a61 * k1[i] + a62 * k2[i] + a63 * k3[i] + a64 * k4[i] + a65 * k5[i]);
        f(*t + c6 * hs, yt, k6, params);
        for (int i = 0; i < dim; i++) {
            yn[i] = y[i] + hs * (b1 * k1[i] + b3 * k3[i] + b4 * k4[i] + b5 * k5[i] + b6 * k6[i]);
            ye[i] = hs * (e1 * k1[i] + e3 * k3[i] + e4 * k4[i] + e5 * k5[i] + e6 * k6[i]);
        }
        f(*t + hs, yn, dn, params);
        double r = DBL_MIN;
        for (int i = 0; i < dim; i++) {
            double tol = epsabs + epsrel * (fabs(yn[i]) + fabs(hs * dn[i]));
            double ri = fabs(ye[i]) / tol;
            if (ri > r) r = ri;
        }
        if (r > 1.1) {                       /* reject, shrink */
            double s = 0.9 / pow(r, 1.0 / 5);
            h = hs * (s < 0.2 ? 0.2 : s);
            continue;
        }
        *t = last ? t1 : *t + hs;
        memcpy(y, yn, sizeof(double) * dim);
        if (r < 0.5) {
            double s = 0.9 / pow(r, 1.0 / 6);
            if (s > 5) s = 5;
            if (s < 1) s = 1;
            h = hs * s;
        } else {
            h = hs;
        }
    }
    return 0;
}
