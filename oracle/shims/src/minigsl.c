/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * mini-GSL: the handful of GSL entry points the reference's hot path uses
 * (SURVEY.md section 8c lists the 18 symbols), restated from their published
 * algorithms because GSL is not installed in this image:
 *   - gsl_integration_qag      cosmology.c:292,485  factors.c:439  powerspectrum.c:271
 *   - gsl_odeiv2 (rkf45)       cosmology.c:330-348
 *   - gsl_rng_ranlxd1          initialcondition.c:153-263  store.c:697-718  utils.c:19-28
 *   - gsl_interp_linear        gravity.c:515-517  FDinterp.c
 */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <float.h>
#include <gsl/gsl_math.h>
#include <gsl/gsl_errno.h>
#include <gsl/gsl_integration.h>
#include <gsl/gsl_odeiv2.h>
#include <gsl/gsl_rng.h>
#include <gsl/gsl_interp.h>

/* ------------------------------------------------------------------ errno */
static gsl_error_handler_t *current_handler = NULL;
gsl_error_handler_t *gsl_set_error_handler(gsl_error_handler_t *h)
{ gsl_error_handler_t *old = current_handler; current_handler = h; return old; }
gsl_error_handler_t *gsl_set_error_handler_off(void)
{ gsl_error_handler_t *old = current_handler; current_handler = NULL; return old; }

/* ------------------------------------------------------------ integration */
#define GL_MAXN 64
typedef struct { int n; double x[GL_MAXN], w[GL_MAXN]; } gl_rule;

static void gl_make(gl_rule *r, int n)
{
    /* Gauss-Legendre nodes on [-1,1] by Newton iteration on P_n */
    r->n = n;
    for (int i = 0; i < n; i++) {
        double x = cos(M_PI * (i + 0.75) / (n + 0.5));
        double pp = 0;
        for (int it = 0; it < 100; it++) {
            double p0 = 1, p1 = x;
            for (int k = 2; k <= n; k++) {
                double p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
                p0 = p1; p1 = p2;
            }
            pp = n * (x * p1 - p0) / (x * x - 1);
            double dx = p1 / pp;
            x -= dx;
            if (fabs(dx) < 1e-16) break;
        }
        {   /* recompute derivative at the converged node */
            double p0 = 1, p1 = x;
            for (int k = 2; k <= n; k++) {
                double p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
                p0 = p1; p1 = p2;
            }
            pp = n * (x * p1 - p0) / (x * x - 1);
        }
        r->x[i] = x;
        r->w[i] = 2.0 / ((1 - x * x) * pp * pp);
    }
}

static double gl_apply(const gl_rule *r, const gsl_function *f, double a, double b)
{
    double c = 0.5 * (a + b), h = 0.5 * (b - a), s = 0;
    for (int i = 0; i < r->n; i++) s += r->w[i] * GSL_FN_EVAL(f, c + h * r->x[i]);
    return s * h;
}

gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n)
{
    gsl_integration_workspace *w = malloc(sizeof(*w));
    w->limit = n;
    return w;
}
void gsl_integration_workspace_free(gsl_integration_workspace *w) { free(w); }

typedef struct { double a, b, r, e; } qag_seg;

int gsl_integration_qag(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        size_t limit, int key, gsl_integration_workspace *w, double *result, double *abserr)
{
    static gl_rule lo[7], hi[7];
    static int made[7];
    (void) w;
    if (key < 1) key = 1;
    if (key > 6) key = 6;
    if (!made[key]) {
        #pragma omp critical(minigsl_qag_tables)
        if (!made[key]) {
            int nk = 10 * key + 1 + (key == 1 ? 4 : 0);   /* 15,21,31,41,51,61 */
            gl_make(&lo[key], nk / 2);
            gl_make(&hi[key], nk);
            made[key] = 1;
        }
    }
    size_t cap = 64, n = 1;
    qag_seg *s = malloc(cap * sizeof(*s));
    s[0].a = a; s[0].b = b;
    s[0].r = gl_apply(&hi[key], f, a, b);
    s[0].e = fabs(s[0].r - gl_apply(&lo[key], f, a, b));
    double tot = s[0].r, err = s[0].e;
    int status = GSL_SUCCESS;
    while (1) {
        double tol = fmax(epsabs, epsrel * fabs(tot));
        if (err <= tol) break;
        /* cannot do better than round-off of the sum */
        if (err <= 50 * DBL_EPSILON * fabs(tot)) break;
        if (n >= limit) { status = GSL_EMAXITER; break; }
        size_t worst = 0;
        for (size_t i = 1; i < n; i++) if (s[i].e > s[worst].e) worst = i;
        if (n + 1 > cap) { cap *= 2; s = realloc(s, cap * sizeof(*s)); }
        double m = 0.5 * (s[worst].a + s[worst].b);
        qag_seg l = { s[worst].a, m, 0, 0 }, r = { m, s[worst].b, 0, 0 };
        l.r = gl_apply(&hi[key], f, l.a, l.b); l.e = fabs(l.r - gl_apply(&lo[key], f, l.a, l.b));
        r.r = gl_apply(&hi[key], f, r.a, r.b); r.e = fabs(r.r - gl_apply(&lo[key], f, r.a, r.b));
        s[worst] = l; s[n++] = r;
        tot = 0; err = 0;
        for (size_t i = 0; i < n; i++) { tot += s[i].r; err += s[i].e; }
    }
    free(s);
    *result = tot; *abserr = err;
    if (status != GSL_SUCCESS && current_handler)
        current_handler("qag: maximum number of subdivisions reached", __FILE__, __LINE__, status);
    return status;
}

/* -------------------------------------------------------------------- ode */
static const gsl_odeiv2_step_type rkf45_type = { "rkf45" };
const gsl_odeiv2_step_type *gsl_odeiv2_step_rkf45 = &rkf45_type;

gsl_odeiv2_driver *gsl_odeiv2_driver_alloc_standard_new(const gsl_odeiv2_system *sys,
        const gsl_odeiv2_step_type *T, double hstart, double epsabs, double epsrel, double a_y, double a_dydt)
{
    (void) T;
    gsl_odeiv2_driver *d = malloc(sizeof(*d));
    d->sys = sys; d->h = hstart; d->epsabs = epsabs; d->epsrel = epsrel; d->a_y = a_y; d->a_dydt = a_dydt;
    return d;
}
void gsl_odeiv2_driver_free(gsl_odeiv2_driver *d) { free(d); }

#define ODE_MAXDIM 16
int gsl_odeiv2_driver_apply(gsl_odeiv2_driver *d, double *t, double t1, double y[])
{
    const gsl_odeiv2_system *sys = d->sys;
    const size_t n = sys->dimension;
    if (n > ODE_MAXDIM) return GSL_FAILURE;
    double sign = (t1 >= *t) ? 1.0 : -1.0;
    if (sign * d->h < 0) return GSL_FAILURE;      /* GSL refuses to integrate against the sign of h */
    double h = d->h;
    double k1[ODE_MAXDIM], k2[ODE_MAXDIM], k3[ODE_MAXDIM], k4[ODE_MAXDIM], k5[ODE_MAXDIM], k6[ODE_MAXDIM];
    double yt[ODE_MAXDIM], yn[ODE_MAXDIM], ye[ODE_MAXDIM], dn[ODE_MAXDIM];
    long nsteps = 0;
    while (sign * (t1 - *t) > 0) {
        int final = 0;
        double h0 = h;
        if (sign * (*t + h0 - t1) > 0) { h0 = t1 - *t; final = 1; }
        while (1) {
            double tt = *t;
            sys->function(tt, y, k1, sys->params);
            for (size_t i = 0; i < n; i++) yt[i] = y[i] + h0 * (1.0 / 4) * k1[i];
            sys->function(tt + h0 / 4, yt, k2, sys->params);
            for (size_t i = 0; i < n; i++) yt[i] = y[i] + h0 * (3.0 / 32 * k1[i] + 9.0 / 32 * k2[i]);
            sys->function(tt + 3 * h0 / 8, yt, k3, sys->params);
            for (size_t i = 0; i < n; i++) yt[i] = y[i] + h0 * (1932.0 / 2197 * k1[i] - 7200.0 / 2197 * k2[i] + 7296.0 / 2197 * k3[i]);
            sys->function(tt + 12 * h0 / 13, yt, k4, sys->params);
            for (size_t i = 0; i < n; i++) yt[i] = y[i] + h0 * (439.0 / 216 * k1[i] - 8.0 * k2[i] + 3680.0 / 513 * k3[i] - 845.0 / 4104 * k4[i]);
            sys->function(tt + h0, yt, k5, sys->params);
            for (size_t i = 0; i < n; i++) yt[i] = y[i] + h0 * (-8.0 / 27 * k1[i] + 2.0 * k2[i] - 3544.0 / 2565 * k3[i] + 1859.0 / 4104 * k4[i] - 11.0 / 40 * k5[i]);
            sys->function(tt + h0 / 2, yt, k6, sys->params);
            for (size_t i = 0; i < n; i++) {
                yn[i] = y[i] + h0 * (16.0 / 135 * k1[i] + 6656.0 / 12825 * k3[i] + 28561.0 / 56430 * k4[i] - 9.0 / 50 * k5[i] + 2.0 / 55 * k6[i]);
                ye[i] = h0 * (1.0 / 360 * k1[i] - 128.0 / 4275 * k3[i] - 2197.0 / 75240 * k4[i] + 1.0 / 50 * k5[i] + 2.0 / 55 * k6[i]);
            }
            sys->function(tt + h0, yn, dn, sys->params);
            /* GSL std_control_hadjust, ord = 5, S = 0.9 */
            double rmax = DBL_MIN;
            for (size_t i = 0; i < n; i++) {
                double D0 = d->epsrel * (d->a_y * fabs(yn[i]) + d->a_dydt * fabs(h0 * dn[i])) + d->epsabs;
                double r = fabs(ye[i]) / fabs(D0);
                if (r > rmax) rmax = r;
            }
            if (rmax > 1.1) {
                double r = 0.9 / pow(rmax, 1.0 / 5);
                if (r < 0.2) r = 0.2;
                h0 *= r; final = 0;
                if (++nsteps > 10000000) return GSL_FAILURE;
                continue;                              /* retry with the smaller step */
            }
            *t = final ? t1 : tt + h0;
            memcpy(y, yn, n * sizeof(double));
            if (rmax < 0.5) {
                double r = 0.9 / pow(rmax, 1.0 / 6);
                if (r > 5) r = 5;
                if (r < 1) r = 1;
                h = h0 * r;
            } else {
                h = h0;
            }
            break;
        }
        if (++nsteps > 10000000) return GSL_FAILURE;
    }
    d->h = h;
    return GSL_SUCCESS;
}

/* -------------------------------------------------------------------- rng */
static const gsl_rng_type ranlxd1_type = { "ranlxd1", 202 };
static const gsl_rng_type ranlxd2_type = { "ranlxd2", 397 };
const gsl_rng_type *gsl_rng_ranlxd1 = &ranlxd1_type;
const gsl_rng_type *gsl_rng_ranlxd2 = &ranlxd2_type;

static const int rlx_next[12] = { 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 0 };
static const double rlx_one_bit = 1.0 / 281474976710656.0;   /* 2^-48 */

#define RANLUX_STEP(x1, x2, i1, i2, i3) \
    x1 = xdbl[i1] - xdbl[i2];            \
    if (x2 < 0) { x1 -= rlx_one_bit; x2 += 1; } \
    xdbl[i3] = x2

static void rlx_increment_state(gsl_rng *state)
{
    int k, kmax;
    double y1, y2, y3;
    double *xdbl = state->xdbl;
    double carry = state->carry;
    unsigned int ir = state->ir, jr = state->jr;

    for (k = 0; ir > 0; ++k) {
        y1 = xdbl[jr] - xdbl[ir];
        y2 = y1 - carry;
        if (y2 < 0) { carry = rlx_one_bit; y2 += 1; } else { carry = 0; }
        xdbl[ir] = y2;
        ir = rlx_next[ir];
        jr = rlx_next[jr];
    }
    kmax = state->pr - 12;
    for (; k <= kmax; k += 12) {
        y1 = xdbl[7] - xdbl[0];
        y1 -= carry;
        RANLUX_STEP(y2, y1, 8, 1, 0);
        RANLUX_STEP(y3, y2, 9, 2, 1);
        RANLUX_STEP(y1, y3, 10, 3, 2);
        RANLUX_STEP(y2, y1, 11, 4, 3);
        RANLUX_STEP(y3, y2, 0, 5, 4);
        RANLUX_STEP(y1, y3, 1, 6, 5);
        RANLUX_STEP(y2, y1, 2, 7, 6);
        RANLUX_STEP(y3, y2, 3, 8, 7);
        RANLUX_STEP(y1, y3, 4, 9, 8);
        RANLUX_STEP(y2, y1, 5, 10, 9);
        RANLUX_STEP(y3, y2, 6, 11, 10);
        if (y3 < 0) { carry = rlx_one_bit; y3 += 1; } else { carry = 0; }
        xdbl[11] = y3;
    }
    kmax = state->pr;
    for (; k < kmax; ++k) {
        y1 = xdbl[jr] - xdbl[ir];
        y2 = y1 - carry;
        if (y2 < 0) { carry = rlx_one_bit; y2 += 1; } else { carry = 0; }
        xdbl[ir] = y2;
        ir = rlx_next[ir];
        jr = rlx_next[jr];
    }
    state->ir = ir; state->ir_old = ir; state->jr = jr; state->carry = carry;
}

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T)
{
    gsl_rng *r = calloc(1, sizeof(*r));
    r->type = T;
    gsl_rng_set(r, 0);
    return r;
}
void gsl_rng_free(gsl_rng *r) { free(r); }

void gsl_rng_set(gsl_rng *state, unsigned long int s)
{
    int ibit, jbit, i, k, l, xbit[31];
    double x, y;
    long int seed;
    if (s == 0) s = 1;
    seed = s;
    i = seed & 0x7FFFFFFFUL;
    for (k = 0; k < 31; ++k) { xbit[k] = i % 2; i /= 2; }
    ibit = 0; jbit = 18;
    for (k = 0; k < 12; ++k) {
        x = 0;
        for (l = 1; l <= 48; ++l) {
            y = (double) ((xbit[ibit] + 1) % 2);
            x += x + y;
            xbit[ibit] = (xbit[ibit] + xbit[jbit]) % 2;
            ibit = (ibit + 1) % 31;
            jbit = (jbit + 1) % 31;
        }
        state->xdbl[k] = rlx_one_bit * x;
    }
    state->carry = 0;
    state->ir = 11;
    state->jr = 7;
    state->ir_old = 0;
    state->pr = state->type->luxury;
}

double gsl_rng_uniform(gsl_rng *state)
{
    int ir = state->ir;
    state->ir = rlx_next[ir];
    if (state->ir == state->ir_old) rlx_increment_state(state);
    return state->xdbl[state->ir];
}

/* ----------------------------------------------------------------- interp */
static const gsl_interp_type linear_type = { "linear" };
const gsl_interp_type *gsl_interp_linear = &linear_type;
gsl_interp *gsl_interp_alloc(const gsl_interp_type *T, size_t n)
{ gsl_interp *p = malloc(sizeof(*p)); p->type = T; p->size = n; return p; }
int gsl_interp_init(gsl_interp *obj, const double xa[], const double ya[], size_t size)
{ (void) xa; (void) ya; obj->size = size; return GSL_SUCCESS; }
void gsl_interp_free(gsl_interp *p) { free(p); }
gsl_interp_accel *gsl_interp_accel_alloc(void) { return calloc(1, sizeof(gsl_interp_accel)); }
void gsl_interp_accel_free(gsl_interp_accel *a) { free(a); }
double gsl_interp_eval(const gsl_interp *obj, const double xa[], const double ya[], double x, gsl_interp_accel *a)
{
    (void) a;
    size_t lo = 0, hi = obj->size - 1;
    while (hi - lo > 1) { size_t m = (lo + hi) / 2; if (xa[m] > x) hi = m; else lo = m; }
    double dx = xa[lo + 1] - xa[lo];
    return ya[lo] + (x - xa[lo]) / dx * (ya[lo + 1] - ya[lo]);
}
