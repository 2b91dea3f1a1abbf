/* fastpm_b200 -- the mass-assignment windows other than CIC (reference: libfastpm/painter.c:17-125), usable from the C host layer and
 * from device code.  x is the distance to the mesh point in cells, invh = 1 / (support / 2).
 *
 * The Lanczos window of the reference goes through a 16384-entry table with a spacing of 1e-3 (__cached__, painter.c:65-83): for
 * 1e-3 < x < 16.384 the value is the function at 1e-3 * (int)(|x| / 1e-3), outside (and for every negative x) the function
 * itself.  That quantisation is part of the reference's result, so it is restated here (evaluated directly: table[i] = f(1e-3 * i)). */
#ifndef FASTPM_B200_WINDOW_H
#define FASTPM_B200_WINDOW_H
#include <math.h>

#ifdef __CUDACC__
#define FPM_WINDOW_HD __host__ __device__ __forceinline__
#else
#define FPM_WINDOW_HD static inline
#endif

enum { FPM_WINDOW_CIC = 0, FPM_WINDOW_LINEAR = 1, FPM_WINDOW_QUAD = 2, FPM_WINDOW_LANCZOS = 3 };     /* = FastPMPainterType, painter.h */
#define FPM_WINDOW_MAX_SUPPORT 8

FPM_WINDOW_HD double fpm_window_linear(double x, double invh) { return 1.0 - fabs(x * invh); }

FPM_WINDOW_HD double fpm_window_quad(double x, double invh)
{
    x = fabs(x) * invh;
    if (x <= 0.5) return 0.75 - x * x;
    x = 1.5 - x;
    return (x * x) * 0.5;
}

FPM_WINDOW_HD double fpm_window_sinc(double x)
{
    x *= 3.1415927;
    if (x < 1e-5 && x > -1e-5) {
        const double x2 = x * x;
        return 1.0 - x2 / 6. + x2 * x2 / 120.;
    }
    return sin(x) / x;
}

FPM_WINDOW_HD double fpm_window_cached_sinc(double x)
{
    const double dx = 1e-3;
    const double tablemax = dx * 16384, tablemin = dx * 1;
    if (x > tablemin && x < tablemax) {
        const int i = (int) (fabs(x) / dx);
        return fpm_window_sinc(dx * i);
    }
    return fpm_window_sinc(x);
}

FPM_WINDOW_HD double fpm_window_lanczos(double x, double invh) { return fpm_window_cached_sinc(x) * fpm_window_cached_sinc(x * invh); }

FPM_WINDOW_HD double fpm_window_eval(int type, double x, double invh)
{
    return type == FPM_WINDOW_LINEAR ? fpm_window_linear(x, invh) : (type == FPM_WINDOW_QUAD ? fpm_window_quad(x, invh) : fpm_window_lanczos(x, invh));
}

/* ---- the derivatives of the windows (painter.c:20-27, 43-59, 85-125), used by a painter made with fastpm_painter_init_diff */
FPM_WINDOW_HD double fpm_window_linear_diff(double x, double invh) { return x < 0 ? 1 * invh : -1 * invh; }

FPM_WINDOW_HD double fpm_window_quad_diff(double x, double invh)
{
    double factor;
    x *= invh;
    if (x < 0) { x = -x; factor = -1 * invh; }
    else factor = +1 * invh;
    if (x < 0.5) return factor * (-2 * x);
    return factor * (-(1.5 - x));
}

FPM_WINDOW_HD double fpm_window_dsinc(double x)
{
    x *= 3.1415927;
    double r = 3.1415927;
    if (x < 1e-5 && x > -1e-5) {
        const double xx = x * x, xxxx = xx * xx;
        r *= -x / 3 + x * xx / 30 - xxxx * x / 840 + xxxx * xx * x / 45360;
    } else {
        r *= 1 / x * cos(x) - 1 / (x * x) * sin(x);
    }
    return r;
}

/* _lanczos_diff (painter.c:117-125) looks all four factors up through ONE table and one "filled" flag: whichever function is asked
 * for first fills it -- the sinc -- and the two dsinc factors then read sinc values from it whenever 1e-3 < x < 16.384; only outside
 * that range (every negative argument in particular) they are the true derivative.  That is what the reference computes, so that is
 * what is restated here. */
FPM_WINDOW_HD double fpm_window_cached_dsinc_as_in_reference(double x)
{
    const double dx = 1e-3;
    const double tablemax = dx * 16384, tablemin = dx * 1;
    if (x > tablemin && x < tablemax) {
        const int i = (int) (fabs(x) / dx);
        return fpm_window_sinc(dx * i);
    }
    return fpm_window_dsinc(x);
}

FPM_WINDOW_HD double fpm_window_lanczos_diff(double x, double invh)
{
    const double u1 = fpm_window_cached_sinc(x), u2 = fpm_window_cached_dsinc_as_in_reference(x);
    const double v1 = fpm_window_cached_sinc(x * invh), v2 = fpm_window_cached_dsinc_as_in_reference(x * invh) * invh;
    return u1 * v2 + u2 * v1;
}

FPM_WINDOW_HD double fpm_window_diff_eval(int type, double x, double invh)
{
    return type == FPM_WINDOW_LINEAR ? fpm_window_linear_diff(x, invh) : (type == FPM_WINDOW_QUAD ? fpm_window_quad_diff(x, invh) : fpm_window_lanczos_diff(x, invh));
}
#endif
