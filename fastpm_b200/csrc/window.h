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
#endif
