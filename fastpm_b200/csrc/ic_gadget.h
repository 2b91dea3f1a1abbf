// fastpm_b200 -- Gaussian white noise in k-space with the Gadget / N-GenIC seeding scheme, the reference's default initial
// condition generator (pmic_fill_gaussian_gadget, libfastpm/initialcondition.c:145-273).
//
// The scheme is made for parallel generation: a master RANLUX stream hands one seed to every (kx, ky) column in a fixed
// order that does not depend on the mesh decomposition (fpm_gadget_seed_table, host, O(N^2) draws); every column then owns
// two generators -- its own and that of its Hermitian partner column -- and walks kz = 0 .. N/2 drawing (phase, amplitude)
// pairs from both, so that the kz = 0 and kz = N/2 planes come out Hermitian (fpm_gadget_fill_column: one CUDA thread per
// column, csrc/kspace.cu; the same function runs on the CPU in tests/emul/ic_emul.cpp).
// Output per mode: sqrt(-log(u)) * exp(i * 2 pi * v), i.e. unit-variance complex white noise (variance 1/2 per component).
#pragma once
#include "ranlux.h"
#include <math.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846264338328
#endif

// SETSEED of the reference writes the seed of column (i, j) into four tables at (i or N-i, j or N-j); only two of them are
// ever read back (GETSEED with d1 == d2): `self`[i][j], the seed of the column, and `conj`[i][j], the seed of the column
// whose mirror image (N-i, N-j) is (i, j).  Both are filled in the reference's own order (initialcondition.c:164-173), so a
// column visited twice keeps the later seed exactly like there.
inline void fpm_gadget_seed_table(int n, int seed, unsigned int *self, unsigned int *conj)
{
    FpmRanlux g;
    fpm_ranlux_seed(g, (unsigned long long) (unsigned int) seed);
    for (size_t q = 0; q < (size_t) n * n; q++) { self[q] = 0; conj[q] = 0; }
    auto set = [&](int i, int j) {
        const unsigned int s = (unsigned int) (0x7fffffff * fpm_ranlux_uniform(g));
        const int ci = (n - i) % n, cj = (n - j) % n;
        self[(size_t) i * n + j] = s;
        conj[(size_t) ci * n + cj] = s;
    };
    for (int i = 0; i < n / 2; i++) {
        for (int j = 0; j < i; j++) set(i, j);
        for (int j = 0; j < i + 1; j++) set(j, i);
        for (int j = 0; j < i; j++) set(n - 1 - i, j);
        for (int j = 0; j < i + 1; j++) set(n - 1 - j, i);
        for (int j = 0; j < i; j++) set(i, n - 1 - j);
        for (int j = 0; j < i + 1; j++) set(j, n - 1 - i);
        for (int j = 0; j < i; j++) set(n - 1 - i, n - 1 - j);
        for (int j = 0; j < i + 1; j++) set(n - 1 - j, n - 1 - i);
    }
}

// one (phase, amplitude) draw: SAMPLE, initialcondition.c:136-142
FPM_RLX_HD void fpm_gadget_sample(FpmRanlux &g, double &ampl, double &phase)
{
    phase = fpm_ranlux_uniform(g) * 2 * M_PI;
    do ampl = fpm_ranlux_uniform(g); while (ampl == 0);
}

// Column (i = kx, j = ky): writes modes kz = 0 .. n/2 to row[kz] (float2 = re, im).  initialcondition.c:187-262.
struct FpmFloat2 { float x, y; };
template <typename F2>
FPM_RLX_HD void fpm_gadget_fill_column(int n, int i, int j, unsigned int seed_self, unsigned int seed_conj_table, F2 *row)
{
    const int ci = (n - i) % n, cj = (n - j) % n;
    // the column whose kz = 0 / N/2 modes are the Hermitian images of another column's uses that column's generator
    const bool mirrored = (ci == i && cj < j) || (ci < i && cj != j) || (ci < i && cj == j);
    FpmRanlux lower, self;
    fpm_ranlux_seed(lower, mirrored ? seed_conj_table : seed_self);
    fpm_ranlux_seed(self, seed_self);
    const int h = n / 2;
    for (int k = 0; k <= h; k++) {
        const bool use_conj = mirrored && (k == 0 || k == h);
        double ampl, phase;
        if (use_conj) {
            fpm_gadget_sample(self, ampl, phase);
            fpm_gadget_sample(lower, ampl, phase);
        } else {
            fpm_gadget_sample(lower, ampl, phase);
            fpm_gadget_sample(self, ampl, phase);
        }
        ampl = sqrt(-log(ampl));
        float re = (float) (ampl * cos(phase)), im = (float) (ampl * sin(phase));
        if (use_conj) im *= -1;
        if (ci == i && cj == j && (n - k) % n == k) im = 0;       // self-conjugate mode: real
        if (i == 0 && j == 0 && k == 0) { re = 0; im = 0; }         // zero mean
        row[k].x = re; row[k].y = im;
    }
}
