// fastpm_b200 -- Luescher's RANLUX in double precision at luxury level 1 ("ranlxd1"), the generator FastPM draws its initial
// conditions from (gsl_rng_ranlxd1 in libfastpm/initialcondition.c:153-263, store.c:697-718).
//
// Written from the published algorithm (M. Luescher, Comput. Phys. Commun. 79 (1994) 100; the double-precision variant and
// the seeding used by GSL): a subtract-with-borrow recurrence x_n = x_(n-12+7) - x_(n-12) - c (mod 1) on 48-bit mantissas
// held in doubles; after every 12 numbers handed out the state is advanced by p = 202 steps in total.  All state values are
// multiples of 2^-48 in [0, 1), so every operation below is exact and host and device produce identical streams.
// Usable from host code (seed tables, CPU emulation in tests/emul) and from device code (one generator per thread).
#pragma once

#ifdef __CUDACC__
#define FPM_RLX_HD __host__ __device__ __forceinline__
#else
#define FPM_RLX_HD inline
#endif

struct FpmRanlux {
    double x[12];
    double carry;
    int ir, jr, ir_old;
};

FPM_RLX_HD int fpm_ranlux_next(int i) { return i == 11 ? 0 : i + 1; }

// gsl_rng_set for ranlxd: 31 seed bits feed a linear-feedback bit stream (taps 0 and 18), 48 inverted bits per state word
FPM_RLX_HD void fpm_ranlux_seed(FpmRanlux &g, unsigned long long seed)
{
    int bits[31];
    if (seed == 0) seed = 1;
    unsigned int s = (unsigned int) (seed & 0x7FFFFFFFull);
    for (int k = 0; k < 31; k++) { bits[k] = (int) (s & 1u); s >>= 1; }
    int ib = 0, jb = 18;
    for (int k = 0; k < 12; k++) {
        double v = 0;
        for (int l = 0; l < 48; l++) {
            v += v + (double) (1 - bits[ib]);
            bits[ib] = (bits[ib] + bits[jb]) & 1;
            ib = ib == 30 ? 0 : ib + 1;
            jb = jb == 30 ? 0 : jb + 1;
        }
        g.x[k] = v * (1.0 / 281474976710656.0);      // 2^-48
    }
    g.carry = 0;
    g.ir = 11; g.jr = 7; g.ir_old = 0;
}

// p = 202 subtract-with-borrow steps (luxury level 1 of the double-precision generator)
FPM_RLX_HD void fpm_ranlux_advance(FpmRanlux &g)
{
    const double one_bit = 1.0 / 281474976710656.0;
    int ir = g.ir, jr = g.jr;
    double carry = g.carry;
    for (int k = 0; k < 202; k++) {
        double y = (g.x[jr] - g.x[ir]) - carry;
        if (y < 0) { carry = one_bit; y += 1; } else carry = 0;
        g.x[ir] = y;
        ir = fpm_ranlux_next(ir);
        jr = fpm_ranlux_next(jr);
    }
    g.ir = ir; g.ir_old = ir; g.jr = jr; g.carry = carry;
}

// gsl_rng_uniform: the next number in [0, 1)
FPM_RLX_HD double fpm_ranlux_uniform(FpmRanlux &g)
{
    g.ir = fpm_ranlux_next(g.ir);
    if (g.ir == g.ir_old) fpm_ranlux_advance(g);
    return g.x[g.ir];
}
