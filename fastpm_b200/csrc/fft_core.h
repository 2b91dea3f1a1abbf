// fastpm_b200 -- shared-memory FFT building blocks.
//
// Everything here is written as "one phase of one CTA, for thread `tid` of `nthr`", with the
// CTA-wide barrier implied BETWEEN phases.  The CUDA kernels (fft.cu) call a phase with
// (threadIdx.x, blockDim.x) and then __syncthreads(); the host emulation used by the CPU tests
// (tests/emul/fft_emul.cpp) calls the same functions in a loop over tid.  This keeps the index
// arithmetic (digit reversal, twiddle indices, tile addressing) testable without a GPU.
//
// Transform: in-place decimation-in-frequency, mixed radix {2,3,4,5}, forward sign exp(-2 pi i jk/n).
// A tile holds `ncol` independent columns: element e of column c lives at s[e*kp + c] (kp >= ncol).
// After the last stage, frequency k sits at position pos with rev[pos] == k  (inv[k] == pos).
// Inverse transforms are done as conj(FFT(conj(x))).
#pragma once

#ifdef __CUDACC__
#define FPM_HD __host__ __device__ __forceinline__
#else
#define FPM_HD inline
#include <math.h>
#ifndef __VECTOR_TYPES_H__       /* a host build that has not pulled in CUDA's vector types */
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif
#endif

#define FPM_FFT_MAX_STAGES 16

struct FpmFftDev {
    int n;                         // transform length
    int nstage;
    int radix[FPM_FFT_MAX_STAGES];
    const float2 *tw;              // [n]  exp(-2 pi i t / n)
    const int *rev;                // [n]  position -> frequency
    const int *inv;                // [n]  frequency -> position
};

FPM_HD float2 fpm_cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
FPM_HD float2 fpm_cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
FPM_HD float2 fpm_csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
FPM_HD float2 fpm_conj(float2 a) { return make_float2(a.x, -a.y); }

// One DIF stage over a tile.  `ncur` is the current sub-transform length (n / product of earlier
// radices), r the radix of this stage.  Butterfly bf (0 <= bf < n/r) of column c touches
// elements blk*ncur + p + k*m, k < r, with m = ncur/r, blk = bf / m, p = bf % m.
// Work items are (bf, c) flattened with c fastest so that a warp touches contiguous shared memory.
FPM_HD void fpm_fft_stage(int tid, int nthr, float2 *s, int kp, int ncol, const FpmFftDev &t, int ncur, int r)
{
    const int n = t.n;
    const int m = ncur / r;
    const int twstep = n / ncur;           // omega_ncur^x == tw[x * twstep]
    const int nbf = n / r;
    const int nwork = nbf * ncol;
    for (int w = tid; w < nwork; w += nthr) {
        const int c = w % ncol;
        const int bf = w / ncol;
        const int blk = bf / m;
        const int p = bf - blk * m;
        float2 *base = s + (size_t) (blk * ncur + p) * kp + c;
        const size_t step = (size_t) m * kp;
        if (r == 4) {
            float2 a0 = base[0], a1 = base[step], a2 = base[2 * step], a3 = base[3 * step];
            float2 t0 = fpm_cadd(a0, a2), t1 = fpm_csub(a0, a2), t2 = fpm_cadd(a1, a3);
            float2 d = fpm_csub(a1, a3);
            float2 t3 = make_float2(d.y, -d.x);                  // (a1 - a3) * (-i)
            float2 y0 = fpm_cadd(t0, t2), y1 = fpm_cadd(t1, t3), y2 = fpm_csub(t0, t2), y3 = fpm_csub(t1, t3);
            if (p != 0) {
                const int ti = p * twstep;
                y1 = fpm_cmul(y1, t.tw[ti]);
                y2 = fpm_cmul(y2, t.tw[2 * ti]);
                y3 = fpm_cmul(y3, t.tw[3 * ti]);
            }
            base[0] = y0; base[step] = y1; base[2 * step] = y2; base[3 * step] = y3;
        } else if (r == 2) {
            float2 a0 = base[0], a1 = base[step];
            float2 y0 = fpm_cadd(a0, a1), y1 = fpm_csub(a0, a1);
            if (p != 0) y1 = fpm_cmul(y1, t.tw[p * twstep]);
            base[0] = y0; base[step] = y1;
        } else if (r == 3) {
            const float sn = 0.86602540378443864676f;
            float2 a0 = base[0], a1 = base[step], a2 = base[2 * step];
            float2 sm = fpm_cadd(a1, a2), d = fpm_csub(a1, a2);
            float2 mm = make_float2(a0.x - 0.5f * sm.x, a0.y - 0.5f * sm.y);
            // -i * sn * d = (sn*d.y, -sn*d.x)
            float2 y1 = make_float2(mm.x + sn * d.y, mm.y - sn * d.x);
            float2 y2 = make_float2(mm.x - sn * d.y, mm.y + sn * d.x);
            float2 y0 = fpm_cadd(a0, sm);
            if (p != 0) {
                const int ti = p * twstep;
                y1 = fpm_cmul(y1, t.tw[ti]);
                y2 = fpm_cmul(y2, t.tw[2 * ti]);
            }
            base[0] = y0; base[step] = y1; base[2 * step] = y2;
        } else {   // r == 5
            const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
            const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
            float2 a0 = base[0], a1 = base[step], a2 = base[2 * step], a3 = base[3 * step], a4 = base[4 * step];
            float2 p14 = fpm_cadd(a1, a4), m14 = fpm_csub(a1, a4), p23 = fpm_cadd(a2, a3), m23 = fpm_csub(a2, a3);
            float2 y0 = make_float2(a0.x + p14.x + p23.x, a0.y + p14.y + p23.y);
            float2 ra = make_float2(a0.x + c1 * p14.x + c2 * p23.x, a0.y + c1 * p14.y + c2 * p23.y);
            float2 rb = make_float2(a0.x + c2 * p14.x + c1 * p23.x, a0.y + c2 * p14.y + c1 * p23.y);
            float2 ia = make_float2(s1 * m14.x + s2 * m23.x, s1 * m14.y + s2 * m23.y);
            float2 ib = make_float2(s2 * m14.x - s1 * m23.x, s2 * m14.y - s1 * m23.y);
            // y1 = ra - i*ia ; y4 = ra + i*ia ; y2 = rb - i*ib ; y3 = rb + i*ib      (-i*z = (z.y, -z.x))
            float2 y1 = make_float2(ra.x + ia.y, ra.y - ia.x), y4 = make_float2(ra.x - ia.y, ra.y + ia.x);
            float2 y2 = make_float2(rb.x + ib.y, rb.y - ib.x), y3 = make_float2(rb.x - ib.y, rb.y + ib.x);
            if (p != 0) {
                const int ti = p * twstep;
                y1 = fpm_cmul(y1, t.tw[ti]);
                y2 = fpm_cmul(y2, t.tw[2 * ti]);
                y3 = fpm_cmul(y3, t.tw[3 * ti]);
                y4 = fpm_cmul(y4, t.tw[4 * ti]);
            }
            base[0] = y0; base[step] = y1; base[2 * step] = y2; base[3 * step] = y3; base[4 * step] = y4;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Real <-> half-complex untangling around a half-length (h = N/2) complex transform.
// twN is the length-N table exp(-2 pi i k / N); only k <= h is used.

// forward: X[k], k = 0..h, of the real sequence whose even/odd samples were packed as z_j = x_2j + i x_2j+1
// and transformed (Z stored digit-reversed in the tile, column c).
FPM_HD float2 fpm_untangle_fwd(const float2 *s, int kp, int c, const FpmFftDev &th, const float2 *twN, int k)
{
    const int h = th.n;
    const int k1 = (k == h) ? 0 : k;
    const int k2 = (k == 0 || k == h) ? 0 : h - k;
    float2 z1 = s[(size_t) th.inv[k1] * kp + c];
    float2 z2 = fpm_conj(s[(size_t) th.inv[k2] * kp + c]);
    float2 e = make_float2(0.5f * (z1.x + z2.x), 0.5f * (z1.y + z2.y));
    float2 d = fpm_csub(z1, z2);
    float2 o = make_float2(0.5f * d.y, -0.5f * d.x);            // (z1 - z2) / (2i)
    float2 w = twN[k];
    return fpm_cadd(e, fpm_cmul(w, o));
}

// backward, pair step: given X[k] and X[h-k] (k in 0..h/2) produce conj(Z[k]) and conj(Z[h-k]) where
// Z[k] = (X[k] + conj X[h-k]) + i exp(+2 pi i k/N) (X[k] - conj X[h-k]);  the conjugates feed the
// forward core (inverse by conjugation).  Z[h] is not produced (k = 0 pairs with X[h] only for input).
FPM_HD void fpm_tangle_bwd_pair(float2 xk, float2 xhk, float2 wk /* twN[k] */, float2 *zk_conj, float2 *zhk_conj)
{
    // Z[k]
    {
        float2 b = fpm_conj(xhk);
        float2 sm = fpm_cadd(xk, b), d = fpm_csub(xk, b);
        float2 wd = fpm_cmul(fpm_conj(wk), d);                    // exp(+2 pi i k/N) * d
        float2 z = make_float2(sm.x - wd.y, sm.y + wd.x);         // sm + i*wd
        *zk_conj = fpm_conj(z);
    }
    // Z[h-k]: exp(+2 pi i (h-k)/N) = -conj(exp(+2 pi i k/N)) = -wk
    {
        float2 b = fpm_conj(xk);
        float2 sm = fpm_cadd(xhk, b), d = fpm_csub(xhk, b);
        float2 wd = fpm_cmul(make_float2(-wk.x, -wk.y), d);
        float2 z = make_float2(sm.x - wd.y, sm.y + wd.x);
        *zhk_conj = fpm_conj(z);
    }
}

// ------------------------------------------------------------------------------------------
// Host-side plan helper (also used by the emulation): factorisation and index tables.
#include <vector>
#include <cmath>
struct FpmFftHostPlan {
    int n = 0;
    std::vector<int> radix;
    std::vector<float2> tw;
    std::vector<int> rev, inv;
    bool ok = false;

    explicit FpmFftHostPlan(int n_) : n(n_)
    {
        int r = n;
        std::vector<int> f;
        while (r % 4 == 0) { f.push_back(4); r /= 4; }
        while (r % 2 == 0) { f.push_back(2); r /= 2; }
        while (r % 3 == 0) { f.push_back(3); r /= 3; }
        while (r % 5 == 0) { f.push_back(5); r /= 5; }
        ok = (r == 1) && n >= 1 && (int) f.size() <= FPM_FFT_MAX_STAGES;
        if (!ok) return;
        radix = f;
        tw.resize(n);
        for (int t = 0; t < n; t++) {
            double a = -2.0 * M_PI * (double) t / (double) n;
            tw[t] = make_float2((float) std::cos(a), (float) std::sin(a));
        }
        rev.resize(n); inv.resize(n);
        for (int pos = 0; pos < n; pos++) {
            // pos = q1*m1 + q2*m2 + ... ; k = q1 + r1*(q2 + r2*(q3 + ...))
            int rem = pos, m = n, k = 0, mult = 1;
            for (size_t j = 0; j < radix.size(); j++) {
                m /= radix[j];
                int q = rem / m; rem -= q * m;
                k += q * mult; mult *= radix[j];
            }
            rev[pos] = k; inv[k] = pos;
        }
    }
};
