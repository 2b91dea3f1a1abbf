// CPU emulation of the CTA phases in fastpm_b200/csrc/fft_core.h (no GPU needed).
// Checks the mixed-radix DIF core, digit-reversal tables and the real<->half-complex
// untangling against a naive O(n^2) double-precision DFT.  Prints "OK" lines; exit code 0 on success.
#include "../../fastpm_b200/csrc/fft_core.h"
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef std::complex<double> cd;

static void run_core(std::vector<float2> &s, int kp, int ncol, const FpmFftHostPlan &pl, int nthr)
{
    FpmFftDev t; t.n = pl.n; t.nstage = (int) pl.radix.size();
    for (int j = 0; j < t.nstage; j++) t.radix[j] = pl.radix[j];
    t.tw = pl.tw.data(); t.rev = pl.rev.data(); t.inv = pl.inv.data();
    int ncur = pl.n;
    for (int j = 0; j < t.nstage; j++) {
        for (int tid = 0; tid < nthr; tid++) fpm_fft_stage(tid, nthr, s.data(), kp, ncol, t, ncur, t.radix[j]);
        ncur /= t.radix[j];
    }
}

static double test_complex(int n, int ncol, int kp, int nthr)
{
    FpmFftHostPlan pl(n);
    if (!pl.ok) { printf("plan failed n=%d\n", n); exit(1); }
    std::vector<float2> s((size_t) n * kp);
    std::vector<cd> x((size_t) n * ncol);
    srand(n * 7 + ncol);
    for (int e = 0; e < n; e++) for (int c = 0; c < ncol; c++) {
        double re = rand() / (double) RAND_MAX - 0.5, im = rand() / (double) RAND_MAX - 0.5;
        x[(size_t) e * ncol + c] = cd(re, im);
        s[(size_t) e * kp + c] = make_float2((float) re, (float) im);
    }
    run_core(s, kp, ncol, pl, nthr);
    double maxerr = 0, norm = 0;
    for (int c = 0; c < ncol; c += (ncol > 2 ? ncol - 1 : 1))
        for (int k = 0; k < n; k++) {
            cd acc = 0;
            for (int e = 0; e < n; e++) acc += x[(size_t) e * ncol + c] * std::polar(1.0, -2.0 * M_PI * (double) ((long long) e * k % n) / n);
            float2 got = s[(size_t) pl.inv[k] * kp + c];
            maxerr = fmax(maxerr, std::abs(acc - cd(got.x, got.y)));
            norm = fmax(norm, std::abs(acc));
        }
    return maxerr / norm;
}

static double test_real(int N, int ncol, int nthr)
{
    const int h = N / 2, kp = ncol + 1;
    FpmFftHostPlan ph(h), pN(N);
    FpmFftDev th; th.n = h; th.nstage = (int) ph.radix.size();
    for (int j = 0; j < th.nstage; j++) th.radix[j] = ph.radix[j];
    th.tw = ph.tw.data(); th.rev = ph.rev.data(); th.inv = ph.inv.data();
    std::vector<double> x((size_t) N * ncol);
    std::vector<float2> s((size_t) (h + 1) * kp);
    srand(N);
    for (int c = 0; c < ncol; c++) for (int j = 0; j < N; j++) x[(size_t) c * N + j] = rand() / (double) RAND_MAX - 0.5;
    for (int c = 0; c < ncol; c++) for (int j = 0; j < h; j++)
        s[(size_t) j * kp + c] = make_float2((float) x[(size_t) c * N + 2 * j], (float) x[(size_t) c * N + 2 * j + 1]);
    run_core(s, kp, ncol, ph, nthr);
    // forward untangle
    std::vector<float2> X((size_t) (h + 1) * ncol);
    double maxerr = 0, norm = 0;
    for (int c = 0; c < ncol; c++) for (int k = 0; k <= h; k++) {
        float2 g = fpm_untangle_fwd(s.data(), kp, c, th, pN.tw.data(), k);
        X[(size_t) k * ncol + c] = g;
        cd acc = 0;
        for (int j = 0; j < N; j++) acc += x[(size_t) c * N + j] * std::polar(1.0, -2.0 * M_PI * (double) ((long long) j * k % N) / N);
        maxerr = fmax(maxerr, std::abs(acc - cd(g.x, g.y))); norm = fmax(norm, std::abs(acc));
    }
    double fwd = maxerr / norm;
    // backward: tangle pairs -> core -> conj -> compare with N * x
    for (int c = 0; c < ncol; c++) for (int k = 0; k <= h / 2; k++) {
        float2 zk, zhk;
        fpm_tangle_bwd_pair(X[(size_t) k * ncol + c], X[(size_t) (h - k) * ncol + c], pN.tw[k], &zk, &zhk);
        s[(size_t) k * kp + c] = zk;
        if (k != 0 && k != h - k) s[(size_t) (h - k) * kp + c] = zhk;
    }
    run_core(s, kp, ncol, ph, nthr);
    maxerr = 0;
    for (int c = 0; c < ncol; c++) for (int j = 0; j < h; j++) {
        float2 g = s[(size_t) ph.inv[j] * kp + c];
        double e0 = g.x - N * x[(size_t) c * N + 2 * j], e1 = -g.y - N * x[(size_t) c * N + 2 * j + 1];
        maxerr = fmax(maxerr, fmax(fabs(e0), fabs(e1)));
    }
    double bwd = maxerr / (0.5 * N);
    return fmax(fwd, bwd);
}

int main()
{
    int fail = 0;
    int sizes[] = { 4, 6, 8, 10, 12, 16, 20, 30, 32, 48, 60, 64, 96, 128, 192, 256, 360, 384, 512, 768, 1024 };
    for (int n : sizes) {
        double e = test_complex(n, 8, 8, 64);
        printf("complex n=%d relerr=%.3g %s\n", n, e, e < 2e-6 ? "OK" : "FAIL");
        if (!(e < 2e-6)) fail = 1;
    }
    { double e = test_complex(2048, 4, 5, 96); printf("complex n=2048 relerr=%.3g %s\n", e, e < 2e-6 ? "OK" : "FAIL"); if (!(e < 2e-6)) fail = 1; }
    int rsizes[] = { 8, 12, 16, 24, 64, 192, 256, 384, 1024 };
    for (int n : rsizes) {
        double e = test_real(n, 4, 32);
        printf("real N=%d relerr=%.3g %s\n", n, e, e < 2e-6 ? "OK" : "FAIL");
        if (!(e < 2e-6)) fail = 1;
    }
    return fail;
}
