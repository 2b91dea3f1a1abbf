// CPU run of the row (z) pass KERNEL SOURCE, fastpm_b200/csrc/fft_zrow.cu, for every mesh size of the fast path including the
// one no GPU test reaches (N = 4096): forward real rows -> half-complex rows against a naive double-precision DFT, then the
// backward kernel on that output must give N times the input back (unnormalised round trip, pmpfft.c:373,391).
#include "cuda_emul.h"
#include <complex>
#include <cstdio>
#include <cstdlib>
#include "../../fastpm_b200/csrc/fft_zrow.cu"

typedef std::complex<double> cd;

template <int R1, int R2, int R3, bool PREFETCH>
static int run(int n, int nrows, unsigned grid)
{
    using C = TmaCfg<R1, R2, R3>;
    const int h = n / 2, pitch_c = ((h + 1 + 15) / 16) * 16;
    std::vector<float> buf((size_t) nrows * 2 * pitch_c, 0.f), orig;
    std::vector<float2> twH(h), twN(n);
    for (int t = 0; t < h; t++) twH[t] = make_float2((float) cos(-2 * M_PI * t / h), (float) sin(-2 * M_PI * t / h));
    for (int t = 0; t < n; t++) twN[t] = make_float2((float) cos(-2 * M_PI * t / n), (float) sin(-2 * M_PI * t / n));
    srand(n);
    for (int r = 0; r < nrows; r++) for (int z = 0; z < n; z++) buf[(size_t) r * 2 * pitch_c + z] = rand() / (float) RAND_MAX - 0.5f;
    orig = buf;
    const float scale = 0.5f;
    ZRowArgs a = { buf.data(), buf.data(), (size_t) nrows, pitch_c, scale, twH.data(), twN.data(), 0 };
    const size_t smem = (size_t) (PREFETCH ? 2 : 1) * 8 * (C::N + 2) * sizeof(float2);
    fpm_emul_launch(grid, C::T * 8, smem, [&]() { fft_zrow_kernel<R1, R2, R3, true, PREFETCH, 1>(a); });
    double err = 0, norm = 0;
    for (int r = 0; r < nrows; r += (nrows > 8 ? 5 : 1)) {
        const float2 *row = reinterpret_cast<const float2 *>(buf.data() + (size_t) r * 2 * pitch_c);
        for (int k = 0; k <= h; k += (n > 512 ? 13 : 1)) {
            cd acc = 0;
            for (int z = 0; z < n; z++) acc += (double) (scale * orig[(size_t) r * 2 * pitch_c + z]) * std::polar(1.0, -2 * M_PI * (double) ((long long) z * k % n) / n);
            err = std::max(err, std::abs(cd(row[k].x, row[k].y) - acc));
            norm = std::max(norm, std::abs(acc));
        }
        if (row[h + 1].x != 0.f || row[h + 1].y != 0.f) err = 1e9;          // the pad element the kernel clears
    }
    a.scale = 1.f;
    fpm_emul_launch(grid, C::T * 8, smem, [&]() { fft_zrow_kernel<R1, R2, R3, false, PREFETCH, 1>(a); });
    double err2 = 0;
    for (int r = 0; r < nrows; r++) for (int z = 0; z < n; z++)
        err2 = std::max(err2, (double) fabsf(buf[(size_t) r * 2 * pitch_c + z] - n * scale * orig[(size_t) r * 2 * pitch_c + z]));
    const double rel = err / norm, rel2 = err2 / (n * scale * 0.5);
    const bool ok = rel < 2e-6 && rel2 < 4e-6;
    printf("zrow N=%4d prefetch=%d rows=%d grid=%u  forward relerr=%.2e  round trip relerr=%.2e %s\n", n, (int) PREFETCH, nrows, grid, rel, rel2, ok ? "OK" : "FAIL");
    return ok ? 0 : 1;
}

int main()
{
    int bad = 0;
    bad += run<8, 8, 4, true>(512, 40, 2);           // 5 tiles over 2 persistent CTAs
    bad += run<24, 4, 4, true>(768, 24, 2);          // rows of 384 = 24 * 4 * 4 complex
    bad += run<8, 8, 8, true>(1024, 24, 2);
    bad += run<24, 8, 4, true>(1536, 24, 2);         // rows of 768 = 24 * 8 * 4 complex: radix-24 first stage
    bad += run<16, 16, 4, true>(2048, 24, 2);
    bad += run<16, 16, 4, false>(2048, 16, 1);
    bad += run<16, 16, 8, false>(4096, 16, 1);
    return bad;
}
