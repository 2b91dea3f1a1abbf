// CPU run of the strided FFT pass KERNEL SOURCE, fastpm_b200/csrc/fft_tma.cu (see cuda_emul.h), for every mesh size of the
// fast path:
//   1. in-place pass, forward and inverse, against a naive double-precision DFT of the tile columns;
//   2. the fused gravity kernel (Green's function x i k_d, zeroed self-conjugate modes) against fpm_apply_transfer of
//      csrc/mesh.cuh -- the reference's arithmetic -- followed by the same DFT;
//   3. the multi-destination (slab transpose) store path: 2 ranks, staging blocks + own rows straight to their final place.
#include "cuda_emul.h"
#include <complex>
#include <cstdio>
#include <cstdlib>
#include "../../fastpm_b200/csrc/fft_tma.cu"

typedef std::complex<double> cd;

struct Tables { std::vector<float> k, kk, kf, kkf, kkf2; FpmKTables kt; std::vector<float2> tw; };
static void make_tables(int n, Tables &t)
{
    t.k.resize(n); t.kk.resize(n); t.kf.resize(n); t.kkf.resize(n); t.kkf2.resize(n); t.tw.resize(n);
    const double L = 100.0, cell = L / n;
    for (int i = 0; i < n; i++) {
        int ii = i >= n / 2 ? i - n : i;
        const float k = (float) (ii * 2 * M_PI / L), w = (float) (k * cell);
        t.k[i] = k; t.kk[i] = k * k;
        t.kf[i] = (float) (1 / cell * (1 / 6.0 * (8 * sin((double) w) - sin(2 * (double) w))));
        t.kkf[i] = t.kk[i]; t.kkf2[i] = t.kk[i];
        t.tw[i] = make_float2((float) cos(-2 * M_PI * i / n), (float) sin(-2 * M_PI * i / n));
    }
    t.kt.k = t.k.data(); t.kt.kk = t.kk.data(); t.kt.k_finite = t.kf.data(); t.kt.kk_finite = t.kkf.data(); t.kt.kk_finite2 = t.kkf2.data(); t.kt.n = n;
}

static CUtensorMap make_tmap(const float2 *src, int n, int pitch_c, int K)
{
    CUtensorMap m;
    memset(&m, 0, sizeof(m));
    FpmEmulTmap *e = reinterpret_cast<FpmEmulTmap *>(&m);
    e->base = reinterpret_cast<const float *>(src);
    e->gstr_bytes[0] = (uint64_t) pitch_c * 8; e->gstr_bytes[1] = (uint64_t) n * pitch_c * 8;
    e->box[0] = 2 * K; e->box[1] = n < 256 ? n : 256; e->box[2] = 1;
    return m;
}

// naive DFT of column (plane o, kz) of `in` ([plane][row][pitch_c]) with optional per-element transfer, sign = -1 forward / +1 inverse
static cd column_dft(const std::vector<float2> &in, int n, int pitch_c, int o, int kz, int kout, int sign,
                     const FpmTransferSpec *xf, const FpmKTables *kt, int outer0)
{
    cd acc = 0;
    for (int r = 0; r < n; r++) {
        float2 v = in[((size_t) o * n + r) * pitch_c + kz];
        if (xf) v = fpm_apply_transfer(*xf, *kt, v, r, outer0 + o, kz);
        acc += cd(v.x, v.y) * std::polar(1.0, sign * 2 * M_PI * (double) ((long long) r * kout % n) / n);
    }
    return acc;
}

// what: bit 0 plain passes, bit 1 force kernel (all directions; bit 3: one direction only), bit 2 slab transpose
template <int R1, int R2, int R3, int K>
static int run(int nouter, unsigned grid, int what_mask)
{
    using C = TmaCfg<R1, R2, R3>;
    const int n = C::N, h = n / 2, pitch_c = ((h + 1 + 15) / 16) * 16;
    Tables tb; make_tables(n, tb);
    const size_t plane = (size_t) n * pitch_c;
    std::vector<float2> in(plane * nouter), out(plane * nouter);
    srand(n + K);
    for (auto &v : in) v = make_float2(rand() / (float) RAND_MAX - 0.5f, rand() / (float) RAND_MAX - 0.5f);
    const size_t smem = (size_t) n * K * 12 + (size_t) n * 8;
    int bad = 0;
    auto base_args = [&](float2 *dst, int conj) {
        TmaPassArgs a;
        memset(&a, 0, sizeof(a));
        a.dst[0] = dst; a.rows_per_rank = n; a.dst_estride = pitch_c; a.dst_ostride = plane; a.dst_ooffset = 0; a.self_rank = -1;
        a.ntile_k = (h + 1 + K - 1) / K; a.nouter = nouter; a.conj = conj; a.chunk = 1; a.early = 3; a.outer0 = 0; a.tw = tb.tw.data();
        a.xfer.active = 0; a.kt = tb.kt;
        return a;
    };
    auto check = [&](const char *what, const std::vector<float2> &res, int sign, const FpmTransferSpec *xf, int outer0, size_t estride, size_t ostride) {
        double err = 0, norm = 0;
        const int kzs[] = { 0, 1, K - 1, K, h / 2 + 3, h - 1, h };
        for (int o = 0; o < nouter; o++) for (int kz : kzs) for (int kout = 0; kout < n; kout += (n > 256 ? n / 37 : 1)) {
            const cd want = column_dft(in, n, pitch_c, o, kz, kout, sign, xf, &tb.kt, outer0);
            const float2 g = res[(size_t) kout * estride + (size_t) o * ostride + kz];
            err = std::max(err, std::abs(cd(g.x, g.y) - want)); norm = std::max(norm, std::abs(want));
        }
        const double rel = norm > 0 ? err / norm : err;      // a plane whose gradient factor is exactly 0 gives an all-zero result
        printf("tile N=%4d K=%2d %-28s relerr=%.2e %s\n", n, K, what, rel, rel < 3e-6 ? "OK" : "FAIL");
        if (!(rel < 3e-6)) bad++;
    };
    // 1. in-place layout, forward and inverse
    for (int conj = 0; conj < 2 && (what_mask & 1); conj++) {
        TmaPassArgs a = base_args(out.data(), conj);
        CUtensorMap tm = make_tmap(in.data(), n, pitch_c, K);
        fpm_emul_launch(grid, C::T * K, smem, [&]() { fft_tma_kernel<R1, R2, R3, K, false>(tm, a); });
        check(conj ? "inverse, plain" : "forward, plain", out, conj ? +1 : -1, NULL, 0, pitch_c, plane);
    }
    // 2. inverse with the force kernel for the three gradient directions; planes ky = 0 .. and around the Nyquist plane
    for (int dir = 0; dir < 3 && (what_mask & 10); dir++) for (int outer0 : { 0, h - 1 }) {
        if ((what_mask & 8) && !(what_mask & 2) && !(dir == 0 && outer0 == h - 1)) continue;
        TmaPassArgs a = base_args(out.data(), 1);
        a.outer0 = outer0;
        a.xfer.active = 1; a.xfer.potorder = 0; a.xfer.negate = 1; a.xfer.ngrad = 1; a.xfer.graddir[0] = dir; a.xfer.gradorder = 1;
        a.xfer.zero_selfconj = 1; a.xfer.scale = 1.0;
        CUtensorMap tm = make_tmap(in.data(), n, pitch_c, K);
        fpm_emul_launch(grid, C::T * K, smem, [&]() { fft_tma_kernel<R1, R2, R3, K, false>(tm, a); });
        char what[64]; snprintf(what, sizeof what, "inverse, force d=%d ky0=%d", dir, outer0);
        check(what, out, +1, &a.xfer, outer0, pitch_c, plane);
    }
    // 3. slab transpose on "2 ranks": rows of rank 1 into a staging block [row][plane][kz], rows of rank 0 (self) to the final
    //    transposed place [row][x0 + plane][kz]
    if (what_mask & 4) {
        const int per = n / 2, x0 = 3;
        std::vector<float2> stage((size_t) 2 * per * nouter * pitch_c), fin((size_t) per * (x0 + nouter) * pitch_c);
        TmaPassArgs a = base_args(NULL, 0);
        a.dst[0] = stage.data(); a.dst[1] = stage.data() + (size_t) per * nouter * pitch_c;
        a.rows_per_rank = per; a.dst_estride = (size_t) nouter * pitch_c; a.dst_ostride = pitch_c; a.dst_ooffset = 0;
        a.self_rank = 0; a.self_dst = fin.data(); a.self_estride = (size_t) (x0 + nouter) * pitch_c; a.self_ostride = pitch_c; a.self_ooffset = x0;
        a.chunk = 2;
        CUtensorMap tm = make_tmap(in.data(), n, pitch_c, K);
        fpm_emul_launch(grid, C::T * K, smem, [&]() { fft_tma_kernel<R1, R2, R3, K, true>(tm, a); });
        std::vector<float2> res(plane * nouter);        // reassemble [row][plane][kz] -> in-place layout for the checker
        for (int kout = 0; kout < n; kout++) for (int o = 0; o < nouter; o++) for (int kz = 0; kz < pitch_c; kz++) {
            const int d = kout / per, kl = kout % per;
            res[(size_t) kout * pitch_c + (size_t) o * plane + kz] =
                d == 0 ? fin[((size_t) kl * (x0 + nouter) + x0 + o) * pitch_c + kz] : stage[(size_t) per * nouter * pitch_c + ((size_t) kl * nouter + o) * pitch_c + kz];
        }
        check("forward, 2-rank transpose", res, -1, NULL, 0, pitch_c, plane);
    }
    return bad;
}

int main(int argc, char **argv)
{
    const bool full = argc > 1 && !strcmp(argv[1], "full");      // ~2.5 min; the default subset (~40 s) is what pytest runs
    int bad = 0;
    bad += run<8, 8, 8, 16>(2, 3, 7);
    bad += run<24, 8, 4, 16>(2, 3, full ? 7 : 13);    // N = 768
    bad += run<16, 16, 4, 16>(2, 3, full ? 7 : 12);
    bad += run<24, 8, 8, 8>(2, 3, full ? 7 : 13);     // N = 1536: radix-24 first stage; plain, one force direction, 2-rank transpose
    bad += run<16, 16, 8, 8>(1, 3, full ? 7 : 9);
    bad += run<16, 16, 16, 4>(1, 5, full ? 7 : 4);
    return bad;
}
