// CPU run of the generic FFT pass KERNEL SOURCES of fastpm_b200/csrc/fft.cu (any Nmesh = 2^a 3^b 5^c: the meshes of
// variable-mesh runs such as 768 or 1536, vpm.c): the strided tile pass with its digit-reversed, multi-destination store and
// the forward / backward row passes, against naive double-precision DFTs (see cuda_emul.h).
#include "cuda_emul.h"
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../fastpm_b200/csrc/fft.cu"

typedef std::complex<double> cd;

static FpmFftDev dev_plan(const FpmFftHostPlan &hp)
{
    FpmFftDev d;
    d.n = hp.n; d.nstage = (int) hp.radix.size();
    for (int j = 0; j < d.nstage; j++) d.radix[j] = hp.radix[j];
    d.tw = hp.tw.data(); d.rev = hp.rev.data(); d.inv = hp.inv.data();
    return d;
}

template <int K>
static int tile_case(int n, int nouter, int conj)
{
    FpmFftHostPlan hp(n);
    if (!hp.ok) { printf("plan failed for %d\n", n); return 1; }
    const int h = n / 2, pitch_c = ((h + 1 + 15) / 16) * 16;
    const size_t plane = (size_t) n * pitch_c;
    std::vector<float2> in(plane * nouter), out0(plane * nouter), out1(plane * nouter);
    srand(n + K);
    for (auto &v : in) v = make_float2(rand() / (float) RAND_MAX - 0.5f, rand() / (float) RAND_MAX - 0.5f);
    TilePassArgs a;
    memset(&a, 0, sizeof(a));
    // rows [0, n/2) to destination 0, [n/2, n) to destination 1 (two-rank slab transpose), transposed layout [row][plane][kz]
    a.src = in.data(); a.src_estride = pitch_c; a.src_ostride = plane;
    a.dst[0] = out0.data(); a.dst[1] = out1.data(); a.rows_per_rank = n / 2; a.dst_estride = (size_t) nouter * pitch_c; a.dst_ostride = pitch_c; a.dst_ooffset = 0;
    a.self_rank = -1; a.ntile_k = (h + 1 + K - 1) / K; a.conj = conj; a.outer0 = 0; a.t = dev_plan(hp); a.xfer.active = 0;
    const size_t work = (size_t) n * K / 4;
    const unsigned thr = work >= 2048 ? 512 : (work >= 512 ? 256 : 128);
    fpm_emul_launch((unsigned) (nouter * a.ntile_k), thr, (size_t) n * K * sizeof(float2), [&]() { fft_tile_kernel<K>(a); });
    double err = 0, norm = 0;
    for (int o = 0; o < nouter; o++) for (int kz : { 0, 1, h / 2, h }) for (int kout = 0; kout < n; kout += (n > 200 ? n / 41 : 1)) {
        cd acc = 0;
        for (int r = 0; r < n; r++) {
            const float2 v = in[((size_t) o * n + r) * pitch_c + kz];
            acc += cd(v.x, v.y) * std::polar(1.0, (conj ? 2 : -2) * M_PI * (double) ((long long) r * kout % n) / n);
        }
        const std::vector<float2> &dst = kout < n / 2 ? out0 : out1;
        const float2 g = dst[((size_t) (kout % (n / 2)) * nouter + o) * pitch_c + kz];
        err = std::max(err, std::abs(cd(g.x, g.y) - acc)); norm = std::max(norm, std::abs(acc));
    }
    const double rel = err / norm;
    printf("generic tile N=%4d K=%2d %s relerr=%.2e %s\n", n, K, conj ? "inverse" : "forward", rel, rel < 3e-6 ? "OK" : "FAIL");
    return rel < 3e-6 ? 0 : 1;
}

template <int R>
static int row_case(int n, int nrows)
{
    FpmFftHostPlan hN(n), hH(n / 2);
    if (!hN.ok || !hH.ok) { printf("plan failed for %d\n", n); return 1; }
    const int h = n / 2, pitch_c = ((h + 1 + 15) / 16) * 16;
    std::vector<float> buf((size_t) nrows * 2 * pitch_c, 0.f), orig;
    srand(n);
    for (int r = 0; r < nrows; r++) for (int z = 0; z < n; z++) buf[(size_t) r * 2 * pitch_c + z] = rand() / (float) RAND_MAX - 0.5f;
    orig = buf;
    ZPassArgs a;
    a.src = buf.data(); a.dst = buf.data(); a.nrows = (size_t) nrows; a.pitch_c = pitch_c; a.scale = 1.f; a.th = dev_plan(hH); a.twN = hN.tw.data();
    const size_t zwork = (size_t) h * R / 4;
    const unsigned thr = zwork >= 2048 ? 512 : (zwork >= 512 ? 256 : 128);
    const size_t smem = (size_t) (h + 1) * (R + 1) * sizeof(float2);
    const unsigned grid = (unsigned) ((nrows + R - 1) / R);
    fpm_emul_launch(grid, thr, smem, [&]() { fft_zfwd_kernel<R>(a); });
    double err = 0, norm = 0;
    for (int r = 0; r < nrows; r += 3) {
        const float2 *row = reinterpret_cast<const float2 *>(buf.data() + (size_t) r * 2 * pitch_c);
        for (int k = 0; k <= h; k += (n > 200 ? 11 : 1)) {
            cd acc = 0;
            for (int z = 0; z < n; z++) acc += (double) orig[(size_t) r * 2 * pitch_c + z] * std::polar(1.0, -2 * M_PI * (double) ((long long) z * k % n) / n);
            err = std::max(err, std::abs(cd(row[k].x, row[k].y) - acc)); norm = std::max(norm, std::abs(acc));
        }
    }
    fpm_emul_launch(grid, thr, smem, [&]() { fft_zbwd_kernel<R>(a); });
    double err2 = 0;
    for (int r = 0; r < nrows; r++) for (int z = 0; z < n; z++)
        err2 = std::max(err2, (double) fabsf(buf[(size_t) r * 2 * pitch_c + z] - n * orig[(size_t) r * 2 * pitch_c + z]));
    const double rel = err / norm, rel2 = err2 / (n * 0.5);
    const bool ok = rel < 3e-6 && rel2 < 6e-6;
    printf("generic rows N=%4d R=%2d forward relerr=%.2e round trip relerr=%.2e %s\n", n, R, rel, rel2, ok ? "OK" : "FAIL");
    return ok ? 0 : 1;
}

int main()
{
    int bad = 0;
    bad += tile_case<16>(48, 2, 0);
    bad += tile_case<16>(160, 1, 1);
    bad += tile_case<8>(768, 1, 0);          // 2^8 * 3
    bad += tile_case<8>(1536, 1, 1);         // 2^9 * 3: B = 3 on nc = 512 (BASELINE.json config 4)
    bad += tile_case<8>(1280, 1, 0);         // 2^8 * 5
    bad += row_case<16>(48, 20);
    bad += row_case<16>(160, 17);
    bad += row_case<16>(768, 16);
    bad += row_case<8>(1536, 9);
    return bad;
}
