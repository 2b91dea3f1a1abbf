// CPU run of the k-space KERNEL SOURCES of fastpm_b200/csrc/kspace.cu (see cuda_emul.h): the stand-alone transfer sweep
// (Green's function, gradients; reference order of operations, mesh.cuh), the CIC deconvolution, and the P(k) binning kernel
// with its warp-segmented shuffle reduction and per-warp histograms (warp shuffles are emulated lane by lane).
//
//   kspace_emul <op> <in.bin> <out.bin>
// in (all ops): int32 n, float64 L, float32 tables[5][n] (k, kk, k_finite, kk_finite, kk_finite2), float64 decic[n],
//               complex64 dk[n][n][pitch_c] in the device layout [ky][kx][kz];  then per op:
//   transfer: int32 potorder, negate, ngrad, dir0, dir1, gradorder, zero_selfconj      -> complex64 out (device layout)
//   decic                                                                               -> complex64 out
//   pgd:      float64 alpha, kl, ks; int32 dir                                          -> complex64 out
//   radial:   int32 mode; float64 param                                                 -> complex64 out
//   unitamp                                                                             -> complex64 out
//   induce:   int32 size; float64 k[size], p[size]                                      -> complex64 out
//   pk:       int32 decic                                                               -> float64 sums[3*(n/2) + 1]
#include "cuda_emul.h"
#include <cstdio>
#include <cstdlib>
#include <string>
#include <cuda_runtime.h>
#include <math.h>
static inline void sincospif(float x, float *s, float *c) { *s = sinf((float) M_PI * x); *c = cosf((float) M_PI * x); }     // whitenoise kernel only
#include "../../fastpm_b200/csrc/kspace.cu"

template <typename T> static std::vector<T> many(FILE *f, size_t n) { std::vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != n) exit(2); return v; }
template <typename T> static T one(FILE *f) { T v; if (fread(&v, sizeof(T), 1, f) != 1) exit(2); return v; }

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    const std::string op = argv[1];
    FILE *in = fopen(argv[2], "rb"), *out = fopen(argv[3], "wb");
    if (!in || !out) return 2;
    const int n = one<int32_t>(in);
    const double L = one<double>(in);
    std::vector<float> tab = many<float>(in, (size_t) 5 * n);
    std::vector<double> dtab = many<double>(in, n);
    FpmGeom g;
    memset(&g, 0, sizeof(g));
    g.n = n; g.nranks = 1; g.nxl = n; g.nyl = n; g.pitch_c = ((n / 2 + 1 + 15) / 16) * 16; g.pitch_r = 2 * g.pitch_c;
    g.boxsize = L; g.cellsize = L / n; g.inv_cellsize = 1.0 / g.cellsize;
    FpmKTables kt;
    kt.k = tab.data(); kt.kk = tab.data() + n; kt.k_finite = tab.data() + 2 * n; kt.kk_finite = tab.data() + 3 * n; kt.kk_finite2 = tab.data() + 4 * n; kt.n = n;
    const size_t total = (size_t) n * n * g.pitch_c;
    std::vector<float2> dk = many<float2>(in, total), res(total);
    if (op == "transfer") {
        FpmTransferSpec s;
        memset(&s, 0, sizeof(s));
        s.active = 1; s.potorder = one<int32_t>(in); s.negate = one<int32_t>(in); s.ngrad = one<int32_t>(in);
        s.graddir[0] = one<int32_t>(in); s.graddir[1] = one<int32_t>(in); s.gradorder = one<int32_t>(in); s.zero_selfconj = one<int32_t>(in); s.scale = 1.0;
        fpm_emul_launch(3, 256, 0, [&]() { transfer_kernel(g, kt, s, dk.data(), res.data(), total); });
        fwrite(res.data(), sizeof(float2), total, out);
    } else if (op == "pgd") {
        // double alpha, kl, ks; int32 dir: the PGD potential sweep, then the gradient i*k_finite[dir] (what fpm_mesh_c2r fuses)
        const double alpha = one<double>(in), kl = one<double>(in), ks = one<double>(in);
        const int dir = one<int32_t>(in);
        std::vector<float2> pot(total);
        fpm_emul_launch(3, 256, 0, [&]() { pgd_transfer_kernel(g, kt, alpha, kl * kl, ks * ks * ks * ks, dk.data(), pot.data(), total); });
        FpmTransferSpec s;
        memset(&s, 0, sizeof(s));
        s.active = 1; s.potorder = -1; s.ngrad = 1; s.graddir[0] = dir; s.gradorder = 1; s.zero_selfconj = 1; s.scale = 1.0;
        fpm_emul_launch(3, 256, 0, [&]() { transfer_kernel(g, kt, s, pot.data(), res.data(), total); });
        fwrite(res.data(), sizeof(float2), total, out);
    } else if (op == "induce") {
        // int32 size, float64 k[size], p[size]: fastpm_ic_induce_correlation with the tabulated P(k), in place
        const int size = one<int32_t>(in);
        std::vector<double> tk = many<double>(in, size), tp = many<double>(in, size);
        res = dk;
        fpm_emul_launch(3, 256, 0, [&]() { induce_kernel(g, kt, res.data(), total, tk.data(), tp.data(), size, L * L * L); });
        fwrite(res.data(), sizeof(float2), total, out);
    } else if (op == "unitamp") {
        // fastpm_ic_remove_variance: in place
        res = dk;
        fpm_emul_launch(3, 256, 0, [&]() { remove_variance_kernel(g, res.data(), total); });
        fwrite(res.data(), sizeof(float2), total, out);
    } else if (op == "radial") {
        // int32 mode, float64 param: the radial softening sweeps (low pass, gaussian36)
        const int mode = one<int32_t>(in);
        const double param = one<double>(in);
        fpm_emul_launch(3, 256, 0, [&]() { radial_transfer_kernel(g, kt, mode, param, dk.data(), res.data(), total); });
        fwrite(res.data(), sizeof(float2), total, out);
    } else if (op == "decic") {
        fpm_emul_launch(3, 256, 0, [&]() { decic_kernel(g, dtab.data(), dk.data(), res.data(), total); });
        fwrite(res.data(), sizeof(float2), total, out);
    } else if (op == "pk") {
        const int decic = one<int32_t>(in), nbins = n / 2;
        const double k0 = 2 * M_PI / L;
        std::vector<double> geom((size_t) 2 * nbins, 0.0), data((size_t) nbins + 1, 0.0), sums((size_t) 3 * nbins + 1);
        fpm_emul_launch(2, 32 * PK_WARPS, sizeof(double) * 2 * nbins * PK_WARPS, [&]() { powerspectrum_kernel<true>(g, dtab.data(), 0, nullptr, k0, geom.data()); });
        fpm_emul_launch(3, 32 * PK_WARPS, sizeof(double) * (nbins + 1) * PK_WARPS, [&]() { powerspectrum_kernel<false>(g, dtab.data(), decic, dk.data(), k0, data.data()); });
        gridDim.x = 1; blockDim.x = 256; blockIdx.x = 0;
        for (unsigned t = 0; t < 256; t++) { threadIdx.x = t; pk_assemble_kernel(geom.data(), data.data(), nbins, sums.data()); }
        fwrite(sums.data(), sizeof(double), sums.size(), out);
    } else if (op == "pk_rows") {
        // the row-streaming P(k) kernel (needs (n/2) % 64 == 0): same output layout as "pk"
        const int decic = one<int32_t>(in), nbins = n / 2, h = n / 2;
        const double k0 = 2 * M_PI / L;
        std::vector<double> geom((size_t) 2 * nbins, 0.0), data((size_t) nbins + 1, 0.0), sums((size_t) 3 * nbins + 1);
        fpm_emul_launch(2, 32 * PK_WARPS, sizeof(double) * 2 * nbins * PK_WARPS, [&]() { powerspectrum_kernel<true>(g, dtab.data(), 0, nullptr, k0, geom.data()); });
        const size_t smem_rows = sizeof(double) * (size_t) (h + 66) + sizeof(double) * (size_t) (nbins + 2) + sizeof(float2) * (size_t) PKR_WARPS * (h + 66);
        fpm_emul_launch(3, 32 * PKR_WARPS, smem_rows, [&]() { powerspectrum_rows_kernel(g, dtab.data(), decic, dk.data(), data.data()); });
        gridDim.x = 1; blockDim.x = 256; blockIdx.x = 0;
        for (unsigned t = 0; t < 256; t++) { threadIdx.x = t; pk_assemble_kernel(geom.data(), data.data(), nbins, sums.data()); }
        fwrite(sums.data(), sizeof(double), sums.size(), out);
    } else return 2;
    fclose(out);
    return 0;
}
