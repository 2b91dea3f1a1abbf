// CPU run of the particle KERNEL SOURCES, fastpm_b200/csrc/paint.cu and csrc/particles.cu (CIC deposit with the vector
// reductions, the fused periodic wrap and the Lagrangian-brick traversal; CIC readout, one canvas and three at once; kick, drift, fused K-K-D-D update; wrap).
// These kernels have no barriers: CUDA threads run one after the other, atomics are plain updates, so the deposit happens in
// particle order -- exactly the reference's serial order (painter.c:320-339 with one OpenMP thread), which makes even the float32
// mesh comparable bit for bit.  Built with -ffp-contract=off (the kernels are compiled with -fmad=false).
//
//   particles_emul <op> <in.bin> <out.bin>     all arrays raw little-endian; layout per op documented at each reader below
#include "cuda_emul.h"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <string>
#include <cuda_runtime.h>
// real atomics: the tile kernels run with one OS thread per CUDA thread (the others one thread after the other, where these are plain updates)
static inline float atomicAdd(float *p, float v)
{
    unsigned int *u = reinterpret_cast<unsigned int *>(p), o = __atomic_load_n(u, __ATOMIC_RELAXED), n;
    float of;
    do { memcpy(&of, &o, 4); const float nf = of + v; memcpy(&n, &nf, 4); } while (!__atomic_compare_exchange_n(u, &o, n, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return of;
}
static inline void atomicAdd(float2 *p, float2 v) { atomicAdd(&p->x, v.x); atomicAdd(&p->y, v.y); }
static inline void atomicAdd(float4 *p, float4 v) { atomicAdd(&p->x, v.x); atomicAdd(&p->y, v.y); atomicAdd(&p->z, v.z); atomicAdd(&p->w, v.w); }
static inline int atomicExch(int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicMin(int *p, int v) { int o = __atomic_load_n(p, __ATOMIC_RELAXED); while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { } return o; }
static inline int atomicMax(int *p, int v) { int o = __atomic_load_n(p, __ATOMIC_RELAXED); while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { } return o; }
#include "../../fastpm_b200/csrc/paint.cu"
#include "../../fastpm_b200/csrc/particles.cu"

template <typename Kernel>
static void launch_seq(unsigned grid, unsigned block, Kernel kernel)
{
    gridDim.x = grid; blockDim.x = block;
    for (unsigned b = 0; b < grid; b++) { blockIdx.x = b; for (unsigned t = 0; t < block; t++) { threadIdx.x = t; kernel(); } }
}

struct Reader {
    FILE *f;
    explicit Reader(const char *fn) : f(fopen(fn, "rb")) { if (!f) { fprintf(stderr, "cannot open %s\n", fn); exit(2); } }
    template <typename T> T one() { T v; if (fread(&v, sizeof(T), 1, f) != 1) exit(2); return v; }
    template <typename T> std::vector<T> many(size_t n) { std::vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != n) exit(2); return v; }
};
template <typename T> static void dump(FILE *f, const std::vector<T> &v) { fwrite(v.data(), sizeof(T), v.size(), f); }

static FpmGeom geom(int n, double L)
{
    FpmGeom g;
    memset(&g, 0, sizeof(g));
    g.n = n; g.nranks = 1; g.rank = 0; g.nxl = n; g.x0 = 0; g.nyl = n; g.y0 = 0;
    g.pitch_c = ((n / 2 + 1 + 15) / 16) * 16; g.pitch_r = 2 * g.pitch_c;
    g.boxsize = L; g.cellsize = L / n; g.inv_cellsize = 1.0 / g.cellsize;
    return g;
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: particles_emul op in out\n"); return 2; }
    const std::string op = argv[1];
    Reader in(argv[2]);
    FILE *out = fopen(argv[3], "wb");
    if (op == "paint" || op == "readout") {
        // in: int32 n, int32 lag_nc (0 linear, else brick walk), int32 wrap, int32 vec, float64 L, float64 M0, int64 np, x[np][3] f64, (readout: canvas f32)
        const int n = in.one<int32_t>(), lag_nc = in.one<int32_t>(), wrap = in.one<int32_t>(), vec = in.one<int32_t>();
        const double L = in.one<double>(), M0 = in.one<double>();
        const long long np = in.one<int64_t>();
        std::vector<double> x = in.many<double>((size_t) 3 * np);
        const FpmGeom g = geom(n, L);
        std::vector<float> canvas((size_t) n * n * g.pitch_r, 0.f);
        const unsigned grid = (unsigned) ((np + 255) / 256);
        int nbrick = 0;
        if (lag_nc) nbrick = (int) ((np / (4LL * lag_nc * lag_nc)) * (4LL * lag_nc * lag_nc) / 256);
        int bad = 0;
        if (op == "paint") {
            auto k = [&]() {
                if (wrap) { if (vec == 4) cic_paint_kernel<4, true>(g, canvas.data(), x.data(), nullptr, M0, nullptr, 1, np, &bad, lag_nc, nbrick, 0, 0);
                            else cic_paint_kernel<0, true>(g, canvas.data(), x.data(), nullptr, M0, nullptr, 1, np, &bad, lag_nc, nbrick, 0, 0); }
                else { if (vec == 4) cic_paint_kernel<4, false>(g, canvas.data(), x.data(), nullptr, M0, nullptr, 1, np, nullptr, lag_nc, nbrick, 0, 0);
                       else if (vec == 2) cic_paint_kernel<2, false>(g, canvas.data(), x.data(), nullptr, M0, nullptr, 1, np, nullptr, lag_nc, nbrick, 0, 0);
                       else cic_paint_kernel<0, false>(g, canvas.data(), x.data(), nullptr, M0, nullptr, 1, np, nullptr, lag_nc, nbrick, 0, 0); }
            };
            launch_seq(grid, 256, k);
            std::vector<float> dense((size_t) n * n * n);            // unpadded [x][y][z]
            for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) memcpy(&dense[((size_t) i * n + j) * n], &canvas[((size_t) i * n + j) * g.pitch_r], sizeof(float) * n);
            dump(out, dense);
            dump(out, x);                                              // positions after the fused wrap
            std::vector<int32_t> flag(1, bad); dump(out, flag);
        } else {
            std::vector<float> dense = in.many<float>((size_t) n * n * n);
            for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) memcpy(&canvas[((size_t) i * n + j) * g.pitch_r], &dense[((size_t) i * n + j) * n], sizeof(float) * n);
            std::vector<float> res((size_t) np, 0.f);
            launch_seq(grid, 256, [&]() { cic_readout_kernel(g, canvas.data(), x.data(), res.data(), 1, 1.0, np, lag_nc, nbrick, 0, 0, nullptr, nullptr); });
            dump(out, res);
        }
    } else if (op == "tpaint" || op == "treadout") {
        // the shared-memory tile kernels (one OS thread per CUDA thread).  in: as for paint / readout, plus int32 nranks, rank after M0's
        // np; lag_nc a multiple of 8, np a multiple of 8 lag_nc^2.  Several "ranks": x-slab geometry of one of them, canvas with the
        // halo plane, particles of that slab only.  out: the canvas rows / readout values, then uint64 stats[2] (particles on the
        // global path, CTAs without a tile)
        const int n = in.one<int32_t>(), lag_nc = in.one<int32_t>(), wrap = in.one<int32_t>(), vec = in.one<int32_t>();
        const double L = in.one<double>(), M0 = in.one<double>();
        const long long np = in.one<int64_t>();
        const int nranks = in.one<int32_t>(), rank = in.one<int32_t>();
        (void) vec;
        std::vector<double> x = in.many<double>((size_t) 3 * np);
        FpmGeom g = geom(n, L);
        g.nranks = nranks; g.rank = rank; g.nxl = n / nranks; g.x0 = rank * g.nxl;
        const int planes = g.nxl + (nranks > 1 ? 1 : 0);
        std::vector<float> canvas((size_t) planes * n * g.pitch_r, 0.f);
        const unsigned grid = (unsigned) (np / FPM_TILE_THREADS);
        unsigned long long stats[2] = { 0, 0 };
        int bad = 0;
        if (op == "tpaint") {
            fpm_emul_launch(grid, FPM_TILE_THREADS, FPM_TILE_CAP * sizeof(float), [&]() {
                if (wrap) cic_paint_tile_kernel<true>(g, canvas.data(), x.data(), nullptr, M0, nullptr, 1, &bad, lag_nc, stats);
                else cic_paint_tile_kernel<false>(g, canvas.data(), x.data(), nullptr, M0, nullptr, 1, nullptr, lag_nc, stats);
            });
            std::vector<float> dense((size_t) planes * n * n);
            for (int i = 0; i < planes; i++) for (int j = 0; j < n; j++) memcpy(&dense[((size_t) i * n + j) * n], &canvas[((size_t) i * n + j) * g.pitch_r], sizeof(float) * n);
            dump(out, dense);
            dump(out, x);
            std::vector<int32_t> flag(1, bad); dump(out, flag);
        } else {
            std::vector<float> dense = in.many<float>((size_t) planes * n * n);
            for (int i = 0; i < planes; i++) for (int j = 0; j < n; j++) memcpy(&canvas[((size_t) i * n + j) * g.pitch_r], &dense[((size_t) i * n + j) * n], sizeof(float) * n);
            std::vector<float> res((size_t) np, 0.f);
            fpm_emul_launch(grid, FPM_TILE_THREADS, FPM_TILE_CAP * sizeof(float), [&]() { cic_readout_tile_kernel(g, canvas.data(), x.data(), res.data(), 1, 1.0, lag_nc, stats); });
            dump(out, res);
        }
        std::vector<uint64_t> st(stats, stats + 2); dump(out, st);
    } else if (op == "wpaint" || op == "wreadout") {
        // in: int32 n, int32 type, int32 support, int32 diffdir, int32 nslab, float64 L, float64 M0, int64 np, x[np][3] f64,
        //     (wreadout: dense canvas f32[n^3])
        // nslab > 1: the mesh is cut into x-slabs as on several GPUs; every slab deposits its own particles into its planes and its
        // halo block, and the halo planes are exchanged as csrc/comm.cu does (fpm_halo_add_wide_from / fpm_halo_fetch_wide_from)
        const int n = in.one<int32_t>(), type = in.one<int32_t>();
        int support = in.one<int32_t>();
        const int diffdir = in.one<int32_t>(), nslab = in.one<int32_t>();
        const double L = in.one<double>(), M0 = in.one<double>();
        const long long np = in.one<int64_t>();
        std::vector<double> x = in.many<double>((size_t) 3 * np);
        FpmGeom g = geom(n, L);
        if (type == FPM_WINDOW_LINEAR || type == FPM_WINDOW_CIC) support = 2; else if (type == FPM_WINDOW_QUAD) support = 3;
        const bool cic = type == FPM_WINDOW_CIC;
        WindowSpec w = { type, support, cic ? 0 : (support - 1) / 2, diffdir, 0, 0, (!cic && support % 2) ? 0.5 : 0, 1 / (0.5 * support) };
        w.hl = w.left; w.hr = support - 1 - w.left + ((!cic && support % 2) ? 1 : 0);
        const size_t pl = (size_t) n * g.pitch_r;
        std::vector<float> canvas((size_t) n * pl, 0.f);
        const int nxl = n / nslab;
        // particles of every slab, in their original order (owner: floor(x / h) mod N / nxl, comm.cu classify_kernel)
        std::vector<std::vector<long long>> mine(nslab);
        for (long long i = 0; i < np; i++) {
            int ix = (int) floor(x[3 * i] * g.inv_cellsize);
            ix %= n; if (ix < 0) ix += n;
            mine[ix / nxl].push_back(i);
        }
        std::vector<std::vector<float>> halo(nslab, std::vector<float>((size_t) (w.hl + w.hr) * pl + 1, 0.f));
        std::vector<float> res((size_t) np, 0.f);
        int outside = 0;
        if (op == "wreadout") {
            std::vector<float> dense = in.many<float>((size_t) n * n * n);
            for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) memcpy(&canvas[((size_t) i * n + j) * g.pitch_r], &dense[((size_t) i * n + j) * n], sizeof(float) * n);
            // fetch: left part <- the previous slab's last hl planes, right part <- the next slab's first hr planes
            if (nslab > 1) for (int r = 0; r < nslab; r++) {
                const int prev = (r - 1 + nslab) % nslab, next = (r + 1) % nslab;
                memcpy(halo[r].data(), &canvas[(size_t) (prev * nxl + nxl - w.hl) * pl], sizeof(float) * w.hl * pl);
                memcpy(halo[r].data() + (size_t) w.hl * pl, &canvas[(size_t) (next * nxl) * pl], sizeof(float) * w.hr * pl);
            }
        }
        for (int r = 0; r < nslab; r++) {
            if (nslab > 1) { g.nranks = nslab; g.rank = r; g.nxl = nxl; g.x0 = r * nxl; }
            const long long npr = nslab > 1 ? (long long) mine[r].size() : np;
            std::vector<double> xr;
            if (nslab > 1) { xr.resize((size_t) 3 * npr); for (long long i = 0; i < npr; i++) for (int d = 0; d < 3; d++) xr[3 * i + d] = x[3 * mine[r][i] + d]; }
            const double *xp = nslab > 1 ? xr.data() : x.data();
            float *cv = canvas.data() + (size_t) (nslab > 1 ? r * nxl : 0) * pl;
            float *hp = nslab > 1 ? halo[r].data() : nullptr;
            const unsigned grid = (unsigned) ((npr + 127) / 128);
            if (npr == 0) continue;
            if (op == "wpaint") {
                launch_seq(grid, 128, [&]() { window_paint_kernel(g, w, cv, hp, xp, nullptr, M0, nullptr, 1, npr, &outside); });
            } else {
                std::vector<float> rr((size_t) npr, 0.f);
                launch_seq(grid, 128, [&]() { window_readout_kernel(g, w, cv, hp, xp, rr.data(), 1, npr, &outside); });
                for (long long i = 0; i < npr; i++) res[nslab > 1 ? mine[r][i] : i] = rr[i];
            }
        }
        if (outside) { fprintf(stderr, "%d planes outside slab + halo\n", outside); return 3; }
        if (op == "wpaint") {
            // add: my first hr planes += the previous slab's right part, my last hl planes += the next slab's left part
            if (nslab > 1) for (int r = 0; r < nslab; r++) {
                const int prev = (r - 1 + nslab) % nslab, next = (r + 1) % nslab;
                for (size_t i = 0; i < (size_t) w.hr * pl; i++) canvas[(size_t) (r * nxl) * pl + i] += halo[prev][(size_t) w.hl * pl + i];
                for (size_t i = 0; i < (size_t) w.hl * pl; i++) canvas[(size_t) (r * nxl + nxl - w.hl) * pl + i] += halo[next][i];
            }
            std::vector<float> dense((size_t) n * n * n);
            for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) memcpy(&dense[((size_t) i * n + j) * n], &canvas[((size_t) i * n + j) * g.pitch_r], sizeof(float) * n);
            dump(out, dense);
        } else {
            dump(out, res);
        }
    } else if (op == "readout3") {
        // in: int32 n, int32 lag_nc, float64 L, int64 np, x[np][3] f64, three dense canvases f32[n^3]
        // out: three separate readouts into a stride-3 column (f32[np][3]), then the one-pass kernel (f32[np][3])
        const int n = in.one<int32_t>(), lag_nc = in.one<int32_t>();
        const double L = in.one<double>();
        const long long np = in.one<int64_t>();
        std::vector<double> x = in.many<double>((size_t) 3 * np);
        const FpmGeom g = geom(n, L);
        std::vector<float> canvas[3];
        for (int d = 0; d < 3; d++) {
            std::vector<float> dense = in.many<float>((size_t) n * n * n);
            canvas[d].assign((size_t) n * n * g.pitch_r, 0.f);
            for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) memcpy(&canvas[d][((size_t) i * n + j) * g.pitch_r], &dense[((size_t) i * n + j) * n], sizeof(float) * n);
        }
        const unsigned grid = (unsigned) ((np + 255) / 256);
        int nbrick = 0;
        if (lag_nc) nbrick = (int) ((np / (4LL * lag_nc * lag_nc)) * (4LL * lag_nc * lag_nc) / 256);
        std::vector<float> sep((size_t) 3 * np, -7.f), one((size_t) 3 * np, -9.f);
        for (int d = 0; d < 3; d++)
            launch_seq(grid, 256, [&]() { cic_readout_kernel(g, canvas[d].data(), x.data(), sep.data() + d, 3, 1.0, np, lag_nc, nbrick, 0, 0, nullptr, nullptr); });
        launch_seq(grid, 256, [&]() { cic_readout3_kernel(g, canvas[0].data(), canvas[1].data(), canvas[2].data(), x.data(), one.data(), np, lag_nc, nbrick); });
        dump(out, sep);
        dump(out, one);
    } else if (op == "pgddrift") {
        // in: int64 np, int32 drift_mode, drift f64[5] (dyyy da1 da2 Dv1 Dv2), f64 dyyy_last, x f64[3np], v, dx1, dx2, pgdc f32[3np]
        // out: x after drift_kernel followed by pgd_shift_kernel (factors.c:75-114 with a pgdc column)
        const long long np = in.one<int64_t>();
        const int dmode = in.one<int32_t>();
        std::vector<double> df = in.many<double>(5);
        const double dyyy_last = in.one<double>();
        std::vector<double> x = in.many<double>((size_t) 3 * np);
        std::vector<float> v = in.many<float>((size_t) 3 * np), d1 = in.many<float>((size_t) 3 * np), d2 = in.many<float>((size_t) 3 * np), pg = in.many<float>((size_t) 3 * np);
        DriftArgs a = { x.data(), x.data(), v.data(), d1.data(), d2.data(), df[0], df[1], df[2], df[3], df[4], dmode, 3 * np };
        launch_seq(5, 256, [&]() { drift_kernel(a); });
        launch_seq(5, 256, [&]() { pgd_shift_kernel(x.data(), pg.data(), df[0], dyyy_last, 3 * np); });
        dump(out, x);
    } else if (op == "update") {
        // in: int64 np, int32 cola, int32 drift_mode, kick f64[5] (dda q1 q2 Dv1 Dv2), drift f64[5] (dyyy da1 da2 Dv1 Dv2),
        //     x f64[3np], v f32[3np], acc f32[3np], dx1 f32[3np], dx2 f32[3np]
        // out: separate kernels K, D then K, D/2, D/2 (x, v), then the same five operations as ONE fused pass (x, v)
        const long long np = in.one<int64_t>();
        const int cola = in.one<int32_t>(), dmode = in.one<int32_t>();
        std::vector<double> kf = in.many<double>(5), df = in.many<double>(5);
        std::vector<double> x0 = in.many<double>((size_t) 3 * np);
        std::vector<float> v0 = in.many<float>((size_t) 3 * np), acc = in.many<float>((size_t) 3 * np), d1 = in.many<float>((size_t) 3 * np), d2 = in.many<float>((size_t) 3 * np);
        std::vector<double> x = x0; std::vector<float> v = v0;
        const unsigned grid = 7;
        auto kick = [&]() { KickArgs a = { v.data(), v.data(), acc.data(), d1.data(), d2.data(), kf[0], kf[1], kf[2], kf[3], kf[4], cola, 3 * np };
                            launch_seq(grid, 256, [&]() { kick_kernel(a); }); };
        auto drift = [&](double s) { DriftArgs a = { x.data(), x.data(), v.data(), d1.data(), d2.data(), s * df[0], s * df[1], s * df[2], df[3], df[4], dmode, 3 * np };
                                     launch_seq(grid, 256, [&]() { drift_kernel(a); }); };
        kick(); drift(1.0);
        dump(out, x); dump(out, v);                                    // after one kick + one drift (what the oracle does)
        kick(); drift(0.5); drift(0.5);
        dump(out, x); dump(out, v);
        FusedArgs fa;
        memset(&fa, 0, sizeof(fa));
        std::vector<double> xf = x0; std::vector<float> vf = v0;
        fa.x = xf.data(); fa.v = vf.data(); fa.acc = acc.data(); fa.dx1 = d1.data(); fa.dx2 = d2.data(); fa.n3 = 3 * np; fa.nops = 5;
        fa.any_kick = 1; fa.any_drift = 1; fa.any_dx = cola || dmode >= 2;
        const double scl[5] = { 0, 1.0, 0, 0.5, 0.5 };
        for (int j = 0; j < 5; j++) {
            const bool isk = (j == 0 || j == 2);
            fa.ops[j].kind = isk ? 0 : 1; fa.ops[j].mode = isk ? cola : dmode;
            for (int q = 0; q < 5; q++) fa.ops[j].f[q] = isk ? kf[q] : (q < 3 ? scl[j] * df[q] : df[q]);
        }
        launch_seq(grid, 256, [&]() { fused_update_kernel(fa); });
        dump(out, xf); dump(out, vf);
    } else if (op == "wrap") {
        // in: int64 n3, float64 L, x f64[n3]; out: x, int32 flag
        const long long n3 = in.one<int64_t>();
        const double L = in.one<double>();
        std::vector<double> x = in.many<double>((size_t) n3);
        int bad = 0;
        launch_seq(3, 256, [&]() { wrap_kernel(x.data(), n3, L, &bad); });
        dump(out, x);
        std::vector<int32_t> flag(1, bad); dump(out, flag);
    } else { fprintf(stderr, "unknown op %s\n", op.c_str()); return 2; }
    fclose(out);
    return 0;
}
