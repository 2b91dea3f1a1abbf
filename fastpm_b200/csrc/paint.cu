// fastpm_b200 -- CIC mass deposition and force readout for sm_100a.
// Reference: libfastpm/painter-cic.c:34-110 (cic_paint_tuned), :113-190 (cic_readout_tuned),
// driven per particle by painter.c:320-339 / :358-374.
//
// Arithmetic follows the reference: cell coordinate, cell index and the eight weights are evaluated
// in double (positions need > 24 mantissa bits at Nmesh = 2048), D[1]/T[1] carry the particle weight,
// products are formed in the order  w_z * w_x * w_y, the mesh is float32.  Readout sums the eight
// products in double, in the reference's order, and rounds once to float (store.c:79-91), so it is
// bit-identical to the reference for an identical mesh.  Paint differs from the reference only in the
// order in which float32 additions reach a cell (the reference itself is unordered: `omp atomic`).
//
// Particles are kept in (nearly) Lagrangian order, which is spatially coherent, so that the 8 scatter
// /gather addresses of a warp fall into a few L2-resident mesh planes; red.global.add.f32 is used for
// the scatter.  x-slab decomposition: a particle of this rank has floor(x/h) in [x0, x0+nxl); its +1
// plane may be the halo plane nxl, which is exchanged with the next rank (single GPU: periodic wrap).
#include "common.cuh"
#include "mesh.cuh"
#include "window.h"
#include <stdlib.h>

// Traversal order.  A store filled by fastpm_store_fill (store.c:756-793) holds particle (i, j, k) of the nc^3 Lagrangian grid
// at index (i*nc + j)*nc + k and, on one GPU, is never permuted.  Walking it linearly makes the ~300k particles in flight at
// any moment one thin i-plane whose mesh footprint (displacements of +-10 cells) spans ~20 full mesh planes -- more than
// the L2 at N = 2048.  With the hint `lag_nc` set, a CTA takes a 4 x 8 x 8 brick of the Lagrangian grid and consecutive CTAs
// walk k, then j, then i, so that the CTAs in flight cover a compact 4 x ~70 x nc slab (footprint ~40 MB at N = 2048) and
// successive waves reuse each other's mesh lines in L2.  Any permutation of thread -> particle is a valid traversal: the
// hint affects speed only.
__device__ __forceinline__ long long cic_particle_index(int lag_nc, int nbrick_blocks, long long np)
{
    const int b = blockIdx.x;
    if (b >= nbrick_blocks) {            // no hint, or the tail beyond the last complete group of 4 i-planes: linear
        const long long i = (long long) b * blockDim.x + threadIdx.x;
        return i < np ? i : -1;
    }
    const int nb = lag_nc >> 3;
    const int bk = b % nb, bj = (b / nb) % nb, bi = b / (nb * nb);
    const int tk = threadIdx.x & 7, tj = (threadIdx.x >> 3) & 7, ti = threadIdx.x >> 6;
    return ((long long) (bi * 4 + ti) * lag_nc + (bj * 8 + tj)) * lag_nc + (bk * 8 + tk);
}

struct CicIndex {
    int lx0, lx1, j0, j1, k0, k1;      // lx*: local plane index or -1 when outside this rank
    double D[3], T[3];
};

__device__ __forceinline__ void cic_setup(const FpmGeom &g, const double *pos, CicIndex &c)
{
    const int n = g.n;
    int I[3];
    #pragma unroll
    for (int d = 0; d < 3; d++) {
        const double xyz = pos[d] * g.inv_cellsize;
        const double fl = floor(xyz);
        I[d] = (int) fl;
        c.D[d] = xyz - (double) I[d];
        c.T[d] = 1. - c.D[d];
    }
    int I1[3];
    #pragma unroll
    for (int d = 0; d < 3; d++) {
        I1[d] = I[d] + 1;
        I[d] %= n; if (I[d] < 0) I[d] += n;
        I1[d] %= n; if (I1[d] < 0) I1[d] += n;
    }
    c.j0 = I[1]; c.j1 = I1[1]; c.k0 = I[2]; c.k1 = I1[2];
    if (g.nranks == 1) {
        c.lx0 = I[0]; c.lx1 = I1[0];
    } else {
        // planes [x0, x0+nxl] are addressable (the last one is the halo); periodic in the global index
        int l0 = I[0] - g.x0; if (l0 < 0) l0 += n;
        int l1 = I1[0] - g.x0; if (l1 < 0) l1 += n;
        c.lx0 = (l0 <= g.nxl) ? l0 : -1;
        c.lx1 = (l1 <= g.nxl) ? l1 : -1;
    }
}

// Adds (w0, w1) to row[k0], row[k1].  VEC = 2 / 4: when both cells fall into one aligned float2 / float4 of the row the pair goes
// out as ONE vector reduction (red.global.add.v2/v4.f32, sm_90+), halving the number of reductions the LSU has to issue; the
// padding lanes of a float4 add +0.0f, which leaves the cell unchanged.  Each element is still an independent float32 atomic
// add, so the result is the same as with scalar atomics (up to the order of additions, which is unordered anyway).
template <int VEC>
__device__ __forceinline__ void cic_add_pair(float *row, int k0, int k1, float w0, float w1)
{
    if (VEC == 4 && k1 == k0 + 1 && (k0 & 3) != 3) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int q = k0 & 3;
        if (q == 0) { v.x = w0; v.y = w1; } else if (q == 1) { v.y = w0; v.z = w1; } else { v.z = w0; v.w = w1; }
        atomicAdd(reinterpret_cast<float4 *>(row + (k0 & ~3)), v);
        return;
    }
    if (VEC >= 2 && k1 == k0 + 1 && (k0 & 1) == 0) {
        atomicAdd(reinterpret_cast<float2 *>(row + k0), make_float2(w0, w1));
        return;
    }
    atomicAdd(row + k0, w0);
    atomicAdd(row + k1, w1);
}

// WRAP: fastpm_store_wrap (store.c:447-475) folded into the deposit -- the same operations on the same values as the
// stand-alone wrap kernel (particles.cu), the position is written back only where it changed, and the pass over x that
// the reference spends on wrapping disappears.
template <int VEC, bool WRAP>
__global__ void __launch_bounds__(256) cic_paint_kernel(const FpmGeom g, float *__restrict__ canvas,
        double *__restrict__ x, const float *__restrict__ mass, double M0, const float *__restrict__ field,
        int field_stride, long long np, int *__restrict__ bad, int lag_nc, int nbrick_blocks)
{
    const long long i = cic_particle_index(lag_nc, nbrick_blocks, np);
    if (i < 0) return;
    double pos[3] = { x[3 * i], x[3 * i + 1], x[3 * i + 2] };
    if (WRAP) {
        const double L = g.boxsize;
        #pragma unroll
        for (int d = 0; d < 3; d++) {
            const double xi = pos[d];
            if (xi >= 0 && xi < L) continue;                         // remainder() + fold leave such a value unchanged (x == L goes to 0)
            const double nwrap = (double) abs((int) (xi / L));       // integer abs() on the truncated ratio, store.c:453
            double x1 = remainder(xi, L);
            while (x1 < 0) x1 += L;
            while (x1 > L) x1 -= L;
            if (nwrap > 10000) atomicExch(bad, 1);
            pos[d] = x1;
            x[3 * i + d] = x1;
        }
    }
    CicIndex c;
    cic_setup(g, pos, c);
    double weight = mass ? M0 + (double) mass[i] : M0;        // fastpm_store_get_mass, store.c:120-128
    if (field) weight *= (double) field[i * field_stride];
    c.D[1] *= weight; c.T[1] *= weight;
    const size_t pr = (size_t) g.pitch_r, pl = (size_t) g.n * pr;
    if (c.lx0 >= 0) {
        float *p0 = canvas + (size_t) c.lx0 * pl;
        cic_add_pair<VEC>(p0 + c.j0 * pr, c.k0, c.k1, (float) (c.T[2] * c.T[0] * c.T[1]), (float) (c.D[2] * c.T[0] * c.T[1]));
        cic_add_pair<VEC>(p0 + c.j1 * pr, c.k0, c.k1, (float) (c.T[2] * c.T[0] * c.D[1]), (float) (c.D[2] * c.T[0] * c.D[1]));
    }
    if (c.lx1 >= 0) {
        float *p1 = canvas + (size_t) c.lx1 * pl;
        cic_add_pair<VEC>(p1 + c.j0 * pr, c.k0, c.k1, (float) (c.T[2] * c.D[0] * c.T[1]), (float) (c.D[2] * c.D[0] * c.T[1]));
        cic_add_pair<VEC>(p1 + c.j1 * pr, c.k0, c.k1, (float) (c.T[2] * c.D[0] * c.D[1]), (float) (c.D[2] * c.D[0] * c.D[1]));
    }
}

__global__ void __launch_bounds__(256) cic_readout_kernel(const FpmGeom g, const float *__restrict__ canvas,
        const double *__restrict__ x, float *__restrict__ out, int out_stride, double prescale, long long np, int lag_nc, int nbrick_blocks)
{
    const long long i = cic_particle_index(lag_nc, nbrick_blocks, np);
    if (i < 0) return;
    double pos[3] = { x[3 * i], x[3 * i + 1], x[3 * i + 2] };
    CicIndex c;
    cic_setup(g, pos, c);
    const size_t pr = (size_t) g.pitch_r, pl = (size_t) g.n * pr;
    double value = 0;
    // a uniform pre-scale reproduces fastpm_apply_multiply_transfer on the real field (pm2lpt.c:133)
    #define CELL(ptr) (prescale == 1.0 ? (double) __ldg(ptr) : (double) (float) ((double) __ldg(ptr) * prescale))
    if (c.lx0 >= 0) {
        const float *p0 = canvas + (size_t) c.lx0 * pl;
        value += CELL(p0 + c.j0 * pr + c.k0) * (c.T[2] * c.T[0] * c.T[1]);
        value += CELL(p0 + c.j0 * pr + c.k1) * (c.D[2] * c.T[0] * c.T[1]);
        value += CELL(p0 + c.j1 * pr + c.k0) * (c.T[2] * c.T[0] * c.D[1]);
        value += CELL(p0 + c.j1 * pr + c.k1) * (c.D[2] * c.T[0] * c.D[1]);
    }
    if (c.lx1 >= 0) {
        const float *p1 = canvas + (size_t) c.lx1 * pl;
        value += CELL(p1 + c.j0 * pr + c.k0) * (c.T[2] * c.D[0] * c.T[1]);
        value += CELL(p1 + c.j0 * pr + c.k1) * (c.D[2] * c.D[0] * c.T[1]);
        value += CELL(p1 + c.j1 * pr + c.k0) * (c.T[2] * c.D[0] * c.D[1]);
        value += CELL(p1 + c.j1 * pr + c.k1) * (c.D[2] * c.D[0] * c.D[1]);
    }
    #undef CELL
    out[i * out_stride] = (float) value;
}

// The three force components in ONE pass over the particles (x read once, acc written as whole 12-byte elements): each
// component is the same sum of the same eight products, in the same order, as cic_readout_kernel gives for its canvas.
// Needs the three inverse transforms resident at the same time (two more meshes); opt-in, see csrc/host/gravity.c.
__global__ void __launch_bounds__(256) cic_readout3_kernel(const FpmGeom g, const float *__restrict__ c0, const float *__restrict__ c1,
        const float *__restrict__ c2, const double *__restrict__ x, float *__restrict__ out, long long np, int lag_nc, int nbrick_blocks)
{
    const long long i = cic_particle_index(lag_nc, nbrick_blocks, np);
    if (i < 0) return;
    double pos[3] = { x[3 * i], x[3 * i + 1], x[3 * i + 2] };
    CicIndex c;
    cic_setup(g, pos, c);
    const size_t pr = (size_t) g.pitch_r, pl = (size_t) g.n * pr;
    const float *canv[3] = { c0, c1, c2 };
    double w[8];
    w[0] = c.T[2] * c.T[0] * c.T[1]; w[1] = c.D[2] * c.T[0] * c.T[1]; w[2] = c.T[2] * c.T[0] * c.D[1]; w[3] = c.D[2] * c.T[0] * c.D[1];
    w[4] = c.T[2] * c.D[0] * c.T[1]; w[5] = c.D[2] * c.D[0] * c.T[1]; w[6] = c.T[2] * c.D[0] * c.D[1]; w[7] = c.D[2] * c.D[0] * c.D[1];
    #pragma unroll
    for (int d = 0; d < 3; d++) {
        double value = 0;
        if (c.lx0 >= 0) {
            const float *p0 = canv[d] + (size_t) c.lx0 * pl;
            value += (double) __ldg(p0 + c.j0 * pr + c.k0) * w[0];
            value += (double) __ldg(p0 + c.j0 * pr + c.k1) * w[1];
            value += (double) __ldg(p0 + c.j1 * pr + c.k0) * w[2];
            value += (double) __ldg(p0 + c.j1 * pr + c.k1) * w[3];
        }
        if (c.lx1 >= 0) {
            const float *p1 = canv[d] + (size_t) c.lx1 * pl;
            value += (double) __ldg(p1 + c.j0 * pr + c.k0) * w[4];
            value += (double) __ldg(p1 + c.j0 * pr + c.k1) * w[5];
            value += (double) __ldg(p1 + c.j1 * pr + c.k0) * w[6];
            value += (double) __ldg(p1 + c.j1 * pr + c.k1) * w[7];
        }
        out[3 * i + d] = (float) value;
    }
}

// ------------------------------------------------------------------ the generic windows (painter.c:176-317): linear, quadratic, Lanczos
// One particle per thread, support^3 mesh points.  _fill_k: per axis the window at the `support` points starting at
// floor(x/h + shift) - left, normalised to sum 1; the deposit / gather then runs x-outermost, z-innermost with the weight
// ((1 * kx) * ky) * kz, all in double like the reference.  One GPU only (the x-halo of these windows is wider than one plane).
struct WindowSpec { int type, support, left; double shift, invh; };

__device__ __forceinline__ void window_fill(const FpmGeom &g, const WindowSpec &w, const double pos[3], int ipos[3], double k[3][FPM_WINDOW_MAX_SUPPORT])
{
    #pragma unroll
    for (int d = 0; d < 3; d++) {
        const double gpos = pos[d] * g.inv_cellsize;
        ipos[d] = (int) floor(gpos + w.shift) - w.left;
        const double dx = gpos - ipos[d];
        double sum = 0;
        for (int i = 0; i < w.support; i++) {
            k[d][i] = fpm_window_eval(w.type, dx - i, w.invh);
            sum += k[d][i];
        }
        for (int i = 0; i < w.support; i++) k[d][i] /= sum;
    }
}

__device__ __forceinline__ int window_wrap(int t, int n)
{
    while (t >= n) t -= n;
    while (t < 0) t += n;
    return t;
}

__global__ void __launch_bounds__(128) window_paint_kernel(const FpmGeom g, const WindowSpec w, float *__restrict__ canvas,
        const double *__restrict__ x, const float *__restrict__ mass, double M0, const float *__restrict__ field, int field_stride, long long np)
{
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const double pos[3] = { x[3 * i], x[3 * i + 1], x[3 * i + 2] };
    int ipos[3];
    double k[3][FPM_WINDOW_MAX_SUPPORT];
    window_fill(g, w, pos, ipos, k);
    double weight = mass ? M0 + (double) mass[i] : M0;
    if (field) weight *= (double) field[i * field_stride];
    const size_t pr = (size_t) g.pitch_r, pl = (size_t) g.n * pr;
    for (int a = 0; a < w.support; a++) {
        const int ix = window_wrap(ipos[0] + a, g.n);
        for (int b = 0; b < w.support; b++) {
            const int iy = window_wrap(ipos[1] + b, g.n);
            float *row = canvas + (size_t) ix * pl + (size_t) iy * pr;
            for (int c = 0; c < w.support; c++) {
                const int iz = window_wrap(ipos[2] + c, g.n);
                double kernel = 1.0;
                kernel *= k[0][a]; kernel *= k[1][b]; kernel *= k[2][c];
                atomicAdd(row + iz, (float) (weight * kernel));
            }
        }
    }
}

__global__ void __launch_bounds__(128) window_readout_kernel(const FpmGeom g, const WindowSpec w, const float *__restrict__ canvas,
        const double *__restrict__ x, float *__restrict__ out, int out_stride, long long np)
{
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const double pos[3] = { x[3 * i], x[3 * i + 1], x[3 * i + 2] };
    int ipos[3];
    double k[3][FPM_WINDOW_MAX_SUPPORT];
    window_fill(g, w, pos, ipos, k);
    const size_t pr = (size_t) g.pitch_r, pl = (size_t) g.n * pr;
    double value = 0;
    for (int a = 0; a < w.support; a++) {
        const int ix = window_wrap(ipos[0] + a, g.n);
        for (int b = 0; b < w.support; b++) {
            const int iy = window_wrap(ipos[1] + b, g.n);
            const float *row = canvas + (size_t) ix * pl + (size_t) iy * pr;
            for (int c = 0; c < w.support; c++) {
                const int iz = window_wrap(ipos[2] + c, g.n);
                double kernel = 1.0;
                kernel *= k[0][a]; kernel *= k[1][b]; kernel *= k[2][c];
                value += kernel * (double) __ldg(row + iz);
            }
        }
    }
    out[i * (long long) out_stride] = (float) value;
}

// adds the received halo plane into local plane 0 (multi-GPU paint epilogue)
__global__ void plane_add_kernel(float *__restrict__ dst, const float *__restrict__ src, size_t nfloats)
{
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; i < nfloats; i += stride) dst[i] += src[i];
}

#ifndef FPM_EMULATE          // host side: not part of the CPU emulation of the kernels (tests/emul/particles_emul.cpp)
// see cic_particle_index(): a store laid out as i-planes of nc x nc particles (fastpm_store_fill order; on several GPUs the
// local slab of it, which migration perturbs only slightly) is walked in Lagrangian bricks over its complete groups of 4 planes
static int g_lag_nc = 0, g_lag_force = 0;
void fpm_set_lagrangian_hint(int nc)
{
    static int off = -1;
    if (off < 0) off = getenv("FASTPM_B200_NO_BRICKS") ? 1 : 0;
    g_lag_force = nc < 0;                    // negative: use the bricks whatever the mesh size (tests)
    if (nc < 0) nc = -nc;
    g_lag_nc = (nc > 0 && nc % 8 == 0 && !off) ? nc : 0;
}
// Bricks pay off only when the linear walk's footprint (~24 mesh planes) no longer fits the 126 MB L2: measured on B200, they
// cost 15 % at N = 1024 (4.3 MB planes, linear walk already L2 resident) and gain 25 % at N = 2048 (17 MB planes).
// Returns the number of leading 256-particle blocks that use the brick mapping (0: linear walk).
static int fpm_lagrangian_hint(long long np, const FpmGeom &g, int *lag_nc)
{
    const size_t plane_bytes = (size_t) g.n * g.pitch_r * sizeof(float);
    *lag_nc = 0;
    if (!g_lag_nc || (plane_bytes <= ((size_t) 6 << 20) && !g_lag_force)) return 0;
    const long long group = 4LL * g_lag_nc * g_lag_nc;          // particles in 4 i-planes
    const long long ngroups = np / group;
    if (ngroups == 0) return 0;
    *lag_nc = g_lag_nc;
    return (int) (ngroups * group / 256);
}

// wrap_bad != NULL: wrap the positions on the way (x is then written where it changed)
int fpm_paint_launch(const FpmMesh *m, float *canvas, const double *x, const float *mass, double M0,
                     const float *field, int field_stride, long long np, int *wrap_bad, cudaStream_t st)
{
    if (np <= 0) return 0;
    const unsigned grid = (unsigned) ((np + 255) / 256);
    static int vec = -1;          // FASTPM_B200_PAINT_VEC = 0 | 2 | 4 (default): width of the vector reductions
    if (vec < 0) { const char *e = getenv("FASTPM_B200_PAINT_VEC"); vec = e ? atoi(e) : 4; }
    double *xw = const_cast<double *>(x);
    int lag_nc = 0;
    const int nbrick = fpm_lagrangian_hint(np, m->geom, &lag_nc);
    #define PAINT_LAUNCH(V, W) cic_paint_kernel<V, W><<<grid, 256, 0, st>>>(m->geom, canvas, xw, mass, M0, field, field_stride, np, wrap_bad, lag_nc, nbrick)
    if (nbrick > 0) fpm_path_counter[FPM_PATH_PAINT_BRICKS]++;
    if (fpm_prof_on) fpm_prof_begin(FPM_K_PAINT, st);
    if (wrap_bad) { if (vec >= 4) PAINT_LAUNCH(4, true); else if (vec >= 2) PAINT_LAUNCH(2, true); else PAINT_LAUNCH(0, true); }
    else { if (vec >= 4) PAINT_LAUNCH(4, false); else if (vec >= 2) PAINT_LAUNCH(2, false); else PAINT_LAUNCH(0, false); }
    if (fpm_prof_on) fpm_prof_end(FPM_K_PAINT, st);
    #undef PAINT_LAUNCH
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_readout_launch(const FpmMesh *m, const float *canvas, const double *x, float *out, int out_stride,
                       double prescale, long long np, cudaStream_t st)
{
    if (np <= 0) return 0;
    const unsigned grid = (unsigned) ((np + 255) / 256);
    int lag_nc = 0;
    const int nbrick = fpm_lagrangian_hint(np, m->geom, &lag_nc);
    if (nbrick > 0) fpm_path_counter[FPM_PATH_READOUT_BRICKS]++;
    FPM_TIMED(FPM_K_READOUT, st, (cic_readout_kernel<<<grid, 256, 0, st>>>(m->geom, canvas, x, out, out_stride, prescale, np, lag_nc, nbrick)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_readout3_launch(const FpmMesh *m, const float *c0, const float *c1, const float *c2, const double *x, float *out,
                        long long np, cudaStream_t st)
{
    if (np <= 0) return 0;
    const unsigned grid = (unsigned) ((np + 255) / 256);
    int lag_nc = 0;
    const int nbrick = fpm_lagrangian_hint(np, m->geom, &lag_nc);
    FPM_TIMED(FPM_K_READOUT, st, (cic_readout3_kernel<<<grid, 256, 0, st>>>(m->geom, c0, c1, c2, x, out, np, lag_nc, nbrick)));
    FPM_CHECK_LAUNCH();
    return 0;
}

static int window_spec(const FpmMesh *m, int type, int support, WindowSpec *w)
{
    // fastpm_painter_init, painter.c:128-174
    if (type == FPM_WINDOW_LINEAR) support = 2;
    else if (type == FPM_WINDOW_QUAD) support = 3;
    else if (type != FPM_WINDOW_LANCZOS) { fpm_set_error("window type %d", type); return -1; }
    if (support < 1 || support > FPM_WINDOW_MAX_SUPPORT) { fpm_set_error("window support %d (1..%d on the device)", support, FPM_WINDOW_MAX_SUPPORT); return -1; }
    if (m->geom.nranks > 1) { fpm_set_error("the linear / quadratic / Lanczos windows run on one GPU only"); return -1; }
    w->type = type; w->support = support; w->left = (support - 1) / 2;
    w->shift = support % 2 == 0 ? 0 : 0.5; w->invh = 1 / (0.5 * support);
    return 0;
}

int fpm_window_paint_launch(const FpmMesh *m, int type, int support, float *canvas, const double *x, const float *mass, double M0,
                            const float *field, int field_stride, long long np, cudaStream_t st)
{
    WindowSpec w;
    if (window_spec(m, type, support, &w)) return -1;
    if (np <= 0) return 0;
    FPM_TIMED(FPM_K_PAINT, st, (window_paint_kernel<<<(unsigned) ((np + 127) / 128), 128, 0, st>>>(m->geom, w, canvas, x, mass, M0, field, field_stride, np)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_window_readout_launch(const FpmMesh *m, int type, int support, const float *canvas, const double *x, float *out, int out_stride,
                              long long np, cudaStream_t st)
{
    WindowSpec w;
    if (window_spec(m, type, support, &w)) return -1;
    if (np <= 0) return 0;
    FPM_TIMED(FPM_K_READOUT, st, (window_readout_kernel<<<(unsigned) ((np + 127) / 128), 128, 0, st>>>(m->geom, w, canvas, x, out, out_stride, np)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_plane_add_launch(float *dst, const float *src, size_t nfloats, cudaStream_t st)
{
    FPM_TIMED(FPM_K_HALO, st, (plane_add_kernel<<<148 * 8, 256, 0, st>>>(dst, src, nfloats)));
    FPM_CHECK_LAUNCH();
    return 0;
}
#endif
