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
// Round 2: the bricks are walked in SUPER-BLOCKS of sb_i x sb_j bricks (times all of k): k fastest, then j inside the block, then
// i inside the block, then the next block.  A mesh plane is touched by every brick whose particles reach it -- with the plain
// k, j, i order that is several consecutive groups of i-planes, a full j-sweep (more than the L2) apart, so the plane came in from
// DRAM several times per deposit; inside a super-block (default 8 x 4 bricks: 32 x 32 x nc particles, ~35 MB of mesh at B = 2) the
// revisits hit L2.  sb packs (sb_i << 16 | sb_j); 0 = the plain order.
__device__ __forceinline__ long long cic_particle_index(int lag_nc, int nbrick_blocks, long long np, int sb = 0)
{
    const int b = blockIdx.x;
    if (b >= nbrick_blocks) {            // no hint, or the tail beyond the last complete group of 4 i-planes: linear
        const long long i = (long long) b * blockDim.x + threadIdx.x;
        return i < np ? i : -1;
    }
    const int nb = lag_nc >> 3;
    int bk, bj, bi;
    if (sb == 0) {
        bk = b % nb; bj = (b / nb) % nb; bi = b / (nb * nb);
    } else {
        const int sbi = sb >> 16, sbj = sb & 0xffff;
        const int per = sbi * sbj * nb;                  // bricks per super-block
        const int s = b / per, r = b - s * per;
        bk = r % nb;
        const int q = r / nb, bjl = q % sbj, bil = q / sbj;
        const int nsj = nb / sbj, sj = s % nsj, si = s / nsj;
        bj = sj * sbj + bjl; bi = si * sbi + bil;
    }
    const int tk = threadIdx.x & 7, tj = (threadIdx.x >> 3) & 7, ti = threadIdx.x >> 6;
    return ((long long) (bi * 4 + ti) * lag_nc + (bj * 8 + tj)) * lag_nc + (bk * 8 + tk);
}

struct CicIndex {
    int u[3];                          // floor(x / h) before the periodic wrap
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
        c.u[d] = I[d];
        c.D[d] = xyz - (double) I[d];
        c.T[d] = 1. - c.D[d];
    }
    // periodic images (painter-cic.c:65-70 loops `while (I >= N) I -= N; while (I < 0) I += N`): a wrapped position gives
    // I in [0, n], so one conditional subtraction serves nearly every particle; anything further out takes the division
    int I1[3];
    #pragma unroll
    for (int d = 0; d < 3; d++) {
        int i0 = I[d];
        if ((unsigned) i0 >= (unsigned) n) {
            if (i0 == n) i0 = 0;
            else { i0 %= n; if (i0 < 0) i0 += n; }
        }
        I[d] = i0;
        I1[d] = (i0 + 1 == n) ? 0 : i0 + 1;
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
        int field_stride, long long np, int *__restrict__ bad, int lag_nc, int nbrick_blocks, long long first, int sb)
{
    long long i = cic_particle_index(lag_nc, nbrick_blocks, np - first, sb);
    if (i < 0) return;
    i += first;
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

// pack0 / pack1 != NULL: `out` is a column of 3-float rows; the row is written WHOLE as { pack0[i], pack1[i], value }.  The three
// force components then cost 4 + 4 + (8 read, 12 written) bytes per particle instead of three strided 4-byte stores into 12-byte
// rows, each of which dirtied the whole 32-byte sector (measured at nc = 1024: 12.9 GB written and 12.9 GB of fill reads per pass
// for 4.3 GB of results, gpurun_out/r02_traffic_particles_nc1024.csv).
__global__ void __launch_bounds__(256) cic_readout_kernel(const FpmGeom g, const float *__restrict__ canvas,
        const double *__restrict__ x, float *__restrict__ out, int out_stride, double prescale, long long np, int lag_nc, int nbrick_blocks,
        long long first, int sb, const float *__restrict__ pack0, const float *__restrict__ pack1)
{
    long long i = cic_particle_index(lag_nc, nbrick_blocks, np - first, sb);
    if (i < 0) return;
    i += first;
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
    if (pack0) {
        float *o = out + 3 * i;
        o[0] = pack0[i]; o[1] = pack1[i]; o[2] = (float) value;
    } else {
        out[i * out_stride] = (float) value;
    }
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

// ------------------------------------------------------------------ shared-memory mesh tiles (round 2)
// Measured on B200 (scripts/ubench/atomics.cu): a float atomicAdd on shared memory costs 0.19 cycles per lane and SM, a
// red.global.add (any vector width) 1.2 - 1.6 when the mesh lines sit in L2 and 3 - 15 when they do not -- the deposit above is
// bound by the latter.  With the Lagrangian hint a CTA of 512 threads takes an 8 x 8 x 8 brick of the particle grid.  Such a
// brick stays compact under the displacement field, so the CTA
//   1. finds the box of mesh cells its particles touch: mean cell offset from thread 0's cell (a few stray particles barely move
//      it), then min / max over the particles within FPM_TILE_REACH cells of the mean, z origin aligned to 4 cells;
//   2. deposit: accumulates the 8 weights of every in-box particle into a zeroed tile in shared memory (atomicAdd on shared
//      memory) and flushes the tile ONCE, row by row, as aligned red.global.add.v4.f32 -- whole 16-byte groups, consecutive lanes
//      on consecutive addresses, all-zero groups skipped;  gather: copies the box into the tile with aligned 16-byte loads and
//      reads the 8 cells of every in-box particle from shared memory;
//   3. particles outside the box (and whole CTAs whose box exceeds the tile) take the global path of the kernels above.
// Arithmetic is unchanged: same double weights in the same product order, the gather sums the same eight products in the same
// order (bit-identical results); the deposit adds the same float32 increments to each cell in another order.
// Stores that are not in (nearly) Lagrangian order simply fail step 1 and run at the speed of the kernels above.
#define FPM_TILE_THREADS 512
#define FPM_TILE_CAP 14336          // floats of the tile: 56 KB, 4 CTAs per SM
#define FPM_TILE_REACH 15           // first attempt: cells from the mean offset; a box too large for the tile is retried with 10

struct TileFrame {
    int ref[3];                      // unwrapped cell of thread 0's particle (x: local plane on several GPUs)
    int sum[3];
    int lo[3], hi[3];                // min / max offset from ref over the in-reach particles
    int org[3];                      // coordinate of tile cell (0, 0, 0); org[2] is a multiple of 4
    int ext[3];                      // tile extents, ext[2] a multiple of 4; ext[0] == 0: no tile for this CTA
    int nin;
};

#ifndef FPM_EMULATE
__device__ __forceinline__ int warp_sum_i(int v) { return __reduce_add_sync(0xffffffffu, v); }
__device__ __forceinline__ int warp_min_i(int v) { return __reduce_min_sync(0xffffffffu, v); }
__device__ __forceinline__ int warp_max_i(int v) { return __reduce_max_sync(0xffffffffu, v); }
#else
static inline int warp_sum_i(int v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }
static inline int warp_min_i(int v) { for (int o = 16; o > 0; o >>= 1) { const int w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; } return v; }
static inline int warp_max_i(int v) { for (int o = 16; o > 0; o >>= 1) { const int w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; } return v; }
#endif

// particle of this thread: 8 x 8 x 8 bricks of the Lagrangian grid, consecutive CTAs along k, then j, then i
__device__ __forceinline__ long long tile_particle_index(int lag_nc)
{
    const int b = blockIdx.x, nb = lag_nc >> 3;
    const int bk = b % nb, bj = (b / nb) % nb, bi = b / (nb * nb);
    const int tk = threadIdx.x & 7, tj = (threadIdx.x >> 3) & 7, ti = threadIdx.x >> 6;
    return ((long long) (bi * 8 + ti) * lag_nc + (bj * 8 + tj)) * lag_nc + (bk * 8 + tk);
}

__device__ __forceinline__ int tile_wrap(int v, int n)
{
    while (v >= n) v -= n;
    while (v < 0) v += n;
    return v;
}

// Steps 1 of the comment above.  u: this thread's unwrapped cell (x already local on several GPUs), valid: the particle may use
// the tile at all.  Returns whether it does; d[] = its offset from f->ref.
__device__ __forceinline__ bool tile_setup(const FpmGeom &g, const int u[3], bool valid, TileFrame *f, int d[3])
{
    const int n = g.n, tid = threadIdx.x, lane = tid & 31, half = n >> 1;
    if (tid == 0) {
        f->ref[0] = u[0]; f->ref[1] = u[1]; f->ref[2] = u[2];
        f->sum[0] = f->sum[1] = f->sum[2] = 0;
        f->ext[0] = valid ? 1 : 0;
        f->nin = 0;
    }
    __syncthreads();
    #pragma unroll
    for (int a = 0; a < 3; a++) {
        d[a] = u[a] - f->ref[a];
        if (a > 0 || g.nranks == 1) { if (d[a] > half) d[a] -= n; else if (d[a] < -half) d[a] += n; }
    }
    {
        const int s0 = warp_sum_i(valid ? d[0] : 0), s1 = warp_sum_i(valid ? d[1] : 0), s2 = warp_sum_i(valid ? d[2] : 0);
        if (lane == 0) { atomicAdd(&f->sum[0], s0); atomicAdd(&f->sum[1], s1); atomicAdd(&f->sum[2], s2); }
    }
    __syncthreads();
    if (f->ext[0] == 0) return false;                                // thread 0's particle is not usable as the reference
    const int m0 = f->sum[0] / FPM_TILE_THREADS, m1 = f->sum[1] / FPM_TILE_THREADS, m2 = f->sum[2] / FPM_TILE_THREADS;
    bool in = false;
    for (int reach = FPM_TILE_REACH; ; reach = 10) {
        __syncthreads();                                             // everybody has read the previous attempt's verdict
        if (tid == 0) {
            #pragma unroll
            for (int a = 0; a < 3; a++) { f->lo[a] = 0x7fffffff; f->hi[a] = -0x7fffffff; }
        }
        __syncthreads();
        in = valid && abs(d[0] - m0) <= reach && abs(d[1] - m1) <= reach && abs(d[2] - m2) <= reach;
        #pragma unroll
        for (int a = 0; a < 3; a++) {
            const int lo = warp_min_i(in ? d[a] : 0x7fffffff), hi = warp_max_i(in ? d[a] : -0x7fffffff);
            if (lane == 0 && lo <= hi) { atomicMin(&f->lo[a], lo); atomicMax(&f->hi[a], hi); }
        }
        __syncthreads();
        if (tid == 0) {
            if (f->lo[0] > f->hi[0]) f->ext[0] = 0;                  // nobody in reach
            else {
                f->org[0] = f->ref[0] + f->lo[0]; f->ext[0] = f->hi[0] - f->lo[0] + 2;
                f->org[1] = f->ref[1] + f->lo[1]; f->ext[1] = f->hi[1] - f->lo[1] + 2;
                const int z0 = f->ref[2] + f->lo[2], z0a = z0 - (((z0 % 4) + 4) % 4);
                f->org[2] = z0a; f->ext[2] = ((f->ref[2] + f->hi[2] + 2 - z0a) + 3) & ~3;
                // several GPUs: the planes of the tile must exist locally (0 .. nxl, the last one being the halo plane)
                if (g.nranks > 1 && (f->org[0] < 0 || f->org[0] + f->ext[0] - 1 > g.nxl)) f->ext[0] = 0;
                else if ((long long) f->ext[0] * f->ext[1] * f->ext[2] > FPM_TILE_CAP) f->ext[0] = (reach > 10) ? -1 : 0;
            }
        }
        __syncthreads();
        if (f->ext[0] >= 0) break;                                   // a tile, or none at all
    }
    return in && f->ext[0] > 0;
}

// rows of the tile are walked by the warps in groups: RPW rows per warp and pass, LPR lanes (16-byte groups) per row
struct TileRows {
    int nq, lpr, rpw, nrow, tx, ty, r, sx, sy, step, q;
    __device__ __forceinline__ TileRows(const TileFrame *f)
    {
        nq = f->ext[2] >> 2;
        lpr = nq <= 8 ? 8 : (nq <= 16 ? 16 : 32);
        rpw = 32 / lpr;
        nrow = f->ext[0] * f->ext[1];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
        q = lane % lpr;
        r = warp * rpw + lane / lpr;
        tx = r / f->ext[1]; ty = r - tx * f->ext[1];
        step = nwarp * rpw;
        sx = step / f->ext[1]; sy = step - sx * f->ext[1];
    }
    __device__ __forceinline__ bool active() const { return r < nrow && q < nq; }
    __device__ __forceinline__ bool more() const { return (r - (int) ((threadIdx.x & 31) / lpr)) < nrow; }     // warp-uniform
    __device__ __forceinline__ void next(const TileFrame *f)
    {
        r += step; tx += sx; ty += sy;
        if (ty >= f->ext[1]) { ty -= f->ext[1]; tx++; }
    }
};

__device__ __forceinline__ size_t tile_row_offset(const FpmGeom &g, const TileFrame *f, int tx, int ty)
{
    const int px = g.nranks == 1 ? tile_wrap(f->org[0] + tx, g.n) : f->org[0] + tx;
    const int py = tile_wrap(f->org[1] + ty, g.n);
    return (size_t) px * ((size_t) g.n * g.pitch_r) + (size_t) py * g.pitch_r;
}

template <bool WRAP>
__global__ void __launch_bounds__(FPM_TILE_THREADS) cic_paint_tile_kernel(const FpmGeom g, float *__restrict__ canvas,
        double *__restrict__ x, const float *__restrict__ mass, double M0, const float *__restrict__ field,
        int field_stride, int *__restrict__ bad, int lag_nc, unsigned long long *__restrict__ stats)
{
    FPM_DYN_SMEM(tile_raw, 16);                   // FPM_TILE_CAP floats (more than the 48 KB a static array may have)
    float *tile = reinterpret_cast<float *>(tile_raw);
    __shared__ TileFrame frame;
    const long long i = tile_particle_index(lag_nc);
    double pos[3] = { x[3 * i], x[3 * i + 1], x[3 * i + 2] };
    if (WRAP) {
        const double L = g.boxsize;
        #pragma unroll
        for (int d = 0; d < 3; d++) {
            const double xi = pos[d];
            if (xi >= 0 && xi < L) continue;
            const double nwrap = (double) abs((int) (xi / L));
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
    double weight = mass ? M0 + (double) mass[i] : M0;
    if (field) weight *= (double) field[i * field_stride];
    c.D[1] *= weight; c.T[1] *= weight;
    int u[3] = { g.nranks == 1 ? c.u[0] : c.lx0, c.u[1], c.u[2] };
    const bool valid = g.nranks == 1 || (c.lx0 >= 0 && c.lx0 < g.nxl);
    int d[3];
    const bool in = tile_setup(g, u, valid, &frame, d);
    const int ex = frame.ext[0], ey = frame.ext[1], ez = frame.ext[2];
    if (ex > 0) {
        float4 *t4 = reinterpret_cast<float4 *>(tile);
        const int nf4 = (ex * ey * ez) >> 2;
        for (int q = threadIdx.x; q < nf4; q += FPM_TILE_THREADS) t4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
    }
    if (in) {
        float *p = tile + ((d[0] - frame.lo[0]) * ey + (d[1] - frame.lo[1])) * ez + (frame.ref[2] + d[2] - frame.org[2]);
        const int sy = ez, sx = ey * ez;
        atomicAdd(p, (float) (c.T[2] * c.T[0] * c.T[1]));
        atomicAdd(p + 1, (float) (c.D[2] * c.T[0] * c.T[1]));
        atomicAdd(p + sy, (float) (c.T[2] * c.T[0] * c.D[1]));
        atomicAdd(p + sy + 1, (float) (c.D[2] * c.T[0] * c.D[1]));
        atomicAdd(p + sx, (float) (c.T[2] * c.D[0] * c.T[1]));
        atomicAdd(p + sx + 1, (float) (c.D[2] * c.D[0] * c.T[1]));
        atomicAdd(p + sx + sy, (float) (c.T[2] * c.D[0] * c.D[1]));
        atomicAdd(p + sx + sy + 1, (float) (c.D[2] * c.D[0] * c.D[1]));
    } else {
        const size_t pr = (size_t) g.pitch_r, pl = (size_t) g.n * pr;
        if (c.lx0 >= 0) {
            float *p0 = canvas + (size_t) c.lx0 * pl;
            cic_add_pair<4>(p0 + c.j0 * pr, c.k0, c.k1, (float) (c.T[2] * c.T[0] * c.T[1]), (float) (c.D[2] * c.T[0] * c.T[1]));
            cic_add_pair<4>(p0 + c.j1 * pr, c.k0, c.k1, (float) (c.T[2] * c.T[0] * c.D[1]), (float) (c.D[2] * c.T[0] * c.D[1]));
        }
        if (c.lx1 >= 0) {
            float *p1 = canvas + (size_t) c.lx1 * pl;
            cic_add_pair<4>(p1 + c.j0 * pr, c.k0, c.k1, (float) (c.T[2] * c.D[0] * c.T[1]), (float) (c.D[2] * c.D[0] * c.T[1]));
            cic_add_pair<4>(p1 + c.j1 * pr, c.k0, c.k1, (float) (c.T[2] * c.D[0] * c.D[1]), (float) (c.D[2] * c.D[0] * c.D[1]));
        }
    }
    if (stats) {
        const int nout = warp_sum_i(in ? 0 : 1);
        if ((threadIdx.x & 31) == 0 && nout) atomicAdd(stats, (unsigned long long) nout);
        if (threadIdx.x == 0 && ex <= 0) atomicAdd(stats + 1, 1ull);
    }
    if (ex <= 0) return;
    __syncthreads();
    const float4 *t4 = reinterpret_cast<const float4 *>(tile);
    for (TileRows w(&frame); w.more(); w.next(&frame)) {
        if (!w.active()) continue;
        const float4 v = t4[(w.r * ez >> 2) + w.q];
        if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
        const int pz = tile_wrap(frame.org[2] + 4 * w.q, g.n);
        atomicAdd(reinterpret_cast<float4 *>(canvas + tile_row_offset(g, &frame, w.tx, w.ty) + pz), v);
    }
}

__global__ void __launch_bounds__(FPM_TILE_THREADS) cic_readout_tile_kernel(const FpmGeom g, const float *__restrict__ canvas,
        const double *__restrict__ x, float *__restrict__ out, int out_stride, double prescale, int lag_nc, unsigned long long *__restrict__ stats)
{
    FPM_DYN_SMEM(tile_raw, 16);
    float *tile = reinterpret_cast<float *>(tile_raw);
    __shared__ TileFrame frame;
    const long long i = tile_particle_index(lag_nc);
    double pos[3] = { x[3 * i], x[3 * i + 1], x[3 * i + 2] };
    CicIndex c;
    cic_setup(g, pos, c);
    int u[3] = { g.nranks == 1 ? c.u[0] : c.lx0, c.u[1], c.u[2] };
    const bool valid = g.nranks == 1 || (c.lx0 >= 0 && c.lx0 < g.nxl);
    int d[3];
    const bool in = tile_setup(g, u, valid, &frame, d);
    const int ex = frame.ext[0], ey = frame.ext[1], ez = frame.ext[2];
    #define CELLV(val) (prescale == 1.0 ? (val) : (float) ((double) (val) * prescale))
    if (ex > 0) {
        float4 *t4 = reinterpret_cast<float4 *>(tile);
        for (TileRows w(&frame); w.more(); w.next(&frame)) {
            if (!w.active()) continue;
            const int pz = tile_wrap(frame.org[2] + 4 * w.q, g.n);
            float4 v = __ldg(reinterpret_cast<const float4 *>(canvas + tile_row_offset(g, &frame, w.tx, w.ty) + pz));
            v.x = CELLV(v.x); v.y = CELLV(v.y); v.z = CELLV(v.z); v.w = CELLV(v.w);
            t4[(w.r * ez >> 2) + w.q] = v;
        }
        __syncthreads();
    }
    if (stats) {
        const int nout = warp_sum_i(in ? 0 : 1);
        if ((threadIdx.x & 31) == 0 && nout) atomicAdd(stats, (unsigned long long) nout);
        if (threadIdx.x == 0 && ex <= 0) atomicAdd(stats + 1, 1ull);
    }
    double value = 0;
    if (in) {
        const float *p = tile + ((d[0] - frame.lo[0]) * ey + (d[1] - frame.lo[1])) * ez + (frame.ref[2] + d[2] - frame.org[2]);
        const int sy = ez, sx = ey * ez;
        value += (double) p[0] * (c.T[2] * c.T[0] * c.T[1]);
        value += (double) p[1] * (c.D[2] * c.T[0] * c.T[1]);
        value += (double) p[sy] * (c.T[2] * c.T[0] * c.D[1]);
        value += (double) p[sy + 1] * (c.D[2] * c.T[0] * c.D[1]);
        value += (double) p[sx] * (c.T[2] * c.D[0] * c.T[1]);
        value += (double) p[sx + 1] * (c.D[2] * c.D[0] * c.T[1]);
        value += (double) p[sx + sy] * (c.T[2] * c.D[0] * c.D[1]);
        value += (double) p[sx + sy + 1] * (c.D[2] * c.D[0] * c.D[1]);
    } else {
        const size_t pr = (size_t) g.pitch_r, pl = (size_t) g.n * pr;
        if (c.lx0 >= 0) {
            const float *p0 = canvas + (size_t) c.lx0 * pl;
            value += (double) CELLV(__ldg(p0 + c.j0 * pr + c.k0)) * (c.T[2] * c.T[0] * c.T[1]);
            value += (double) CELLV(__ldg(p0 + c.j0 * pr + c.k1)) * (c.D[2] * c.T[0] * c.T[1]);
            value += (double) CELLV(__ldg(p0 + c.j1 * pr + c.k0)) * (c.T[2] * c.T[0] * c.D[1]);
            value += (double) CELLV(__ldg(p0 + c.j1 * pr + c.k1)) * (c.D[2] * c.T[0] * c.D[1]);
        }
        if (c.lx1 >= 0) {
            const float *p1 = canvas + (size_t) c.lx1 * pl;
            value += (double) CELLV(__ldg(p1 + c.j0 * pr + c.k0)) * (c.T[2] * c.D[0] * c.T[1]);
            value += (double) CELLV(__ldg(p1 + c.j0 * pr + c.k1)) * (c.D[2] * c.D[0] * c.T[1]);
            value += (double) CELLV(__ldg(p1 + c.j1 * pr + c.k0)) * (c.T[2] * c.D[0] * c.D[1]);
            value += (double) CELLV(__ldg(p1 + c.j1 * pr + c.k1)) * (c.D[2] * c.D[0] * c.D[1]);
        }
    }
    #undef CELLV
    out[i * out_stride] = (float) value;
}

// ------------------------------------------------------------------ the generic windows (painter.c:176-317): linear, quadratic, Lanczos
// One particle per thread, support^3 mesh points.  _fill_k: per axis the window at the `support` points starting at
// floor(x/h + shift) - left, normalised to sum 1; the deposit / gather then runs x-outermost, z-innermost with the weight
// ((1 * kx) * ky) * kz, all in double like the reference.
//   * diffdir >= 0 (fastpm_painter_init_diff, painter.c:178-205): along that axis the window is replaced by its derivative times
//     1/cellsize, the normalisation staying that of the window itself.  The CIC window goes through these kernels as well then
//     (cic_paint_tuned / cic_readout_tuned with D = 1/cellsize, T = -1/cellsize, painter-cic.c:57-60,137-140), with the tuned
//     routines' own order of products.
//   * several GPUs: planes outside the slab [x0, x0 + nxl) live in a separate halo block, hl planes for x0 - hl .. x0 - 1 followed
//     by hr planes for x0 + nxl .. x0 + nxl + hr - 1 (hl = left, hr = support - 1 - left, one more for the odd supports whose
//     base cell is floor(x + 0.5)); comm.cu adds / fetches them (the reference: ghost particles, pmghosts.c:45-78).
struct WindowSpec { int type, support, left, diffdir, hl, hr; double shift, invh; };

__device__ __forceinline__ void window_fill(const FpmGeom &g, const WindowSpec &w, const double pos[3], int ipos[3], double k[3][FPM_WINDOW_MAX_SUPPORT])
{
    #pragma unroll
    for (int d = 0; d < 3; d++) {
        const double gpos = pos[d] * g.inv_cellsize;
        if (w.type == FPM_WINDOW_CIC) {
            ipos[d] = (int) floor(gpos);
            const double D = gpos - ipos[d];
            k[d][0] = d == w.diffdir ? -g.inv_cellsize : 1. - D;
            k[d][1] = d == w.diffdir ? g.inv_cellsize : D;
            continue;
        }
        ipos[d] = (int) floor(gpos + w.shift) - w.left;
        const double dx = gpos - ipos[d];
        double sum = 0;
        for (int i = 0; i < w.support; i++) {
            k[d][i] = fpm_window_eval(w.type, dx - i, w.invh);
            sum += k[d][i];
            if (d == w.diffdir) k[d][i] = fpm_window_diff_eval(w.type, dx - i, w.invh) * g.inv_cellsize;
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

// plane ix (global, before the periodic wrap) of the canvas as this rank holds it; NULL when the slab and its halo do not reach it
template <typename F>
__device__ __forceinline__ F *window_plane(const FpmGeom &g, const WindowSpec &w, F *canvas, F *halo, int ix, size_t pl)
{
    if (g.nranks == 1) return canvas + (size_t) window_wrap(ix, g.n) * pl;
    int lx = ix - g.x0;
    if (lx >= g.nxl + w.hr) lx -= g.n;                // the periodic image nearest to the slab
    else if (lx < -w.hl) lx += g.n;
    if (lx >= 0 && lx < g.nxl) return canvas + (size_t) lx * pl;
    if (halo == nullptr) return nullptr;
    if (lx < 0 && lx >= -w.hl) return halo + (size_t) (lx + w.hl) * pl;
    if (lx >= g.nxl && lx < g.nxl + w.hr) return halo + (size_t) (w.hl + lx - g.nxl) * pl;
    return nullptr;
}

__global__ void __launch_bounds__(128) window_paint_kernel(const FpmGeom g, const WindowSpec w, float *__restrict__ canvas, float *__restrict__ halo,
        const double *__restrict__ x, const float *__restrict__ mass, double M0, const float *__restrict__ field, int field_stride, long long np,
        int *__restrict__ outside)
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
    const bool cic = w.type == FPM_WINDOW_CIC;
    for (int a = 0; a < w.support; a++) {
        float *plane = window_plane(g, w, canvas, halo, ipos[0] + a, pl);
        if (plane == nullptr) { if (outside) atomicAdd(outside, 1); continue; }
        for (int b = 0; b < w.support; b++) {
            const int iy = window_wrap(ipos[1] + b, g.n);
            float *row = plane + (size_t) iy * pr;
            const double wy = k[1][b] * weight;                 // painter-cic.c:79-80: the weight is folded into the y factors
            for (int c = 0; c < w.support; c++) {
                const int iz = window_wrap(ipos[2] + c, g.n);
                if (cic) {
                    atomicAdd(row + iz, (float) (k[2][c] * k[0][a] * wy));
                } else {
                    double kernel = 1.0;
                    kernel *= k[0][a]; kernel *= k[1][b]; kernel *= k[2][c];
                    atomicAdd(row + iz, (float) (weight * kernel));
                }
            }
        }
    }
}

__global__ void __launch_bounds__(128) window_readout_kernel(const FpmGeom g, const WindowSpec w, const float *__restrict__ canvas,
        const float *__restrict__ halo, const double *__restrict__ x, float *__restrict__ out, int out_stride, long long np, int *__restrict__ outside)
{
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    const double pos[3] = { x[3 * i], x[3 * i + 1], x[3 * i + 2] };
    int ipos[3];
    double k[3][FPM_WINDOW_MAX_SUPPORT];
    window_fill(g, w, pos, ipos, k);
    const size_t pr = (size_t) g.pitch_r, pl = (size_t) g.n * pr;
    const bool cic = w.type == FPM_WINDOW_CIC;
    double value = 0;
    for (int a = 0; a < w.support; a++) {
        const float *plane = window_plane(g, w, canvas, halo, ipos[0] + a, pl);
        if (plane == nullptr) { if (outside) atomicAdd(outside, 1); continue; }
        for (int b = 0; b < w.support; b++) {
            const int iy = window_wrap(ipos[1] + b, g.n);
            const float *row = plane + (size_t) iy * pr;
            for (int c = 0; c < w.support; c++) {
                const int iz = window_wrap(ipos[2] + c, g.n);
                if (cic) {
                    value += (double) __ldg(row + iz) * (k[2][c] * k[0][a] * k[1][b]);
                } else {
                    double kernel = 1.0;
                    kernel *= k[0][a]; kernel *= k[1][b]; kernel *= k[2][c];
                    value += kernel * (double) __ldg(row + iz);
                }
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
int fpm_get_lagrangian_hint(void) { return g_lag_nc; }
// Bricks pay off only when the linear walk's footprint (~24 mesh planes) no longer fits the 126 MB L2: measured on B200, they
// cost 15 % at N = 1024 (4.3 MB planes, linear walk already L2 resident) and gain 25 % at N = 2048 (17 MB planes).
// Returns the number of leading 256-particle blocks that use the brick mapping (0: linear walk).
static int fpm_superblock(int lag_nc, long long ngroups)
{
    static int sbi = -1, sbj = -1;       // FASTPM_B200_SUPERBLOCK=IxJ bricks (default 8x4; 1x1 = the plain k, j, i order of round 1)
    if (sbi < 0) {
        sbi = 8; sbj = 4;
        const char *e = getenv("FASTPM_B200_SUPERBLOCK");
        if (e) { int a = 0, b = 0; if (sscanf(e, "%dx%d", &a, &b) == 2 && a > 0 && b > 0) { sbi = a; sbj = b; } }
    }
    const int nb = lag_nc >> 3;
    int bi = sbi, bj = sbj;
    while (bj > 1 && nb % bj) bj--;
    while (bi > 1 && ngroups % bi) bi--;
    return (bi == 1 && bj == 1) ? 0 : ((bi << 16) | bj);
}
static int fpm_lagrangian_hint(long long np, const FpmGeom &g, int *lag_nc, int *sb = nullptr)
{
    if (sb) *sb = 0;
    const size_t plane_bytes = (size_t) g.n * g.pitch_r * sizeof(float);
    *lag_nc = 0;
    if (!g_lag_nc || (plane_bytes <= ((size_t) 6 << 20) && !g_lag_force)) return 0;
    const long long group = 4LL * g_lag_nc * g_lag_nc;          // particles in 4 i-planes
    const long long ngroups = np / group;
    if (ngroups == 0) return 0;
    *lag_nc = g_lag_nc;
    if (sb) *sb = fpm_superblock(g_lag_nc, ngroups);
    return (int) (ngroups * group / 256);
}

// Shared-memory tiles (cic_paint_tile_kernel / cic_readout_tile_kernel): for the leading complete groups of 8 i-planes of a store
// with the Lagrangian hint, when a brick of 8 particles per side spans at most ~21 mesh cells (Nmesh / nc <= 2.5) and rows can be
// moved in aligned 16-byte groups.  OFF by default (FASTPM_B200_TILES=1 switches them on): measured on B200 at nc = 1024 / N = 2048
// (round 2, gpurun_out/r02b_bench_*.json) the deposit takes 88 ms through the tiles against 42 ms with global reductions and the
// gather 85 ms against 21 ms -- see DESIGN.md section 3 for why.
static unsigned long long *g_tile_stats = nullptr;      // [4] device counters when FASTPM_B200_TILE_STATS is set: particles that took
                                                        // the global path and CTAs without a tile, for the deposit and for the gather
static long long fpm_tile_particles(long long np, const FpmGeom &g)
{
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("FASTPM_B200_TILES");
        on = (e && atoi(e) != 0) ? 1 : 0;
        if (on && getenv("FASTPM_B200_TILE_STATS") && cudaMalloc(&g_tile_stats, 4 * sizeof(unsigned long long)) == cudaSuccess)
            cudaMemset(g_tile_stats, 0, 4 * sizeof(unsigned long long));
    }
    if (!on || !g_lag_nc || g.n % 4 != 0 || 2 * g.n > 5 * g_lag_nc) return 0;
    const long long group = 8LL * g_lag_nc * g_lag_nc;
    return (np / group) * group;
}
int fpm_tile_stats_fetch(unsigned long long out[4])
{
    for (int i = 0; i < 4; i++) out[i] = 0;
    if (!g_tile_stats) return 0;
    FPM_CUDA_OK(cudaMemcpy(out, g_tile_stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return 0;
}
static int tile_attrs()
{
    static bool done = false;
    if (done) return 0;
    FPM_CUDA_OK(cudaFuncSetAttribute(cic_paint_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (FPM_TILE_CAP * sizeof(float))));
    FPM_CUDA_OK(cudaFuncSetAttribute(cic_paint_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (FPM_TILE_CAP * sizeof(float))));
    FPM_CUDA_OK(cudaFuncSetAttribute(cic_readout_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (FPM_TILE_CAP * sizeof(float))));
    done = true;
    return 0;
}

// wrap_bad != NULL: wrap the positions on the way (x is then written where it changed)
int fpm_paint_launch(const FpmMesh *m, float *canvas, const double *x, const float *mass, double M0,
                     const float *field, int field_stride, long long np, int *wrap_bad, cudaStream_t st)
{
    if (np <= 0) return 0;
    static int vec = -1;          // FASTPM_B200_PAINT_VEC = 0 | 2 | 4 (default): width of the vector reductions
    if (vec < 0) { const char *e = getenv("FASTPM_B200_PAINT_VEC"); vec = e ? atoi(e) : 4; }
    double *xw = const_cast<double *>(x);
    // leading complete groups of 8 i-planes: shared-memory tiles; the rest (and everything without the hint): global reductions
    const long long ntile = fpm_tile_particles(np, m->geom);
    if (fpm_prof_on) fpm_prof_begin(FPM_K_PAINT, st);
    if (ntile > 0) {
        if (tile_attrs()) return -1;
        const unsigned tgrid = (unsigned) (ntile / FPM_TILE_THREADS);
        fpm_path_counter[FPM_PATH_PAINT_TILES]++;
        if (wrap_bad) cic_paint_tile_kernel<true><<<tgrid, FPM_TILE_THREADS, FPM_TILE_CAP * sizeof(float), st>>>(m->geom, canvas, xw, mass, M0, field, field_stride, wrap_bad, g_lag_nc, g_tile_stats);
        else cic_paint_tile_kernel<false><<<tgrid, FPM_TILE_THREADS, FPM_TILE_CAP * sizeof(float), st>>>(m->geom, canvas, xw, mass, M0, field, field_stride, wrap_bad, g_lag_nc, g_tile_stats);
        fpm_launch_counter++;
    }
    const long long nrest = np - ntile;
    if (nrest > 0) {
        const unsigned grid = (unsigned) ((nrest + 255) / 256);
        int lag_nc = 0, sb = 0;
        const int nbrick = fpm_lagrangian_hint(nrest, m->geom, &lag_nc, &sb);
        #define PAINT_LAUNCH(V, W) cic_paint_kernel<V, W><<<grid, 256, 0, st>>>(m->geom, canvas, xw, mass, M0, field, field_stride, np, wrap_bad, lag_nc, nbrick, ntile, sb)
        if (nbrick > 0) fpm_path_counter[FPM_PATH_PAINT_BRICKS]++;
        if (wrap_bad) { if (vec >= 4) PAINT_LAUNCH(4, true); else if (vec >= 2) PAINT_LAUNCH(2, true); else PAINT_LAUNCH(0, true); }
        else { if (vec >= 4) PAINT_LAUNCH(4, false); else if (vec >= 2) PAINT_LAUNCH(2, false); else PAINT_LAUNCH(0, false); }
        #undef PAINT_LAUNCH
    }
    if (fpm_prof_on) fpm_prof_end(FPM_K_PAINT, st);
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_readout_launch(const FpmMesh *m, const float *canvas, const double *x, float *out, int out_stride,
                       double prescale, long long np, cudaStream_t st, const float *pack0 = nullptr, const float *pack1 = nullptr)
{
    if (np <= 0) return 0;
    const long long ntile = pack0 ? 0 : fpm_tile_particles(np, m->geom);
    if (fpm_prof_on) fpm_prof_begin(FPM_K_READOUT, st);
    if (ntile > 0) {
        if (tile_attrs()) return -1;
        fpm_path_counter[FPM_PATH_READOUT_TILES]++;
        cic_readout_tile_kernel<<<(unsigned) (ntile / FPM_TILE_THREADS), FPM_TILE_THREADS, FPM_TILE_CAP * sizeof(float), st>>>(m->geom, canvas, x, out, out_stride, prescale, g_lag_nc, g_tile_stats ? g_tile_stats + 2 : nullptr);
        fpm_launch_counter++;
    }
    const long long nrest = np - ntile;
    if (nrest > 0) {
        const unsigned grid = (unsigned) ((nrest + 255) / 256);
        // the gather keeps the plain brick order: measured at nc = 1024 (gpurun_out/r02g_sb_*.json) super-blocks take 10 % off the
        // deposit (47.6 -> 42.8 ms) and nothing off the gather (26.6 -> 27.1 ms)
        int lag_nc = 0, sb = 0;
        const int nbrick = fpm_lagrangian_hint(nrest, m->geom, &lag_nc, nullptr);
        if (nbrick > 0) fpm_path_counter[FPM_PATH_READOUT_BRICKS]++;
        cic_readout_kernel<<<grid, 256, 0, st>>>(m->geom, canvas, x, out, out_stride, prescale, np, lag_nc, nbrick, ntile, sb, pack0, pack1);
    }
    if (fpm_prof_on) fpm_prof_end(FPM_K_READOUT, st);
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
    fpm_path_counter[FPM_PATH_READOUT3]++;
    FPM_TIMED(FPM_K_READOUT, st, (cic_readout3_kernel<<<grid, 256, 0, st>>>(m->geom, c0, c1, c2, x, out, np, lag_nc, nbrick)));
    FPM_CHECK_LAUNCH();
    return 0;
}

// halo planes a window needs below / above a slab: the lowest cell of a particle at x0 <= x/h < x0 + nxl is floor(x/h + shift) - left
// >= x0 - left, the highest floor(x/h + shift) - left + support - 1 <= x0 + nxl - 1 + (support - 1 - left) + (shift > 0)
int fpm_window_halo(int type, int support, int *left, int *right)
{
    if (type == FPM_WINDOW_CIC || type == FPM_WINDOW_LINEAR) support = 2;
    else if (type == FPM_WINDOW_QUAD) support = 3;
    const int l = type == FPM_WINDOW_CIC ? 0 : (support - 1) / 2;
    *left = l;
    *right = support - 1 - l + ((type != FPM_WINDOW_CIC && support % 2) ? 1 : 0);
    return 0;
}

static int window_spec(const FpmMesh *m, int type, int support, int diffdir, bool have_halo, WindowSpec *w)
{
    // fastpm_painter_init, painter.c:128-174
    if (type == FPM_WINDOW_LINEAR || type == FPM_WINDOW_CIC) support = 2;
    else if (type == FPM_WINDOW_QUAD) support = 3;
    else if (type != FPM_WINDOW_LANCZOS) { fpm_set_error("window type %d", type); return -1; }
    if (support < 1 || support > FPM_WINDOW_MAX_SUPPORT) { fpm_set_error("window support %d (1..%d on the device)", support, FPM_WINDOW_MAX_SUPPORT); return -1; }
    if (diffdir < -1 || diffdir > 2) { fpm_set_error("window derivative direction %d", diffdir); return -1; }
    w->type = type; w->support = support; w->left = type == FPM_WINDOW_CIC ? 0 : (support - 1) / 2; w->diffdir = diffdir;
    w->shift = (type != FPM_WINDOW_CIC && support % 2) ? 0.5 : 0; w->invh = 1 / (0.5 * support);
    fpm_window_halo(type, support, &w->hl, &w->hr);
    if (m->geom.nranks > 1) {
        if (!have_halo) { fpm_set_error("the linear / quadratic / Lanczos / derivative windows need their halo block on several GPUs"); return -1; }
        if (w->hl > m->geom.nxl || w->hr > m->geom.nxl) { fpm_set_error("window halo of %d + %d planes on slabs of %d planes", w->hl, w->hr, m->geom.nxl); return -1; }
    }
    return 0;
}

// particles whose window left the slab and its halo (a store that was not decomposed): counted, reported by the next call
static int *g_window_outside = nullptr;
static int window_outside_check(cudaStream_t st)
{
    if (!g_window_outside) {
        FPM_CUDA_OK(cudaMalloc(&g_window_outside, sizeof(int)));
        FPM_CUDA_OK(cudaMemsetAsync(g_window_outside, 0, sizeof(int), st));
        return 0;
    }
    int n = 0;
    FPM_CUDA_OK(cudaMemcpyAsync(&n, g_window_outside, sizeof(int), cudaMemcpyDeviceToHost, st));
    FPM_CUDA_OK(cudaStreamSynchronize(st));
    if (n) {
        FPM_CUDA_OK(cudaMemsetAsync(g_window_outside, 0, sizeof(int), st));
        fpm_set_error("window paint / readout: %d mesh planes of particles outside this rank's slab and halo (store not decomposed?)", n);
        return -1;
    }
    return 0;
}

int fpm_window_paint_launch(const FpmMesh *m, int type, int support, int diffdir, float *canvas, float *halo, const double *x, const float *mass, double M0,
                            const float *field, int field_stride, long long np, cudaStream_t st)
{
    WindowSpec w;
    if (window_spec(m, type, support, diffdir, halo != nullptr, &w)) return -1;
    if (np <= 0) return 0;
    int *outside = nullptr;
    if (m->geom.nranks > 1) { if (window_outside_check(st)) return -1; outside = g_window_outside; }
    FPM_TIMED(FPM_K_PAINT, st, (window_paint_kernel<<<(unsigned) ((np + 127) / 128), 128, 0, st>>>(m->geom, w, canvas, halo, x, mass, M0, field, field_stride, np, outside)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_window_readout_launch(const FpmMesh *m, int type, int support, int diffdir, const float *canvas, const float *halo, const double *x, float *out,
                              int out_stride, long long np, cudaStream_t st)
{
    WindowSpec w;
    if (window_spec(m, type, support, diffdir, halo != nullptr, &w)) return -1;
    if (np <= 0) return 0;
    int *outside = nullptr;
    if (m->geom.nranks > 1) { if (window_outside_check(st)) return -1; outside = g_window_outside; }
    FPM_TIMED(FPM_K_READOUT, st, (window_readout_kernel<<<(unsigned) ((np + 127) / 128), 128, 0, st>>>(m->geom, w, canvas, halo, x, out, out_stride, np, outside)));
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
