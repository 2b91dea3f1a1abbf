// fastpm_b200 -- slab-decomposed 3-D real<->complex FFT for sm_100a, replacing pm_r2c / pm_c2r
// (reference: libfastpm/pmpfft.c:370-399, which delegates to PFFT/FFTW).
//
// Three passes per transform, each one read + one write of the local mesh (6*S bytes per transform):
//   forward  F1 z-pass  real rows -> half-complex rows, in place, with the input scale folded in
//            F2 y-pass  tile [y:N][kz:K] of plane x  -> written TRANSPOSED to cplx[ky][x][kz]
//                       (on several GPUs the store goes straight to the peer that owns ky)
//            F3 x-pass  tile [kx:N][kz:K] of plane ky, in place
//   backward B1 x-pass  tile of plane ky (optionally multiplied by the gravity kernel, transfer.cuh),
//                       inverse over kx, written transposed back to real-layout [x][ky][kz]
//            B2 y-pass  in place,  B3 z-pass half-complex rows -> real rows, in place.
// All strided tiles are [N rows][K contiguous complex]: K*8 B = 64..128 B contiguous per row, rows
// 128 B aligned because pitch_c is a multiple of 16 complex.
#include "common.cuh"
#include "fft_core.h"
#include "mesh.cuh"
#include <vector>
#include <stdlib.h>
#include <string.h>

// ------------------------------------------------------------------ tile pass
struct TilePassArgs {
    const float2 *src;
    size_t src_estride;     // elements between successive FFT inputs of one column
    size_t src_ostride;     // elements between outer indices
    float2 *dst[FPM_MAX_RANKS];
    int rows_per_rank;      // output rows [d*rows_per_rank, (d+1)*rows_per_rank) go to dst[d]
    size_t dst_estride;
    size_t dst_ostride;
    int dst_ooffset;        // added to the outer index on the destination side
    int self_rank, self_ooffset;        // staged transpose: rows of rank self_rank (>= 0) go straight to self_dst
    float2 *self_dst;
    size_t self_estride, self_ostride;
    int ntile_k;            // tiles along kz
    int conj;               // 1: inverse transform by conjugation
    int outer0;             // global index of outer 0 (ky0 for B1: needed by the transfer)
    FpmFftDev t;
    FpmTransferSpec xfer;   // applied on load when xfer.active (B1 only: rows are kx, outer is ky)
    FpmKTables kt;
};

template <int K>
__global__ void __launch_bounds__(512) fft_tile_kernel(const TilePassArgs a)
{
    FPM_DYN_SMEM(smem_raw, 16);
    float2 *smem = reinterpret_cast<float2 *>(smem_raw);
    const int n = a.t.n;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int o = blockIdx.x / a.ntile_k;
    const int kz0 = (blockIdx.x - o * a.ntile_k) * K;

    // ---- load (coalesced: K consecutive complex per row)
    {
        const float2 *src = a.src + (size_t) o * a.src_ostride + kz0;
        const int c = tid % K;
        const int rstep = nthr / K;
        #pragma unroll 4
        for (int row = tid / K; row < n; row += rstep) {
            float2 v = __ldg(src + (size_t) row * a.src_estride + c);
            if (a.xfer.active) v = fpm_apply_transfer(a.xfer, a.kt, v, row, a.outer0 + o, kz0 + c);
            if (a.conj) v.y = -v.y;
            smem[row * K + c] = v;
        }
    }
    __syncthreads();
    // ---- transform
    {
        int ncur = n;
        for (int j = 0; j < a.t.nstage; j++) {
            const int r = a.t.radix[j];
            fpm_fft_stage(tid, nthr, smem, K, K, a.t, ncur, r);
            ncur /= r;
            __syncthreads();
        }
    }
    // ---- store (digit-reversed row order; each row segment is still K contiguous complex)
    {
        const int c = tid % K;
        const int rstep = nthr / K;
        const size_t obase = (size_t) (a.dst_ooffset + o) * a.dst_ostride + kz0 + c;
        #pragma unroll 4
        for (int pos = tid / K; pos < n; pos += rstep) {
            const int k = __ldg(a.t.rev + pos);
            const int d = k / a.rows_per_rank;
            const int kl = k - d * a.rows_per_rank;
            float2 v = smem[pos * K + c];
            if (a.conj) v.y = -v.y;
            if (d == a.self_rank) a.self_dst[(size_t) kl * a.self_estride + (size_t) (a.self_ooffset + o) * a.self_ostride + kz0 + c] = v;
            else a.dst[d][(size_t) kl * a.dst_estride + obase] = v;
        }
    }
}

// ------------------------------------------------------------------ z passes
struct ZPassArgs {
    const float *src;      // rows of pitch_r floats == pitch_c complex
    float *dst;            // may be the same buffer (in place) or another one with the same row layout
    size_t nrows;
    int pitch_c;
    float scale;           // forward only: input is multiplied by this
    FpmFftDev th;          // half-length plan
    const float2 *twN;     // length-N table
};

// forward: R rows per CTA; tile [h][R+1]
template <int R>
__global__ void __launch_bounds__(512) fft_zfwd_kernel(const ZPassArgs a)
{
    FPM_DYN_SMEM(smem_raw, 16);
    float2 *smem = reinterpret_cast<float2 *>(smem_raw);
    constexpr int KP = R + 1;
    const int h = a.th.n;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const size_t row0 = (size_t) blockIdx.x * R;
    const int nr = (int) ((a.nrows - row0) < (size_t) R ? (a.nrows - row0) : (size_t) R);

    for (int w = tid; w < R * h; w += nthr) {
        const int r = w / h, j = w - r * h;
        float2 v = make_float2(0.f, 0.f);
        if (r < nr) {
            v = reinterpret_cast<const float2 *>(a.src + (row0 + r) * (size_t) (2 * a.pitch_c))[j];
            v.x *= a.scale; v.y *= a.scale;
        }
        smem[j * KP + r] = v;
    }
    __syncthreads();
    int ncur = h;
    for (int j = 0; j < a.th.nstage; j++) {
        const int r = a.th.radix[j];
        fpm_fft_stage(tid, nthr, smem, KP, R, a.th, ncur, r);
        ncur /= r;
        __syncthreads();
    }
    for (int w = tid; w < R * (h + 1); w += nthr) {
        const int r = w / (h + 1), k = w - r * (h + 1);
        if (r < nr) {
            float2 x = fpm_untangle_fwd(smem, KP, r, a.th, a.twN, k);
            reinterpret_cast<float2 *>(a.dst + (row0 + r) * (size_t) (2 * a.pitch_c))[k] = x;
        }
    }
}

// backward: half-complex rows -> real rows (unnormalised), tile [h+1][R+1]
template <int R>
__global__ void __launch_bounds__(512) fft_zbwd_kernel(const ZPassArgs a)
{
    FPM_DYN_SMEM(smem_raw, 16);
    float2 *smem = reinterpret_cast<float2 *>(smem_raw);
    constexpr int KP = R + 1;
    const int h = a.th.n;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const size_t row0 = (size_t) blockIdx.x * R;
    const int nr = (int) ((a.nrows - row0) < (size_t) R ? (a.nrows - row0) : (size_t) R);

    for (int w = tid; w < R * (h + 1); w += nthr) {
        const int r = w / (h + 1), k = w - r * (h + 1);
        float2 v = make_float2(0.f, 0.f);
        if (r < nr) v = reinterpret_cast<const float2 *>(a.src + (row0 + r) * (size_t) (2 * a.pitch_c))[k];
        smem[k * KP + r] = v;
    }
    __syncthreads();
    // pairs (k, h-k), k = 0..h/2, column r
    const int npair = h / 2 + 1;
    for (int w = tid; w < R * npair; w += nthr) {
        const int k = w / R, r = w - k * R;
        float2 xk = smem[k * KP + r], xhk = smem[(h - k) * KP + r];
        float2 zk, zhk;
        fpm_tangle_bwd_pair(xk, xhk, a.twN[k], &zk, &zhk);
        smem[k * KP + r] = zk;
        if (k != 0 && k != h - k) smem[(h - k) * KP + r] = zhk;
    }
    __syncthreads();
    int ncur = h;
    for (int j = 0; j < a.th.nstage; j++) {
        const int r = a.th.radix[j];
        fpm_fft_stage(tid, nthr, smem, KP, R, a.th, ncur, r);
        ncur /= r;
        __syncthreads();
    }
    for (int w = tid; w < R * h; w += nthr) {
        const int r = w / h, j = w - r * h;
        if (r < nr) {
            float2 z = smem[a.th.inv[j] * KP + r];
            reinterpret_cast<float2 *>(a.dst + (row0 + r) * (size_t) (2 * a.pitch_c))[j] = make_float2(z.x, -z.y);
        }
    }
}

#ifndef FPM_EMULATE          // host side: not part of the CPU emulation of the kernels (tests/emul/fft_generic_emul.cpp)
// ------------------------------------------------------------------ plan
struct FpmFftPlan {
    int n;
    FpmFftDev tN, tH;       // device-pointer versions
    float2 *d_twN = nullptr, *d_twH = nullptr;
    int *d_revN = nullptr, *d_invN = nullptr, *d_revH = nullptr, *d_invH = nullptr;
    int K, R, thr_tile, thr_z;
    int pitch_c;
    size_t smem_tile, smem_z;
};

static int upload_plan(const FpmFftHostPlan &hp, FpmFftDev *dev, float2 **d_tw, int **d_rev, int **d_inv)
{
    FPM_CUDA_OK(cudaMalloc(d_tw, sizeof(float2) * hp.n));
    FPM_CUDA_OK(cudaMalloc(d_rev, sizeof(int) * hp.n));
    FPM_CUDA_OK(cudaMalloc(d_inv, sizeof(int) * hp.n));
    FPM_CUDA_OK(cudaMemcpy(*d_tw, hp.tw.data(), sizeof(float2) * hp.n, cudaMemcpyHostToDevice));
    FPM_CUDA_OK(cudaMemcpy(*d_rev, hp.rev.data(), sizeof(int) * hp.n, cudaMemcpyHostToDevice));
    FPM_CUDA_OK(cudaMemcpy(*d_inv, hp.inv.data(), sizeof(int) * hp.n, cudaMemcpyHostToDevice));
    dev->n = hp.n; dev->nstage = (int) hp.radix.size();
    for (int j = 0; j < dev->nstage; j++) dev->radix[j] = hp.radix[j];
    dev->tw = *d_tw; dev->rev = *d_rev; dev->inv = *d_inv;
    return 0;
}

template <int K> static int set_tile_attr(size_t smem)
{
    FPM_CUDA_OK(cudaFuncSetAttribute(fft_tile_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    return 0;
}
template <int R> static int set_z_attr(size_t smem)
{
    FPM_CUDA_OK(cudaFuncSetAttribute(fft_zfwd_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    FPM_CUDA_OK(cudaFuncSetAttribute(fft_zbwd_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    return 0;
}

int fpm_fft_plan_create(int n, FpmFftPlan **out)
{
    if (n < 4 || (n & 1)) { fpm_set_error("Nmesh must be even and >= 4, got %d", n); return -1; }
    FpmFftHostPlan hN(n), hH(n / 2);
    if (!hN.ok || !hH.ok) { fpm_set_error("Nmesh = %d has a prime factor other than 2, 3, 5", n); return -1; }
    FpmFftPlan *p = new FpmFftPlan();
    p->n = n;
    p->pitch_c = ((n / 2 + 1 + 15) / 16) * 16;
    if (upload_plan(hN, &p->tN, &p->d_twN, &p->d_revN, &p->d_invN)) return -1;
    if (upload_plan(hH, &p->tH, &p->d_twH, &p->d_revH, &p->d_invH)) return -1;
    p->K = n <= 512 ? 16 : (n <= 2048 ? 8 : 4);
    p->R = n <= 1024 ? 16 : 8;
    if ((size_t) n * p->K * sizeof(float2) > 200 * 1024) { fpm_set_error("Nmesh = %d too large for the shared-memory tile", n); return -1; }
    p->smem_tile = (size_t) n * p->K * sizeof(float2);
    p->smem_z = (size_t) (n / 2 + 1) * (p->R + 1) * sizeof(float2);
    size_t work = (size_t) n * p->K / 4;             // radix-4 butterflies per stage
    p->thr_tile = work >= 2048 ? 512 : (work >= 512 ? 256 : 128);
    size_t zwork = (size_t) (n / 2) * p->R / 4;
    p->thr_z = zwork >= 2048 ? 512 : (zwork >= 512 ? 256 : 128);
    // the opt-in dynamic shared memory limit is a property of the kernel, not of a plan: raise it once, to the
    // device maximum, for every instantiation (several meshes of different size coexist: vpm.c, lptpm, basepm)
    static bool attrs_done = false;
    if (!attrs_done) {
        const size_t lim = 227 * 1024;
        if (set_tile_attr<16>(lim) || set_tile_attr<8>(lim) || set_tile_attr<4>(lim) || set_z_attr<16>(lim) || set_z_attr<8>(lim)) return -1;
        attrs_done = true;
    }
    *out = p;
    return 0;
}

void fpm_fft_plan_destroy(FpmFftPlan *p)
{
    if (!p) return;
    cudaFree(p->d_twN); cudaFree(p->d_twH); cudaFree(p->d_revN); cudaFree(p->d_invN); cudaFree(p->d_revH); cudaFree(p->d_invH);
    delete p;
}

int fpm_fft_tma_supported(int n);
int fpm_fft_tma_tile_k(int n);
int fpm_fft_tma_pass_from_tile(int n, const TilePassArgs &a, int pitch_c, int nouter, cudaStream_t st);

static int g_fft_generic = -1;
void fpm_fft_force_generic(int on) { g_fft_generic = on ? 1 : 0; }
static int use_tma(int n)
{
    if (g_fft_generic < 0) { const char *e = getenv("FASTPM_B200_FFT"); g_fft_generic = (e && !strcmp(e, "generic")) ? 1 : 0; }
    return !g_fft_generic && fpm_fft_tma_supported(n);
}

static int launch_tile(const FpmFftPlan *p, const TilePassArgs &a, int nouter, cudaStream_t st)
{
    if (use_tma(p->n) && a.src_estride == (size_t) p->pitch_c && a.src_ostride == (size_t) p->n * p->pitch_c)
        return fpm_fft_tma_pass_from_tile(p->n, a, p->pitch_c, nouter, st);
    const unsigned grid = (unsigned) ((size_t) nouter * a.ntile_k);
    if (grid == 0) return 0;
    fpm_path_counter[FPM_PATH_FFT_TILE_GENERIC]++;
    if (fpm_prof_on) fpm_prof_begin(FPM_K_FFT_TILE, st);
    if (p->K == 16) fft_tile_kernel<16><<<grid, p->thr_tile, p->smem_tile, st>>>(a);
    else if (p->K == 8) fft_tile_kernel<8><<<grid, p->thr_tile, p->smem_tile, st>>>(a);
    else fft_tile_kernel<4><<<grid, p->thr_tile, p->smem_tile, st>>>(a);
    if (fpm_prof_on) fpm_prof_end(FPM_K_FFT_TILE, st);
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_fft_zrow_supported(int n, size_t nrows);
int fpm_fft_zrow_pass(int n, const float *src, float *dst, size_t nrows, int pitch_c, float scale,
                      const float2 *twH, const float2 *twN, int forward, cudaStream_t st);

static int launch_z(const FpmFftPlan *p, const ZPassArgs &a, int forward, cudaStream_t st)
{
    if (use_tma(p->n) && fpm_fft_zrow_supported(p->n, a.nrows))
        return fpm_fft_zrow_pass(p->n, a.src, a.dst, a.nrows, a.pitch_c, a.scale, a.th.tw, a.twN, forward, st);
    const unsigned grid = (unsigned) ((a.nrows + p->R - 1) / p->R);
    if (grid == 0) return 0;
    fpm_path_counter[FPM_PATH_FFT_Z_GENERIC]++;
    if (fpm_prof_on) fpm_prof_begin(FPM_K_FFT_Z, st);
    if (p->R == 16) {
        if (forward) fft_zfwd_kernel<16><<<grid, p->thr_z, p->smem_z, st>>>(a); else fft_zbwd_kernel<16><<<grid, p->thr_z, p->smem_z, st>>>(a);
    } else {
        if (forward) fft_zfwd_kernel<8><<<grid, p->thr_z, p->smem_z, st>>>(a); else fft_zbwd_kernel<8><<<grid, p->thr_z, p->smem_z, st>>>(a);
    }
    if (fpm_prof_on) fpm_prof_end(FPM_K_FFT_Z, st);
    FPM_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------ staged slab transpose (several GPUs)
// The transposing pass writes rows of K*8 = 64..128 contiguous bytes; sent straight over NVLink such small stores reach about
// half of the link rate.  Staged: the pass writes into a LOCAL buffer laid out [destination rank][its row][my plane][kz],
// chunk of planes by chunk, and each finished chunk is pushed to its owners by the copy engines as large 2-D copies on a second
// stream while the next chunk is being transformed -- NVLink runs at its bulk rate, overlapped with the arithmetic.
// Two sets of events / staging meshes (`set` 0 and 1): the pushes of one transform may still be in flight while the next
// transform's transposing pass fills the other staging mesh (pipelined inverse transforms of the force components, host/gravity.c).
extern unsigned long long fpm_comm_bytes[4];      // comm.cu
static cudaStream_t g_copy_stream = nullptr;
static cudaEvent_t g_ev_chunk[2][16], g_ev_done[2];
static int g_push_pending[2] = { 0, 0 };
// waits (on the compute stream) for the pushes of `set` issued by the last staged_transpose
static int staged_transpose_wait(int set, cudaStream_t st)
{
    if (!g_push_pending[set]) return 0;
    // the exposed tail of the pushes: time between the end of the work queued before this point and the arrival of the last copy
    if (fpm_prof_on) fpm_prof_begin(FPM_K_PUSH, st);
    FPM_CUDA_OK(cudaStreamWaitEvent(st, g_ev_done[set], 0));
    if (fpm_prof_on) fpm_prof_end(FPM_K_PUSH, st);
    g_push_pending[set] = 0;
    return 0;
}
static int staged_transpose(FpmMesh *m, TilePassArgs a, const float2 *src, float *const *final_peers, int nouter, int off0, int set, cudaStream_t st)
{
    const FpmGeom &g = m->geom;
    const FpmFftPlan *p = m->plan;
    const int G = g.nranks, n = g.n, per = n / G;          // rows per destination rank
    const size_t pc = (size_t) g.pitch_c, plane = (size_t) n * pc;
    if (!g_copy_stream) {
        FPM_CUDA_OK(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
        for (int s2 = 0; s2 < 2; s2++) {
            for (int i = 0; i < 16; i++) FPM_CUDA_OK(cudaEventCreateWithFlags(&g_ev_chunk[s2][i], cudaEventDisableTiming));
            FPM_CUDA_OK(cudaEventCreateWithFlags(&g_ev_done[s2], cudaEventDisableTiming));
        }
    }
    fpm_path_counter[FPM_PATH_STAGED_TRANSPOSE]++;
    float2 *stage = reinterpret_cast<float2 *>(set ? m->stage2 : m->stage);
    const size_t blk = (size_t) per * nouter * pc;          // one destination's block: [per rows][nouter planes][pitch_c]
    // my own rows keep the direct destination the caller set up (final layout); everybody else's go to their staging block
    a.self_rank = g.rank; a.self_dst = a.dst[g.rank]; a.self_estride = a.dst_estride; a.self_ostride = a.dst_ostride;
    const int self_ooffset0 = a.dst_ooffset;
    for (int d = 0; d < G; d++) a.dst[d] = stage + (size_t) d * blk;
    a.rows_per_rank = per; a.dst_estride = (size_t) nouter * pc; a.dst_ostride = pc;
    // chunks of planes: every finished chunk is pushed while the next is transformed; the last chunk's push is the exposed tail
    static int want_chunks = -1;       // FASTPM_B200_TRANSPOSE_CHUNKS (default 8; round 1 used 4: the tail was 12 ms per step on 2 GPUs)
    if (want_chunks < 0) { const char *e = getenv("FASTPM_B200_TRANSPOSE_CHUNKS"); want_chunks = e ? atoi(e) : 8; if (want_chunks < 1) want_chunks = 1; if (want_chunks > 16) want_chunks = 16; }
    int nch = nouter >= 16 * want_chunks ? want_chunks : (nouter >= 64 ? 4 : 1);
    while (nouter % nch) nch--;
    const int cp = nouter / nch;
    const int outer0 = a.outer0;
    for (int ch = 0; ch < nch; ch++) {
        a.src = src + (size_t) ch * cp * plane;
        a.dst_ooffset = ch * cp;
        a.self_ooffset = self_ooffset0 + ch * cp;
        a.outer0 = outer0 + ch * cp;
        if (launch_tile(p, a, cp, st)) return -1;
        FPM_CUDA_OK(cudaEventRecord(g_ev_chunk[set][ch], st));
        FPM_CUDA_OK(cudaStreamWaitEvent(g_copy_stream, g_ev_chunk[set][ch], 0));
        for (int dd = 0; dd < G - 1; dd++) {
            const int d = (g.rank + 1 + dd) % G;            // start with the neighbour: spreads the traffic over the links
            // my planes [off0 + ch*cp, +cp) of every row kl that rank d owns: final[d][kl][off0 + ch*cp ..][kz]
            float2 *dst = reinterpret_cast<float2 *>(final_peers[d]) + (size_t) (off0 + ch * cp) * pc;
            const float2 *s2 = stage + (size_t) d * blk + (size_t) ch * cp * pc;
            FPM_CUDA_OK(cudaMemcpy2DAsync(dst, plane * sizeof(float2), s2, (size_t) nouter * pc * sizeof(float2),
                                          (size_t) cp * pc * sizeof(float2), per, cudaMemcpyDeviceToDevice, g_copy_stream));
            fpm_comm_bytes[0] += (unsigned long long) cp * pc * sizeof(float2) * per;
        }
    }
    FPM_CUDA_OK(cudaEventRecord(g_ev_done[set], g_copy_stream));
    g_push_pending[set] = 1;
    return 0;
}

// ------------------------------------------------------------------ transforms
// Forward: real -> cplx, scaled by `scale`.  `work` receives the z- and y-pass intermediate (pass `real`
// itself to transform in place and destroy the input).  cplx_peers[d] is rank d's k-space buffer
// (cplx_peers[rank] is the local one); with nranks == 1 only entry 0 is used.
int fpm_fft_r2c(FpmMesh *m, const float *real_in, float *work, float *const *cplx_peers, float scale, cudaStream_t st)
{
    const FpmGeom &g = m->geom;
    const FpmFftPlan *p = m->plan;
    const int n = g.n;
    const size_t plane = (size_t) n * g.pitch_c;
    // F1
    ZPassArgs z; z.src = real_in; z.dst = work; z.nrows = (size_t) g.nxl * n; z.pitch_c = g.pitch_c; z.scale = scale; z.th = p->tH; z.twN = p->d_twN;
    if (launch_z(p, z, 1, st)) return -1;
    float *real = work;
    // several GPUs: nobody may still be reading the k-space buffers the next pass stores into
    if (g.nranks > 1 && m->barrier && m->barrier(m, st)) return -1;
    // F2: outer = local x plane, rows = y; destination [ky][x][kz]
    TilePassArgs a = {};
    a.self_rank = -1;
    a.src = reinterpret_cast<const float2 *>(real); a.src_estride = g.pitch_c; a.src_ostride = plane;
    for (int d = 0; d < g.nranks; d++) a.dst[d] = reinterpret_cast<float2 *>(cplx_peers[d]);
    a.rows_per_rank = n / g.nranks; a.dst_estride = plane; a.dst_ostride = g.pitch_c; a.dst_ooffset = g.x0;
    a.ntile_k = (n / 2 + 1 + p->K - 1) / p->K; a.conj = 0; a.outer0 = g.x0; a.t = p->tN; a.xfer.active = 0; a.kt = m->ktab;
    if (g.nranks > 1 && m->stage) { if (staged_transpose(m, a, a.src, cplx_peers, g.nxl, g.x0, 0, st) || staged_transpose_wait(0, st)) return -1; }
    else {
        if (launch_tile(p, a, g.nxl, st)) return -1;
        if (g.nranks > 1) fpm_comm_bytes[3] += (unsigned long long) g.nxl * n * g.pitch_c * sizeof(float2) / g.nranks * (g.nranks - 1);
    }
    if (g.nranks > 1 && m->barrier && m->barrier(m, st)) return -1;
    // F3: in place on the local k-space buffer: outer = local ky plane, rows = kx
    TilePassArgs b = a;
    b.src = reinterpret_cast<const float2 *>(cplx_peers[g.rank]); b.src_estride = g.pitch_c; b.src_ostride = plane;
    b.dst[0] = reinterpret_cast<float2 *>(cplx_peers[g.rank]); b.rows_per_rank = n; b.dst_estride = g.pitch_c; b.dst_ostride = plane; b.dst_ooffset = 0;
    b.outer0 = g.y0;
    if (launch_tile(p, b, g.nyl, st)) return -1;
    return 0;
}

// Backward: cplx (preserved) -> real field, optionally multiplied by a k-space kernel on the way in.
// real_peers[d] is rank d's work buffer for the x- and y-pass; the last (z) pass writes `real_out`, which may
// be real_peers[rank] itself or any other buffer (e.g. cplx, for the in-place public pm_c2r).
// Two halves, so that on several GPUs the slab transpose of one field can travel while another field is finished and read out:
//   begin   barrier, B1 (x-pass, transposing, kernel fused in) into the peers' buffers -- with a staging mesh the copy-engine
//           pushes of `set` are left in flight;
//   finish  waits for those pushes, barrier, B2 (y-pass) and B3 (z-pass) in place.
int fpm_fft_c2r_begin(FpmMesh *m, const float *cplx, float *const *real_peers, const FpmTransferSpec *xfer, int set, cudaStream_t st)
{
    const FpmGeom &g = m->geom;
    const FpmFftPlan *p = m->plan;
    const int n = g.n;
    const size_t plane = (size_t) n * g.pitch_c;
    // B1: outer = local ky plane, rows = kx; destination [x][ky][kz]
    TilePassArgs a = {};
    a.self_rank = -1;
    a.src = reinterpret_cast<const float2 *>(cplx); a.src_estride = g.pitch_c; a.src_ostride = plane;
    for (int d = 0; d < g.nranks; d++) a.dst[d] = reinterpret_cast<float2 *>(real_peers[d]);
    a.rows_per_rank = n / g.nranks; a.dst_estride = plane; a.dst_ostride = g.pitch_c; a.dst_ooffset = g.y0;
    a.ntile_k = (n / 2 + 1 + p->K - 1) / p->K; a.conj = 1; a.outer0 = g.y0; a.t = p->tN; a.kt = m->ktab;
    if (xfer) a.xfer = *xfer; else a.xfer.active = 0;
    if (g.nranks > 1 && m->barrier && m->barrier(m, st)) return -1;
    if (g.nranks > 1 && (set ? m->stage2 : m->stage)) { if (staged_transpose(m, a, a.src, real_peers, g.nyl, g.y0, set, st)) return -1; }
    else {
        if (launch_tile(p, a, g.nyl, st)) return -1;
        if (g.nranks > 1) fpm_comm_bytes[3] += (unsigned long long) g.nyl * n * g.pitch_c * sizeof(float2) / g.nranks * (g.nranks - 1);
    }
    return 0;
}

int fpm_fft_c2r_finish(FpmMesh *m, float *const *real_peers, float *real_out, int set, cudaStream_t st)
{
    const FpmGeom &g = m->geom;
    const FpmFftPlan *p = m->plan;
    const int n = g.n;
    const size_t plane = (size_t) n * g.pitch_c;
    if (g.nranks > 1 && staged_transpose_wait(set, st)) return -1;
    if (g.nranks > 1 && m->barrier && m->barrier(m, st)) return -1;
    // B2: in place, outer = local x plane, rows = ky
    float *real = real_peers[g.rank];
    TilePassArgs b = {};
    b.self_rank = -1; b.conj = 1; b.t = p->tN; b.kt = m->ktab; b.xfer.active = 0;
    b.ntile_k = (n / 2 + 1 + p->K - 1) / p->K;
    b.src = reinterpret_cast<const float2 *>(real); b.src_estride = g.pitch_c; b.src_ostride = plane;
    b.dst[0] = reinterpret_cast<float2 *>(real);
    b.rows_per_rank = n; b.dst_estride = g.pitch_c; b.dst_ostride = plane; b.dst_ooffset = 0; b.outer0 = g.x0;
    if (launch_tile(p, b, g.nxl, st)) return -1;
    // B3
    ZPassArgs z; z.src = real; z.dst = real_out ? real_out : real; z.nrows = (size_t) g.nxl * n; z.pitch_c = g.pitch_c; z.scale = 1.f; z.th = p->tH; z.twN = p->d_twN;
    if (launch_z(p, z, 0, st)) return -1;
    return 0;
}

int fpm_fft_c2r(FpmMesh *m, const float *cplx, float *const *real_peers, float *real_out, const FpmTransferSpec *xfer, cudaStream_t st)
{
    if (fpm_fft_c2r_begin(m, cplx, real_peers, xfer, 0, st)) return -1;
    return fpm_fft_c2r_finish(m, real_peers, real_out, 0, st);
}

// adapter: the generic tile-pass description -> the TMA kernel's arguments (fft_tma.cu)
#include "fft_tma_args.h"
int fpm_fft_tma_pass_from_tile(int n, const TilePassArgs &a, int pitch_c, int nouter, cudaStream_t st)
{
    TmaPassArgs t;
    for (int d = 0; d < FPM_MAX_RANKS; d++) t.dst[d] = a.dst[d];
    t.rows_per_rank = a.rows_per_rank; t.dst_estride = a.dst_estride; t.dst_ostride = a.dst_ostride; t.dst_ooffset = a.dst_ooffset;
    t.self_rank = a.self_rank; t.self_ooffset = a.self_ooffset; t.self_dst = a.self_dst; t.self_estride = a.self_estride; t.self_ostride = a.self_ostride;
    t.ntile_k = 0; t.nouter = nouter; t.conj = a.conj; t.outer0 = a.outer0; t.tw = a.t.tw; t.xfer = a.xfer; t.kt = a.kt;
    return fpm_fft_tma_pass(n, a.src, pitch_c, nouter, t, st);
}
#endif
