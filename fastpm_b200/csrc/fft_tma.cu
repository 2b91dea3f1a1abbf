// fastpm_b200 -- the strided FFT pass for power-of-two meshes, Blackwell style.
//
// One persistent CTA per SM walks over tiles [N rows][K contiguous complex] of the mesh:
//   * the NEXT tile is fetched by TMA (cp.async.bulk.tensor, 256-row boxes) into shared buffer A and
//     signalled on an mbarrier while the current tile is being transformed, so HBM reads, the
//     arithmetic and the stores of consecutive tiles overlap inside one CTA;
//   * the transform itself lives in registers: every thread owns E = R1 elements of one column and runs
//     radix-R1 / R2 / R3 butterflies (N = R1*R2*R3, each <= 16) with two shared-memory exchanges in
//     between (buffer B, re and im in separate rounds so that A + B fit in 227 KB at N = 2048/4096);
//   * results go from registers straight to global memory, K*8 B contiguous per row, in natural
//     frequency order (k = q1 + R1*q2 + R1*R2*q3), optionally to a peer GPU's buffer (slab transpose).
// The gravity kernel (Green's function x i k_d, mesh.cuh) is applied as the tile leaves buffer A in the
// first pass of an inverse transform, exactly as in the generic path (fft.cu), which remains the
// fallback for the other mesh sizes with factors 3 or 5 and the cross-check for this one.  N = 1536 = 24 * 8 * 8 (the 3x mesh of
// an nc = 512 run, BASELINE configs[3]) has a radix-24 first stage (fft_reg.cuh).
#include "common.cuh"
#include "fft_core.h"
#include "mesh.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "fft_tma_args.h"

// selects dst[d] without indexing the parameter array dynamically (that would force a local-memory copy)
__device__ __forceinline__ float2 *pick_dst(const TmaPassArgs &a, int d)
{
    float2 *p = a.dst[0];
    #pragma unroll
    for (int i = 1; i < FPM_MAX_RANKS; i++) p = (d == i) ? a.dst[i] : p;
    return p;
}

#include "fft_reg.cuh"

// MULTI: the output rows are spread over several destination buffers (slab transpose on several GPUs).  A separate
// instantiation: the extra addressing of that path costs the one-GPU kernel registers and 30 % of its speed otherwise.
template <int R1, int R2, int R3, int K, bool MULTI>
__global__ void __launch_bounds__(TmaCfg<R1, R2, R3>::T * K, (TmaCfg<R1, R2, R3>::T * K <= 512 && R1 <= 16) ? 2 : 1)
fft_tma_kernel(const __grid_constant__ CUtensorMap tmap, const TmaPassArgs a)
{
    using C = TmaCfg<R1, R2, R3>;
    constexpr int N = C::N, E = C::E, T = C::T, M1 = C::M1, M2 = C::M2;
    constexpr int BOX = N < 256 ? N : 256;
    FPM_DYN_SMEM(smem_raw, 1024);
    const float2 *A = reinterpret_cast<const float2 *>(smem_raw);            // [N][K] complex, written by TMA
    float *B = reinterpret_cast<float *>(smem_raw + (size_t) N * K * 8);     // [N][K] floats, swizzled rows
    float2 *TW = reinterpret_cast<float2 *>(smem_raw + (size_t) N * K * 12); // [N] exp(-2 pi i t / N)
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int c = tid % K;
    const int t = tid / K;                      // 0 .. T-1
    const int ntiles = a.nouter * a.ntile_k;
    constexpr uint32_t tile_bytes = (uint32_t) N * K * 8u;

    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    for (int i = tid; i < N; i += T * K) TW[i] = __ldg(a.tw + i);
    __syncthreads();

    // A CTA may take `chunk` kz-adjacent tiles in a row before it jumps ahead (so that the two 64-byte halves of a line are
    // written by the same SM close in time); which passes gain from it is measured, see fpm_fft_tma_pass below.
    const int chunk = a.chunk;
    const int jump = 1 + ((int) gridDim.x - 1) * chunk;
    auto next_tile = [&](int tl) { return (tl % chunk != chunk - 1) ? tl + 1 : tl + jump; };
    int tile = blockIdx.x * chunk;
    if (tid == 0 && tile < ntiles) {
        const int o = tile / a.ntile_k, kz0 = (tile - o * a.ntile_k) * K;
        mbar_expect_tx(&bar, tile_bytes);
        #pragma unroll 1
        for (int r0 = 0; r0 < N; r0 += BOX) tma_load_3d(const_cast<float2 *>(A) + (size_t) r0 * K, &tmap, &bar, 2 * kz0, r0, o);
    }
    uint32_t phase = 0;

    // per-thread constant addresses
    const float2 *Ard = A + t * K + c;                                                   // + k*M1*K
    Fft3<R1, R2, R3, K, true> fx(B, t, c, TW);
    constexpr bool single = !MULTI;
    const int h = N / 2;

    // The gravity kernel, specialised for the tile pass (operations of fpm_apply_transfer, mesh.cuh, i.e. of
    // transfer.c:154-186 + gravity.c:17,21-64), kept to ~20 instructions per mode:
    //   * s = kk[ix] + kk[iy] + kk[iz] is a sum of three floats whose exponents differ by < 2^20: exact in double in any
    //     order; it is split into a float pair (s_hi, s_lo);
    //   * the reference's (float)((double) v * (1 / s)) is the correctly rounded float quotient v / s up to double
    //     rounding; here q = v * rcp(s_hi), corrected once with the exact remainder v - q * (s_hi + s_lo): the same float
    //     except when v / s lies within ~2^-46 (relative) of a rounding boundary -- about one mode in 4 million then
    //     differs by one float ulp, far below the round-off of the FFT itself;
    //   * the i*k_d product of two floats rounded to float is exactly what (float)((double) v * (double) kf) yields;
    //   * the modes the reference sets to zero (s == 0; gradient of a self-conjugate mode) are patched after the loop.
    const bool xf = a.xfer.active;

    #pragma unroll 1
    for (; tile < ntiles; tile = next_tile(tile)) {
        const int o = tile / a.ntile_k, kz0 = (tile - o * a.ntile_k) * K;
        float2 v[E];

        // ---- tile has landed: pull this thread's E elements (rows t + k*M1) out of A, then let the next tile's TMA
        //      start at once: everything below overlaps with that load
        mbar_wait(&bar, phase);
        phase ^= 1;
        #pragma unroll
        for (int k = 0; k < E; k++) v[k] = Ard[k * M1 * K];
        auto prefetch_next = [&]() {
            if (tid == 0) {
                const int nxt = next_tile(tile);
                if (nxt < ntiles) {
                    const int o2 = nxt / a.ntile_k, kz2 = (nxt - o2 * a.ntile_k) * K;
                    mbar_expect_tx(&bar, tile_bytes);
                    #pragma unroll 1
                    for (int r0 = 0; r0 < N; r0 += BOX) tma_load_3d(const_cast<float2 *>(A) + (size_t) r0 * K, &tmap, &bar, 2 * kz2, r0, o2);
                }
            }
        };
        if (a.early == 1) { __syncthreads(); prefetch_next(); }

        const int iy = a.outer0 + o, iz = kz0 + c;
        // Per-row factors kk[ix], kf[ix] with ix = t + k * M1 come from the reference-ordered float tables through L1.  The gradient
        // along the rows (direction 0) has its own loop with a compile-time stride: with a run-time stride (0 for the uniform
        // directions, M1 for this one) the pass took 27.2 ms at N = 2048 instead of 21.2.  Measured alternatives (round 2,
        // gpurun_out/r02f_passes_x*.txt): one interleaved { kk, kf } 8-byte load per element 27.5 ms, tables permuted to [t][k]
        // and read 16 bytes at a time 21.6 ms -- neither beats this form.
        if (xf && iz <= h) {
            const float *kkt = a.xfer.potorder == 1 ? a.kt.kk_finite : (a.xfer.potorder == 2 ? a.kt.kk_finite2 : a.kt.kk);
            const float *kft = a.xfer.gradorder == 0 ? a.kt.k : a.kt.k_finite;
            const int ngrad = a.xfer.ngrad;
            const bool has_pot = a.xfer.potorder >= 0, has_scale = a.xfer.scale != 1.0, negate = a.xfer.negate != 0;
            const bool sc_yz = (iy == 0 || iy == h) && (iz == 0 || iz == h);
            if (has_pot) {
                const double yz = (double) __ldg(kkt + iy) + (double) __ldg(kkt + iz);
                #pragma unroll
                for (int k = 0; k < E; k++) {
                    const float rowkk = __ldg(kkt + t + k * M1);
                    const double sd = yz + (double) rowkk;
                    const float s_hi = (float) sd;
                    const float s_lo = (float) (sd - (double) s_hi);
                    const float r = fpm_rcp_approx(s_hi);
                    const float qx = __fmul_rn(v[k].x, r), qy = __fmul_rn(v[k].y, r);
                    const float rx = __fmaf_rn(-qx, s_lo, __fmaf_rn(-qx, s_hi, v[k].x));
                    const float ry = __fmaf_rn(-qy, s_lo, __fmaf_rn(-qy, s_hi, v[k].y));
                    v[k].x = __fmaf_rn(rx, r, qx);
                    v[k].y = __fmaf_rn(ry, r, qy);
                }
                if (iy == 0 && iz == 0 && t == 0) v[0] = make_float2(0.f, 0.f);          // s == 0 (transfer.c:176-181)
            }
            if (negate) {
                #pragma unroll
                for (int k = 0; k < E; k++) { v[k].x = -v[k].x; v[k].y = -v[k].y; }
            }
            #pragma unroll
            for (int g = 0; g < 2; g++) {
                if (g < ngrad) {
                    // factor of element k: direction 0 runs along the rows of the tile (one factor per element), 1 and 2 are uniform
                    const int dir = a.xfer.graddir[g];
                    if (dir == 0) {
                        #pragma unroll
                        for (int k = 0; k < E; k++) {
                            const float f = __ldg(kft + t + k * M1);
                            const float re = -__fmul_rn(v[k].y, f), im = __fmul_rn(v[k].x, f);
                            v[k].x = re; v[k].y = im;
                        }
                    } else {
                        const float f = __ldg(kft + (dir == 1 ? iy : iz));
                        #pragma unroll
                        for (int k = 0; k < E; k++) {
                            const float re = -__fmul_rn(v[k].y, f), im = __fmul_rn(v[k].x, f);
                            v[k].x = re; v[k].y = im;
                        }
                    }
                    if (a.xfer.zero_selfconj && sc_yz && t == 0) {                           // gravity.c:48-56: ix in {0, N/2}
                        v[0] = make_float2(0.f, 0.f);
                        v[E / 2] = make_float2(0.f, 0.f);
                    }
                }
            }
            if (has_scale) {
                #pragma unroll
                for (int k = 0; k < E; k++) { v[k].x = (float) ((double) v[k].x * a.xfer.scale); v[k].y = (float) ((double) v[k].y * a.xfer.scale); }
            }
        }
        if (a.conj) {
            #pragma unroll
            for (int k = 0; k < E; k++) v[k].y = -v[k].y;
        }

        // ---- three register stages with two exchanges through B
        fx.run(v, [&]() { if (a.early == 0) prefetch_next(); }, [&]() { if (a.early == 3) prefetch_next(); });

        if (a.early == 5) prefetch_next();
        // ---- store: frequency kf = q1 + R1*q2 + R1*R2*q3, K*8 B contiguous per row
        const size_t obase = (size_t) (a.dst_ooffset + o) * a.dst_ostride + kz0 + c;
        if constexpr (single) {
            float2 *d0 = a.dst[0] + obase;
            const size_t s3 = (size_t) (R1 * R2) * a.dst_estride;
            #pragma unroll
            for (int i = 0; i < E / R3; i++) {
                const int b = t + i * T, q1 = b / R2, q2 = b - q1 * R2;
                float2 *p = d0 + (size_t) (q1 + R1 * q2) * a.dst_estride;
                #pragma unroll
                for (int q3 = 0; q3 < R3; q3++) {
                    float2 y = v[i * R3 + q3];
                    if (a.conj) y.y = -y.y;
                    p[q3 * s3] = y;
                }
            }
        } else {
            const size_t obase_self = (size_t) (a.self_ooffset + o) * a.self_ostride + kz0 + c;
            #pragma unroll
            for (int i = 0; i < E / R3; i++) {
                const int b = t + i * T, q1 = b / R2, q2 = b - q1 * R2;
                #pragma unroll
                for (int q3 = 0; q3 < R3; q3++) {
                    const int kf = q1 + R1 * q2 + R1 * R2 * q3;
                    const int d = kf / a.rows_per_rank, kl = kf - d * a.rows_per_rank;
                    float2 y = v[i * R3 + q3];
                    if (a.conj) y.y = -y.y;
                    if (d == a.self_rank) a.self_dst[(size_t) kl * a.self_estride + obase_self] = y;
                    else pick_dst(a, d)[(size_t) kl * a.dst_estride + obase] = y;
                }
            }
        }
    }
}

// ------------------------------------------------------------------ host side
#ifndef FPM_EMULATE          // not part of the CPU emulation of the kernel (tests/emul/tma_emul.cpp)
static PFN_cuTensorMapEncodeTiled get_encode()
{
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        void *p = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled) p;
    }
    return fn;
}

template <int R1, int R2, int R3, int K>
static int launch_cfg(const CUtensorMap &tmap, const TmaPassArgs &a, int nsm, cudaStream_t st)
{
    using C = TmaCfg<R1, R2, R3>;
    const size_t smem = (size_t) C::N * K * 8 + (size_t) C::N * K * 4 + (size_t) C::N * 8;
    static bool attr = false;
    if (!attr) {
        FPM_CUDA_OK(cudaFuncSetAttribute(fft_tma_kernel<R1, R2, R3, K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        FPM_CUDA_OK(cudaFuncSetAttribute(fft_tma_kernel<R1, R2, R3, K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr = true;
    }
    const int ntiles = a.nouter * a.ntile_k;
    const int per_sm = (C::T * K <= 512 && R1 <= 16) ? 2 : 1;        // 24 elements per thread: one 512-thread CTA per SM, 128 registers
    const int grid = ntiles < nsm * per_sm ? ntiles : nsm * per_sm;
    if (grid <= 0) return 0;
    fpm_path_counter[a.rows_per_rank == C::N ? FPM_PATH_FFT_TMA : FPM_PATH_FFT_TMA_MULTI]++;
    if (fpm_prof_on) fpm_prof_begin(FPM_K_FFT_TILE, st);
    if (a.rows_per_rank == C::N) fft_tma_kernel<R1, R2, R3, K, false><<<grid, C::T * K, smem, st>>>(tmap, a);
    else fft_tma_kernel<R1, R2, R3, K, true><<<grid, C::T * K, smem, st>>>(tmap, a);
    if (fpm_prof_on) fpm_prof_end(FPM_K_FFT_TILE, st);
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_fft_tma_supported(int n) { return n == 512 || n == 768 || n == 1024 || n == 1536 || n == 2048 || n == 4096; }
// tile width (complex per row): wide tiles give 128 B row segments with one 1024-thread CTA per SM, narrow ones two
// 512-thread CTAs per SM whose phases (shared-memory exchange, arithmetic, stores) interleave
static int g_narrow = -1;
int fpm_fft_tma_tile_k(int n)
{
    if (g_narrow < 0) { const char *e = getenv("FASTPM_B200_FFT_NARROW"); g_narrow = e ? atoi(e) : 0; }
    if (n == 4096) return 4;
    if (n == 1536) return 8;          // 24 * 8 * 8: 24 elements per thread, 64 threads per column
    if (n == 2048) return g_narrow ? 4 : 8;
    if (n == 1024) return g_narrow ? 8 : 16;
    return 16;
}

// One strided pass over `nouter` planes of `src` ([plane][N rows][pitch_c complex]).
int fpm_fft_tma_pass(int n, const float2 *src, int pitch_c, int nouter, const TmaPassArgs &args, cudaStream_t st)
{
    PFN_cuTensorMapEncodeTiled enc = get_encode();
    if (!enc) { fpm_set_error("cuTensorMapEncodeTiled is not available from the driver"); return -1; }
    const int K = fpm_fft_tma_tile_k(n);
    CUtensorMap tmap;
    cuuint64_t gdim[3] = { (cuuint64_t) 2 * pitch_c, (cuuint64_t) n, (cuuint64_t) nouter };
    cuuint64_t gstr[2] = { (cuuint64_t) pitch_c * 8, (cuuint64_t) n * pitch_c * 8 };
    cuuint32_t box[3] = { (cuuint32_t) (2 * K), (cuuint32_t) (n < 256 ? n : 256), 1 };
    cuuint32_t estr[3] = { 1, 1, 1 };
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *) src, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fpm_set_error("cuTensorMapEncodeTiled failed with code %d", (int) r); return -1; }
    static int nsm = 0;
    if (!nsm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev); if (nsm <= 0) nsm = 148; }
    TmaPassArgs a = args;
    static int early = -1;        // FASTPM_B200_TMA_EARLY: where the next tile's TMA is issued (diagnostic; 4 = measured best per pass)
    if (early < 0) { const char *e = getenv("FASTPM_B200_TMA_EARLY"); early = e ? atoi(e) : 4; }
    // When the next tile's TMA is issued: 0 at the first barrier of exchange 1; 1 right after the tile has been read (one more
    // barrier); 3 at the first barrier of exchange 2; 5 just before the stores.  Measured on B200 (scripts/fft_passes.py): a
    // load that overlaps the register/shared-memory phases only from exchange 2 on is best for the in-place passes and for the
    // pass with the k-space kernel (N = 2048: 14.6 ms against 18-20 ms at position 0); the scattered stores of a transposing
    // pass collide with an in-flight load, so at N = 2048 the plain transposing pass issues it last (22 ms against 24-27 ms).
    const bool transposing = args.dst_estride != (size_t) pitch_c;
    a.early = early == 4 ? ((transposing && !args.xfer.active && n >= 2048) ? 5 : 3) : early;
    a.nouter = nouter;
    a.ntile_k = (n / 2 + 1 + K - 1) / K;
    // kz-adjacent tiles a CTA processes back to back, so that the two 64-byte halves of an output line meet in L2 microseconds
    // apart instead of whenever the neighbouring CTA gets there (ncu at N = 1536: the transposing passes READ 1.8 x the mesh from
    // DRAM -- half-written lines are completed from memory -- and only 1.02 x in place).  Measured (scripts/fft_passes.py,
    // gpurun_out/r02w_passes_*): N = 1536 transposing 10.7 -> 7.8 ms plain, 10.2 -> 9.3 ms with the force kernel at 4 tiles, in
    // place 6.3 -> 6.9 (worse); N = 2048 nothing to gain (22.2 -> 21.5 / 22.2 plain, the other passes 6-30 % slower).  Hence: 4 for
    // the transposing passes at N = 1536, 1 everywhere else; FASTPM_B200_TMA_CHUNK overrides.
    static int chunk = -1;
    if (chunk < 0) { const char *e = getenv("FASTPM_B200_TMA_CHUNK"); chunk = e ? atoi(e) : 0; }
    a.chunk = chunk > 0 ? chunk : ((n == 1536 && transposing) ? 4 : 1);
    switch (n) {
        case 512: return launch_cfg<8, 8, 8, 16>(tmap, a, nsm, st);
        case 768: return launch_cfg<24, 8, 4, 16>(tmap, a, nsm, st);
        case 1536: return launch_cfg<24, 8, 8, 8>(tmap, a, nsm, st);
        case 1024: return K == 8 ? launch_cfg<16, 16, 4, 8>(tmap, a, nsm, st) : launch_cfg<16, 16, 4, 16>(tmap, a, nsm, st);
        case 2048: return K == 4 ? launch_cfg<16, 16, 8, 4>(tmap, a, nsm, st) : launch_cfg<16, 16, 8, 8>(tmap, a, nsm, st);
        case 4096: return launch_cfg<16, 16, 16, 4>(tmap, a, nsm, st);
    }
    fpm_set_error("fpm_fft_tma_pass: unsupported N = %d", n);
    return -1;
}
#endif
