// fastpm_b200 -- the contiguous (z) passes of the FFT for power-of-two meshes: real rows <-> half-complex rows.
//
// A persistent CTA walks over tiles of 8 adjacent rows.  Rows travel between HBM and shared memory as whole-row bulk
// copies (cp.async.bulk, 4-16 KB each, signalled on an mbarrier / tracked as bulk groups), so every DRAM access is a
// full-line burst no matter how the threads are mapped; the next tile's rows are fetched while the current tile is
// transformed.  The 8 rows play the part of the 8 "columns" of the strided tile pass (fft_tma.cu): the same
// register-resident three-stage transform (fft_reg.cuh) of length H = N/2 runs on the packed sequence
// z_j = x_2j + i x_2j+1, lanes running across rows (staged rows are padded by 2 complex so that these accesses are
// bank-conflict free).
//   forward : rows -> A | read (x scale) -> 3 stages -> Z to O (natural order) -> untangle X_k = E_k + w_N^k O_k,
//             k = 0..H -> O -> rows
//   backward: rows -> A | read X_e and X_(H-e), tangle in registers -> 3 stages (inverse by conjugation) -> O -> rows
// The generic shared-memory version (fft.cu: fft_zfwd_kernel / fft_zbwd_kernel) remains for other mesh sizes.
#include "common.cuh"
#include "fft_core.h"
#include "mesh.cuh"
#include "fft_reg.cuh"
#include <stdlib.h>

struct ZRowArgs {
    const float *src;
    float *dst;
    size_t nrows;
    int pitch_c;
    float scale;
    const float2 *twH;      // [H] exp(-2 pi i t / H)
    const float2 *twN;      // [N] exp(-2 pi i t / N)
    int late;               // 1: the next tile's rows are requested at the first barrier of exchange 2 instead of exchange 1
};

// PREFETCH: input rows (A) and the exchange / output staging buffer (O) are separate, the next tile is loaded while this
// one is transformed.  Otherwise one buffer serves both and overlap comes from several CTAs per SM.
template <int R1, int R2, int R3, bool FWD, bool PREFETCH, int MINB>
__global__ void __launch_bounds__(TmaCfg<R1, R2, R3>::T * 8, MINB)
fft_zrow_kernel(const ZRowArgs a)
{
    using C = TmaCfg<R1, R2, R3>;
    constexpr int K = 8;
    constexpr int H = C::N, E = C::E, T = C::T, M1 = C::M1;
    constexpr int P = H + 2;                                    // complex pitch of a staged row (16-byte multiple, banks spread)
    constexpr uint32_t IN_BYTES = (FWD ? H : H + 2) * 8, OUT_BYTES = (FWD ? H + 2 : H) * 8;
    FPM_DYN_SMEM(smem_raw, 128);
    float2 *A = reinterpret_cast<float2 *>(smem_raw);          // [8][P]
    float2 *O = PREFETCH ? A + K * P : A;                       // [8][P]; the exchange buffer B aliases it
    float *B = reinterpret_cast<float *>(O);
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int c = tid % K;                                      // row within the tile
    const int t = tid / K;
    const int ntiles = (int) (a.nrows / K);
    const size_t pitch_r = (size_t) (2 * a.pitch_c);

    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue_load = [&](int tile) {
        mbar_expect_tx(&bar, K * IN_BYTES);
        const float *g = a.src + (size_t) tile * K * pitch_r;
        #pragma unroll 1
        for (int r = 0; r < K; r++) bulk_load_1d(A + r * P, g + r * pitch_r, IN_BYTES, &bar);
    };

    int tile = blockIdx.x;
    if (PREFETCH && tid == 0 && tile < ntiles) issue_load(tile);
    uint32_t phase = 0;

    Fft3<R1, R2, R3, K> fx(B, t, c, a.twH);
    const float2 *Arow = A + c * P;
    float2 *Orow = O + c * P;

    #pragma unroll 1
    for (; tile < ntiles; tile += gridDim.x) {
        if (!PREFETCH && tid == 0) {
            bulk_wait_read0();                                  // the previous tile's stores have left the shared buffer
            issue_load(tile);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;

        float2 v[E];
        if (FWD) {
            #pragma unroll
            for (int k = 0; k < E; k++) {
                const float2 x = Arow[t + k * M1];
                v[k] = make_float2(x.x * a.scale, x.y * a.scale);
            }
        } else {
            #pragma unroll
            for (int k = 0; k < E; k++) {
                const int e = t + k * M1;
                const float2 xk = Arow[e], xh = Arow[H - e], w = __ldg(a.twN + e);
                // Z_e = (X_e + conj X_(H-e)) + i exp(+2 pi i e/N) (X_e - conj X_(H-e)); the conjugate feeds the forward core
                const float2 sm = make_float2(xk.x + xh.x, xk.y - xh.y), d = make_float2(xk.x - xh.x, xk.y + xh.y);
                const float2 wd = make_float2(w.x * d.x + w.y * d.y, w.x * d.y - w.y * d.x);       // conj(w) * d
                v[k] = make_float2(sm.x - wd.y, -(sm.y + wd.x));
            }
        }
        if (PREFETCH && tid == 0) bulk_wait_read0();            // O (= B) is free again before anybody writes it

        auto prefetch_next = [&]() {
            if (PREFETCH && tid == 0) {
                const int nxt = tile + gridDim.x;
                if (nxt < ntiles) issue_load(nxt);
            }
        };
        fx.run(v, [&]() { if (a.late == 0) prefetch_next(); }, [&]() { if (a.late) prefetch_next(); });

        // ---- third exchange: frequency kf = q1 + R1*q2 + R1*R2*q3 of row c -> O[c][kf]
        __syncthreads();
        #pragma unroll
        for (int i = 0; i < E / R3; i++) {
            const int b = t + i * T, q1 = b / R2, q2 = b - q1 * R2;
            float2 *zw = Orow + q1 + R1 * q2;
            #pragma unroll
            for (int q3 = 0; q3 < R3; q3++) {
                const float2 y = v[i * R3 + q3];
                zw[q3 * R1 * R2] = FWD ? y : make_float2(y.x, -y.y);
            }
        }
        if (FWD) {
            __syncthreads();
            float2 xH = make_float2(0.f, 0.f);
            #pragma unroll
            for (int j = 0; j < E; j++) {
                const int k = t + j * T;
                const int m = (k == 0) ? 0 : H - k;
                const float2 z1 = Orow[k];
                float2 z2 = Orow[m];
                z2.y = -z2.y;                                    // conj Z_(H-k)
                const float2 e = make_float2(0.5f * (z1.x + z2.x), 0.5f * (z1.y + z2.y));
                const float2 d = make_float2(z1.x - z2.x, z1.y - z2.y);
                const float2 o = make_float2(0.5f * d.y, -0.5f * d.x);
                const float2 w = __ldg(a.twN + k);
                v[j] = make_float2(e.x + (w.x * o.x - w.y * o.y), e.y + (w.x * o.y + w.y * o.x));
                if (k == 0) xH = make_float2(e.x - o.x, e.y - o.y);                // w_N^H = -1
            }
            __syncthreads();
            #pragma unroll
            for (int j = 0; j < E; j++) Orow[t + j * T] = v[j];
            if (t == 0) { Orow[H] = xH; Orow[H + 1] = make_float2(0.f, 0.f); }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            float *g = a.dst + (size_t) tile * K * pitch_r;
            #pragma unroll 1
            for (int r = 0; r < K; r++) bulk_store_1d(g + r * pitch_r, O + r * P, OUT_BYTES);
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait0();
}

#ifndef FPM_EMULATE          // host side: not part of the CPU emulation of the kernel (tests/emul/zrow_emul.cpp)
static int zrow_sm_count()
{
    static int nsm = 0;
    if (!nsm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev); if (nsm <= 0) nsm = 148; }
    return nsm;
}

template <int R1, int R2, int R3, bool PREFETCH, int MINB>
static int launch_zrow(const ZRowArgs &a, int forward, cudaStream_t st)
{
    using C = TmaCfg<R1, R2, R3>;
    const size_t smem = (size_t) (PREFETCH ? 2 : 1) * 8 * (C::N + 2) * sizeof(float2);
    static bool attr = false;
    if (!attr) {
        FPM_CUDA_OK(cudaFuncSetAttribute(fft_zrow_kernel<R1, R2, R3, true, PREFETCH, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        FPM_CUDA_OK(cudaFuncSetAttribute(fft_zrow_kernel<R1, R2, R3, false, PREFETCH, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr = true;
    }
    const size_t ntiles = a.nrows / 8;
    const size_t resident = (size_t) zrow_sm_count() * MINB;
    const unsigned grid = (unsigned) (ntiles < resident ? ntiles : resident);
    if (grid == 0) return 0;
    fpm_path_counter[FPM_PATH_FFT_ZROW]++;
    if (fpm_prof_on) fpm_prof_begin(FPM_K_FFT_Z, st);
    if (forward) fft_zrow_kernel<R1, R2, R3, true, PREFETCH, MINB><<<grid, C::T * 8, smem, st>>>(a);
    else fft_zrow_kernel<R1, R2, R3, false, PREFETCH, MINB><<<grid, C::T * 8, smem, st>>>(a);
    if (fpm_prof_on) fpm_prof_end(FPM_K_FFT_Z, st);
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_fft_zrow_supported(int n, size_t nrows)
{
    return (n == 512 || n == 768 || n == 1024 || n == 1536 || n == 2048 || n == 4096) && (nrows % 8 == 0) && nrows / 8 < ((size_t) 1 << 31);
}

int fpm_fft_zrow_pass(int n, const float *src, float *dst, size_t nrows, int pitch_c, float scale,
                      const float2 *twH, const float2 *twN, int forward, cudaStream_t st)
{
    static int late = -1;
    if (late < 0) { const char *e = getenv("FASTPM_B200_ZROW_LATE"); late = e ? atoi(e) : 0; }
    ZRowArgs a = { src, dst, nrows, pitch_c, scale, twH, twN, late };
    static int single = -1;       // FASTPM_B200_ZROW_SINGLE=1: one shared buffer per CTA (more CTAs per SM) at N = 2048
    if (single < 0) { const char *e = getenv("FASTPM_B200_ZROW_SINGLE"); single = e ? atoi(e) : 0; }
    switch (n) {
        case 512: return launch_zrow<8, 8, 4, true, 3>(a, forward, st);
        case 768: return launch_zrow<24, 4, 4, true, 4>(a, forward, st);            // rows of 384 = 24 * 4 * 4 complex
        case 1024: return launch_zrow<8, 8, 8, true, 2>(a, forward, st);
        case 1536: return launch_zrow<24, 8, 4, true, 2>(a, forward, st);          // rows of 768 = 24 * 8 * 4 complex
        case 2048: return single ? launch_zrow<16, 16, 4, false, 2>(a, forward, st) : launch_zrow<16, 16, 4, true, 1>(a, forward, st);
        case 4096: return launch_zrow<16, 16, 8, false, 1>(a, forward, st);
    }
    fpm_set_error("fpm_fft_zrow_pass: unsupported N = %d", n);
    return -1;
}
#endif
