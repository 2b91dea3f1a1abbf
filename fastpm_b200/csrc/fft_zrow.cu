// fastpm_b200 -- the contiguous (z) passes of the FFT for power-of-two meshes: real rows <-> half-complex rows.
//
// A CTA takes 8 adjacent rows; the 8 rows play the part of the 8 "columns" of the strided tile pass (fft_tma.cu),
// so the same register-resident three-stage transform (fft_reg.cuh) of length H = N/2 runs on the packed sequence
// z_j = x_2j + i x_2j+1, with lanes running ACROSS rows: every warp-level load or store touches 8 rows x 32 B
// (one full sector per row, every byte used), 16 independent loads in flight per thread, two CTAs per SM.
//   forward : load (x scale) -> 3 stages -> Z to shared (natural order) -> untangle X_k = E_k + w_N^k O_k, k = 0..H
//   backward: load X_e and X_(H-e), tangle in registers -> 3 stages (inverse by conjugation) -> shared -> real rows
// The generic shared-memory version (fft.cu: fft_zfwd_kernel / fft_zbwd_kernel) remains for other mesh sizes.
#include "common.cuh"
#include "fft_core.h"
#include "mesh.cuh"
#include "fft_reg.cuh"

struct ZRowArgs {
    const float *src;
    float *dst;
    size_t nrows;
    int pitch_c;
    float scale;
    const float2 *twH;      // [H] exp(-2 pi i t / H)
    const float2 *twN;      // [N] exp(-2 pi i t / N)
};

template <int R1, int R2, int R3, bool FWD>
__global__ void __launch_bounds__(TmaCfg<R1, R2, R3>::T * 8, (TmaCfg<R1, R2, R3>::T * 8 <= 512) ? 2 : 1)
fft_zrow_kernel(const ZRowArgs a)
{
    using C = TmaCfg<R1, R2, R3>;
    constexpr int K = 8;
    constexpr int H = C::N, E = C::E, T = C::T, M1 = C::M1;
    constexpr int SHX = Log2Of<R1>::v;
    static_assert(T % (2 << SHX) == 0, "Z swizzle assumption");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *Z = reinterpret_cast<float2 *>(smem_raw);          // [H][8] complex, row-swizzled; B aliases its first half
    float *B = reinterpret_cast<float *>(smem_raw);

    const int tid = threadIdx.x;
    const int c = tid % K;                                      // row within the tile
    const int t = tid / K;
    const size_t row = (size_t) blockIdx.x * K + c;
    const float2 *in = reinterpret_cast<const float2 *>(a.src + row * (size_t) (2 * a.pitch_c));
    float2 *out = reinterpret_cast<float2 *>(a.dst + row * (size_t) (2 * a.pitch_c));

    Fft3<R1, R2, R3, K> fx(B, t, c, a.twH);
    float2 v[E];
    if (FWD) {
        #pragma unroll
        for (int k = 0; k < E; k++) {
            float2 x = __ldg(in + t + k * M1);
            v[k] = make_float2(x.x * a.scale, x.y * a.scale);
        }
    } else {
        #pragma unroll
        for (int k = 0; k < E; k++) {
            const int e = t + k * M1;
            const float2 xk = __ldg(in + e), xh = __ldg(in + (H - e)), w = __ldg(a.twN + e);
            // Z_e = (X_e + conj X_(H-e)) + i exp(+2 pi i e/N) (X_e - conj X_(H-e)); the conjugate feeds the forward core
            const float2 sm = make_float2(xk.x + xh.x, xk.y - xh.y), d = make_float2(xk.x - xh.x, xk.y + xh.y);
            const float2 wd = make_float2(w.x * d.x + w.y * d.y, w.x * d.y - w.y * d.x);       // conj(w) * d
            v[k] = make_float2(sm.x - wd.y, -(sm.y + wd.x));
        }
    }

    fx.run(v, []() {});

    // ---- third exchange: frequency kf = q1 + R1*q2 + R1*R2*q3 -> Z row kf ^ ((kf >> log2 R1) & 1)
    __syncthreads();
    #pragma unroll
    for (int i = 0; i < E / R3; i++) {
        const int b = t + i * T, q1 = b / R2, q2 = b - q1 * R2;
        float2 *zw = Z + ((q1 ^ (q2 & 1)) + R1 * q2) * K + c;
        #pragma unroll
        for (int q3 = 0; q3 < R3; q3++) zw[q3 * R1 * R2 * K] = v[i * R3 + q3];
    }
    __syncthreads();
    const float2 *zr = Z + (t ^ ((t >> SHX) & 1)) * K + c;       // + j*T*K : frequency k = t + j*T
    if (FWD) {
        #pragma unroll
        for (int j = 0; j < E; j++) {
            const int k = t + j * T;
            const int m = (k == 0) ? 0 : H - k;
            const float2 z1 = zr[j * T * K];
            float2 z2 = Z[(m ^ ((m >> SHX) & 1)) * K + c];
            z2.y = -z2.y;                                        // conj Z_(H-k)
            const float2 e = make_float2(0.5f * (z1.x + z2.x), 0.5f * (z1.y + z2.y));
            const float2 d = make_float2(z1.x - z2.x, z1.y - z2.y);
            const float2 o = make_float2(0.5f * d.y, -0.5f * d.x);
            const float2 w = __ldg(a.twN + k);
            out[k] = make_float2(e.x + (w.x * o.x - w.y * o.y), e.y + (w.x * o.y + w.y * o.x));
            if (k == 0) out[H] = make_float2(e.x - o.x, e.y - o.y);        // w_N^H = -1
        }
    } else {
        #pragma unroll
        for (int j = 0; j < E; j++) {
            const float2 z = zr[j * T * K];
            out[t + j * T] = make_float2(z.x, -z.y);
        }
    }
}

template <int R1, int R2, int R3>
static int launch_zrow(const ZRowArgs &a, int forward, cudaStream_t st)
{
    using C = TmaCfg<R1, R2, R3>;
    const size_t smem = (size_t) C::N * 8 * sizeof(float2);
    static bool attr = false;
    if (!attr) {
        FPM_CUDA_OK(cudaFuncSetAttribute(fft_zrow_kernel<R1, R2, R3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        FPM_CUDA_OK(cudaFuncSetAttribute(fft_zrow_kernel<R1, R2, R3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr = true;
    }
    const unsigned grid = (unsigned) (a.nrows / 8);
    if (grid == 0) return 0;
    if (fpm_prof_on) fpm_prof_begin(FPM_K_FFT_Z, st);
    if (forward) fft_zrow_kernel<R1, R2, R3, true><<<grid, C::T * 8, smem, st>>>(a);
    else fft_zrow_kernel<R1, R2, R3, false><<<grid, C::T * 8, smem, st>>>(a);
    if (fpm_prof_on) fpm_prof_end(FPM_K_FFT_Z, st);
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_fft_zrow_supported(int n, size_t nrows) { return (n == 512 || n == 1024 || n == 2048 || n == 4096) && (nrows % 8 == 0); }

int fpm_fft_zrow_pass(int n, const float *src, float *dst, size_t nrows, int pitch_c, float scale,
                      const float2 *twH, const float2 *twN, int forward, cudaStream_t st)
{
    ZRowArgs a = { src, dst, nrows, pitch_c, scale, twH, twN };
    switch (n) {
        case 512: return launch_zrow<8, 8, 4>(a, forward, st);
        case 1024: return launch_zrow<8, 8, 8>(a, forward, st);
        case 2048: return launch_zrow<16, 16, 4>(a, forward, st);
        case 4096: return launch_zrow<16, 16, 8>(a, forward, st);
    }
    fpm_set_error("fpm_fft_zrow_pass: unsupported N = %d", n);
    return -1;
}
