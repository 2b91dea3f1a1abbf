// fastpm_b200 -- register-resident FFT building blocks shared by the TMA tile pass (fft_tma.cu) and the row
// (z) passes (fft_zrow.cu): PTX wrappers for mbarrier / TMA, the unrolled radix-2^k register FFT, and the
// three-stage transform of one column held E = R1 elements per thread with two shared-memory exchanges.
#pragma once
#include "common.cuh"
#include <cuda.h>

#ifndef FPM_EMULATE
// ------------------------------------------------------------------ small PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
// reciprocal estimate, within one ulp (MUFU.RCP)
__device__ __forceinline__ float fpm_rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// 1-D bulk copies (TMA without a tensor map): global -> shared signalled on an mbarrier, shared -> global in bulk groups.
// Addresses 16-byte aligned, sizes multiples of 16 bytes.
__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void *gdst, const void *smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// makes this thread's generic-proxy shared-memory writes visible to the async proxy (bulk stores)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#else
// ------------------------------------------------------------------ CPU emulation of the wrappers (tests/emul/*.cpp)
// Every CUDA thread is an OS thread, __syncthreads() a pthread barrier, shared memory an ordinary array.  Asynchronous copies
// are done synchronously by the issuing thread; the mbarrier keeps a completed-phase count (low word) and the bytes still
// expected (high word), so that waiting threads really wait for the data.  A "tensor map" is a plain descriptor.
struct FpmEmulTmap { const float *base; uint64_t gstr_bytes[2]; uint32_t box[3]; };      // lives in the bytes of a CUtensorMap
inline void mbar_init(uint64_t *bar, int) { __atomic_store_n(bar, 0ull, __ATOMIC_SEQ_CST); }
inline float fpm_rcp_approx(float x) { return 1.0f / x; }
inline void mbar_fence_init() {}
inline void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { __atomic_fetch_add(bar, (uint64_t) bytes << 32, __ATOMIC_SEQ_CST); }
inline void fpm_emul_complete_tx(uint64_t *bar, uint32_t bytes)
{
    uint64_t v = __atomic_sub_fetch(bar, (uint64_t) bytes << 32, __ATOMIC_SEQ_CST);
    if ((v >> 32) == 0) __atomic_fetch_add(bar, 1ull, __ATOMIC_SEQ_CST);                   // phase complete
}
inline void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (((uint32_t) __atomic_load_n(bar, __ATOMIC_SEQ_CST) & 1u) == parity) sched_yield();
}
inline void tma_load_3d(void *smem_dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1, int c2)
{
    const FpmEmulTmap *t = reinterpret_cast<const FpmEmulTmap *>(tmap);
    float *d = (float *) smem_dst;
    for (uint32_t z = 0; z < t->box[2]; z++)
        for (uint32_t y = 0; y < t->box[1]; y++) {
            const float *src = t->base + ((size_t) (c2 + z) * t->gstr_bytes[1] + (size_t) (c1 + y) * t->gstr_bytes[0]) / 4 + c0;
            memcpy(d, src, t->box[0] * 4);
            d += t->box[0];
        }
    fpm_emul_complete_tx(bar, t->box[0] * t->box[1] * t->box[2] * 4);
}
inline void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) { memcpy(smem_dst, gsrc, bytes); fpm_emul_complete_tx(bar, bytes); }
inline void bulk_store_1d(void *gdst, const void *smem_src, uint32_t bytes) { memcpy(gdst, smem_src, bytes); }
inline void bulk_commit() {}
inline void bulk_wait_read0() {}
inline void bulk_wait0() {}
inline void fence_async_smem() {}
#endif

// ------------------------------------------------------------------ packed complex arithmetic
// Blackwell issues FADD2 / FMUL2 / FFMA2 on a float2 held in an aligned register pair, with per-half negate and swap / broadcast
// operand modifiers: a complex add or subtract is ONE instruction and a complex product two (no moves), against 2 and 4 scalar
// ones.  Measured on B200 (round 2, scripts/ubench/atomics.cu and scripts/fft_passes.py): FFMA2 issues at HALF the rate of FFMA
// (2.26 against 2 x 1.06 cycles per warp and SMSP), so the packed forms save issue slots but no pipe time, and the N = 2048 passes
// ran 7 % SLOWER with them (in place 15.5 ms against 14.5; transposing 22.8 against 22.0; rows 19.1 / 17.1 against 18.0 / 16.0):
// the passes are bound by shared-memory / LSU wavefronts, not by FP32 issue.  Kept as an opt-in (-DFPM_PACKED_F32) for reference.
#if !defined(FPM_EMULATE) && defined(FPM_PACKED_F32)
__device__ __forceinline__ float2 c2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 c2sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 c2mul(float2 d, float2 w)
{
    const float2 t = __fmul2_rn(make_float2(d.x, d.x), w);
    return __ffma2_rn(make_float2(d.y, d.y), make_float2(-w.y, w.x), t);
}
#else
__device__ __forceinline__ float2 c2add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 c2sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 c2mul(float2 d, float2 w) { return make_float2(d.x * w.x - d.y * w.y, d.x * w.y + d.y * w.x); }
#endif

// ------------------------------------------------------------------ register FFT (radix-2 DIF, unrolled)
// twiddle exp(-2 pi i idx/16), idx = 0..7, as compile-time constants
__device__ __forceinline__ float2 w16(int idx)
{
    constexpr float c[8] = { 1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                             0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f };
    constexpr float s[8] = { 0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f,
                             -1.0f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f };
    return make_float2(c[idx], s[idx]);
}

// in-place DIF on R registers; afterwards X[q] sits in v[bitrev_R(q)]
template <int R>
__device__ __forceinline__ void fft_reg(float2 (&v)[R])
{
    #pragma unroll
    for (int half = R / 2; half >= 1; half >>= 1) {
        #pragma unroll
        for (int blk = 0; blk < R; blk += 2 * half) {
            #pragma unroll
            for (int j = 0; j < half; j++) {
                const float2 a = v[blk + j], b = v[blk + j + half];
                v[blk + j] = c2add(a, b);
                const float2 d = c2sub(a, b);
                const int idx = j * (8 / half);              // exp(-2 pi i j / (2 half)) = w16(j * 16 / (2 half))
                if (idx == 0) v[blk + j + half] = d;
                else if (idx == 4) v[blk + j + half] = make_float2(d.y, -d.x);
                else v[blk + j + half] = c2mul(d, w16(idx));
            }
        }
    }
}
template <int R> __device__ __forceinline__ constexpr int bitrev(int q)
{
    int r = 0;
    for (int b = 1; b < R; b <<= 1) { r = (r << 1) | (q & 1); q >>= 1; }
    return r;
}

// ---- radix 24 = 3 x 8 (mesh sizes 3 * 2^k: N = 1536 = 24 * 8 * 8, rows of N/2 = 768 = 24 * 8 * 4)
// One radix-3 DIF step over the three thirds of the array, twiddled by w24^(j r), then a radix-8 transform of each third:
// X[3 p + r] ends up in v[8 r + bitrev8(p)].
__device__ __forceinline__ float2 w24(int idx)          // exp(-2 pi i idx / 24), idx = 0 .. 14
{
    constexpr float c[15] = { 1.0f, 0.96592582628906829f, 0.86602540378443865f, 0.70710678118654752f, 0.5f, 0.25881904510252076f, 0.0f,
                              -0.25881904510252076f, -0.5f, -0.70710678118654752f, -0.86602540378443865f, -0.96592582628906829f, -1.0f,
                              -0.96592582628906829f, -0.86602540378443865f };
    constexpr float s[15] = { 0.0f, -0.25881904510252076f, -0.5f, -0.70710678118654752f, -0.86602540378443865f, -0.96592582628906829f, -1.0f,
                              -0.96592582628906829f, -0.86602540378443865f, -0.70710678118654752f, -0.5f, -0.25881904510252076f, 0.0f,
                              0.25881904510252076f, 0.5f };
    return make_float2(c[idx], s[idx]);
}
template <> __device__ __forceinline__ void fft_reg<24>(float2 (&v)[24])
{
    constexpr float S3 = 0.86602540378443865f;           // sin(2 pi / 3)
    #pragma unroll
    for (int j = 0; j < 8; j++) {
        const float2 a = v[j], b = v[j + 8], c = v[j + 16];
        const float2 sum = c2add(b, c), dif = c2sub(b, c);
        v[j] = c2add(a, sum);
        const float2 m = make_float2(a.x - 0.5f * sum.x, a.y - 0.5f * sum.y);
        const float2 r = make_float2(S3 * dif.y, -S3 * dif.x);            // -i sin(2 pi / 3) (b - c)
        const float2 y1 = c2add(m, r), y2 = c2sub(m, r);                  // a + w3 b + w3^2 c,  a + w3^2 b + w3 c
        v[j + 8] = j == 0 ? y1 : c2mul(y1, w24(j));
        v[j + 16] = j == 0 ? y2 : c2mul(y2, w24(2 * j));
    }
    #pragma unroll
    for (int r = 0; r < 3; r++) {
        float2 w[8];
        #pragma unroll
        for (int j = 0; j < 8; j++) w[j] = v[8 * r + j];
        fft_reg<8>(w);
        #pragma unroll
        for (int j = 0; j < 8; j++) v[8 * r + j] = w[j];
    }
}
// where fft_reg<R> leaves output q
template <int R> __device__ __forceinline__ constexpr int fft_pos(int q) { return R == 24 ? 8 * (q % 3) + bitrev<8>(q / 3) : bitrev<R>(q); }

// ------------------------------------------------------------------ the kernel

template <int R1, int R2, int R3> struct TmaCfg {
    static constexpr int N = R1 * R2 * R3;
    static constexpr int E = R1;               // elements per thread
    static constexpr int T = N / E;            // threads per column
    static constexpr int M1 = N / R1;          // = T
    static constexpr int M2 = M1 / R2;         // = R3
};
template <int R> struct Log2Of { static constexpr int v = (R >= 16) ? 4 : (R >= 8 ? 3 : (R >= 4 ? 2 : (R >= 2 ? 1 : 0))); };

// Exchange buffer B: floats [N][K]; physical row = row ^ ((row >> log2 R3) & MASK), MASK = 32/K - 1, keeps every
// access pattern of the three stages on distinct banks.  Because MASK < R3 <= M1 the swizzle only ever touches
// bits that come from a single index of each pattern, so all addresses below are "per-thread base + immediate":
//   X1 write   rows q*M1 + t            -> q*M1 + swz(t)
//   X1 read /  rows q1*M1 + k*M2 + p2   -> q1*M1 + k*M2 + (p2 ^ (k & MASK))
//   X2 write
//   X2 read    rows b*R3 + k            -> b*R3 + (k ^ (b & MASK))
// Input: v[k] = element t + k*M1 of the column.  Output: v[i*R3 + q3] = frequency q1 + R1*q2 + R1*R2*q3 with
// (q1, q2) = ((t + i*T) / R2, (t + i*T) % R2).
// TWS: the twiddle table lives in shared memory (copied there once per CTA) instead of being read through L1
template <int R1, int R2, int R3, int K, bool TWS = false>
struct Fft3 {
    using C = TmaCfg<R1, R2, R3>;
    static constexpr int N = C::N, E = C::E, T = C::T, M1 = C::M1, M2 = C::M2;
    static constexpr int SH = Log2Of<R3>::v;
    static constexpr int MASK = 32 / K - 1;
    static_assert(R3 > 1 && M2 == R3, "three-stage configurations only");
    static_assert(MASK < R3 && (T % (MASK + 1)) == 0 && (1 << SH) * (MASK + 1) <= M1 && (T % M2) == 0, "swizzle assumptions");

    float *Bx1w, *Bx2, *Bx3;
    int pm[MASK + 1];
    int bm, tw2i, t;
    const float2 *tw;

    __device__ __forceinline__ Fft3(float *B, int t_, int c, const float2 *tw_) : t(t_), tw(tw_)
    {
        Bx1w = B + (t ^ ((t >> SH) & MASK)) * K + c;                    // + q*M1*K
        const int p2 = t % M2, q1_0 = t / M2;                           // butterfly b = t + i*T -> (q1_0 + i*T/M2, p2)
        Bx2 = B + (q1_0 * M1) * K + c;                                  // + i*(T/M2)*M1*K + k*M2*K + pm[k & MASK]
        #pragma unroll
        for (int m = 0; m <= MASK; m++) pm[m] = (p2 ^ m) * K;
        Bx3 = B + (t * R3) * K + c;                                     // + i*T*R3*K + ((k ^ bm))*K
        bm = t & MASK;
        tw2i = p2 * R1;                                                 // w_M1^(q p2) = tw[q*p2*R1]
    }

    template <typename Hook>
    __device__ __forceinline__ void run(float2 (&v)[E], Hook after_first_barrier) { run(v, after_first_barrier, []() {}); }

    // hooks: called by every thread right after the first barrier of exchange 1 / of exchange 2
    template <typename Hook, typename Hook2>
    __device__ __forceinline__ void run(float2 (&v)[E], Hook after_first_barrier, Hook2 after_x2_barrier)
    {
        const float2 *tw1 = tw;
        // ---- stage 1: radix R1 over rows t + k*M1; output q goes to row q*M1 + t, times w_N^(q t)
        fft_reg<R1>(v);
        #pragma unroll
        for (int q = 1; q < R1; q++) {
            const float2 w = TWS ? tw1[q * t] : __ldg(tw1 + q * t);
            v[fft_pos<R1>(q)] = c2mul(v[fft_pos<R1>(q)], w);
        }

        // ---- exchange 1 (B), then stage 2
        float2 u[E];
        #pragma unroll
        for (int half = 0; half < 2; half++) {
            __syncthreads();                    // B free (and, in round 0: every thread is done reading A)
            if (half == 0) after_first_barrier();
            #pragma unroll
            for (int q = 0; q < R1; q++) {
                const float2 y = v[fft_pos<R1>(q)];
                Bx1w[q * M1 * K] = half ? y.y : y.x;
            }
            __syncthreads();
            #pragma unroll
            for (int i = 0; i < E / R2; i++) {
                #pragma unroll
                for (int k = 0; k < R2; k++) {
                    const float val = Bx2[i * (T / M2) * M1 * K + k * M2 * K + pm[k & MASK]];
                    if (half) u[i * R2 + k].y = val; else u[i * R2 + k].x = val;
                }
            }
        }
        #pragma unroll
        for (int i = 0; i < E / R2; i++) {
            float2 w2[R2];
            #pragma unroll
            for (int k = 0; k < R2; k++) w2[k] = u[i * R2 + k];
            fft_reg<R2>(w2);
            #pragma unroll
            for (int q = 1; q < R2; q++) {
                const float2 w = TWS ? tw[q * tw2i] : __ldg(tw + q * tw2i);
                w2[bitrev<R2>(q)] = c2mul(w2[bitrev<R2>(q)], w);
            }
            #pragma unroll
            for (int q = 0; q < R2; q++) u[i * R2 + q] = w2[bitrev<R2>(q)];      // natural order: u[i*R2 + q2]
        }

        // ---- exchange 2, then stage 3 (radix R3, no twiddles) on butterflies b = t + i*T = q1*R2 + q2
        #pragma unroll
        for (int half = 0; half < 2; half++) {
            __syncthreads();
            if (half == 0) after_x2_barrier();
            #pragma unroll
            for (int i = 0; i < E / R2; i++) {
                #pragma unroll
                for (int q = 0; q < R2; q++)
                    Bx2[i * (T / M2) * M1 * K + q * M2 * K + pm[q & MASK]] = half ? u[i * R2 + q].y : u[i * R2 + q].x;
            }
            __syncthreads();
            #pragma unroll
            for (int i = 0; i < E / R3; i++) {
                #pragma unroll
                for (int k = 0; k < R3; k++) {
                    // k = kh*(MASK+1) + kl: (k ^ bm) = kh*(MASK+1) + (kl ^ bm)
                    const float val = Bx3[i * T * R3 * K + (k & ~MASK) * K + ((k & MASK) ^ bm) * K];
                    if (half) v[i * R3 + k].y = val; else v[i * R3 + k].x = val;
                }
            }
        }
        #pragma unroll
        for (int i = 0; i < E / R3; i++) {
            float2 w3[R3];
            #pragma unroll
            for (int k = 0; k < R3; k++) w3[k] = v[i * R3 + k];
            fft_reg<R3>(w3);
            #pragma unroll
            for (int q = 0; q < R3; q++) v[i * R3 + q] = w3[bitrev<R3>(q)];
        }

    }
};
