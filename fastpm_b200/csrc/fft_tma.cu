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
// fallback for mesh sizes with factors 3 or 5 and the cross-check for this one.
#include "common.cuh"
#include "fft_core.h"
#include "mesh.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>

#include "fft_tma_args.h"

// selects dst[d] without indexing the parameter array dynamically (that would force a local-memory copy)
__device__ __forceinline__ float2 *pick_dst(const TmaPassArgs &a, int d)
{
    float2 *p = a.dst[0];
    #pragma unroll
    for (int i = 1; i < FPM_MAX_RANKS; i++) p = (d == i) ? a.dst[i] : p;
    return p;
}

// ------------------------------------------------------------------ small PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
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
                v[blk + j] = make_float2(a.x + b.x, a.y + b.y);
                const float2 d = make_float2(a.x - b.x, a.y - b.y);
                const int idx = j * (8 / half);              // exp(-2 pi i j / (2 half)) = w16(j * 16 / (2 half))
                if (idx == 0) v[blk + j + half] = d;
                else if (idx == 4) v[blk + j + half] = make_float2(d.y, -d.x);
                else {
                    const float2 w = w16(idx);
                    v[blk + j + half] = make_float2(d.x * w.x - d.y * w.y, d.x * w.y + d.y * w.x);
                }
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

// ------------------------------------------------------------------ the kernel
template <int R1, int R2, int R3> struct TmaCfg {
    static constexpr int N = R1 * R2 * R3;
    static constexpr int E = R1;               // elements per thread
    static constexpr int T = N / E;            // threads per column
    static constexpr int M1 = N / R1;          // = T
    static constexpr int M2 = M1 / R2;         // = R3
};

// Exchange buffer B: floats [N][K]; physical row = row ^ ((row >> log2 R3) & MASK), MASK = 32/K - 1, keeps every
// access pattern of the three stages on distinct banks.  Because MASK < R3 <= M1 the swizzle only ever touches
// bits that come from a single index of each pattern, so all addresses below are "per-thread base + immediate":
//   X1 write   rows q*M1 + t            -> q*M1 + swz(t)
//   X1 read /  rows q1*M1 + k*M2 + p2   -> q1*M1 + k*M2 + (p2 ^ (k & MASK))
//   X2 write
//   X2 read    rows b*R3 + k            -> b*R3 + (k ^ (b & MASK))
template <int R3> struct Log2R3 { static constexpr int v = (R3 >= 16) ? 4 : (R3 >= 8 ? 3 : (R3 >= 4 ? 2 : (R3 >= 2 ? 1 : 0))); };

struct XferThread { float kf_y, kf_z; bool sc_yz; bool pad; };

template <int R1, int R2, int R3, int K>
__global__ void __launch_bounds__(TmaCfg<R1, R2, R3>::T * K, 1)
fft_tma_kernel(const __grid_constant__ CUtensorMap tmap, const TmaPassArgs a)
{
    using C = TmaCfg<R1, R2, R3>;
    constexpr int N = C::N, E = C::E, T = C::T, M1 = C::M1, M2 = C::M2;
    constexpr int BOX = N < 256 ? N : 256;
    constexpr int SH = Log2R3<R3>::v;
    constexpr int MASK = 32 / K - 1;
    static_assert(R3 > 1 && M2 == R3, "three-stage configurations only");
    static_assert(MASK < R3 && (T % (MASK + 1)) == 0 && (1 << SH) * (MASK + 1) <= M1, "swizzle assumptions");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const float2 *A = reinterpret_cast<const float2 *>(smem_raw);            // [N][K] complex, written by TMA
    float *B = reinterpret_cast<float *>(smem_raw + (size_t) N * K * 8);     // [N][K] floats, swizzled rows
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int c = tid % K;
    const int t = tid / K;                      // 0 .. T-1
    const int ntiles = a.nouter * a.ntile_k;
    constexpr uint32_t tile_bytes = (uint32_t) N * K * 8u;

    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int tile = blockIdx.x;
    if (tid == 0 && tile < ntiles) {
        const int o = tile / a.ntile_k, kz0 = (tile - o * a.ntile_k) * K;
        mbar_expect_tx(&bar, tile_bytes);
        #pragma unroll 1
        for (int r0 = 0; r0 < N; r0 += BOX) tma_load_3d(const_cast<float2 *>(A) + (size_t) r0 * K, &tmap, &bar, 2 * kz0, r0, o);
    }
    uint32_t phase = 0;

    // per-thread constant addresses
    const float2 *Ard = A + t * K + c;                                                   // + k*M1*K
    float *Bx1w = B + (t ^ ((t >> SH) & MASK)) * K + c;                                   // + q*M1*K
    // stage-2 butterflies b = t + i*T -> (q1, p2); q1 = b / M2 = t / M2 + i*(T/M2), p2 = t % M2 for every i
    const int p2 = t % M2, q1_0 = t / M2;
    float *Bx2 = B + (q1_0 * M1) * K + c;                                                 // + i*(T/M2)*M1*K + k*M2*K + pm
    int pm[MASK + 1];
    #pragma unroll
    for (int m = 0; m <= MASK; m++) pm[m] = (p2 ^ m) * K;
    // stage-3 butterflies b = t + i*T: rows b*R3 + (k ^ (b & MASK)), b & MASK = t & MASK
    float *Bx3 = B + (t * R3) * K + c;                                                    // + i*T*R3*K + ((k ^ bm))*K
    const int bm = t & MASK;
    const float2 *tw1 = a.tw;                                                             // w_N^(q t)     = tw[q*t]
    const int tw2i = p2 * R1;                                                             // w_M1^(q p2)  = tw[q*p2*R1]
    const bool single = (a.rows_per_rank == N);
    const int h = N / 2;

    #pragma unroll 1
    for (; tile < ntiles; tile += gridDim.x) {
        const int o = tile / a.ntile_k, kz0 = (tile - o * a.ntile_k) * K;
        float2 v[E];

        XferThread xt;
        const float *kkt = a.kt.kk, *kft = a.kt.k;
        if (a.xfer.active) {
            const int iy = a.outer0 + o, iz = kz0 + c;
            kkt = a.xfer.potorder == 1 ? a.kt.kk_finite : (a.xfer.potorder == 2 ? a.kt.kk_finite2 : a.kt.kk);
            kft = a.xfer.gradorder == 0 ? a.kt.k : a.kt.k_finite;
            xt.pad = iz > h;
            const int izc = xt.pad ? 0 : iz;
            xt.kf_y = __ldg(kft + iy); xt.kf_z = __ldg(kft + izc);
            xt.sc_yz = (iy == 0 || iy == h) && (iz == 0 || iz == h);
        } else { xt.pad = true; xt.kf_y = 0; xt.kf_z = 0; xt.sc_yz = false; }
        // sum_d kk[i_d] is accumulated x, y, z in double like the reference (transfer.c:171-174): the kx (row) term
        // changes per element, the ky and kz terms are constants of this thread for the whole tile
        const double kky = a.xfer.active ? (double) __ldg(kkt + a.outer0 + o) : 0.0;
        const double kkz = a.xfer.active ? (double) __ldg(kkt + ((kz0 + c) > h ? 0 : (kz0 + c))) : 0.0;

        // ---- tile has landed: pull this thread's E elements (rows t + k*M1) out of A
        mbar_wait(&bar, phase);
        phase ^= 1;
        #pragma unroll
        for (int k = 0; k < E; k++) {
            float2 x = Ard[k * M1 * K];
            if (a.xfer.active && !xt.pad) {
                const int ix = t + k * M1;
                if (a.xfer.potorder >= 0) {
                    double sum = 0;
                    sum += (double) __ldg(kkt + ix); sum += kky; sum += kkz;
                    if (sum != 0) {
                        const double inv = __drcp_rn(sum);
                        x.x = (float) ((double) x.x * inv);
                        x.y = (float) ((double) x.y * inv);
                    } else { x.x = 0.f; x.y = 0.f; }
                }
                if (a.xfer.negate) { x.x = -x.x; x.y = -x.y; }
                for (int g = 0; g < a.xfer.ngrad; g++) {
                    const int dir = a.xfer.graddir[g];
                    if (a.xfer.zero_selfconj && xt.sc_yz && (ix == 0 || ix == h)) { x.x = 0.f; x.y = 0.f; }
                    else {
                        const double f = (double) (dir == 0 ? __ldg(kft + ix) : (dir == 1 ? xt.kf_y : xt.kf_z));
                        const float re = (float) (-((double) x.y * f));
                        const float im = (float) ((double) x.x * f);
                        x.x = re; x.y = im;
                    }
                }
                if (a.xfer.scale != 1.0) { x.x = (float) ((double) x.x * a.xfer.scale); x.y = (float) ((double) x.y * a.xfer.scale); }
            }
            if (a.conj) x.y = -x.y;
            v[k] = x;
        }

        // ---- stage 1: radix R1 over rows t + k*M1; output q goes to row q*M1 + t, times w_N^(q t)
        fft_reg<R1>(v);
        #pragma unroll
        for (int q = 1; q < R1; q++) {
            const float2 w = __ldg(tw1 + q * t);
            const float2 y = v[bitrev<R1>(q)];
            v[bitrev<R1>(q)] = make_float2(y.x * w.x - y.y * w.y, y.x * w.y + y.y * w.x);
        }

        // ---- exchange 1 (B), then stage 2
        float2 u[E];
        #pragma unroll
        for (int half = 0; half < 2; half++) {
            __syncthreads();                    // B free (and, in round 0: every thread is done reading A)
            if (half == 0 && tid == 0) {
                const int nxt = tile + gridDim.x;
                if (nxt < ntiles) {
                    const int o2 = nxt / a.ntile_k, kz2 = (nxt - o2 * a.ntile_k) * K;
                    mbar_expect_tx(&bar, tile_bytes);
                    #pragma unroll 1
                    for (int r0 = 0; r0 < N; r0 += BOX) tma_load_3d(const_cast<float2 *>(A) + (size_t) r0 * K, &tmap, &bar, 2 * kz2, r0, o2);
                }
            }
            #pragma unroll
            for (int q = 0; q < R1; q++) {
                const float2 y = v[bitrev<R1>(q)];
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
                const float2 w = __ldg(a.tw + q * tw2i);
                const float2 y = w2[bitrev<R2>(q)];
                w2[bitrev<R2>(q)] = make_float2(y.x * w.x - y.y * w.y, y.x * w.y + y.y * w.x);
            }
            #pragma unroll
            for (int q = 0; q < R2; q++) u[i * R2 + q] = w2[bitrev<R2>(q)];      // natural order: u[i*R2 + q2]
        }

        // ---- exchange 2, then stage 3 (radix R3, no twiddles) on butterflies b = t + i*T = q1*R2 + q2
        #pragma unroll
        for (int half = 0; half < 2; half++) {
            __syncthreads();
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

        // ---- store: frequency kf = q1 + R1*q2 + R1*R2*q3, K*8 B contiguous per row
        const size_t obase = (size_t) (a.dst_ooffset + o) * a.dst_ostride + kz0 + c;
        if (single) {
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
            #pragma unroll
            for (int i = 0; i < E / R3; i++) {
                const int b = t + i * T, q1 = b / R2, q2 = b - q1 * R2;
                #pragma unroll
                for (int q3 = 0; q3 < R3; q3++) {
                    const int kf = q1 + R1 * q2 + R1 * R2 * q3;
                    const int d = kf / a.rows_per_rank, kl = kf - d * a.rows_per_rank;
                    float2 y = v[i * R3 + q3];
                    if (a.conj) y.y = -y.y;
                    pick_dst(a, d)[(size_t) kl * a.dst_estride + obase] = y;
                }
            }
        }
    }
}

// ------------------------------------------------------------------ host side
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
    const size_t smem = (size_t) C::N * K * 8 + (size_t) C::N * K * 4;
    static bool attr = false;
    if (!attr) {
        FPM_CUDA_OK(cudaFuncSetAttribute(fft_tma_kernel<R1, R2, R3, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr = true;
    }
    const int ntiles = a.nouter * a.ntile_k;
    const int grid = ntiles < nsm ? ntiles : nsm;
    if (grid <= 0) return 0;
    if (fpm_prof_on) fpm_prof_begin(FPM_K_FFT_TILE, st);
    fft_tma_kernel<R1, R2, R3, K><<<grid, C::T * K, smem, st>>>(tmap, a);
    if (fpm_prof_on) fpm_prof_end(FPM_K_FFT_TILE, st);
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_fft_tma_supported(int n) { return n == 512 || n == 1024 || n == 2048 || n == 4096; }
int fpm_fft_tma_tile_k(int n) { return n == 4096 ? 4 : (n == 2048 ? 8 : 16); }

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
    a.nouter = nouter;
    a.ntile_k = (n / 2 + 1 + K - 1) / K;
    switch (n) {
        case 512: return launch_cfg<8, 8, 8, 16>(tmap, a, nsm, st);
        case 1024: return launch_cfg<16, 16, 4, 16>(tmap, a, nsm, st);
        case 2048: return launch_cfg<16, 16, 8, 8>(tmap, a, nsm, st);
        case 4096: return launch_cfg<16, 16, 16, 4>(tmap, a, nsm, st);
    }
    fpm_set_error("fpm_fft_tma_pass: unsupported N = %d", n);
    return -1;
}
