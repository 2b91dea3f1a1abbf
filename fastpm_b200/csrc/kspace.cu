// fastpm_b200 -- whole-mesh sweeps that are not folded into an FFT pass: stand-alone k-space kernels,
// CIC deconvolution, P(k) binning, power-spectrum colouring of white noise, 2LPT source terms.
// Reference: libfastpm/transfer.c:78-113 (decic), :154-220 (laplace / multiply), :189-210 (any_transfer);
// powerspectrum.c:35-124; initialcondition.c:56-64 (induce_correlation); pm2lpt.c:97-122 (2LPT source);
// powerspectrum.c:392-433 (funck_eval, log-log interpolation of the input table).
//
// k-space layout of this build: cplx[ky_local][kx][kz] with pitch_c complex per row (see common.cuh).
#include "common.cuh"
#include "mesh.cuh"
#include "ic_gadget.h"
#include <vector>
#include <string.h>
#include <stdlib.h>

// ------------------------------------------------------------------ generic mode iterator
// one thread per complex element of the local k-space block, kz fastest (coalesced)
struct ModeIdx { int ix, iy, iz; size_t off; bool valid; };

__device__ __forceinline__ ModeIdx mode_from_linear(const FpmGeom &g, size_t t)
{
    ModeIdx m;
    const int hc = g.n / 2 + 1;
    const int pc = g.pitch_c;
    m.iz = (int) (t % pc);
    const size_t row = t / pc;
    m.ix = (int) (row % g.n);
    const int iyl = (int) (row / g.n);
    m.iy = iyl + g.y0;
    m.off = t;
    m.valid = (m.iz < hc) && (iyl < g.nyl);
    return m;
}

__global__ void __launch_bounds__(256) transfer_kernel(const FpmGeom g, const FpmKTables kt, const FpmTransferSpec s,
        const float2 *__restrict__ from, float2 *__restrict__ to, size_t total)
{
    size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        ModeIdx m = mode_from_linear(g, t);
        if (!m.valid) continue;
        to[m.off] = fpm_apply_transfer(s, kt, from[m.off], m.ix, m.iy, m.iz);
    }
}

// pgdcorrection.c:28-59 apply_pgdpot_transfer: to = (float)(alpha * exp(-kl^2/kk - kk^2/ks^4) / kk * from), kk = k_x^2 + k_y^2 + k_z^2
// summed in double from the float tables in that order; 0 where kk == 0.
__global__ void __launch_bounds__(256) pgd_transfer_kernel(const FpmGeom g, const FpmKTables kt, const double alpha, const double kl2,
        const double ks4, const float2 *__restrict__ from, float2 *__restrict__ to, size_t total)
{
    size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        ModeIdx m = mode_from_linear(g, t);
        if (!m.valid) continue;
        double kk = 0;
        kk += (double) kt.kk[m.ix]; kk += (double) kt.kk[m.iy]; kk += (double) kt.kk[m.iz];
        float2 v = from[m.off];
        if (kk > 0) {
            const double fac = alpha * exp(-kl2 / kk - kk * kk / ks4) / kk;
            v.x = (float) (fac * (double) v.x);
            v.y = (float) (fac * (double) v.y);
        } else {
            v.x = 0.f; v.y = 0.f;
        }
        to[m.off] = v;
    }
}

// fastpm_ic_remove_variance, initialcondition.c:66-99: every mode keeps its phase and gets unit amplitude,
// (cos, sin)(atan2(im, re)) in double; (0, 0) stays (0, 0).
__global__ void __launch_bounds__(256) remove_variance_kernel(const FpmGeom g, float2 *__restrict__ dk, size_t total)
{
    size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        ModeIdx m = mode_from_linear(g, t);
        if (!m.valid) continue;
        const float2 v = dk[m.off];
        const double a = v.x, b = v.y;
        float2 o = make_float2(0.f, 0.f);
        if (!(a == 0 && b == 0)) {
            const double phase = atan2(b, a);
            o.x = (float) cos(phase);
            o.y = (float) sin(phase);
        }
        dk[m.off] = o;
    }
}

// Radial force softening (gravity.c:244-270): mode 0 = sharp low pass, factor 1 where kk < param (= kth^2) else 0
// (fastpm_apply_lowpass_transfer, transfer.c:43-66); mode 1 = exp(-36 (k / k_nyquist)^36), param = k_nyquist (gaussian36,
// gravity.c:104-109 through fastpm_apply_any_transfer, transfer.c:189-211).  kk summed in double from the float tables.
__global__ void __launch_bounds__(256) radial_transfer_kernel(const FpmGeom g, const FpmKTables kt, const int mode, const double param,
        const float2 *__restrict__ from, float2 *__restrict__ to, size_t total)
{
    size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        ModeIdx m = mode_from_linear(g, t);
        if (!m.valid) continue;
        double kk = 0;
        kk += (double) kt.kk[m.ix]; kk += (double) kt.kk[m.iy]; kk += (double) kt.kk[m.iz];
        double smth;
        if (mode == 0) {
            smth = kk < param ? 1 : 0;
        } else {
            const double x = sqrt(kk) / param;
            smth = exp(-36 * pow(x, 36.0));
        }
        float2 v = from[m.off];
        v.x = (float) ((double) v.x * smth);
        v.y = (float) ((double) v.y * smth);
        to[m.off] = v;
    }
}

// transfer.c:78-113: kernel[d][i] = 1/sinc^2(k h/2) in double (table prepared on the host), product in
// double, one rounding to float per component.
__global__ void __launch_bounds__(256) decic_kernel(const FpmGeom g, const double *__restrict__ dtab,
        const float2 *__restrict__ from, float2 *__restrict__ to, size_t total)
{
    size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        ModeIdx m = mode_from_linear(g, t);
        if (!m.valid) continue;
        double smth = 1.0;
        smth *= dtab[m.ix]; smth *= dtab[m.iy]; smth *= dtab[m.iz];
        float2 v = from[m.off];
        v.x = (float) ((double) v.x * smth);
        v.y = (float) ((double) v.y * smth);
        to[m.off] = v;
    }
}

// powerspectrum.c:35-124.  bins = n/2 integer shells of |i|; per bin: sum w, sum w*|delta|^2, sum w*|k|.
// One warp per k-space row (ky, kx fixed; kz = 0..N/2 contiguous), two adjacent modes per lane: along a row the shell
// index is non-decreasing in kz, so the 64 modes of a warp-wide load fall into a few CONTIGUOUS runs of equal bin.
// Each run is summed with a segmented shuffle reduction and only the head lane of a run touches the CTA's
// shared-memory histogram (double), which is flushed with one global atomic per non-empty bin.
// sum w and sum w*|k| depend on the mesh geometry only: GEOM = true accumulates those two (once per mesh, cached),
// GEOM = false accumulates the data-dependent sum w*|delta|^2 and the all-mode variance.  `decic` folds the
// deconvolution (rounded to float like the reference's in-place sweep, transfer.c:78-113) into the read.
#define PK_WARPS 8
#define PK_UNROLL 2
// Each warp owns a private histogram in shared memory: the head lanes of one warp-wide step hold DISTINCT bins (runs of a
// monotone sequence), so they update it with plain read-modify-writes -- no atomics, no contention between warps (double
// atomicAdd on shared memory is a CAS loop).  Measured: no faster than the shared atomics it replaced -- the kernel is bound
// by instruction issue (fp64, shuffles), see DESIGN.md section 9.
template <bool GEOM>
__global__ void __launch_bounds__(32 * PK_WARPS) powerspectrum_kernel(const FpmGeom g, const double *__restrict__ dtab, int decic,
        const float2 *__restrict__ dk, double k0, double *__restrict__ out /* GEOM: [2][nbins]; else [nbins] + 1 */,
        const float2 *__restrict__ dk2 = nullptr /* cross spectrum: Re(dk conj(dk2)), powerspectrum.c:87-91 */)
{
    FPM_DYN_SMEM(hist_raw, 8);
    double *hist_all = reinterpret_cast<double *>(hist_raw);
    const int nbins = g.n / 2;
    const int nslots = GEOM ? 2 * nbins : nbins + 1;
    const int nwarps = blockDim.x >> 5;                 // <= PK_WARPS: fewer when the per-warp histograms of a large mesh would not fit
    for (int i = threadIdx.x; i < nslots * nwarps; i += blockDim.x) hist_all[i] = 0;
    __syncthreads();
    volatile double *hist = hist_all + (size_t) (threadIdx.x >> 5) * nslots;
    const int n = g.n, h = n / 2, lane = threadIdx.x & 31;
    const size_t nrows = (size_t) g.nyl * n;
    const size_t wstride = (size_t) gridDim.x * nwarps;
    const int nchunk = (h + 1 + 63) / 64;
    double allsum = 0;
    for (size_t row = (size_t) blockIdx.x * nwarps + (threadIdx.x >> 5); row < nrows; row += wstride) {
        const int ix = (int) (row % n), iy = (int) (row / n) + g.y0;
        const int ikx = ix > h ? ix - n : ix, iky = iy > h ? iy - n : iy;
        const int kxy = ikx * ikx + iky * iky;
        const double dxy = (!GEOM && decic) ? dtab[ix] * dtab[iy] : 1.0;      // (1 * d[ix]) * d[iy], the reference's order
        const float4 *src = reinterpret_cast<const float4 *>(dk + row * (size_t) g.pitch_c);
        const float4 *src2 = reinterpret_cast<const float4 *>(dk2 + row * (size_t) g.pitch_c);
        for (int c0 = 0; c0 < nchunk; c0 += PK_UNROLL) {
            float4 vv[PK_UNROLL], vv2[PK_UNROLL];
            if (!GEOM) {
                #pragma unroll
                for (int u = 0; u < PK_UNROLL; u++) {
                    const int iz = (c0 + u) * 64 + 2 * lane;          // pitch_c is a multiple of 16: iz + 1 stays inside the row
                    vv[u] = (c0 + u < nchunk && iz <= h) ? __ldg(src + (iz >> 1)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    vv2[u] = (dk2 != nullptr && c0 + u < nchunk && iz <= h) ? __ldg(src2 + (iz >> 1)) : vv[u];
                }
            }
            #pragma unroll
            for (int u = 0; u < PK_UNROLL; u++) {
                if (c0 + u >= nchunk) break;                         // warp-uniform
                int bins[2];
                double acc[2], acc2[2];
                #pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int iz = (c0 + u) * 64 + 2 * lane + e;
                    const bool inrow = iz <= h;
                    const double w = (iz == 0 || iz == h) ? 1.0 : 2.0;
                    const int kk = kxy + iz * iz;                    // iz <= h: never wrapped
                    const float sf = sqrtf((float) kk);              // kk < 2^24: exact in float; the estimate is off by at most 1
                    int bin = (int) sf;
                    if ((bin + 1) * (bin + 1) <= kk) bin++;
                    if (bin * bin > kk) bin--;
                    if (!inrow || bin > nbins) bin = nbins;          // dropped: beyond the last shell or row padding
                    const bool counted = inrow && bin < nbins && kk != 0;
                    acc[e] = 0; acc2[e] = 0;
                    if (GEOM) {
                        if (counted) { acc[e] = w; acc2[e] = w * (sqrt((double) kk) * k0); }
                    } else {
                        float2 v = e ? make_float2(vv[u].z, vv[u].w) : make_float2(vv[u].x, vv[u].y);
                        if (decic && inrow) {                        // the field IS the deconvolved one: also for the variance
                            const double smth = dxy * dtab[iz];
                            v.x = (float) ((double) v.x * smth);
                            v.y = (float) ((double) v.y * smth);
                        }
                        // the second field of a cross spectrum (never with a pending deconvolution: the launcher sees to that)
                        const float2 v2 = dk2 == nullptr ? v : (e ? make_float2(vv2[u].z, vv2[u].w) : make_float2(vv2[u].x, vv2[u].y));
                        const double p2 = w * ((double) v.x * (double) v2.x + (double) v.y * (double) v2.y);
                        if (inrow) allsum += p2;
                        if (counted) acc[e] = p2;
                    }
                    bins[e] = bin;
                }
                // the lane's two modes: merge when they share a bin, else the first one goes out on its own.  bins[0] of the
                // lanes that do so are strictly increasing across the warp: distinct addresses, plain update.
                int bin = bins[1];
                double s1 = acc[1], s2 = acc2[1];
                if (bins[0] == bin) { s1 += acc[0]; s2 += acc2[0]; }
                else if (bins[0] < nbins && (acc[0] != 0 || acc2[0] != 0)) {
                    hist[bins[0]] += acc[0];
                    if (GEOM) hist[nbins + bins[0]] += acc2[0];
                }
                __syncwarp();
                // segmented reduction over runs of equal bin (contiguous because bin is monotone in iz)
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int ob = __shfl_down_sync(0xffffffffu, bin, o);
                    const double o1 = __shfl_down_sync(0xffffffffu, s1, o);
                    double o2 = 0;
                    if (GEOM) o2 = __shfl_down_sync(0xffffffffu, s2, o);
                    if (lane + o < 32 && ob == bin) { s1 += o1; s2 += o2; }
                }
                // the head lane of every run (distinct bins) adds the run's total
                const int pb = __shfl_up_sync(0xffffffffu, bin, 1);
                if ((lane == 0 || pb != bin) && bin < nbins && (s1 != 0 || s2 != 0)) {
                    hist[bin] += s1;
                    if (GEOM) hist[nbins + bin] += s2;
                }
                __syncwarp();
            }
        }
    }
    if (!GEOM) {
        for (int o = 16; o > 0; o >>= 1) allsum += __shfl_xor_sync(0xffffffffu, allsum, o);
        if (lane == 0) hist[nbins] += allsum;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nslots; i += blockDim.x) {
        double t = 0;
        for (int w = 0; w < nwarps; w++) t += hist_all[(size_t) w * nslots + i];
        if (t != 0) atomicAdd(&out[i], t);
    }
}

// ------------------------------------------------------------------ P(k), row-streaming version (round 2)
// The kernel above spends ~110 instructions per mode (segmented shuffle reductions on doubles every 64 modes) and runs at a
// quarter of the HBM rate.  Here a warp takes one k-space row (ky, kx fixed) at a time and every lane owns a CONTIGUOUS chunk
// of CH = (N/2)/32 modes of it: along kz the shell index is non-decreasing, so a lane walks runs of equal bin, keeps the
// running sum of |delta|^2 in a register and touches the histogram only when the bin changes.
//   * rows travel global -> shared memory with 16-byte cp.async copies (coalesced, no registers); mode m sits at position
//     m + 2*min(m / CH, 31), so that the lanes' 16-byte reads are bank-conflict free; many warps per SM hide the latency;
//   * the bin is tracked incrementally: kk grows by 2 iz + 1 per step, which is at most (bin+1)^2 - bin^2 because bin >= iz,
//     so the bin advances by at most one per step -- one square root per chunk, no loop;
//   * one histogram per CTA in shared memory, updated with atomicAdd(double) when a run ends (a compare-and-swap loop, measured
//     at a fraction of a cycle per lane on B200 -- scripts/ubench/atomics.cu -- and needed only every few modes);
//   * interior modes have weight 2, the two ends of a row weight 1 (powerspectrum.c:94): runs accumulate |delta|^2 (ends: half of
//     it) and are doubled when flushed -- scaling by 2 is exact, so this is the sum of w |delta|^2 in another order.
// The deconvolution (solver.c:471, transfer.c:78-113) is folded into the read exactly as above: the mode is multiplied by
// (d[ix] d[iy]) d[iz] in double and ROUNDED TO FLOAT before it is squared.  Needs (N/2) % 64 == 0; other sizes keep the kernel above.
#define PKR_WARPS 8
#ifndef FPM_EMULATE
__device__ __forceinline__ void pk_cp16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t) __cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void pk_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void pk_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#else
inline void pk_cp16(void *smem_dst, const void *gsrc) { memcpy(smem_dst, gsrc, 16); }
inline void pk_cp_commit() {}
inline void pk_cp_wait_all() {}
#endif

__device__ __forceinline__ double pk_round_to_float(double t)
{
    unsigned long long u;
    memcpy(&u, &t, 8);
    u += 0x0FFFFFFFull + ((u >> 29) & 1ull);
    u &= ~0x1FFFFFFFull;
    memcpy(&t, &u, 8);
    return t;
}

__global__ void __launch_bounds__(32 * PKR_WARPS) powerspectrum_rows_kernel(const FpmGeom g, const double *__restrict__ dtab, int decic,
        const float2 *__restrict__ dk, double *__restrict__ out /* [nbins] + 1 */)
{
    FPM_DYN_SMEM(smem_raw, 16);
    const int n = g.n, h = n / 2, nbins = h, CH = h / 32;
    const int RB = h + 66;                                       // positions of one staged row (h + 1 modes, 2 pad per chunk), even
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *dts = reinterpret_cast<double *>(smem_raw);                      // [RB] deconvolution factors along z, staged like a row
    double *hist = dts + RB;                                                  // [nbins + 1], one per CTA
    float2 *rowbuf = reinterpret_cast<float2 *>(hist + nbins + 2) + (size_t) warp * RB;      // [PKR_WARPS][RB]
    const unsigned chinv = (65536u + CH - 1) / CH;               // m / CH == (m * chinv) >> 16 for m < 8192
    for (int i = threadIdx.x; i <= nbins; i += blockDim.x) hist[i] = 0;
    for (int m = threadIdx.x; m <= h; m += blockDim.x) {
        const int c = (int) ((m * chinv) >> 16);
        dts[m + 2 * (c > 31 ? 31 : c)] = decic ? dtab[m] : 1.0;           // iz = h follows lane 31's chunk
    }
    __syncthreads();
    const int nrows_i = g.nyl * n;                               // < 2^31 for every supported mesh
    const int wstride = (int) gridDim.x * PKR_WARPS;
    const int nq = (h + 2) / 2;                                  // float4 (two modes) per row, including the pair that holds iz = h
    int row = (int) blockIdx.x * PKR_WARPS + warp;
    int ix = row % n, iyl = row / n;
    const int dix = wstride % n, diy = wstride / n;
    double allsum = 0;
    const int iz0 = lane * CH;
    const float2 *src = rowbuf + iz0 + 2 * lane;
    const double *dsrc = dts + iz0 + 2 * lane;
    for (; row < nrows_i; row += wstride) {
        {
            const float4 *gsrc = reinterpret_cast<const float4 *>(dk + (size_t) row * (size_t) g.pitch_c);
            for (int q = lane; q < nq; q += 32) {
                const int m = 2 * q;                             // CH is even: both modes of the pair fall into the same chunk
                const int c = (int) ((m * chinv) >> 16);
                pk_cp16(rowbuf + m + 2 * (c > 31 ? 31 : c), gsrc + q);
            }
            pk_cp_commit();
        }
        const int iy = iyl + g.y0;
        const int ikx = ix > h ? ix - n : ix, iky = iy > h ? iy - n : iy;
        const int kxy = ikx * ikx + iky * iky;
        const double dxy = decic ? dtab[ix] * dtab[iy] : 1.0;           // (1 * d[ix]) * d[iy], the reference's order
        int iz = iz0;
        int kk = kxy + iz * iz;
        int bin;
        {
            const float sf = sqrtf((float) kk);                          // kk < 2^24: exact in float; the estimate is off by at most 1
            bin = (int) sf;
            if ((bin + 1) * (bin + 1) <= kk) bin++;
            if (bin * bin > kk) bin--;
        }
        int next = (bin + 1) * (bin + 1);
        int run_bin = bin;
        double run = 0;
        pk_cp_wait_all();
        __syncwarp();
        // one mode: bin bookkeeping, deconvolution, |delta|^2 (weight applied by the caller)
        auto mode_p2 = [&](float2 v, double dz) -> double {
            const bool up = kk >= next;                                  // at most one shell further than the previous mode
            bin += up ? 1 : 0;
            next += up ? 2 * bin + 1 : 0;
            if (bin != run_bin) {
                if (run_bin < nbins && run != 0) atomicAdd(hist + run_bin, 2.0 * run);
                run = 0; run_bin = bin;
            }
            double ax = (double) v.x, ay = (double) v.y;
            if (decic) {
                // (double) (float) (ax * smth) without the two conversions (the conversion pipe was the busiest unit of this kernel):
                // round the product to a 24-bit significand, nearest-even, on the bit pattern -- identical for every value in the
                // normal float range (|x| between 1.2e-38 and 3.4e38; density modes never leave it)
                const double smth = dxy * dz;
                ax = pk_round_to_float(ax * smth);
                ay = pk_round_to_float(ay * smth);
            }
            kk += 2 * iz + 1;
            iz++;
            return ax * ax + ay * ay;
        };
        {
            // first pair of the chunk: holds iz = 0 (weight 1; the DC mode belongs to no shell) for lane 0
            const float4 vv = *reinterpret_cast<const float4 *>(src);
            const double2 dd = *reinterpret_cast<const double2 *>(dsrc);
            const bool first0 = iz == 0, dc = first0 && kxy == 0;
            double p2 = mode_p2(make_float2(vv.x, vv.y), dd.x);
            if (first0) p2 *= 0.5;
            allsum += p2;
            if (!dc) run += p2;
            p2 = mode_p2(make_float2(vv.z, vv.w), dd.y);
            allsum += p2; run += p2;
        }
        for (int j = 2; j < CH; j += 2) {                                // two modes per 16-byte read (positions are even)
            const float4 vv = *reinterpret_cast<const float4 *>(src + j);
            const double2 dd = *reinterpret_cast<const double2 *>(dsrc + j);
            double p2 = mode_p2(make_float2(vv.x, vv.y), dd.x);
            allsum += p2; run += p2;
            p2 = mode_p2(make_float2(vv.z, vv.w), dd.y);
            allsum += p2; run += p2;
        }
        if (lane == 31) {                                                // iz = h, weight 1
            const double p2 = 0.5 * mode_p2(src[CH], dsrc[CH]);
            allsum += p2; run += p2;
        }
        if (run_bin < nbins && run != 0) atomicAdd(hist + run_bin, 2.0 * run);
        __syncwarp();                                                    // every lane is done with the row buffer
        ix += dix; iyl += diy;
        if (ix >= n) { ix -= n; iyl++; }
    }
    for (int o = 16; o > 0; o >>= 1) allsum += __shfl_xor_sync(0xffffffffu, allsum, o);
    if (lane == 0) atomicAdd(hist + nbins, 2.0 * allsum);
    __syncthreads();
    for (int i = threadIdx.x; i <= nbins; i += blockDim.x)
        if (hist[i] != 0) atomicAdd(&out[i], hist[i]);
}

// out[3*nbins + 1] = geometry sums (cached) and data sums laid out as the callers expect: [sum w][sum w |d|^2][sum w k][variance]
__global__ void pk_assemble_kernel(const double *__restrict__ geom, const double *__restrict__ data, int nbins, double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nbins) { out[i] = geom[i]; out[nbins + i] = data[i]; out[2 * nbins + i] = geom[nbins + i]; }
    if (i == 0) out[3 * nbins] = data[nbins];
}

// buf[i] = (float)(buf[i] * value): fastpm_apply_multiply_transfer, transfer.c:213-220
__global__ void __launch_bounds__(256) scale_kernel(const float *__restrict__ from, float *__restrict__ to, size_t nfloats, double value)
{
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; i < nfloats; i += stride) to[i] = (float) ((double) from[i] * value);
}

// buf[i] = (float)(buf[i] / value): the unit conversion reverted in fastpm_unset_species_snapshot, solver.c:738-742
__global__ void __launch_bounds__(256) divide_kernel(const float *__restrict__ from, float *__restrict__ to, size_t nfloats, double value)
{
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; i < nfloats; i += stride) to[i] = (float) ((double) from[i] / value);
}

// pm2lpt.c:103-105 / :118-120:  source += a*b   or   source -= a*a   (float arithmetic, like the reference)
__global__ void __launch_bounds__(256) muladd_kernel(float *__restrict__ source, const float *__restrict__ a, const float *__restrict__ b,
        size_t nfloats, float sign)
{
    size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; i < nfloats; i += stride) {
        const float prod = __fmul_rn(a[i], b[i]);
        source[i] = sign > 0 ? __fadd_rn(source[i], prod) : __fsub_rn(source[i], prod);
    }
}

// fastpm_funck_eval (powerspectrum.c:392-433): piecewise log-log (or linear where a value is <= 0) interpolation
__device__ double funck_eval(const double *__restrict__ tk, const double *__restrict__ tf, int size, double k)
{
    if (k == 0) return 1;
    int l = 0, r = size - 1;
    while (r - l > 1) {
        const int mid = (r + l) / 2;
        if (k < tk[mid]) r = mid; else l = mid;
    }
    const double k2 = tk[r], k1 = tk[l], f2 = tf[r], f1 = tf[l];
    if (l == r) return tf[l];
    if (f1 <= 0 || f2 <= 0 || k1 == 0 || k2 == 0) {
        double f = (k - k1) * f2 + (k2 - k) * f1;
        return f / (k2 - k1);
    }
    const double lk = log(k), lf1 = log(f1), lf2 = log(f2), lk1 = log(k1), lk2 = log(k2);
    double f = (lk - lk1) * lf2 + (lk2 - lk) * lf1;
    f /= (lk2 - lk1);
    return exp(f);
}

// fastpm_ic_induce_correlation: delta_k *= sqrt(P(|k|)/V)   with |k| = sqrt(sum of float kk[d])
__global__ void __launch_bounds__(256) induce_kernel(const FpmGeom g, const FpmKTables kt, float2 *__restrict__ dk, size_t total,
        const double *__restrict__ tk, const double *__restrict__ tp, int size, double volume)
{
    size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        ModeIdx m = mode_from_linear(g, t);
        if (!m.valid) continue;
        double kk = 0;
        kk += (double) kt.kk[m.ix]; kk += (double) kt.kk[m.iy]; kk += (double) kt.kk[m.iz];
        const double k = sqrt(kk);
        double f = sqrt(funck_eval(tk, tp, size, k));
        f *= sqrt(1.0 / volume);
        float2 v = dk[m.off];
        v.x = (float) ((double) v.x * f);
        v.y = (float) ((double) v.y * f);
        dk[m.off] = v;
    }
}

// ------------------------------------------------------------------ synthetic white noise (bench ICs)
// Counter-based generator (Philox-4x32-10) so that a 1024^3 field needs no serial RANLUX stream; the
// field is Hermitian by construction because it is the r2c transform of a REAL white-noise field:
// this kernel fills the real mesh with unit-variance Gaussians (Box-Muller).  Declared deviation from
// initialcondition.c:145-273, used only for benchmark-size synthetic ICs; parity tests feed both sides
// the oracle's Gadget-scheme delta_k instead.
__device__ __forceinline__ void philox_round(unsigned &c0, unsigned &c1, unsigned &c2, unsigned &c3, unsigned k0, unsigned k1)
{
    const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
    const unsigned h0 = (unsigned) (p0 >> 32), l0 = (unsigned) p0, h1 = (unsigned) (p1 >> 32), l1 = (unsigned) p1;
    c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
}
__global__ void __launch_bounds__(256) whitenoise_kernel(const FpmGeom g, float *__restrict__ real, unsigned long long seed)
{
    const size_t ncell = (size_t) g.nxl * g.n * g.n;
    size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (; 2 * t < ncell; t += stride) {
        const size_t cell = 2 * t;                   // two cells (adjacent in z) per counter
        const int z = (int) (cell % g.n);
        const size_t row = cell / g.n;               // xl*n + y
        const size_t gcell = ((size_t) g.x0 * g.n * g.n + cell) / 2;
        unsigned c0 = (unsigned) gcell, c1 = (unsigned) (gcell >> 32), c2 = 0x5eedu, c3 = 0;
        unsigned k0 = (unsigned) seed, k1 = (unsigned) (seed >> 32);
        #pragma unroll
        for (int r = 0; r < 10; r++) { philox_round(c0, c1, c2, c3, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
        const float u1 = ((float) c0 + 1.0f) * 2.3283064365386963e-10f;     // (0, 1]
        const float u2 = (float) c1 * 2.3283064365386963e-10f;
        const float rad = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        float *p = real + row * g.pitch_r + z;
        p[0] = rad * cs; p[1] = rad * sn;
    }
}

// set one mode (and nothing else): used for delta_k(0,0,0) = 1 (src/fastpm.c:541-544)
__global__ void set_mode_kernel(const FpmGeom g, float2 *dk, int ix, int iy, int iz, float re, float im)
{
    const int iyl = iy - g.y0;
    if (iyl < 0 || iyl >= g.nyl) return;
    dk[((size_t) iyl * g.n + ix) * g.pitch_c + iz] = make_float2(re, im);
}

// ------------------------------------------------------------------ Gadget-scheme Gaussian field (ic_gadget.h)
// One thread per (kx, ky_local) column of this rank's k-space slab; every thread runs its two RANLUX generators (local
// memory) down kz.  The rows are written one float2 at a time (uncoalesced): an initial-condition step, run once.
__global__ void __launch_bounds__(64) gadget_fill_kernel(const FpmGeom g, const unsigned int *__restrict__ self,
        const unsigned int *__restrict__ conj, float2 *__restrict__ dk)
{
    const size_t idx = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t) g.nyl * g.n) return;
    const int i = (int) (idx % g.n), jl = (int) (idx / g.n), j = jl + g.y0;
    const size_t q = (size_t) i * g.n + j;
    fpm_gadget_fill_column(g.n, i, j, self[q], conj[q], dk + ((size_t) jl * g.n + i) * g.pitch_c);
}

#ifndef FPM_EMULATE          // host side: not part of the CPU emulation of the kernels (tests/emul/kspace_emul.cpp)
// ------------------------------------------------------------------ launchers
static inline unsigned sweep_grid(size_t n)
{
    size_t b = (n + 255) / 256;
    const size_t cap = 148 * 16;
    return (unsigned) (b < cap ? (b > 0 ? b : 1) : cap);
}
static inline size_t cplx_total(const FpmGeom &g) { return (size_t) g.nyl * g.n * g.pitch_c; }

int fpm_transfer_launch(const FpmMesh *m, const float *from, float *to, const FpmTransferSpec *s, cudaStream_t st)
{
    const size_t total = cplx_total(m->geom);
    FPM_TIMED(FPM_K_KSPACE, st, (transfer_kernel<<<sweep_grid(total), 256, 0, st>>>(m->geom, m->ktab, *s, (const float2 *) from, (float2 *) to, total)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_pgd_transfer_launch(const FpmMesh *m, const float *from, float *to, double alpha, double kl, double ks, cudaStream_t st)
{
    const size_t total = cplx_total(m->geom);
    const double kl2 = kl * kl, ks4 = ks * ks * ks * ks;              // pgdcorrection.c:35-36
    FPM_TIMED(FPM_K_KSPACE, st, (pgd_transfer_kernel<<<sweep_grid(total), 256, 0, st>>>(m->geom, m->ktab, alpha, kl2, ks4, (const float2 *) from, (float2 *) to, total)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_remove_variance_launch(const FpmMesh *m, float *dk, cudaStream_t st)
{
    const size_t total = cplx_total(m->geom);
    FPM_TIMED(FPM_K_KSPACE, st, (remove_variance_kernel<<<sweep_grid(total), 256, 0, st>>>(m->geom, (float2 *) dk, total)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_radial_transfer_launch(const FpmMesh *m, const float *from, float *to, int mode, double param, cudaStream_t st)
{
    const size_t total = cplx_total(m->geom);
    FPM_TIMED(FPM_K_KSPACE, st, (radial_transfer_kernel<<<sweep_grid(total), 256, 0, st>>>(m->geom, m->ktab, mode, param, (const float2 *) from, (float2 *) to, total)));
    FPM_CHECK_LAUNCH();
    return 0;
}

// the product of one double factor per axis (same kernel as the CIC deconvolution): Gaussian softening and smoothing
int fpm_axis_factors_launch(const FpmMesh *m, const double *d_table, const float *from, float *to, cudaStream_t st)
{
    const size_t total = cplx_total(m->geom);
    FPM_TIMED(FPM_K_KSPACE, st, (decic_kernel<<<sweep_grid(total), 256, 0, st>>>(m->geom, d_table, (const float2 *) from, (float2 *) to, total)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_decic_launch(const FpmMesh *m, const float *from, float *to, cudaStream_t st)
{
    const size_t total = cplx_total(m->geom);
    FPM_TIMED(FPM_K_KSPACE, st, (decic_kernel<<<sweep_grid(total), 256, 0, st>>>(m->geom, m->d_decic, (const float2 *) from, (float2 *) to, total)));
    FPM_CHECK_LAUNCH();
    return 0;
}

// d_out: [3][n/2] = Nmodes, sum w*|d|^2, sum w*k (not yet normalised; multi-GPU callers all-reduce first), then 1 slot:
// the sum of w*|d|^2 over every mode including DC and the corners beyond the last shell
int fpm_powerspectrum_launch(const FpmMesh *m, const float *dk, int decic, double *d_out, cudaStream_t st, const float *dk2)
{
    const FpmGeom &g = m->geom;
    const int nbins = g.n / 2;
    // per-warp histograms: as many warps per CTA as fit (8 up to N = 2048; 4 for the geometry sums of a 4096^3 mesh)
    int w_geom = PK_WARPS, w_data = PK_WARPS;
    while (w_geom > 1 && sizeof(double) * (size_t) (2 * nbins) * w_geom > 227 * 1024) w_geom >>= 1;
    while (w_data > 1 && sizeof(double) * (size_t) (nbins + 1) * w_data > 227 * 1024) w_data >>= 1;
    const size_t smem_geom = sizeof(double) * (size_t) (2 * nbins) * w_geom, smem_data = sizeof(double) * (size_t) (nbins + 1) * w_data;
    static bool attr_done = false;
    if (!attr_done) {
        FPM_CUDA_OK(cudaFuncSetAttribute(powerspectrum_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        FPM_CUDA_OK(cudaFuncSetAttribute(powerspectrum_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = true;
    }
    if (smem_geom > 227 * 1024 || smem_data > 227 * 1024) { fpm_set_error("powerspectrum: too many bins"); return -1; }
    const int ctas_per_sm = smem_data <= 56 * 1024 ? 4 : (smem_data <= 75 * 1024 ? 3 : (smem_data <= 113 * 1024 ? 2 : 1));
    const double k0 = 2 * M_PI / g.boxsize;
    FpmMesh *mm = const_cast<FpmMesh *>(m);           // the per-mesh cache of the geometry sums
    if (!mm->d_pkgeom) {
        FPM_CUDA_OK(cudaMalloc(&mm->d_pkgeom, sizeof(double) * (3 * nbins + 1)));
        FPM_CUDA_OK(cudaMemsetAsync(mm->d_pkgeom, 0, sizeof(double) * (3 * nbins + 1), st));
        FPM_TIMED(FPM_K_PK, st, (powerspectrum_kernel<true><<<148, 32 * w_geom, smem_geom, st>>>(g, m->d_decic, 0, nullptr, k0, mm->d_pkgeom)));
        FPM_CHECK_LAUNCH();
    }
    double *d_data = mm->d_pkgeom + 2 * nbins;        // [nbins] + 1
    FPM_CUDA_OK(cudaMemsetAsync(d_data, 0, sizeof(double) * (nbins + 1), st));
    static int rows_mode = -1;        // FASTPM_B200_PK=generic: the shuffle-reduction kernel for every mesh size (cross-check)
    if (rows_mode < 0) { const char *e = getenv("FASTPM_B200_PK"); rows_mode = (e && !strcmp(e, "generic")) ? 0 : 1; }
    const int h = g.n / 2;
    const size_t smem_rows = sizeof(double) * (size_t) (h + 66) + sizeof(double) * (size_t) (nbins + 2) + sizeof(float2) * (size_t) PKR_WARPS * (h + 66);
    if (dk2 == dk) dk2 = nullptr;
    if (dk2 != nullptr && decic) { fpm_set_error("cross power spectrum with a pending deconvolution"); return -1; }
    if (dk2 == nullptr && rows_mode && h % 64 == 0 && h <= 4096 && smem_rows <= 227 * 1024 && (size_t) g.nyl * g.n < ((size_t) 1 << 31)) {
        static bool attr_rows = false;
        if (!attr_rows) { FPM_CUDA_OK(cudaFuncSetAttribute(powerspectrum_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr_rows = true; }
        int per_sm = (int) ((227 * 1024) / (smem_rows + 1024));
        if (per_sm > 6) per_sm = 6;                       // 48 warps per SM
        if (per_sm < 1) per_sm = 1;
        const size_t nrows = (size_t) g.nyl * g.n;
        size_t grid = (size_t) 148 * per_sm;
        if (grid * PKR_WARPS > nrows) grid = (nrows + PKR_WARPS - 1) / PKR_WARPS;
        fpm_path_counter[FPM_PATH_PK_ROWS]++;
        FPM_TIMED(FPM_K_PK, st, (powerspectrum_rows_kernel<<<(unsigned) grid, 32 * PKR_WARPS, smem_rows, st>>>(g, m->d_decic, decic, (const float2 *) dk, d_data)));
    } else {
        FPM_TIMED(FPM_K_PK, st, (powerspectrum_kernel<false><<<148 * ctas_per_sm, 32 * w_data, smem_data, st>>>(g, m->d_decic, decic, (const float2 *) dk, k0, d_data, (const float2 *) dk2)));
    }
    FPM_CHECK_LAUNCH();
    pk_assemble_kernel<<<(nbins + 255) / 256, 256, 0, st>>>(mm->d_pkgeom, d_data, nbins, d_out);
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_scale_launch(const float *from, float *to, size_t nfloats, double value, cudaStream_t st)
{
    FPM_TIMED(FPM_K_KSPACE, st, (scale_kernel<<<sweep_grid(nfloats), 256, 0, st>>>(from, to, nfloats, value)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_divide_launch(const float *from, float *to, size_t nfloats, double value, cudaStream_t st)
{
    FPM_TIMED(FPM_K_KSPACE, st, (divide_kernel<<<sweep_grid(nfloats), 256, 0, st>>>(from, to, nfloats, value)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_muladd_launch(float *source, const float *a, const float *b, size_t nfloats, int sign, cudaStream_t st)
{
    FPM_TIMED(FPM_K_KSPACE, st, (muladd_kernel<<<sweep_grid(nfloats), 256, 0, st>>>(source, a, b, nfloats, sign > 0 ? 1.f : -1.f)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_induce_launch(const FpmMesh *m, float *dk, const double *d_tk, const double *d_tp, int size, cudaStream_t st)
{
    const FpmGeom &g = m->geom;
    const size_t total = cplx_total(g);
    const double volume = g.boxsize * g.boxsize * g.boxsize;
    FPM_TIMED(FPM_K_KSPACE, st, (induce_kernel<<<sweep_grid(total), 256, 0, st>>>(g, m->ktab, (float2 *) dk, total, d_tk, d_tp, size, volume)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_whitenoise_launch(const FpmMesh *m, float *real, unsigned long long seed, cudaStream_t st)
{
    const FpmGeom &g = m->geom;
    const size_t ncell = (size_t) g.nxl * g.n * g.n;
    FPM_TIMED(FPM_K_OTHER, st, (whitenoise_kernel<<<sweep_grid(ncell / 2), 256, 0, st>>>(g, real, seed)));
    FPM_CHECK_LAUNCH();
    return 0;
}

// pmic_fill_gaussian_gadget (initialcondition.c:145-273): the seed table on the host (a serial stream of ~N^2 draws), the
// columns on the device.  The whole buffer is cleared first like the reference's memset.
int fpm_gadget_fill_launch(const FpmMesh *m, float *dk, int seed, cudaStream_t st)
{
    const FpmGeom &g = m->geom;
    const size_t nn = (size_t) g.n * g.n;
    std::vector<unsigned int> self(nn), conj(nn);
    fpm_gadget_seed_table(g.n, seed, self.data(), conj.data());
    unsigned int *d_tab = NULL;
    FPM_CUDA_OK(cudaMalloc(&d_tab, 2 * nn * sizeof(unsigned int)));
    FPM_CUDA_OK(cudaMemcpyAsync(d_tab, self.data(), nn * sizeof(unsigned int), cudaMemcpyHostToDevice, st));
    FPM_CUDA_OK(cudaMemcpyAsync(d_tab + nn, conj.data(), nn * sizeof(unsigned int), cudaMemcpyHostToDevice, st));
    FPM_CUDA_OK(cudaMemsetAsync(dk, 0, cplx_total(g) * sizeof(float2), st));
    const size_t ncol = (size_t) g.nyl * g.n;
    FPM_TIMED(FPM_K_OTHER, st, (gadget_fill_kernel<<<(unsigned) ((ncol + 63) / 64), 64, 0, st>>>(g, d_tab, d_tab + nn, (float2 *) dk)));
    cudaError_t e = cudaGetLastError();
    cudaStreamSynchronize(st);                       // the host tables and d_tab go away below
    cudaFree(d_tab);
    fpm_launch_counter++;
    if (e != cudaSuccess) { fpm_set_error("gadget_fill_kernel launch failed: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}

int fpm_set_mode_launch(const FpmMesh *m, float *dk, int ix, int iy, int iz, float re, float im, cudaStream_t st)
{
    FPM_TIMED(FPM_K_OTHER, st, (set_mode_kernel<<<1, 1, 0, st>>>(m->geom, (float2 *) dk, ix, iy, iz, re, im)));
    FPM_CHECK_LAUNCH();
    return 0;
}
#endif
