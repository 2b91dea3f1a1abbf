// fastpm_b200 -- streaming particle updates for sm_100a: kick, drift, periodic wrap, 2LPT
// displacement, grid fill and column summaries.
// Reference: libfastpm/factors.c:73-115 (fastpm_drift_one), :148-171 (fastpm_kick_one), :176-197,
// :374-392 (the store loops); store.c:447-475 (wrap), :723-806 (fill), :808-908 (summary);
// pm2lpt.c:168-210 (pm_2lpt_evolve).
//
// The factor tables (32 samples, interpolated at a_f and at the particle time stamp) are looked up on
// the host exactly as the reference does (factors.c:40-71,117-146) and arrive here as scalars.  Each
// expression keeps the reference's types: float columns, double factors, double position, one rounding
// to float where the reference stores a float -- so kick and drift are bit-identical to the reference.
#include "common.cuh"

struct KickArgs {
    float *v_out; const float *v_in; const float *acc; const float *dx1; const float *dx2;
    double dda, q1, q2, Dv1, Dv2;
    int cola;
    long long n3;      // 3 * np
};

__global__ void __launch_bounds__(256) kick_kernel(const KickArgs a)
{
    long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (; i < a.n3; i += stride) {
        float ax = a.acc[i];
        if (a.cola) {
            const double t = (double) a.dx1[i] * a.q1 + (double) a.dx2[i] * a.q2;
            ax = (float) ((double) ax + t);                                  // ax += (...)
        }
        float vo = (float) ((double) a.v_in[i] + (double) ax * a.dda);       // vo[d] = v + ax * dda
        if (a.cola) {
            const double t = (double) a.dx1[i] * a.Dv1 + (double) a.dx2[i] * a.Dv2;
            vo = (float) ((double) vo + t);                                  // vo[d] += (...)
        }
        a.v_out[i] = vo;
    }
}

struct DriftArgs {
    double *x_out; const double *x_in; const float *v; const float *dx1; const float *dx2;
    double dyyy, da1, da2, Dv1, Dv2;
    int mode;          // FastPMForceType: 0 FASTPM, 1 PM, 2 COLA, 3 2LPT, 4 ZA
    long long n3;
};

__global__ void __launch_bounds__(256) drift_kernel(const DriftArgs a)
{
    long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (; i < a.n3; i += stride) {
        double xo;
        const double x = a.x_in[i];
        switch (a.mode) {
            case 3: xo = x + (double) a.dx1[i] * a.da1 + (double) a.dx2[i] * a.da2; break;
            case 4: xo = x + (double) a.dx1[i] * a.da1; break;
            case 2: {
                const double v = (double) a.v[i] - ((double) a.dx1[i] * a.Dv1 + (double) a.dx2[i] * a.Dv2);
                xo = x + v * a.dyyy;
                xo += (double) a.dx1[i] * a.da1 + (double) a.dx2[i] * a.da2;
                break;
            }
            default: xo = x + (double) a.v[i] * a.dyyy; break;
        }
        a.x_out[i] = xo;
    }
}

// The PGD displacement every drift carries when the store has a pgdc column (factors.c:108-113):
//   xo += 0.5 * pgdc * dyyy / dyyy_table[last], added to the drifted position in double (x is stored as double, so adding it in
//   a second pass over x gives the reference's bits).
__global__ void __launch_bounds__(256) pgd_shift_kernel(double *__restrict__ x, const float *__restrict__ pgdc, const double dyyy,
        const double dyyy_last, const long long n3)
{
    long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (; i < n3; i += stride) x[i] = x[i] + 0.5 * (double) pgdc[i] * dyyy / dyyy_last;
}

// A run of consecutive in-place kicks and drifts of one store (the K K D D between two force evaluations, solver.c:283-356)
// applied in ONE pass: every particle component goes through the same operations, in the same order and with the same
// roundings as kick_kernel / drift_kernel above, but v and x stay in registers between them.
struct FpmUpdateOp { int kind; int mode; double f[5]; };    // kind 0 kick: dda q1 q2 Dv1 Dv2 (mode = cola); 1 drift: dyyy da1 da2 Dv1 Dv2
#define FPM_MAX_UPDATE_OPS 8
struct FusedArgs {
    double *x; float *v; const float *acc; const float *dx1; const float *dx2;
    long long n3;
    int nops, any_kick, any_drift, any_dx;
    FpmUpdateOp ops[FPM_MAX_UPDATE_OPS];
};

__global__ void __launch_bounds__(256) fused_update_kernel(const FusedArgs a)
{
    long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (; i < a.n3; i += stride) {
        float v = a.v[i];
        double x = a.any_drift ? a.x[i] : 0.0;
        const float acc = a.any_kick ? a.acc[i] : 0.f;
        const float d1 = a.any_dx ? a.dx1[i] : 0.f, d2 = (a.any_dx && a.dx2) ? a.dx2[i] : 0.f;
        for (int j = 0; j < a.nops; j++) {
            const FpmUpdateOp &op = a.ops[j];
            if (op.kind == 0) {
                float ax = acc;
                if (op.mode) {
                    const double t = (double) d1 * op.f[1] + (double) d2 * op.f[2];
                    ax = (float) ((double) ax + t);
                }
                float vo = (float) ((double) v + (double) ax * op.f[0]);
                if (op.mode) {
                    const double t = (double) d1 * op.f[3] + (double) d2 * op.f[4];
                    vo = (float) ((double) vo + t);
                }
                v = vo;
            } else {
                double xo;
                switch (op.mode) {
                    case 3: xo = x + (double) d1 * op.f[1] + (double) d2 * op.f[2]; break;
                    case 4: xo = x + (double) d1 * op.f[1]; break;
                    case 2: {
                        const double vv = (double) v - ((double) d1 * op.f[3] + (double) d2 * op.f[4]);
                        xo = x + vv * op.f[0];
                        xo += (double) d1 * op.f[1] + (double) d2 * op.f[2];
                        break;
                    }
                    default: xo = x + (double) v * op.f[0]; break;
                }
                x = xo;
            }
        }
        if (a.any_kick) a.v[i] = v;
        if (a.any_drift) a.x[i] = x;
    }
}

// dst[i] = (float) src[i]: the q column of a freshly filled store is the position rounded to float (store.c:784-789)
__global__ void __launch_bounds__(256) cast_f64_f32_kernel(float *__restrict__ dst, const double *__restrict__ src, long long n)
{
    long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = (float) src[i];
}

// x[i][d] += s[d]: the (de-)shift of the particles around the 2LPT readouts when the ICs sit at cell centres (pm2lpt.c:30-34,141-145)
__global__ void __launch_bounds__(256) shift_kernel(double *x, long long n3, double s0, double s1, double s2)
{
    long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (; i < n3; i += stride) {
        const int d = (int) (i % 3);
        x[i] += d == 0 ? s0 : (d == 1 ? s1 : s2);
    }
}

// store.c:447-475: remainder() then fold into [0, L]; a particle further than 10000 boxes away is an error
__global__ void __launch_bounds__(256) wrap_kernel(double *x, long long n3, double L, int *bad)
{
    long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (; i < n3; i += stride) {
        const double xi = x[i];
        // the reference applies integer abs() to a double here (store.c:453): the truncated ratio
        const double nwrap = (double) abs((int) (xi / L));
        double x1 = remainder(xi, L);
        while (x1 < 0) x1 += L;
        while (x1 > L) x1 -= L;
        x[i] = x1;
        if (nwrap > 10000) atomicExch(bad, 1);
    }
}

// pm2lpt.c:192-208
struct LptEvolveArgs {
    double *x; float *v; const float *dx1; const float *dx2;
    double D1, D2, Dv1, Dv2;
    long long n3;
};
__global__ void __launch_bounds__(256) lpt_evolve_kernel(const LptEvolveArgs a)
{
    long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (; i < a.n3; i += stride) {
        a.x[i] += a.D1 * (double) a.dx1[i] + a.D2 * (double) a.dx2[i];
        if (a.v) {
            float v = a.v[i];
            v = (float) ((double) v + (double) a.dx2[i] * a.Dv2);
            v = (float) ((double) v + a.Dv1 * (double) a.dx1[i]);
            a.v[i] = v;
        }
    }
}

// store.c:756-793: one particle per cell of the nc^3 Lagrangian grid owned by this rank,
// id = i*nc^2 + j*nc + k, x = id-derived index * (L/nc) + shift (store.c:676-692)
__global__ void __launch_bounds__(256) fill_grid_kernel(double *x, unsigned long long *id, float *v, int nc, int i0, long long np,
        double scale, double shift)
{
    long long p = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (; p < np; p += stride) {
        const long long plane = (long long) nc * nc;
        const int i = (int) (p / plane) + i0;
        const long long rem = p % plane;
        const int j = (int) (rem / nc), k = (int) (rem % nc);
        if (id) id[p] = (unsigned long long) i * plane + (unsigned long long) j * nc + k;
        x[3 * p + 0] = i * scale + shift;
        x[3 * p + 1] = j * scale + shift;
        x[3 * p + 2] = k * scale + shift;
        if (v) { v[3 * p] = 0.f; v[3 * p + 1] = 0.f; v[3 * p + 2] = 0.f; }
    }
}

#ifndef FPM_EMULATE          // warp-shuffle reduction and host side: not part of the CPU emulation (tests/emul/particles_emul.cpp)
// ------------------------------------------------------------------ summary (min, max, sum, sum of squares)
template <typename T>
__global__ void __launch_bounds__(256) summary_kernel(const T *col, long long np, int ncomp, double *partial)
{
    // partial[block][comp][4]
    __shared__ double sh[8][4 * 9];
    double mn[9], mx[9], s1[9], s2[9];
    for (int d = 0; d < ncomp; d++) { mn[d] = 1e20; mx[d] = -1e20; s1[d] = 0; s2[d] = 0; }
    long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (; i < np; i += stride)
        for (int d = 0; d < ncomp; d++) {
            const double v = (double) col[i * ncomp + d];
            s1[d] += v; s2[d] += v * v; mn[d] = fmin(mn[d], v); mx[d] = fmax(mx[d], v);
        }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = 0; d < ncomp; d++) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[d] = fmin(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmax(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
            s1[d] += __shfl_xor_sync(0xffffffffu, s1[d], o);
            s2[d] += __shfl_xor_sync(0xffffffffu, s2[d], o);
        }
        if (lane == 0) { sh[warp][4 * d] = mn[d]; sh[warp][4 * d + 1] = mx[d]; sh[warp][4 * d + 2] = s1[d]; sh[warp][4 * d + 3] = s2[d]; }
    }
    __syncthreads();
    if (threadIdx.x < ncomp) {
        const int d = threadIdx.x;
        double a = 1e20, b = -1e20, c = 0, e = 0;
        for (int w = 0; w < (int) (blockDim.x >> 5); w++) {
            a = fmin(a, sh[w][4 * d]); b = fmax(b, sh[w][4 * d + 1]); c += sh[w][4 * d + 2]; e += sh[w][4 * d + 3];
        }
        double *out = partial + ((size_t) blockIdx.x * ncomp + d) * 4;
        out[0] = a; out[1] = b; out[2] = c; out[3] = e;
    }
}

// ------------------------------------------------------------------ sort by a dense particle id without sorting
// fastpm_sort_snapshot (libfastpmio/io.c:860-960) orders a catalog by id with a distributed radix sort.  The ids of this path are the
// Lagrangian indices 0 .. n-1 (store.c:676-692), so the sorted position of a row IS its id: one scatter per column replaces the sort.
// counts[0] += rows whose id lies outside [id0, id0 + n), counts[1] += rows with id[i] != id0 + i (0: the store is in id order)
__global__ void __launch_bounds__(256) id_order_kernel(const unsigned long long *__restrict__ id, long long n, unsigned long long id0,
        unsigned long long *counts)
{
    long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long) gridDim.x * blockDim.x;
    unsigned long long outside = 0, displaced = 0;
    for (; i < n; i += stride) {
        const unsigned long long j = id[i] - id0;                   // unsigned: an id below id0 wraps to a huge value
        outside += j >= (unsigned long long) n;
        displaced += j != (unsigned long long) i;
    }
    if (outside) atomicAdd(counts, outside);
    if (displaced) atomicAdd(counts + 1, displaced);
}

// dst[id[i] - id0] = src[i] for rows of `units` elements of T (T = 4-byte words, or bytes for the odd column); reads are coalesced,
// every row is written whole by `units` neighbouring threads
template <typename T>
__global__ void __launch_bounds__(256) permute_by_id_kernel(T *__restrict__ dst, const T *__restrict__ src,
        const unsigned long long *__restrict__ id, long long total, int units, unsigned long long id0)
{
    long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        const long long i = t / units;
        const int w = (int) (t - i * units);
        dst[(long long) (id[i] - id0) * units + w] = src[t];
    }
}

// ------------------------------------------------------------------ sub-sampling and whole-row moves (store.c:380-412, 967-1034)
// mask[i] = fraction >= 1 || rand[i] <= fraction (store.c:975-979; fraction_each: one fraction per particle, store.c:991-995)
__global__ void __launch_bounds__(256) subsample_mask_kernel(const float *__restrict__ rnd, const double *__restrict__ fraction_each, double fraction,
        long long n, unsigned char *__restrict__ mask)
{
    long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const double f = fraction_each ? fraction_each[i] : fraction;
        const double rand_i = rnd[i];
        mask[i] = (f >= 1 || rand_i <= f) ? 1 : 0;
    }
}

// Stable compaction = exclusive prefix sum of the mask.  Every thread owns a contiguous segment of `seg` rows: first pass counts,
// the (few) per-thread counts are scanned on the host, second pass writes dest[i] = rows kept before row i.
__global__ void __launch_bounds__(256) mask_count_kernel(const unsigned char *__restrict__ mask, long long n, long long seg, long long *__restrict__ counts)
{
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long b = t * seg, e = b + seg < n ? b + seg : n;
    long long c = 0;
    for (long long i = b; i < e; i++) c += mask[i] != 0;
    counts[t] = c;
}
__global__ void __launch_bounds__(256) mask_dest_kernel(const unsigned char *__restrict__ mask, long long n, long long seg, const long long *__restrict__ offsets,
        long long *__restrict__ dest)
{
    const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long b = t * seg, e = b + seg < n ? b + seg : n;
    long long at = offsets[t];
    for (long long i = b; i < e; i++) { dest[i] = at; at += mask[i] != 0; }
}

// dst[dest[i]] = src[i] for the rows with a non-zero mask (rows of `units` elements of T)
template <typename T>
__global__ void __launch_bounds__(256) compact_rows_kernel(T *__restrict__ dst, const T *__restrict__ src, const unsigned char *__restrict__ mask,
        const long long *__restrict__ dest, long long total, int units)
{
    long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        const long long i = t / units;
        if (!mask[i]) continue;
        dst[dest[i] * units + (t - i * units)] = src[t];
    }
}

// dst[i] = src[ind[i]] (fastpm_store_permute, store.c:380-399)
template <typename T>
__global__ void __launch_bounds__(256) gather_rows_kernel(T *__restrict__ dst, const T *__restrict__ src, const int *__restrict__ ind, long long total, int units)
{
    long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long) gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        const long long i = t / units;
        dst[t] = src[(long long) ind[i] * units + (t - i * units)];
    }
}

// ------------------------------------------------------------------ launchers
static inline unsigned stream_grid(long long n)
{
    long long b = (n + 255) / 256;
    const long long cap = 148 * 16;
    return (unsigned) (b < cap ? (b > 0 ? b : 1) : cap);
}

int fpm_kick_launch(float *v_out, const float *v_in, const float *acc, const float *dx1, const float *dx2,
                    double dda, double q1, double q2, double Dv1, double Dv2, int cola, long long np, cudaStream_t st)
{
    if (np <= 0) return 0;
    KickArgs a = { v_out, v_in, acc, dx1, dx2, dda, q1, q2, Dv1, Dv2, cola, 3 * np };
    FPM_TIMED(FPM_K_KICK, st, (kick_kernel<<<stream_grid(a.n3), 256, 0, st>>>(a)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_drift_launch(double *x_out, const double *x_in, const float *v, const float *dx1, const float *dx2,
                     double dyyy, double da1, double da2, double Dv1, double Dv2, int mode, long long np, cudaStream_t st)
{
    if (np <= 0) return 0;
    DriftArgs a = { x_out, x_in, v, dx1, dx2, dyyy, da1, da2, Dv1, Dv2, mode, 3 * np };
    FPM_TIMED(FPM_K_DRIFT, st, (drift_kernel<<<stream_grid(a.n3), 256, 0, st>>>(a)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_pgd_shift_launch(double *x, const float *pgdc, double dyyy, double dyyy_last, long long np, cudaStream_t st)
{
    if (np <= 0) return 0;
    FPM_TIMED(FPM_K_DRIFT, st, (pgd_shift_kernel<<<stream_grid(3 * np), 256, 0, st>>>(x, pgdc, dyyy, dyyy_last, 3 * np)));
    FPM_CHECK_LAUNCH();
    return 0;
}

// ops: [nops][7] doubles = kind, mode, f0..f4
int fpm_fused_update_launch(double *x, float *v, const float *acc, const float *dx1, const float *dx2, long long np,
                            int nops, const double *ops, cudaStream_t st)
{
    if (np <= 0 || nops <= 0) return 0;
    if (nops > FPM_MAX_UPDATE_OPS) { fpm_set_error("fused update: at most %d operations", FPM_MAX_UPDATE_OPS); return -1; }
    FusedArgs a;
    a.x = x; a.v = v; a.acc = acc; a.dx1 = dx1; a.dx2 = dx2; a.n3 = 3 * np; a.nops = nops;
    a.any_kick = a.any_drift = a.any_dx = 0;
    for (int j = 0; j < nops; j++) {
        const double *o = ops + 7 * j;
        a.ops[j].kind = (int) o[0]; a.ops[j].mode = (int) o[1];
        for (int q = 0; q < 5; q++) a.ops[j].f[q] = o[2 + q];
        if (a.ops[j].kind == 0) { a.any_kick = 1; if (a.ops[j].mode) a.any_dx = 1; }
        else { a.any_drift = 1; if (a.ops[j].mode >= 2) a.any_dx = 1; }
    }
    if (a.any_kick && !acc) { fpm_set_error("fused update: a kick needs the acc column"); return -1; }
    if (a.any_dx && !dx1) { fpm_set_error("fused update: COLA / LPT operations need the dx1 (and dx2) columns"); return -1; }
    FPM_TIMED(a.any_drift ? FPM_K_DRIFT : FPM_K_KICK, st, (fused_update_kernel<<<stream_grid(a.n3), 256, 0, st>>>(a)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_cast_f64_f32_launch(float *dst, const double *src, long long n, cudaStream_t st)
{
    if (n <= 0) return 0;
    FPM_TIMED(FPM_K_OTHER, st, (cast_f64_f32_kernel<<<stream_grid(n), 256, 0, st>>>(dst, src, n)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_shift_launch(double *x, long long np, double s0, double s1, double s2, cudaStream_t st)
{
    if (np <= 0) return 0;
    FPM_TIMED(FPM_K_OTHER, st, (shift_kernel<<<stream_grid(3 * np), 256, 0, st>>>(x, 3 * np, s0, s1, s2)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_wrap_launch(double *x, long long np, double L, int *d_bad, cudaStream_t st)
{
    if (np <= 0) return 0;
    FPM_TIMED(FPM_K_OTHER, st, (wrap_kernel<<<stream_grid(3 * np), 256, 0, st>>>(x, 3 * np, L, d_bad)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_lpt_evolve_launch(double *x, float *v, const float *dx1, const float *dx2, double D1, double D2, double Dv1, double Dv2,
                          long long np, cudaStream_t st)
{
    if (np <= 0) return 0;
    LptEvolveArgs a = { x, v, dx1, dx2, D1, D2, Dv1, Dv2, 3 * np };
    FPM_TIMED(FPM_K_OTHER, st, (lpt_evolve_kernel<<<stream_grid(a.n3), 256, 0, st>>>(a)));
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_fill_grid_launch(double *x, unsigned long long *id, float *v, int nc, int i0, long long np, double scale, double shift, cudaStream_t st)
{
    if (np <= 0) return 0;
    FPM_TIMED(FPM_K_OTHER, st, (fill_grid_kernel<<<stream_grid(np), 256, 0, st>>>(x, id, v, nc, i0, np, scale, shift)));
    FPM_CHECK_LAUNCH();
    return 0;
}

// dtype: 4 = float32 column, 8 = float64 column.  host_out[comp][4] = {min, max, sum, sumsq}
int fpm_summary_launch(const void *col, int dtype, int ncomp, long long np, double *host_out, cudaStream_t st)
{
    if (ncomp < 1 || ncomp > 9) { fpm_set_error("summary: ncomp %d out of range", ncomp); return -1; }
    const unsigned grid = stream_grid(np);
    double *d_partial = nullptr;
    FPM_CUDA_OK(cudaMalloc(&d_partial, sizeof(double) * 4 * ncomp * grid));
    if (dtype == 4) FPM_TIMED(FPM_K_SUMMARY, st, (summary_kernel<float><<<grid, 256, 0, st>>>((const float *) col, np, ncomp, d_partial)));
    else FPM_TIMED(FPM_K_SUMMARY, st, (summary_kernel<double><<<grid, 256, 0, st>>>((const double *) col, np, ncomp, d_partial)));
    FPM_CHECK_LAUNCH();
    double *h = (double *) malloc(sizeof(double) * 4 * ncomp * grid);
    FPM_CUDA_OK(cudaMemcpyAsync(h, d_partial, sizeof(double) * 4 * ncomp * grid, cudaMemcpyDeviceToHost, st));
    FPM_CUDA_OK(cudaStreamSynchronize(st));
    for (int d = 0; d < ncomp; d++) { host_out[4 * d] = 1e20; host_out[4 * d + 1] = -1e20; host_out[4 * d + 2] = 0; host_out[4 * d + 3] = 0; }
    for (unsigned b = 0; b < grid; b++)
        for (int d = 0; d < ncomp; d++) {
            const double *p = h + ((size_t) b * ncomp + d) * 4;
            if (p[0] < host_out[4 * d]) host_out[4 * d] = p[0];
            if (p[1] > host_out[4 * d + 1]) host_out[4 * d + 1] = p[1];
            host_out[4 * d + 2] += p[2]; host_out[4 * d + 3] += p[3];
        }
    free(h);
    FPM_CUDA_OK(cudaFree(d_partial));
    return 0;
}

int fpm_id_order_launch(const unsigned long long *id, long long n, unsigned long long id0, unsigned long long *host_counts, cudaStream_t st)
{
    host_counts[0] = host_counts[1] = 0;
    if (n <= 0) return 0;
    unsigned long long *d_counts = nullptr;
    FPM_CUDA_OK(cudaMalloc(&d_counts, 2 * sizeof(unsigned long long)));
    FPM_CUDA_OK(cudaMemsetAsync(d_counts, 0, 2 * sizeof(unsigned long long), st));
    FPM_TIMED(FPM_K_OTHER, st, (id_order_kernel<<<stream_grid(n), 256, 0, st>>>(id, n, id0, d_counts)));
    FPM_CHECK_LAUNCH();
    FPM_CUDA_OK(cudaMemcpyAsync(host_counts, d_counts, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    FPM_CUDA_OK(cudaStreamSynchronize(st));
    FPM_CUDA_OK(cudaFree(d_counts));
    return 0;
}

// every id must lie in [id0, id0 + n) (fpm_id_order_launch: counts[0] == 0), or the scatter writes outside dst
int fpm_permute_by_id_launch(void *dst, const void *src, const unsigned long long *id, long long n, unsigned long long id0, int elsize, cudaStream_t st)
{
    if (elsize < 1) { fpm_set_error("permute by id: element size %d", elsize); return -1; }
    if (n <= 0) return 0;
    if (elsize % 4 == 0 && (uintptr_t) dst % 4 == 0 && (uintptr_t) src % 4 == 0) {
        const long long total = n * (elsize / 4);
        FPM_TIMED(FPM_K_OTHER, st, (permute_by_id_kernel<unsigned int><<<stream_grid(total), 256, 0, st>>>((unsigned int *) dst, (const unsigned int *) src, id, total, elsize / 4, id0)));
    } else {
        const long long total = n * elsize;
        FPM_TIMED(FPM_K_OTHER, st, (permute_by_id_kernel<unsigned char><<<stream_grid(total), 256, 0, st>>>((unsigned char *) dst, (const unsigned char *) src, id, total, elsize, id0)));
    }
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_subsample_mask_launch(const float *rnd, const double *fraction_each, double fraction, long long n, unsigned char *mask, cudaStream_t st)
{
    if (n <= 0) return 0;
    FPM_TIMED(FPM_K_OTHER, st, (subsample_mask_kernel<<<stream_grid(n), 256, 0, st>>>(rnd, fraction_each, fraction, n, mask)));
    FPM_CHECK_LAUNCH();
    return 0;
}

// dest may be NULL (count only); *host_total = number of non-zero mask entries
int fpm_mask_scan_launch(const unsigned char *mask, long long n, long long *dest, long long *host_total, cudaStream_t st)
{
    *host_total = 0;
    if (n <= 0) return 0;
    const unsigned grid = stream_grid((n + 63) / 64);            // at least 64 rows per thread
    const long long nthr = (long long) grid * 256, seg = (n + nthr - 1) / nthr;
    long long *d_counts = nullptr;
    FPM_CUDA_OK(cudaMalloc(&d_counts, sizeof(long long) * nthr));
    FPM_TIMED(FPM_K_OTHER, st, (mask_count_kernel<<<grid, 256, 0, st>>>(mask, n, seg, d_counts)));
    FPM_CHECK_LAUNCH();
    long long *h = (long long *) malloc(sizeof(long long) * nthr);
    FPM_CUDA_OK(cudaMemcpyAsync(h, d_counts, sizeof(long long) * nthr, cudaMemcpyDeviceToHost, st));
    FPM_CUDA_OK(cudaStreamSynchronize(st));
    long long at = 0;
    for (long long t = 0; t < nthr; t++) { const long long c = h[t]; h[t] = at; at += c; }
    *host_total = at;
    if (dest) {
        FPM_CUDA_OK(cudaMemcpyAsync(d_counts, h, sizeof(long long) * nthr, cudaMemcpyHostToDevice, st));
        FPM_TIMED(FPM_K_OTHER, st, (mask_dest_kernel<<<grid, 256, 0, st>>>(mask, n, seg, d_counts, dest)));
        FPM_CHECK_LAUNCH();
        FPM_CUDA_OK(cudaStreamSynchronize(st));
    }
    free(h);
    FPM_CUDA_OK(cudaFree(d_counts));
    return 0;
}

static inline bool rows_as_words(const void *a, const void *b, int elsize) { return elsize % 4 == 0 && (uintptr_t) a % 4 == 0 && (uintptr_t) b % 4 == 0; }

int fpm_compact_rows_launch(void *dst, const void *src, const unsigned char *mask, const long long *dest, long long n, int elsize, cudaStream_t st)
{
    if (elsize < 1) { fpm_set_error("compact rows: element size %d", elsize); return -1; }
    if (n <= 0) return 0;
    if (rows_as_words(dst, src, elsize)) {
        const long long total = n * (elsize / 4);
        FPM_TIMED(FPM_K_OTHER, st, (compact_rows_kernel<unsigned int><<<stream_grid(total), 256, 0, st>>>((unsigned int *) dst, (const unsigned int *) src, mask, dest, total, elsize / 4)));
    } else {
        const long long total = n * elsize;
        FPM_TIMED(FPM_K_OTHER, st, (compact_rows_kernel<unsigned char><<<stream_grid(total), 256, 0, st>>>((unsigned char *) dst, (const unsigned char *) src, mask, dest, total, elsize)));
    }
    FPM_CHECK_LAUNCH();
    return 0;
}

int fpm_gather_rows_launch(void *dst, const void *src, const int *ind, long long n, int elsize, cudaStream_t st)
{
    if (elsize < 1) { fpm_set_error("gather rows: element size %d", elsize); return -1; }
    if (n <= 0) return 0;
    if (rows_as_words(dst, src, elsize)) {
        const long long total = n * (elsize / 4);
        FPM_TIMED(FPM_K_OTHER, st, (gather_rows_kernel<unsigned int><<<stream_grid(total), 256, 0, st>>>((unsigned int *) dst, (const unsigned int *) src, ind, total, elsize / 4)));
    } else {
        const long long total = n * elsize;
        FPM_TIMED(FPM_K_OTHER, st, (gather_rows_kernel<unsigned char><<<stream_grid(total), 256, 0, st>>>((unsigned char *) dst, (const unsigned char *) src, ind, total, elsize)));
    }
    FPM_CHECK_LAUNCH();
    return 0;
}
#endif
