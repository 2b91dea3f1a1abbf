// fastpm_b200 -- PM mesh object (the opaque `PM` of the reference API, api/fastpm/libfastpm.h:20)
// and the k-space multiplicative kernels (reference: libfastpm/transfer.c, gravity.c:14-64,174-242,
// pmapi.c:223-275), written so that they can be fused into the load of the inverse x-pass.
#pragma once
#include "common.cuh"
#include "fft_core.h"

struct FpmFftPlan;
struct FpmMesh;

// Per-axis factor tables, all float32 exactly as the reference stores them (pmapi.c:248-270).
// The mesh and box are cubic (PMInit has scalar Nmesh/BoxSize, pmpfft.h:29-35) so one set serves all axes.
struct FpmKTables {
    const float *k;            // k
    const float *kk;           // k^2
    const float *k_finite;     // 4-point central difference: (1/h)(8 sin w - sin 2w)/6
    const float *kk_finite;    // k^2 sinc^2(w/2)
    const float *kk_finite2;   // k^2 (4/3 sinc^2(w/2) - 1/3 sinc^2(w))
    int n;
};

// What to multiply a mode by, in the reference's order of operations and roundings:
//   potorder >= 0 : laplace  v <- (float)(v * (1/sum_d kk[potorder][i_d])), 0 where the sum is 0   (transfer.c:154-186)
//   negate        : v <- -v                                                                      (gravity.c:17)
//   ngrad x       : v <- i * kf[gradorder][i_dir] * v, each product rounded to float             (gravity.c:21-64 / transfer.c:116-151)
//   zero_selfconj : the gradient of a mode whose three indices are all self-conjugate is 0 (gravity.c:48-56;
//                   fastpm_apply_diff_transfer lacks the `else`, transfer.c:142, but is only ever called in
//                   place, so the zero it stored is what it reads: same result)
struct FpmTransferSpec {
    int active;
    int potorder;       // -1: no laplace; 0: kk, 1: kk_finite, 2: kk_finite2
    int negate;
    int ngrad;          // 0, 1 or 2
    int graddir[2];
    int gradorder;      // 0: k, 1: k_finite
    int zero_selfconj;
    double scale;       // final (float)(v * scale) when scale != 1
};

__device__ __forceinline__ float2 fpm_apply_transfer(const FpmTransferSpec &s, const FpmKTables &kt, float2 v, int ix, int iy, int iz)
{
    if (iz > kt.n / 2) return v;                      // padding columns
    if (s.potorder >= 0) {
        const float *kk = s.potorder == 0 ? kt.kk : (s.potorder == 1 ? kt.kk_finite : kt.kk_finite2);
        double sum = 0;
        sum += (double) kk[ix]; sum += (double) kk[iy]; sum += (double) kk[iz];
        if (sum != 0) {
            const double inv = 1 / sum;
            v.x = (float) ((double) v.x * inv);
            v.y = (float) ((double) v.y * inv);
        } else {
            v.x = 0.f; v.y = 0.f;
        }
    }
    if (s.negate) { v.x = -v.x; v.y = -v.y; }
    if (s.ngrad > 0) {
        const int n = kt.n, h = n / 2;
        const bool selfconj = (ix == 0 || ix == h) && (iy == 0 || iy == h) && (iz == 0 || iz == h);
        const float *kf = s.gradorder == 0 ? kt.k : kt.k_finite;
        for (int g = 0; g < s.ngrad; g++) {
            const int dir = s.graddir[g];
            const double f = (double) kf[dir == 0 ? ix : (dir == 1 ? iy : iz)];
            if (s.zero_selfconj && selfconj) {
                v.x = 0.f; v.y = 0.f;
            } else {
                const float re = (float) (-((double) v.y * f));
                const float im = (float) ((double) v.x * f);
                v.x = re; v.y = im;
            }
        }
    }
    if (s.scale != 1.0) {
        v.x = (float) ((double) v.x * s.scale);
        v.y = (float) ((double) v.y * s.scale);
    }
    return v;
}

typedef int (*fpm_barrier_fn)(FpmMesh *m, cudaStream_t st);

struct FpmMesh {
    FpmGeom geom;
    FpmFftPlan *plan;
    FpmKTables ktab;        // device pointers
    float *d_ktab_store;    // one allocation behind ktab
    double *d_decic;        // [n] 1/sinc^2(k h/2) per index (double, as the reference's per-thread table)
    double *d_pkgeom;       // P(k) cache: [2][n/2] sum w, sum w|k| of this rank's modes (geometry only), then [n/2 + 1] scratch
    fpm_barrier_fn barrier; // cross-GPU barrier between a transposing pass and the next (multi-GPU only)
    void *comm;             // opaque communicator (multi-GPU only)
    float *stage;           // multi-GPU: local staging mesh for the slab transpose (NULL: store straight into the peers)
    float *stage2;          // second staging mesh (pipelined inverse transforms: set 1), NULL when there is no room for it
};

int fpm_fft_plan_create(int n, FpmFftPlan **out);
void fpm_fft_plan_destroy(FpmFftPlan *p);
int fpm_fft_r2c(FpmMesh *m, const float *real_in, float *work, float *const *cplx_peers, float scale, cudaStream_t st);
int fpm_fft_c2r(FpmMesh *m, const float *cplx, float *const *real_peers, float *real_out, const FpmTransferSpec *xfer, cudaStream_t st);
int fpm_fft_c2r_begin(FpmMesh *m, const float *cplx, float *const *real_peers, const FpmTransferSpec *xfer, int set, cudaStream_t st);
int fpm_fft_c2r_finish(FpmMesh *m, float *const *real_peers, float *real_out, int set, cudaStream_t st);
