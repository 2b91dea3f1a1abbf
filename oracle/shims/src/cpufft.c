/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * CPU 3-D real<->complex FFT behind the PFFT-API shim (pfft.h).  This restates
 * the arithmetic the reference obtains from PFFT+FFTW3 (not in its tree):
 * single-precision data, twiddles evaluated in double and rounded to float,
 * unnormalised forward (sign -1) and backward (sign +1) DFTs.
 *
 * Algorithm: mixed-radix (2,3,4,5, generic odd) Stockham autosort on batches
 * of VL lines held as split re/im tiles [n][VL] so the inner loop vectorises;
 * the real axis is done as a half-length complex transform plus the usual
 * even/odd untangling.  OpenMP over batches of lines.
 *
 * Also exported (for tests): oracle_fft3_r2c / oracle_fft3_c2r on plain
 * arrays, so the GPU FFT can be compared against it without the reference.
 */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <pfft.h>
#include <fftw3-mpi.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846264338328
#endif

#define VL 16

typedef struct {
    int n;
    int nstage;
    int radix[40];
    float *twr[40];     /* per stage: [m][radix-1] twiddles, forward sign */
    float *twi[40];
} fft1d;

static void fft1d_init(fft1d *p, int n)
{
    p->n = n; p->nstage = 0;
    int r = n;
    while (r > 1) {
        int f;
        if (r % 4 == 0) f = 4;
        else if (r % 2 == 0) f = 2;
        else if (r % 3 == 0) f = 3;
        else if (r % 5 == 0) f = 5;
        else { f = 7; while (r % f) f += 2; }
        p->radix[p->nstage++] = f;
        r /= f;
    }
    int ncur = n;
    for (int s = 0; s < p->nstage; s++) {
        int rad = p->radix[s], m = ncur / rad;
        p->twr[s] = malloc(sizeof(float) * m * (rad - 1) + 64);
        p->twi[s] = malloc(sizeof(float) * m * (rad - 1) + 64);
        for (int q = 0; q < m; q++)
            for (int j = 1; j < rad; j++) {
                double a = -2.0 * M_PI * (double) j * (double) q / (double) ncur;
                p->twr[s][q * (rad - 1) + j - 1] = (float) cos(a);
                p->twi[s][q * (rad - 1) + j - 1] = (float) sin(a);
            }
        ncur = m;
    }
}

static void fft1d_free(fft1d *p)
{
    for (int s = 0; s < p->nstage; s++) { free(p->twr[s]); free(p->twi[s]); }
}

/* One batch of VL transforms.  Data in (ar,ai) as [n][VL]; (br,bi) is work.
 * Returns 0 if the result is in (ar,ai), 1 if in (br,bi).  sign=-1 forward. */
static int fft1d_batch(const fft1d *p, float *restrict ar, float *restrict ai,
                       float *restrict br, float *restrict bi, int sign)
{
    const float sg = (float) sign;
    const float ts = -sg;                   /* twiddle tables hold the forward sign: conjugate for backward */
    int ncur = p->n, s = 1, flip = 0;
    for (int st = 0; st < p->nstage; st++) {
        const int rad = p->radix[st], m = ncur / rad;
        const float *twr = p->twr[st], *twi = p->twi[st];
        float *restrict xr = flip ? br : ar, *restrict xi = flip ? bi : ai;
        float *restrict yr = flip ? ar : br, *restrict yi = flip ? ai : bi;
        for (int pp = 0; pp < m; pp++) {
            const float *wr = twr + pp * (rad - 1), *wi = twi + pp * (rad - 1);
            for (int q = 0; q < s; q++) {
#define IN(k) ((size_t)(q + s * (pp + (k) * m)) * VL)
#define OUT(j) ((size_t)(q + s * (rad * pp + (j))) * VL)
                if (rad == 2) {
                    const float w1r = wr[0], w1i = ts * wi[0];
                    for (int v = 0; v < VL; v++) {
                        float a0r = xr[IN(0) + v], a0i = xi[IN(0) + v];
                        float a1r = xr[IN(1) + v], a1i = xi[IN(1) + v];
                        float dr = a0r - a1r, di = a0i - a1i;
                        yr[OUT(0) + v] = a0r + a1r; yi[OUT(0) + v] = a0i + a1i;
                        yr[OUT(1) + v] = dr * w1r - di * w1i;
                        yi[OUT(1) + v] = dr * w1i + di * w1r;
                    }
                } else if (rad == 4) {
                    const float w1r = wr[0], w1i = ts * wi[0];
                    const float w2r = wr[1], w2i = ts * wi[1];
                    const float w3r = wr[2], w3i = ts * wi[2];
                    for (int v = 0; v < VL; v++) {
                        float a0r = xr[IN(0) + v], a0i = xi[IN(0) + v];
                        float a1r = xr[IN(1) + v], a1i = xi[IN(1) + v];
                        float a2r = xr[IN(2) + v], a2i = xi[IN(2) + v];
                        float a3r = xr[IN(3) + v], a3i = xi[IN(3) + v];
                        float t0r = a0r + a2r, t0i = a0i + a2i;
                        float t1r = a0r - a2r, t1i = a0i - a2i;
                        float t2r = a1r + a3r, t2i = a1i + a3i;
                        /* (a1-a3) * (sign*i): forward (-i): (x+iy)(-i) = y - ix */
                        float ur = a1r - a3r, ui = a1i - a3i;
                        float t3r = -sg * ui, t3i = sg * ur;
                        float b1r = t1r + t3r, b1i = t1i + t3i;
                        float b2r = t0r - t2r, b2i = t0i - t2i;
                        float b3r = t1r - t3r, b3i = t1i - t3i;
                        yr[OUT(0) + v] = t0r + t2r; yi[OUT(0) + v] = t0i + t2i;
                        yr[OUT(1) + v] = b1r * w1r - b1i * w1i; yi[OUT(1) + v] = b1r * w1i + b1i * w1r;
                        yr[OUT(2) + v] = b2r * w2r - b2i * w2i; yi[OUT(2) + v] = b2r * w2i + b2i * w2r;
                        yr[OUT(3) + v] = b3r * w3r - b3i * w3i; yi[OUT(3) + v] = b3r * w3i + b3i * w3r;
                    }
                } else if (rad == 3) {
                    const float w1r = wr[0], w1i = ts * wi[0];
                    const float w2r = wr[1], w2i = ts * wi[1];
                    const float c = -0.5f, sn = sg * 0.86602540378443864676f;   /* exp(sign*2pi i/3) = c + i*sn */
                    for (int v = 0; v < VL; v++) {
                        float a0r = xr[IN(0) + v], a0i = xi[IN(0) + v];
                        float a1r = xr[IN(1) + v], a1i = xi[IN(1) + v];
                        float a2r = xr[IN(2) + v], a2i = xi[IN(2) + v];
                        float sr = a1r + a2r, si = a1i + a2i;
                        float dr = a1r - a2r, di = a1i - a2i;
                        float mr = a0r + c * sr, mi = a0i + c * si;
                        /* i*sn*(d) = (-sn*di, sn*dr) */
                        float b1r = mr - sn * di, b1i = mi + sn * dr;
                        float b2r = mr + sn * di, b2i = mi - sn * dr;
                        yr[OUT(0) + v] = a0r + sr; yi[OUT(0) + v] = a0i + si;
                        yr[OUT(1) + v] = b1r * w1r - b1i * w1i; yi[OUT(1) + v] = b1r * w1i + b1i * w1r;
                        yr[OUT(2) + v] = b2r * w2r - b2i * w2i; yi[OUT(2) + v] = b2r * w2i + b2i * w2r;
                    }
                } else {
                    /* generic radix: direct DFT of size rad */
                    float cr[64], ci[64];
                    for (int j = 0; j < rad; j++) {
                        double a = sign * 2.0 * M_PI * j / rad;
                        cr[j] = (float) cos(a); ci[j] = (float) sin(a);
                    }
                    for (int v = 0; v < VL; v++) {
                        float inr[64], ini[64];
                        for (int k = 0; k < rad; k++) { inr[k] = xr[IN(k) + v]; ini[k] = xi[IN(k) + v]; }
                        for (int j = 0; j < rad; j++) {
                            float accr = 0, acci = 0;
                            for (int k = 0; k < rad; k++) {
                                int t = (j * k) % rad;
                                accr += inr[k] * cr[t] - ini[k] * ci[t];
                                acci += inr[k] * ci[t] + ini[k] * cr[t];
                            }
                            if (j == 0) { yr[OUT(0) + v] = accr; yi[OUT(0) + v] = acci; }
                            else {
                                float wjr = wr[j - 1], wji = ts * wi[j - 1];
                                yr[OUT(j) + v] = accr * wjr - acci * wji;
                                yi[OUT(j) + v] = accr * wji + acci * wjr;
                            }
                        }
                    }
                }
#undef IN
#undef OUT
            }
        }
        flip ^= 1; ncur = m; s *= rad;
    }
    return flip;
}

struct oracle_fft3_plan {
    ptrdiff_t n[3];
    unsigned flags;
    int backward;
    fft1d p0, p1, p2h;          /* axis 0, axis 1, half-length axis 2 */
    float *unr, *uni;           /* untangle twiddles exp(-2 pi i k / n2), k = 0..n2/2 */
    float *scratch;             /* n0*n1*(n2/2+1) complex, for the transposed layouts */
};

static struct oracle_fft3_plan *plan_new(const ptrdiff_t *n, unsigned flags, int backward)
{
    struct oracle_fft3_plan *p = calloc(1, sizeof(*p));
    p->n[0] = n[0]; p->n[1] = n[1]; p->n[2] = n[2];
    p->flags = flags; p->backward = backward;
    fft1d_init(&p->p0, (int) n[0]);
    fft1d_init(&p->p1, (int) n[1]);
    fft1d_init(&p->p2h, (int) (n[2] / 2));
    int h = (int) (n[2] / 2);
    p->unr = malloc(sizeof(float) * (h + 1));
    p->uni = malloc(sizeof(float) * (h + 1));
    for (int k = 0; k <= h; k++) {
        double a = -2.0 * M_PI * k / (double) n[2];
        p->unr[k] = (float) cos(a); p->uni[k] = (float) sin(a);
    }
    p->scratch = NULL;
    return p;
}

static void plan_free(struct oracle_fft3_plan *p)
{
    if (!p) return;
    fft1d_free(&p->p0); fft1d_free(&p->p1); fft1d_free(&p->p2h);
    free(p->unr); free(p->uni); free(p->scratch); free(p);
}

static float *tile_alloc(size_t n) { void *q = NULL; if (posix_memalign(&q, 64, sizeof(float) * n * VL)) abort(); return q; }

/* real axis, forward: rows of n2 reals (pitch 2*(h+1) floats) -> h+1 complex, in place or not */
static void z_forward(const struct oracle_fft3_plan *p, const float *in, float *out)
{
    const int h = (int) (p->n[2] / 2);
    const size_t pitch = 2 * (size_t) (h + 1);
    const size_t nrows = (size_t) p->n[0] * p->n[1];
    const size_t ngroups = (nrows + VL - 1) / VL;
    #pragma omp parallel
    {
        float *ar = tile_alloc(h + 1), *ai = tile_alloc(h + 1), *br = tile_alloc(h + 1), *bi = tile_alloc(h + 1);
        #pragma omp for schedule(static)
        for (size_t g = 0; g < ngroups; g++) {
            size_t r0 = g * VL;
            int nt = (int) ((nrows - r0) < VL ? (nrows - r0) : VL);
            for (int t = 0; t < VL; t++) {
                if (t < nt) {
                    const float *row = in + (r0 + t) * pitch;
                    for (int j = 0; j < h; j++) { ar[j * VL + t] = row[2 * j]; ai[j * VL + t] = row[2 * j + 1]; }
                } else
                    for (int j = 0; j < h; j++) { ar[j * VL + t] = 0; ai[j * VL + t] = 0; }
            }
            int f = fft1d_batch(&p->p2h, ar, ai, br, bi, -1);
            float *zr = f ? br : ar, *zi = f ? bi : ai;
            float *wr = f ? ar : br, *wi = f ? ai : bi;       /* untangled result, k = 0..h */
            for (int k = 0; k <= h; k++) {
                int k1 = (k == h) ? 0 : k, k2 = (h - k) % h;
                float cr = p->unr[k], ci = p->uni[k];
                for (int t = 0; t < VL; t++) {
                    float z1r = zr[k1 * VL + t], z1i = zi[k1 * VL + t];
                    float z2r = zr[k2 * VL + t], z2i = -zi[k2 * VL + t];     /* conj(Z[h-k]) */
                    float er = 0.5f * (z1r + z2r), ei = 0.5f * (z1i + z2i);
                    /* O = (Z1 - conj Z2) / (2i) = (-i/2)(d) = (di/2, -dr/2) */
                    float dr = z1r - z2r, di = z1i - z2i;
                    float or_ = 0.5f * di, oi = -0.5f * dr;
                    wr[k * VL + t] = er + (cr * or_ - ci * oi);
                    wi[k * VL + t] = ei + (cr * oi + ci * or_);
                }
            }
            for (int t = 0; t < nt; t++) {
                float *row = out + (r0 + t) * pitch;
                for (int k = 0; k <= h; k++) { row[2 * k] = wr[k * VL + t]; row[2 * k + 1] = wi[k * VL + t]; }
            }
        }
        free(ar); free(ai); free(br); free(bi);
    }
}

/* real axis, backward: h+1 complex -> n2 reals (unnormalised) */
static void z_backward(const struct oracle_fft3_plan *p, const float *in, float *out)
{
    const int h = (int) (p->n[2] / 2);
    const size_t pitch = 2 * (size_t) (h + 1);
    const size_t nrows = (size_t) p->n[0] * p->n[1];
    const size_t ngroups = (nrows + VL - 1) / VL;
    #pragma omp parallel
    {
        float *ar = tile_alloc(h + 1), *ai = tile_alloc(h + 1), *br = tile_alloc(h + 1), *bi = tile_alloc(h + 1);
        #pragma omp for schedule(static)
        for (size_t g = 0; g < ngroups; g++) {
            size_t r0 = g * VL;
            int nt = (int) ((nrows - r0) < VL ? (nrows - r0) : VL);
            for (int t = 0; t < VL; t++) {
                if (t < nt) {
                    const float *row = in + (r0 + t) * pitch;
                    for (int k = 0; k <= h; k++) { br[k * VL + t] = row[2 * k]; bi[k * VL + t] = row[2 * k + 1]; }
                } else
                    for (int k = 0; k <= h; k++) { br[k * VL + t] = 0; bi[k * VL + t] = 0; }
            }
            for (int k = 0; k < h; k++) {
                float cr = p->unr[k], ci = -p->uni[k];        /* exp(+2 pi i k / n2) */
                for (int t = 0; t < VL; t++) {
                    float x1r = br[k * VL + t], x1i = bi[k * VL + t];
                    float x2r = br[(h - k) * VL + t], x2i = -bi[(h - k) * VL + t];
                    float sr = x1r + x2r, si = x1i + x2i;
                    float dr = x1r - x2r, di = x1i - x2i;
                    /* i * w * d */
                    float wr_ = cr * dr - ci * di, wi_ = cr * di + ci * dr;
                    ar[k * VL + t] = sr - wi_;
                    ai[k * VL + t] = si + wr_;
                }
            }
            int f = fft1d_batch(&p->p2h, ar, ai, br, bi, +1);
            const float *zr = f ? br : ar, *zi = f ? bi : ai;
            for (int t = 0; t < nt; t++) {
                float *row = out + (r0 + t) * pitch;
                for (int j = 0; j < h; j++) { row[2 * j] = zr[j * VL + t]; row[2 * j + 1] = zi[j * VL + t]; }
                row[2 * h] = 0; row[2 * h + 1] = 0;
            }
        }
        free(ar); free(ai); free(br); free(bi);
    }
}

/* complex FFT along axis `ax` (0 or 1) of the [n0][n1][hc] complex array, VL adjacent k per batch */
static void axis_pass_tails(const fft1d *pl, int sign, const float *in, ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t hc, int ax,
                            float *out, int transposed_out, int transposed_in)
{
    /* Layout A (natural): idx(x,y,k) = (x*n1 + y)*hc + k
     * Layout T (PFFT transposed): idx(x,y,k) = (y*hc + k)*n0 + x */
    const int n = pl->n;
    size_t nouter = ax == 0 ? (size_t) n1 : (size_t) n0;
    size_t gper = ((size_t) hc + VL - 1) / VL;
    #pragma omp parallel
    {
        float *ar = tile_alloc(n), *ai = tile_alloc(n), *br = tile_alloc(n), *bi = tile_alloc(n);
        #pragma omp for schedule(static) collapse(2)
        for (size_t o = 0; o < nouter; o++)
            for (size_t g = 0; g < gper; g++) {
                size_t k0 = g * VL;
                int nt = (int) (((size_t) hc - k0) < VL ? ((size_t) hc - k0) : VL);
                for (int e = 0; e < n; e++) {
                    for (int t = 0; t < VL; t++) {
                        if (t < nt) {
                            size_t x = ax == 0 ? (size_t) e : o, y = ax == 0 ? o : (size_t) e, k = k0 + t;
                            size_t idx = transposed_in ? (y * hc + k) * n0 + x : (x * n1 + y) * hc + k;
                            ar[e * VL + t] = in[2 * idx]; ai[e * VL + t] = in[2 * idx + 1];
                        } else { ar[e * VL + t] = 0; ai[e * VL + t] = 0; }
                    }
                }
                int f = fft1d_batch(pl, ar, ai, br, bi, sign);
                const float *rr = f ? br : ar, *ri = f ? bi : ai;
                if (transposed_out) {
                    for (int t = 0; t < nt; t++)
                        for (int e = 0; e < n; e++) {
                            size_t x = ax == 0 ? (size_t) e : o, y = ax == 0 ? o : (size_t) e, k = k0 + t;
                            size_t idx = (y * hc + k) * n0 + x;
                            out[2 * idx] = rr[e * VL + t]; out[2 * idx + 1] = ri[e * VL + t];
                        }
                } else {
                    for (int e = 0; e < n; e++)
                        for (int t = 0; t < nt; t++) {
                            size_t x = ax == 0 ? (size_t) e : o, y = ax == 0 ? o : (size_t) e, k = k0 + t;
                            size_t idx = (x * n1 + y) * hc + k;
                            out[2 * idx] = rr[e * VL + t]; out[2 * idx + 1] = ri[e * VL + t];
                        }
                }
            }
        free(ar); free(ai); free(br); free(bi);
    }
}

static void ensure_scratch(struct oracle_fft3_plan *p)
{
    if (!p->scratch) {
        size_t tot = 2 * (size_t) p->n[0] * p->n[1] * (p->n[2] / 2 + 1);
        p->scratch = malloc(sizeof(float) * tot);
        if (!p->scratch) { fprintf(stderr, "cpufft: out of memory\n"); abort(); }
    }
}

void oracle_fft3_exec_r2c(struct oracle_fft3_plan *p, float *in, float *out)
{
    const ptrdiff_t n0 = p->n[0], n1 = p->n[1], hc = p->n[2] / 2 + 1;
    if (p->flags & PFFT_TRANSPOSED_OUT) {
        float *work = in;
        if (in == out) { ensure_scratch(p); work = p->scratch; }
        z_forward(p, in, work);
        axis_pass_tails(&p->p1, -1, work, n0, n1, hc, 1, work, 0, 0);
        axis_pass_tails(&p->p0, -1, work, n0, n1, hc, 0, out, 1, 0);
        if (work == p->scratch && in == out) { /* result already written to out from scratch */ }
    } else {
        z_forward(p, in, out);
        axis_pass_tails(&p->p1, -1, out, n0, n1, hc, 1, out, 0, 0);
        axis_pass_tails(&p->p0, -1, out, n0, n1, hc, 0, out, 0, 0);
    }
}

void oracle_fft3_exec_c2r(struct oracle_fft3_plan *p, float *in, float *out)
{
    const ptrdiff_t n0 = p->n[0], n1 = p->n[1], hc = p->n[2] / 2 + 1;
    if (p->flags & PFFT_TRANSPOSED_IN) {
        ensure_scratch(p);
        axis_pass_tails(&p->p0, +1, in, n0, n1, hc, 0, p->scratch, 0, 1);
        axis_pass_tails(&p->p1, +1, p->scratch, n0, n1, hc, 1, p->scratch, 0, 0);
        z_backward(p, p->scratch, out);
    } else {
        float *work = in;
        axis_pass_tails(&p->p0, +1, in, n0, n1, hc, 0, work, 0, 0);
        axis_pass_tails(&p->p1, +1, work, n0, n1, hc, 1, work, 0, 0);
        z_backward(p, work, out);
    }
}

struct oracle_fft3_plan *oracle_fft3_plan_new(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int transposed)
{
    ptrdiff_t n[3] = { n0, n1, n2 };
    return plan_new(n, transposed ? (PFFT_TRANSPOSED_OUT | PFFT_TRANSPOSED_IN) : 0, 0);
}
void oracle_fft3_plan_free(struct oracle_fft3_plan *p) { plan_free(p); }

/* ---------------------------------------------------------------- PFFT API */
void pfftf_init(void) {}
void pfftf_cleanup(void) {}
void pfft_init(void) {}
void pfft_cleanup(void) {}

int pfft_create_procmesh(int rnk, MPI_Comm comm, const int *np, MPI_Comm *comm_cart)
{
    for (int i = 0; i < rnk; i++) if (np[i] != 1) { fprintf(stderr, "pfft shim: single rank only\n"); abort(); }
    *comm_cart = comm;
    return 0;
}

ptrdiff_t pfft_local_size_dft_r2c(int rnk_n, const ptrdiff_t *n, MPI_Comm comm_cart, unsigned pfft_flags,
        ptrdiff_t *local_ni, ptrdiff_t *local_i_start, ptrdiff_t *local_no, ptrdiff_t *local_o_start)
{
    (void) comm_cart;
    if (rnk_n != 3) abort();
    for (int d = 0; d < 3; d++) { local_ni[d] = n[d]; local_no[d] = n[d]; local_i_start[d] = 0; local_o_start[d] = 0; }
    local_no[2] = n[2] / 2 + 1;
    if (pfft_flags & PFFT_PADDED_R2C) local_ni[2] = 2 * (n[2] / 2 + 1);
    return n[0] * n[1] * (n[2] / 2 + 1);
}

pfftf_plan pfftf_plan_dft_r2c(int rnk_n, const ptrdiff_t *n, float *in, pfftf_complex *out,
        MPI_Comm comm_cart, int sign, unsigned pfft_flags)
{
    (void) in; (void) out; (void) comm_cart; (void) sign;
    if (rnk_n != 3) abort();
    return plan_new(n, pfft_flags, 0);
}
pfftf_plan pfftf_plan_dft_c2r(int rnk_n, const ptrdiff_t *n, pfftf_complex *in, float *out,
        MPI_Comm comm_cart, int sign, unsigned pfft_flags)
{
    (void) in; (void) out; (void) comm_cart; (void) sign;
    if (rnk_n != 3) abort();
    return plan_new(n, pfft_flags, 1);
}
void pfftf_execute_dft_r2c(const pfftf_plan plan, float *in, pfftf_complex *out)
{ oracle_fft3_exec_r2c(plan, in, (float *) out); }
void pfftf_execute_dft_c2r(const pfftf_plan plan, pfftf_complex *in, float *out)
{ oracle_fft3_exec_c2r(plan, (float *) in, out); }
void pfftf_destroy_plan(pfftf_plan plan) { plan_free(plan); }

/* ------------------------------------------------- FFTW-MPI path: abort stubs */
#define NOFFTW(name) do { fprintf(stderr, "oracle shim: %s: the FFTW-MPI path (-f) is not provided\n", name); abort(); } while (0)
ptrdiff_t fftw_mpi_local_size(int rnk, const ptrdiff_t *n, MPI_Comm comm, ptrdiff_t *a, ptrdiff_t *b)
{ (void) rnk; (void) n; (void) comm; (void) a; (void) b; NOFFTW("fftw_mpi_local_size"); return 0; }
ptrdiff_t fftw_mpi_local_size_transposed(int rnk, const ptrdiff_t *n, MPI_Comm comm, ptrdiff_t *a, ptrdiff_t *b, ptrdiff_t *c, ptrdiff_t *d)
{ (void) rnk; (void) n; (void) comm; (void) a; (void) b; (void) c; (void) d; NOFFTW("fftw_mpi_local_size_transposed"); return 0; }
fftwf_plan fftwf_mpi_plan_dft_r2c(int rnk, const ptrdiff_t *n, float *in, fftwf_complex *out, MPI_Comm comm, unsigned flags)
{ (void) rnk; (void) n; (void) in; (void) out; (void) comm; (void) flags; NOFFTW("fftwf_mpi_plan_dft_r2c"); return NULL; }
fftwf_plan fftwf_mpi_plan_dft_c2r(int rnk, const ptrdiff_t *n, fftwf_complex *in, float *out, MPI_Comm comm, unsigned flags)
{ (void) rnk; (void) n; (void) in; (void) out; (void) comm; (void) flags; NOFFTW("fftwf_mpi_plan_dft_c2r"); return NULL; }
void fftwf_mpi_execute_dft_r2c(fftwf_plan p, float *in, fftwf_complex *out) { (void) p; (void) in; (void) out; NOFFTW("fftwf_mpi_execute_dft_r2c"); }
void fftwf_mpi_execute_dft_c2r(fftwf_plan p, fftwf_complex *in, float *out) { (void) p; (void) in; (void) out; NOFFTW("fftwf_mpi_execute_dft_c2r"); }
void fftwf_destroy_plan(fftwf_plan p) { (void) p; NOFFTW("fftwf_destroy_plan"); }
void fftw_destroy_plan(fftw_plan p) { (void) p; NOFFTW("fftw_destroy_plan"); }
