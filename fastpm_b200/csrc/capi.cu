// fastpm_b200 -- the extern "C" boundary declared in include/fastpm_b200.h.
#include "common.cuh"
#include "mesh.cuh"
#include "ranlux.h"
#include "../../include/fastpm_b200.h"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

unsigned long long fpm_launch_counter = 0;
unsigned long long fpm_path_counter[FPM_PATH_COUNT] = { 0 };
int fpm_debug_sync = 0;

// ------------------------------------------------------------------ per-class event timing (common.cuh)
int fpm_prof_on = 0;
struct ProfPair { cudaEvent_t a, b; int cls; };
static std::vector<ProfPair> g_prof_pairs;
static size_t g_prof_used = 0;
static int g_prof_open = -1;
void fpm_prof_begin(int cls, cudaStream_t st)
{
    if (g_prof_used == g_prof_pairs.size()) {
        ProfPair p; cudaEventCreate(&p.a); cudaEventCreate(&p.b); p.cls = cls;
        g_prof_pairs.push_back(p);
    }
    g_prof_pairs[g_prof_used].cls = cls;
    cudaEventRecord(g_prof_pairs[g_prof_used].a, st);
    g_prof_open = (int) g_prof_used;
}
void fpm_prof_end(int cls, cudaStream_t st)
{
    (void) cls;
    if (g_prof_open < 0) return;
    cudaEventRecord(g_prof_pairs[g_prof_open].b, st);
    g_prof_used++;
    g_prof_open = -1;
}

// launchers defined in the kernel files
int fpm_paint_launch(const FpmMesh *m, float *canvas, const double *x, const float *mass, double M0, const float *field, int field_stride, long long np, int *wrap_bad, cudaStream_t st);
int fpm_readout_launch(const FpmMesh *m, const float *canvas, const double *x, float *out, int out_stride, double prescale, long long np, cudaStream_t st, const float *pack0 = nullptr, const float *pack1 = nullptr);
int fpm_plane_add_launch(float *dst, const float *src, size_t nfloats, cudaStream_t st);
int fpm_readout3_launch(const FpmMesh *m, const float *c0, const float *c1, const float *c2, const double *x, float *out, long long np, cudaStream_t st);
int fpm_window_paint_launch(const FpmMesh *m, int type, int support, int diffdir, float *canvas, float *halo, const double *x, const float *mass, double M0, const float *field, int field_stride, long long np, cudaStream_t st);
int fpm_window_readout_launch(const FpmMesh *m, int type, int support, int diffdir, const float *canvas, const float *halo, const double *x, float *out, int out_stride, long long np, cudaStream_t st);
int fpm_window_halo(int type, int support, int *left, int *right);
int fpm_kick_launch(float *v_out, const float *v_in, const float *acc, const float *dx1, const float *dx2, double dda, double q1, double q2, double Dv1, double Dv2, int cola, long long np, cudaStream_t st);
int fpm_drift_launch(double *x_out, const double *x_in, const float *v, const float *dx1, const float *dx2, double dyyy, double da1, double da2, double Dv1, double Dv2, int mode, long long np, cudaStream_t st);
int fpm_wrap_launch(double *x, long long np, double L, int *d_bad, cudaStream_t st);
int fpm_shift_launch(double *x, long long np, double s0, double s1, double s2, cudaStream_t st);
int fpm_cast_f64_f32_launch(float *dst, const double *src, long long n, cudaStream_t st);
int fpm_subsample_mask_launch(const float *rnd, const double *fraction_each, double fraction, long long n, unsigned char *mask, cudaStream_t st);
int fpm_mask_scan_launch(const unsigned char *mask, long long n, long long *dest, long long *host_total, cudaStream_t st);
int fpm_compact_rows_launch(void *dst, const void *src, const unsigned char *mask, const long long *dest, long long n, int elsize, cudaStream_t st);
int fpm_gather_rows_launch(void *dst, const void *src, const int *ind, long long n, int elsize, cudaStream_t st);
int fpm_id_order_launch(const unsigned long long *id, long long n, unsigned long long id0, unsigned long long *host_counts, cudaStream_t st);
int fpm_permute_by_id_launch(void *dst, const void *src, const unsigned long long *id, long long n, unsigned long long id0, int elsize, cudaStream_t st);
int fpm_fused_update_launch(double *x, float *v, const float *acc, const float *dx1, const float *dx2, long long np, int nops, const double *ops, cudaStream_t st);
int fpm_lpt_evolve_launch(double *x, float *v, const float *dx1, const float *dx2, double D1, double D2, double Dv1, double Dv2, long long np, cudaStream_t st);
int fpm_fill_grid_launch(double *x, unsigned long long *id, float *v, int nc, int i0, long long np, double scale, double shift, cudaStream_t st);
int fpm_summary_launch(const void *col, int dtype, int ncomp, long long np, double *host_out, cudaStream_t st);
int fpm_transfer_launch(const FpmMesh *m, const float *from, float *to, const FpmTransferSpec *s, cudaStream_t st);
int fpm_decic_launch(const FpmMesh *m, const float *from, float *to, cudaStream_t st);
int fpm_radial_transfer_launch(const FpmMesh *m, const float *from, float *to, int mode, double param, cudaStream_t st);
int fpm_remove_variance_launch(const FpmMesh *m, float *dk, cudaStream_t st);
int fpm_axis_factors_launch(const FpmMesh *m, const double *d_table, const float *from, float *to, cudaStream_t st);
int fpm_pgd_transfer_launch(const FpmMesh *m, const float *from, float *to, double alpha, double kl, double ks, cudaStream_t st);
int fpm_pgd_shift_launch(double *x, const float *pgdc, double dyyy, double dyyy_last, long long np, cudaStream_t st);
int fpm_powerspectrum_launch(const FpmMesh *m, const float *dk, int decic, double *d_out, cudaStream_t st, const float *dk2 = nullptr);
int fpm_scale_launch(const float *from, float *to, size_t nfloats, double value, cudaStream_t st);
int fpm_divide_launch(const float *from, float *to, size_t nfloats, double value, cudaStream_t st);
int fpm_muladd_launch(float *source, const float *a, const float *b, size_t nfloats, int sign, cudaStream_t st);
int fpm_induce_launch(const FpmMesh *m, float *dk, const double *d_tk, const double *d_tp, int size, cudaStream_t st);
int fpm_whitenoise_launch(const FpmMesh *m, float *real, unsigned long long seed, cudaStream_t st);
int fpm_set_mode_launch(const FpmMesh *m, float *dk, int ix, int iy, int iz, float re, float im, cudaStream_t st);
void fpm_fft_force_generic(int on);
int fpm_gadget_fill_launch(const FpmMesh *m, float *dk, int seed, cudaStream_t st);
void fpm_set_lagrangian_hint(int nc);
int fpm_get_lagrangian_hint(void);
int fpm_tile_stats_fetch(unsigned long long out[4]);

// ------------------------------------------------------------------ runtime state
static char g_error[1024] = "";
static cudaStream_t g_stream = nullptr;
static int g_device = -1;

extern "C" void fpm_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

extern "C" int fpm_device_init(int device);
static int ensure_init()
{
    if (g_device >= 0) {
        // the caller (or a library it uses) may have made another device current on this thread since the last call
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != g_device) cudaSetDevice(g_device);
        return 0;
    }
    return fpm_device_init(0);
}

cudaStream_t fpm_internal_stream(void) { ensure_init(); return g_stream; }

// ------------------------------------------------------------------ deferred CIC deconvolution
// fastpm_do_force deconvolves delta_k in place (transfer.c:78, solver.c:471) just before the FORCE/after event, whose usual
// consumer is the P(k) handler.  fpm_decic_defer() records the request instead of sweeping the mesh; fpm_powerspectrum* then
// folds the factor into its read (same roundings: the mode is rounded to float before it is squared), and EVERY other entry
// point of this library that is handed the buffer applies the pending sweep first (lazy_touch), so the deferral is not
// observable through the C ABI.  fpm_decic_cancel() drops it (the buffer is about to be released).
static const fpm_mesh *g_lazy_mesh = NULL;
static const char *g_lazy_buf = NULL;
static size_t g_lazy_bytes = 0;
int fpm_lazy_touch(const void *p, size_t bytes)
{
    if (!g_lazy_buf || !p) return 0;
    const char *q = (const char *) p;
    if (q + (bytes ? bytes : 1) <= g_lazy_buf || q >= g_lazy_buf + g_lazy_bytes) return 0;
    const fpm_mesh *m = g_lazy_mesh;
    float *buf = (float *) g_lazy_buf;
    g_lazy_buf = NULL; g_lazy_mesh = NULL; g_lazy_bytes = 0;
    return fpm_decic_launch(m, buf, buf, g_stream);
}
#define LAZY1(p) do { if (fpm_lazy_touch((p), 0)) return -1; } while (0)
// applies the pending deferred deconvolution, whatever buffer it belongs to (fpm_sync_deferred, include/fastpm_b200.h)
int fpm_lazy_flush_all(void) { return g_lazy_buf ? fpm_lazy_touch(g_lazy_buf, 0) : 0; }

extern "C" {

const char *fpm_last_error(void) { return g_error; }
const char *fpm_version(void) { return "fastpm_b200 0.1 (sm_100a)"; }

int fpm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int fpm_device_init(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        fpm_set_error("fastpm_b200 needs a CUDA device (sm_100a); none is visible: %s", cudaGetErrorString(e));
        return -1;
    }
    FPM_CUDA_OK(cudaSetDevice(device));
    if (g_stream && g_device != device) { cudaStreamDestroy(g_stream); g_stream = nullptr; }
    if (!g_stream) FPM_CUDA_OK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    g_device = device;
    { const char *e = getenv("FASTPM_B200_DEBUG_SYNC"); fpm_debug_sync = e ? atoi(e) : 0; }
    return 0;
}

// diagnostic: current device of this thread, the library's device, and the device that owns `ptr` (or -1)
int fpm_debug_state(const void *ptr, int out[4])
{
    out[0] = out[1] = out[2] = out[3] = -1;
    cudaGetDevice(&out[0]);
    out[1] = g_device;
    if (ptr) {
        cudaPointerAttributes at; memset(&at, 0, sizeof(at));
        if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess) { out[2] = at.device; out[3] = (int) at.type; }
        else cudaGetLastError();
    }
    return 0;
}

int fpm_device_mem_info(size_t *free_bytes, size_t *total_bytes)
{
    if (ensure_init()) return -1;
    FPM_CUDA_OK(cudaMemGetInfo(free_bytes, total_bytes));
    return 0;
}

void *fpm_malloc(size_t bytes)
{
    if (ensure_init()) return NULL;
    void *p = NULL;
    // whole multiples of 2 MiB: such blocks are never sub-allocated by the driver, so a CUDA-IPC handle of the block
    // maps exactly this block at offset 0 in a peer process (multi-GPU meshes, migration buffers, barrier flags)
    const size_t gran = (size_t) 2 << 20;
    const size_t rounded = ((bytes ? bytes : 1) + gran - 1) / gran * gran;
    cudaError_t e = cudaMalloc(&p, rounded);
    if (e != cudaSuccess) { fpm_set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); return NULL; }
    return p;
}
void fpm_free(void *ptr) {
    if (ptr && ptr == (void *) g_lazy_buf) { g_lazy_buf = NULL; g_lazy_mesh = NULL; g_lazy_bytes = 0; }
 if (ptr) cudaFree(ptr); }

void *fpm_host_alloc_pinned(size_t bytes)
{
    void *p = NULL;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { fpm_set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); return NULL; }
    return p;
}
void fpm_host_free_pinned(void *ptr) { if (ptr) cudaFreeHost(ptr); }

int fpm_memcpy_h2d(void *dst, const void *src, size_t bytes)
{
    if (fpm_lazy_touch(dst, bytes)) return -1;
    if (ensure_init()) return -1;
    FPM_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_stream));
    FPM_CUDA_OK(cudaStreamSynchronize(g_stream));
    return 0;
}
int fpm_memcpy_d2h(void *dst, const void *src, size_t bytes)
{
    if (fpm_lazy_touch(src, bytes)) return -1;
    if (ensure_init()) return -1;
    FPM_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream));
    FPM_CUDA_OK(cudaStreamSynchronize(g_stream));
    return 0;
}
// ---- copies on a second stream, for callers that overlap the PCIe traffic of a run with its first / last force evaluation
// (bench.py: the end-to-end arm).  A copy is ordered after everything the library stream has queued when it is issued (its
// source is final, its destination no longer in use); nothing the library queues afterwards waits for it until fpm_copy_fence().
static cudaStream_t g_copy_stream = nullptr;
static cudaEvent_t g_copy_ev = nullptr, g_main_ev = nullptr;
static int copy_stream_init()
{
    if (g_copy_stream) return 0;
    FPM_CUDA_OK(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
    FPM_CUDA_OK(cudaEventCreateWithFlags(&g_copy_ev, cudaEventDisableTiming));
    FPM_CUDA_OK(cudaEventCreateWithFlags(&g_main_ev, cudaEventDisableTiming));
    return 0;
}
static int copy_async(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind)
{
    if (ensure_init() || copy_stream_init()) return -1;
    FPM_CUDA_OK(cudaEventRecord(g_main_ev, g_stream));
    FPM_CUDA_OK(cudaStreamWaitEvent(g_copy_stream, g_main_ev, 0));
    FPM_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, kind, g_copy_stream));
    return 0;
}
int fpm_memcpy_h2d_async(void *dst, const void *src, size_t bytes)
{
    if (fpm_lazy_touch(dst, bytes)) return -1;
    return copy_async(dst, src, bytes, cudaMemcpyHostToDevice);
}
int fpm_memcpy_d2h_async(void *dst, const void *src, size_t bytes)
{
    if (fpm_lazy_touch(src, bytes)) return -1;
    return copy_async(dst, src, bytes, cudaMemcpyDeviceToHost);
}
// the library stream waits (on the device) for the copies issued so far
int fpm_copy_fence(void)
{
    if (!g_copy_stream) return 0;
    FPM_CUDA_OK(cudaEventRecord(g_copy_ev, g_copy_stream));
    FPM_CUDA_OK(cudaStreamWaitEvent(g_stream, g_copy_ev, 0));
    return 0;
}
// the same fence, applied in front of the next kick / drift / fused particle update the library launches (uploads of v, dx1, dx2
// that may still travel while a force evaluation -- which reads only the positions -- runs)
static int g_fence_before_update = 0;
int fpm_copy_fence_before_update(void) { g_fence_before_update = 1; return 0; }
static int fence_if_requested()
{
    if (!g_fence_before_update) return 0;
    g_fence_before_update = 0;
    return fpm_copy_fence();
}
// the host waits for them
int fpm_copy_wait(void)
{
    if (!g_copy_stream) return 0;
    FPM_CUDA_OK(cudaStreamSynchronize(g_copy_stream));
    return 0;
}

int fpm_memcpy_d2d(void *dst, const void *src, size_t bytes)
{
    if (fpm_lazy_touch(src, bytes) || fpm_lazy_touch(dst, bytes)) return -1;
    if (ensure_init()) return -1;
    FPM_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, g_stream));
    return 0;
}
int fpm_memset(void *dst, int value, size_t bytes)
{
    if ((const char *) dst == g_lazy_buf && bytes >= g_lazy_bytes) { g_lazy_buf = NULL; g_lazy_mesh = NULL; g_lazy_bytes = 0; }
    else if (fpm_lazy_touch(dst, bytes)) return -1;
    if (ensure_init()) return -1;
    // large clears (pm_clear of a mesh) are a sweep like any kernel: timed under their own class
    const bool timed = fpm_prof_on && bytes >= ((size_t) 1 << 20);
    if (timed) fpm_prof_begin(FPM_K_MEMSET, g_stream);
    FPM_CUDA_OK(cudaMemsetAsync(dst, value, bytes, g_stream));
    if (timed) fpm_prof_end(FPM_K_MEMSET, g_stream);
    return 0;
}
int fpm_sync(void)
{
    if (ensure_init()) return -1;
    FPM_CUDA_OK(cudaStreamSynchronize(g_stream));
    return 0;
}

struct FpmTimer { cudaEvent_t a, b; };
int fpm_timer_create(void **timer)
{
    if (ensure_init()) return -1;
    FpmTimer *t = new FpmTimer();
    FPM_CUDA_OK(cudaEventCreate(&t->a));
    FPM_CUDA_OK(cudaEventCreate(&t->b));
    *timer = t;
    return 0;
}
int fpm_timer_start(void *timer) { FPM_CUDA_OK(cudaEventRecord(((FpmTimer *) timer)->a, g_stream)); return 0; }
int fpm_timer_stop(void *timer) { FPM_CUDA_OK(cudaEventRecord(((FpmTimer *) timer)->b, g_stream)); return 0; }
int fpm_timer_elapsed_ms(void *timer, double *ms)
{
    FpmTimer *t = (FpmTimer *) timer;
    FPM_CUDA_OK(cudaEventSynchronize(t->b));
    float f = 0;
    FPM_CUDA_OK(cudaEventElapsedTime(&f, t->a, t->b));
    *ms = f;
    return 0;
}
void fpm_timer_destroy(void *timer)
{
    FpmTimer *t = (FpmTimer *) timer;
    if (!t) return;
    cudaEventDestroy(t->a); cudaEventDestroy(t->b);
    delete t;
}
uint64_t fpm_kernel_launch_count(void) { return fpm_launch_counter; }
int fpm_path_counts(uint64_t *out, int n)
{
    for (int i = 0; i < n; i++) out[i] = i < FPM_PATH_COUNT ? fpm_path_counter[i] : 0;
    return FPM_PATH_COUNT;
}

int fpm_prof_enable(int on) { fpm_prof_on = on; return 0; }
int fpm_prof_reset(void) { g_prof_used = 0; g_prof_open = -1; return 0; }
// counts[c], total_ms[c] for c < FPM_K_COUNT (paint, readout, fft_tile, fft_z, kick, drift, kspace, pk, summary, other)
int fpm_prof_get(int64_t *counts, double *total_ms, int ncls)
{
    if (ensure_init()) return -1;
    FPM_CUDA_OK(cudaStreamSynchronize(g_stream));
    for (int c = 0; c < ncls; c++) { counts[c] = 0; total_ms[c] = 0; }
    for (size_t i = 0; i < g_prof_used; i++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, g_prof_pairs[i].a, g_prof_pairs[i].b) != cudaSuccess) { cudaGetLastError(); continue; }
        int c = g_prof_pairs[i].cls;
        if (c >= 0 && c < ncls) { counts[c]++; total_ms[c] += ms; }
    }
    return 0;
}

// per-launch list in issue order: cls[i], ms[i]; returns the number of launches recorded (may exceed max)
int fpm_prof_get_launches(int32_t *cls, double *ms, int max)
{
    if (ensure_init()) return -1;
    FPM_CUDA_OK(cudaStreamSynchronize(g_stream));
    for (size_t i = 0; i < g_prof_used && (int) i < max; i++) {
        float f = 0;
        if (cudaEventElapsedTime(&f, g_prof_pairs[i].a, g_prof_pairs[i].b) != cudaSuccess) { cudaGetLastError(); f = 0; }
        cls[i] = g_prof_pairs[i].cls; ms[i] = f;
    }
    return (int) g_prof_used;
}

// ------------------------------------------------------------------ mesh
static double sinc_unnormed(double x)
{
    if (x < 1e-5 && x > -1e-5) {
        double x2 = x * x;
        return 1.0 - x2 / 6. + x2 * x2 / 120.;
    }
    return sin(x) / x;
}

// pm_create_k_factors, pmapi.c:235-275: every quantity goes through float exactly where the reference's does
static void host_ktables(int n, double boxsize, std::vector<float> &tab, std::vector<double> &decic)
{
    tab.assign((size_t) 5 * n, 0.f);
    decic.assign(n, 0.0);
    const double CellSize = boxsize / n;
    for (int ind = 0; ind < n; ind++) {
        int ii = ind;
        if (ii >= n / 2) ii -= n;
        const double MeshtoK = ii * 2 * M_PI / boxsize;                 // pmpfft.c:332-340
        volatile float k = (float) MeshtoK;
        volatile float w = (float) (k * CellSize);
        volatile float ff1 = (float) sinc_unnormed(0.5 * w);
        volatile float ff2 = (float) sinc_unnormed(w);
        volatile float kk = k * k;
        tab[0 * n + ind] = k;
        tab[1 * n + ind] = kk;
        tab[2 * n + ind] = (float) (1 / CellSize * (1 / 6.0 * (8 * sin((double) w) - sin(2 * (double) w))));
        volatile float f11 = ff1 * ff1;
        tab[3 * n + ind] = kk * f11;
        tab[4 * n + ind] = (float) (kk * (4 / 3.0 * ff1 * ff1 - 1 / 3.0 * ff2 * ff2));
        // decic table, transfer.c:88-96
        const double wd = (double) k * boxsize / (double) n;
        const double cic = sinc_unnormed(0.5 * wd);
        decic[ind] = 1.0 / pow(cic, 2);
    }
}

fpm_mesh *fpm_mesh_create(int nmesh, double boxsize, int nranks, int rank)
{
    if (ensure_init()) return NULL;
    if (nranks < 1 || nranks > FPM_MAX_RANKS || rank < 0 || rank >= nranks) { fpm_set_error("bad rank %d / %d", rank, nranks); return NULL; }
    if (nmesh % nranks != 0) { fpm_set_error("Nmesh = %d is not divisible by the number of slabs %d (cf. solver.c:113-121)", nmesh, nranks); return NULL; }
    FpmMesh *m = new FpmMesh();
    memset(m, 0, sizeof(*m));
    FpmGeom &g = m->geom;
    g.n = nmesh; g.nranks = nranks; g.rank = rank;
    g.nxl = nmesh / nranks; g.x0 = rank * g.nxl;
    g.nyl = nmesh / nranks; g.y0 = rank * g.nyl;
    g.pitch_c = ((nmesh / 2 + 1 + 15) / 16) * 16;
    g.pitch_r = 2 * g.pitch_c;
    g.boxsize = boxsize;
    g.cellsize = boxsize / nmesh;
    g.inv_cellsize = 1.0 / g.cellsize;                 // pmpfft.c:150-151
    if (fpm_fft_plan_create(nmesh, &m->plan)) { delete m; return NULL; }
    std::vector<float> tab; std::vector<double> decic;
    host_ktables(nmesh, boxsize, tab, decic);
    if (cudaMalloc(&m->d_ktab_store, sizeof(float) * tab.size()) != cudaSuccess ||
        cudaMalloc(&m->d_decic, sizeof(double) * nmesh) != cudaSuccess) {
        fpm_set_error("mesh tables: cudaMalloc failed"); delete m; return NULL;
    }
    cudaMemcpy(m->d_ktab_store, tab.data(), sizeof(float) * tab.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(m->d_decic, decic.data(), sizeof(double) * nmesh, cudaMemcpyHostToDevice);
    m->ktab.k = m->d_ktab_store;
    m->ktab.kk = m->d_ktab_store + nmesh;
    m->ktab.k_finite = m->d_ktab_store + 2 * nmesh;
    m->ktab.kk_finite = m->d_ktab_store + 3 * nmesh;
    m->ktab.kk_finite2 = m->d_ktab_store + 4 * nmesh;
    m->ktab.n = nmesh;
    return m;
}

void fpm_mesh_destroy(fpm_mesh *m)
{
    if (!m) return;
    fpm_fft_plan_destroy(m->plan);
    cudaFree(m->d_ktab_store); cudaFree(m->d_decic); cudaFree(m->d_pkgeom);
    delete m;
}

static int halo_planes(const FpmGeom &g) { return g.nranks > 1 ? 1 : 0; }

static size_t mesh_alloc_floats(const FpmGeom &g)
{
    size_t planes_r = (size_t) g.nxl + halo_planes(g), planes_c = (size_t) g.nyl;
    size_t p = planes_r > planes_c ? planes_r : planes_c;
    return p * (size_t) g.n * (size_t) g.pitch_r;
}

int fpm_mesh_info(const fpm_mesh *m, int64_t info[16])
{
    const FpmGeom &g = m->geom;
    memset(info, 0, sizeof(int64_t) * 16);
    info[0] = g.n; info[1] = (int64_t) mesh_alloc_floats(g); info[2] = g.pitch_r; info[3] = g.pitch_c;
    info[4] = g.nxl; info[5] = g.x0; info[6] = g.nyl; info[7] = g.y0; info[8] = g.nranks; info[9] = g.rank;
    info[10] = halo_planes(g);
    return 0;
}

int fpm_mesh_ktables_host(const fpm_mesh *m, float *host_out)
{
    FPM_CUDA_OK(cudaMemcpy(host_out, m->d_ktab_store, sizeof(float) * 5 * m->geom.n, cudaMemcpyDeviceToHost));
    return 0;
}

// ------------------------------------------------------------------ paint / readout
int fpm_paint(const fpm_mesh *m, float *canvas, const double *x, int64_t np, double M0, const float *mass, const float *field, int field_stride)
{
    LAZY1(canvas);
    return fpm_paint_launch(m, canvas, x, mass, M0, field, field_stride, np, NULL, g_stream);
}

int fpm_readout(const fpm_mesh *m, const float *canvas, const double *x, int64_t np, float *out, int out_stride, double prescale)
{
    LAZY1(canvas);
    return fpm_readout_launch(m, canvas, x, out, out_stride, prescale, np, g_stream);
}

int fpm_readout_pack3(const fpm_mesh *m, const float *canvas, const double *x, int64_t np, const float *comp0, const float *comp1, float *out3)
{
    LAZY1(canvas);
    if (!comp0 || !comp1) { fpm_set_error("fpm_readout_pack3: the two stored components are required"); return -1; }
    return fpm_readout_launch(m, canvas, x, out3, 3, 1.0, np, g_stream, comp0, comp1);
}

int fpm_readout3(const fpm_mesh *m, const float *canvas0, const float *canvas1, const float *canvas2, const double *x, int64_t np, float *out3)
{
    LAZY1(canvas0); LAZY1(canvas1); LAZY1(canvas2);
    return fpm_readout3_launch(m, canvas0, canvas1, canvas2, x, out3, np, g_stream);
}

int fpm_paint_window(const fpm_mesh *m, int window, int support, float *canvas, const double *x, int64_t np, double M0, const float *mass,
                     const float *field, int field_stride)
{
    LAZY1(canvas);
    return fpm_window_paint_launch(m, window, support, -1, canvas, nullptr, x, mass, M0, field, field_stride, np, g_stream);
}

int fpm_readout_window(const fpm_mesh *m, int window, int support, const float *canvas, const double *x, int64_t np, float *out, int out_stride)
{
    LAZY1(canvas);
    return fpm_window_readout_launch(m, window, support, -1, canvas, nullptr, x, out, out_stride, np, g_stream);
}

// the same with a derivative direction (fastpm_painter_init_diff; window 0 = CIC allowed) and, on several GPUs, the block of halo planes
int fpm_paint_window_ex(const fpm_mesh *m, int window, int support, int diffdir, float *canvas, float *halo, const double *x, int64_t np, double M0,
                        const float *mass, const float *field, int field_stride)
{
    LAZY1(canvas);
    return fpm_window_paint_launch(m, window, support, diffdir, canvas, halo, x, mass, M0, field, field_stride, np, g_stream);
}

int fpm_readout_window_ex(const fpm_mesh *m, int window, int support, int diffdir, const float *canvas, const float *halo, const double *x, int64_t np,
                          float *out, int out_stride)
{
    LAZY1(canvas);
    return fpm_window_readout_launch(m, window, support, diffdir, canvas, halo, x, out, out_stride, np, g_stream);
}

int fpm_window_halo_planes(int window, int support, int *left, int *right) { return fpm_window_halo(window, support, left, right); }

// ------------------------------------------------------------------ FFT
static void to_spec(const fpm_transfer *k, FpmTransferSpec *s)
{
    s->active = k->active; s->potorder = k->potorder; s->negate = k->negate; s->ngrad = k->ngrad;
    s->graddir[0] = k->graddir[0]; s->graddir[1] = k->graddir[1]; s->gradorder = k->gradorder;
    s->zero_selfconj = k->zero_selfconj; s->scale = k->scale;
}

int fpm_r2c_ws(fpm_mesh *m, const float *real, float *work, float *cplx, double scale)
{
    LAZY1(real); LAZY1(work); LAZY1(cplx);
    if (m->geom.nranks != 1) { fpm_set_error("fpm_r2c: multi-GPU meshes go through the communicator entry points"); return -1; }
    if (work == cplx) { fpm_set_error("fpm_r2c: the work buffer and the k-space output must differ (the y-pass transposes)"); return -1; }
    float *peers[FPM_MAX_RANKS] = { cplx };
    return fpm_fft_r2c(m, real, work, peers, (float) scale, g_stream);
}

int fpm_r2c(fpm_mesh *m, float *real, float *cplx, double scale) { return fpm_r2c_ws(m, real, real, cplx, scale); }

int fpm_c2r_ws(fpm_mesh *m, const float *cplx, float *work, float *real, const fpm_transfer *kernel)
{
    LAZY1(cplx); LAZY1(work); LAZY1(real);
    if (m->geom.nranks != 1) { fpm_set_error("fpm_c2r: multi-GPU meshes go through the communicator entry points"); return -1; }
    if (work == cplx) { fpm_set_error("fpm_c2r: the work buffer and the k-space input must differ (the x-pass transposes)"); return -1; }
    float *peers[FPM_MAX_RANKS] = { work };
    FpmTransferSpec s;
    if (kernel && kernel->active) { to_spec(kernel, &s); return fpm_fft_c2r(m, cplx, peers, real, &s, g_stream); }
    return fpm_fft_c2r(m, cplx, peers, real, NULL, g_stream);
}

int fpm_c2r(fpm_mesh *m, const float *cplx, float *real, const fpm_transfer *kernel) { return fpm_c2r_ws(m, cplx, real, real, kernel); }

// fastpm_kernel_type_get_orders, gravity.c:111-171.  enum order: api/fastpm/libfastpm.h:45-51
int fpm_transfer_for_kernel(int kernel_type, int attr, int memb, fpm_transfer *out)
{
    int potorder, gradorder;
    switch (kernel_type) {
        case 0: potorder = 1; gradorder = 1; break;   /* 3_4 */
        case 1: potorder = 1; gradorder = 0; break;   /* 3_2 */
        case 2: potorder = 2; gradorder = 1; break;   /* 5_4 */
        case 3: potorder = 0; gradorder = 1; break;   /* 1_4 */
        case 4: potorder = 0; gradorder = 1; break;   /* 1_4_DIFF0 */
        case 5: potorder = 0; gradorder = 1; break;   /* GADGET */
        case 6: potorder = 0; gradorder = 0; break;   /* EASTWOOD */
        case 7: potorder = 0; gradorder = 0; break;   /* NAIVE */
        default: fpm_set_error("Wrong kernel type %d", kernel_type); return -1;
    }
    memset(out, 0, sizeof(*out));
    out->active = 1; out->potorder = potorder; out->negate = 1; out->gradorder = gradorder; out->zero_selfconj = 1; out->scale = 1.0;
    if (attr == 0) { out->ngrad = 1; out->graddir[0] = memb; }
    else if (attr == 1) { out->ngrad = 0; }
    else { fpm_set_error("Unknown type for gravity attribute %d", attr); return -1; }
    return 0;
}

// 1: always use the generic shared-memory FFT passes (fft.cu); 0: TMA/register passes where supported (fft_tma.cu)
int fpm_fft_set_generic(int on) { fpm_fft_force_generic(on); return 0; }

// ------------------------------------------------------------------ k-space sweeps
int fpm_apply_transfer(const fpm_mesh *m, const float *from, float *to, const fpm_transfer *kernel)
{
    LAZY1(from); LAZY1(to);
    FpmTransferSpec s; to_spec(kernel, &s);
    return fpm_transfer_launch(m, from, to, &s, g_stream);
}
int fpm_apply_decic(const fpm_mesh *m, const float *from, float *to) { LAZY1(from); LAZY1(to); return fpm_decic_launch(m, from, to, g_stream); }
int fpm_apply_pgd_transfer(const fpm_mesh *m, const float *from, float *to, double alpha, double kl, double ks)
{
    LAZY1(from); LAZY1(to);
    if (!(ks > 0)) { fpm_set_error("pgd transfer: ks must be positive"); return -1; }
    return fpm_pgd_transfer_launch(m, from, to, alpha, kl, ks, g_stream);
}
int fpm_remove_variance(const fpm_mesh *m, float *cplx) { LAZY1(cplx); return fpm_remove_variance_launch(m, cplx, g_stream); }
int fpm_apply_radial(const fpm_mesh *m, const float *from, float *to, int mode, double param)
{
    LAZY1(from); LAZY1(to);
    if (mode != 0 && mode != 1) { fpm_set_error("radial transfer: mode %d", mode); return -1; }
    return fpm_radial_transfer_launch(m, from, to, mode, param, g_stream);
}
int fpm_apply_axis_factors(const fpm_mesh *m, const float *from, float *to, const double *factors_host)
{
    LAZY1(from); LAZY1(to);
    const size_t bytes = sizeof(double) * (size_t) m->geom.n;
    double *d = NULL;
    FPM_CUDA_OK(cudaMallocAsync(&d, bytes, g_stream));
    FPM_CUDA_OK(cudaMemcpyAsync(d, factors_host, bytes, cudaMemcpyHostToDevice, g_stream));
    FPM_CUDA_OK(cudaStreamSynchronize(g_stream));          // factors_host may be a stack array of the caller
    const int rc = fpm_axis_factors_launch(m, d, from, to, g_stream);
    FPM_CUDA_OK(cudaFreeAsync(d, g_stream));
    return rc;
}
int fpm_scale(const float *from, float *to, size_t nfloats, double value) { if (fpm_lazy_touch(from, 4 * nfloats) || fpm_lazy_touch(to, 4 * nfloats)) return -1; return fpm_scale_launch(from, to, nfloats, value, g_stream); }
int fpm_divide(const float *from, float *to, size_t nfloats, double value) { if (fpm_lazy_touch(from, 4 * nfloats) || fpm_lazy_touch(to, 4 * nfloats)) return -1; return fpm_divide_launch(from, to, nfloats, value, g_stream); }
int fpm_muladd(float *source, const float *a, const float *b, size_t nfloats, int sign) { if (fpm_lazy_touch(source, 4 * nfloats) || fpm_lazy_touch(a, 4 * nfloats) || fpm_lazy_touch(b, 4 * nfloats)) return -1; return fpm_muladd_launch(source, a, b, nfloats, sign, g_stream); }
int fpm_set_mode(const fpm_mesh *m, float *cplx, int ix, int iy, int iz, float re, float im) { LAZY1(cplx); return fpm_set_mode_launch(m, cplx, ix, iy, iz, re, im, g_stream); }

int fpm_induce_correlation(const fpm_mesh *m, float *cplx, const double *k_host, const double *p_host, int size)
{
    LAZY1(cplx);
    double *d_k = NULL, *d_p = NULL;
    FPM_CUDA_OK(cudaMalloc(&d_k, sizeof(double) * size));
    FPM_CUDA_OK(cudaMalloc(&d_p, sizeof(double) * size));
    FPM_CUDA_OK(cudaMemcpyAsync(d_k, k_host, sizeof(double) * size, cudaMemcpyHostToDevice, g_stream));
    FPM_CUDA_OK(cudaMemcpyAsync(d_p, p_host, sizeof(double) * size, cudaMemcpyHostToDevice, g_stream));
    int rc = fpm_induce_launch(m, cplx, d_k, d_p, size, g_stream);
    cudaStreamSynchronize(g_stream);
    cudaFree(d_k); cudaFree(d_p);
    return rc;
}

int fpm_fill_gaussian_gadget(const fpm_mesh *m, float *cplx, int seed)
{
    if (ensure_init()) return -1;
    LAZY1(cplx);
    return fpm_gadget_fill_launch(m, cplx, seed, g_stream);
}

int fpm_fill_whitenoise(const fpm_mesh *m, float *real, uint64_t seed) { LAZY1(real); return fpm_whitenoise_launch(m, real, seed, g_stream); }

// the device block the shell sums land in: kept between calls (cudaMalloc / cudaFree synchronise the whole device, copy streams included,
// and P(k) is measured in every step)
static double *pk_sums_buffer(int nbins)
{
    static double *d_buf = NULL;
    static int cap = 0;
    if (nbins > cap) {
        if (d_buf) cudaFree(d_buf);
        d_buf = NULL; cap = 0;
        if (cudaMalloc(&d_buf, sizeof(double) * (3 * (size_t) nbins + 1)) != cudaSuccess) { fpm_set_error("P(k) sums: out of device memory"); cudaGetLastError(); return NULL; }
        cap = nbins;
    }
    return d_buf;
}

int fpm_powerspectrum_sums(const fpm_mesh *m, const float *cplx, int decic, double *sums_host)
{
    if ((const char *) cplx == g_lazy_buf && m == g_lazy_mesh && !decic) decic = 1;      // pending deconvolution folded into the read
    else LAZY1(cplx);
    const int nbins = m->geom.n / 2;
    double *d_out = pk_sums_buffer(nbins);
    if (!d_out) return -1;
    int rc = fpm_powerspectrum_launch(m, cplx, decic, d_out, g_stream);
    if (!rc) {
        FPM_CUDA_OK(cudaMemcpyAsync(sums_host, d_out, sizeof(double) * (3 * nbins + 1), cudaMemcpyDeviceToHost, g_stream));
        FPM_CUDA_OK(cudaStreamSynchronize(g_stream));
    }
    return rc;
}

// cross spectrum of two fields, sum of w * Re(d1 conj d2) per shell (powerspectrum.c:87-105); same layout of sums_host
int fpm_cross_powerspectrum_sums(const fpm_mesh *m, const float *cplx1, const float *cplx2, double *sums_host)
{
    if (cplx1 == cplx2) return fpm_powerspectrum_sums(m, cplx1, 0, sums_host);
    LAZY1(cplx1); LAZY1(cplx2);
    const int nbins = m->geom.n / 2;
    double *d_out = pk_sums_buffer(nbins);
    if (!d_out) return -1;
    int rc = fpm_powerspectrum_launch(m, cplx1, 0, d_out, g_stream, cplx2);
    if (!rc) {
        FPM_CUDA_OK(cudaMemcpyAsync(sums_host, d_out, sizeof(double) * (3 * nbins + 1), cudaMemcpyDeviceToHost, g_stream));
        FPM_CUDA_OK(cudaStreamSynchronize(g_stream));
    }
    return rc;
}

int fpm_powerspectrum(const fpm_mesh *m, const float *cplx, int decic, double *k_host, double *p_host, double *nmodes_host)
{
    const int nbins = m->geom.n / 2;
    std::vector<double> sums((size_t) 3 * nbins + 1);
    if (fpm_powerspectrum_sums(m, cplx, decic, sums.data())) return -1;
    const double L = m->geom.boxsize, volume = L * L * L;
    for (int i = 0; i < nbins; i++) {                       // powerspectrum.c:117-123
        const double nm = sums[i];
        nmodes_host[i] = nm;
        if (nm == 0) { k_host[i] = 0; p_host[i] = 0; continue; }
        k_host[i] = sums[2 * nbins + i] / nm;
        p_host[i] = sums[nbins + i] / nm * volume;
    }
    return 0;
}

int fpm_decic_defer(const fpm_mesh *m, float *cplx)
{
    if (ensure_init()) return -1;
    if (g_lazy_buf && fpm_lazy_touch(g_lazy_buf, 0)) return -1;           // one pending buffer at a time
    if (fpm_lazy_touch(cplx, 0)) return -1;
    int64_t info[16];
    fpm_mesh_info(m, info);
    g_lazy_mesh = m; g_lazy_buf = (const char *) cplx; g_lazy_bytes = (size_t) info[1] * sizeof(float);
    return 0;
}
/* Work this library defers is invisible to code that reads the public device pointers with its own CUDA kernels: call this first.
 * Applies the pending in-place deconvolution of fpm_decic_defer (if any) and waits for the stream. */
int fpm_sync_deferred(void)
{
    if (fpm_lazy_flush_all()) return -1;
    return fpm_sync();
}
int fpm_decic_cancel(const float *cplx)
{
    if ((const char *) cplx == g_lazy_buf) { g_lazy_buf = NULL; g_lazy_mesh = NULL; g_lazy_bytes = 0; }
    return 0;
}

// ------------------------------------------------------------------ particles
int fpm_kick(float *v_out, const float *v_in, const float *acc, const float *dx1, const float *dx2, int64_t np,
             int forcemode, double dda, double q1, double q2, double Dv1, double Dv2)
{
    const int cola = (forcemode == 2);
    if (cola && (!dx1 || !dx2)) { fpm_set_error("COLA kick needs the dx1 and dx2 columns (solver.c:84-88)"); return -1; }
    if (fence_if_requested()) return -1;
    return fpm_kick_launch(v_out, v_in, acc, dx1, dx2, dda, q1, q2, Dv1, Dv2, cola, np, g_stream);
}

int fpm_drift(double *x_out, const double *x_in, const float *v, const float *dx1, const float *dx2, int64_t np,
              int forcemode, double dyyy, double da1, double da2, double Dv1, double Dv2)
{
    if (forcemode >= 2 && (!dx1 || (forcemode != 4 && !dx2))) { fpm_set_error("drift mode %d needs the dx1/dx2 columns", forcemode); return -1; }
    if (fence_if_requested()) return -1;
    return fpm_drift_launch(x_out, x_in, v, dx1, dx2, dyyy, da1, da2, Dv1, Dv2, forcemode, np, g_stream);
}

int fpm_pgd_shift(double *x, const float *pgdc, int64_t np, double dyyy, double dyyy_last)
{
    if (ensure_init()) return -1;
    if (dyyy_last == 0) { fpm_set_error("pgd shift: empty drift interval (factors.c:110-111 skips it)"); return -1; }
    return fpm_pgd_shift_launch(x, pgdc, dyyy, dyyy_last, np, g_stream);
}

int fpm_update_fused(double *x, float *v, const float *acc, const float *dx1, const float *dx2, int64_t np, int nops, const double *ops)
{
    if (ensure_init() || fence_if_requested()) return -1;
    return fpm_fused_update_launch(x, v, acc, dx1, dx2, np, nops, ops, g_stream);
}

// The "too far" flag of the previous wrap is examined when the next one is issued (or by fpm_wrap_check), so that
// the integrator never drains the stream just to look at it.
static int *d_wrap_bad = NULL, *h_wrap_bad = NULL;
int fpm_cast_f64_to_f32(float *dst, const double *src, int64_t n)
{
    if (ensure_init()) return -1;
    return fpm_cast_f64_f32_launch(dst, src, n, g_stream);
}

// fastpm_sort_snapshot by a dense id (libfastpmio/io.c:860-960) on the device: see particles.cu
int fpm_id_order_counts(const uint64_t *id, int64_t n, uint64_t id0, uint64_t *host_counts)
{
    if (ensure_init() || fence_if_requested()) return -1;
    unsigned long long c[2];
    if (fpm_id_order_launch((const unsigned long long *) id, n, id0, c, g_stream)) return -1;
    host_counts[0] = c[0]; host_counts[1] = c[1];
    return 0;
}
// sub-sampling and whole-row moves of a store (store.c:380-412, 967-1034): see particles.cu
int fpm_subsample_mask(const float *rand_dev, const double *fraction_each_dev, double fraction, int64_t n, uint8_t *mask)
{
    if (ensure_init() || fence_if_requested()) return -1;
    return fpm_subsample_mask_launch(rand_dev, fraction_each_dev, fraction, n, mask, g_stream);
}
int fpm_mask_scan(const uint8_t *mask, int64_t n, int64_t *dest, int64_t *host_total)
{
    if (ensure_init() || fence_if_requested()) return -1;
    long long total = 0;
    if (fpm_mask_scan_launch(mask, n, (long long *) dest, &total, g_stream)) return -1;
    *host_total = total;
    return 0;
}
int fpm_compact_rows(void *dst, const void *src, const uint8_t *mask, const int64_t *dest, int64_t n, int elsize)
{
    if (ensure_init() || fence_if_requested()) return -1;
    if (dst == src) { fpm_set_error("compact rows: out of place only"); return -1; }
    return fpm_compact_rows_launch(dst, src, mask, (const long long *) dest, n, elsize, g_stream);
}
int fpm_gather_rows(void *dst, const void *src, const int32_t *ind, int64_t n, int elsize)
{
    if (ensure_init() || fence_if_requested()) return -1;
    if (dst == src) { fpm_set_error("gather rows: out of place only"); return -1; }
    return fpm_gather_rows_launch(dst, src, ind, n, elsize, g_stream);
}
int fpm_permute_by_id(void *dst, const void *src, const uint64_t *id, int64_t n, uint64_t id0, int elsize)
{
    if (ensure_init() || fence_if_requested()) return -1;
    if (dst == src) { fpm_set_error("permute by id: out of place only"); return -1; }
    return fpm_permute_by_id_launch(dst, src, (const unsigned long long *) id, n, id0, elsize, g_stream);
}

// _fastpm_store_fill_rand (store.c:694-720): uniform deviates of ONE serial RANLUX stream per rank over the np_upper entries of the
// column.  The seed of rank r is 0x7fffffff times the (8 r)-th deviate of the generator seeded with 1231584 (the fixed seed itself on
// rank 0), truncated to an integer as gsl_rng_set takes it.  A serial stream cannot be split over CUDA threads: the host draws it,
// a chunk at a time, and the chunks are copied up.
int fpm_fill_rand(float *rand_dev, int64_t n, int rank)
{
    if (ensure_init()) return -1;
    FpmRanlux g;
    double seed = 1231584;
    fpm_ranlux_seed(g, (unsigned long long) seed);
    for (int d = 0; d < rank * 8; d++) seed = 0x7fffffff * fpm_ranlux_uniform(g);
    fpm_ranlux_seed(g, (unsigned long long) seed);
    const int64_t chunk = 1 << 22;
    std::vector<float> buf((size_t) (n < chunk ? (n > 0 ? n : 1) : chunk));
    for (int64_t i0 = 0; i0 < n; i0 += chunk) {
        const int64_t m = n - i0 < chunk ? n - i0 : chunk;
        for (int64_t i = 0; i < m; i++) buf[(size_t) i] = (float) fpm_ranlux_uniform(g);
        FPM_CUDA_OK(cudaMemcpyAsync(rand_dev + i0, buf.data(), sizeof(float) * (size_t) m, cudaMemcpyHostToDevice, g_stream));
        FPM_CUDA_OK(cudaStreamSynchronize(g_stream));
    }
    return 0;
}

int fpm_shift_positions(double *x, int64_t np, double s0, double s1, double s2)
{
    if (ensure_init()) return -1;
    return fpm_shift_launch(x, np, s0, s1, s2, g_stream);
}

int fpm_wrap_check(void)
{
    if (!h_wrap_bad) return 0;
    FPM_CUDA_OK(cudaStreamSynchronize(g_stream));
    if (*h_wrap_bad) {
        *h_wrap_bad = 0;
        FPM_CUDA_OK(cudaMemsetAsync(d_wrap_bad, 0, sizeof(int), g_stream));
        fpm_set_error("A particle is too far from the bounds. Wrapping failed. (store.c:460-471)");
        return -1;
    }
    return 0;
}
int fpm_wrap(double *x, int64_t np, double boxsize)
{
    if (ensure_init()) return -1;
    if (!d_wrap_bad) {
        FPM_CUDA_OK(cudaMalloc(&d_wrap_bad, sizeof(int)));
        FPM_CUDA_OK(cudaHostAlloc(&h_wrap_bad, sizeof(int), cudaHostAllocDefault));
        *h_wrap_bad = 0;
        FPM_CUDA_OK(cudaMemsetAsync(d_wrap_bad, 0, sizeof(int), g_stream));
    }
    if (*h_wrap_bad) {
        FPM_CUDA_OK(cudaStreamSynchronize(g_stream));
        *h_wrap_bad = 0;
        FPM_CUDA_OK(cudaMemsetAsync(d_wrap_bad, 0, sizeof(int), g_stream));
        fpm_set_error("A particle is too far from the bounds. Wrapping failed. (store.c:460-471)");
        return -1;
    }
    if (fpm_wrap_launch(x, np, boxsize, d_wrap_bad, g_stream)) return -1;
    FPM_CUDA_OK(cudaMemcpyAsync(h_wrap_bad, d_wrap_bad, sizeof(int), cudaMemcpyDeviceToHost, g_stream));
    return 0;
}

// the device flag of "a particle is too far" for kernels of other files that wrap on the way (comm.cu: classify); the
// caller enqueues its kernel and then fpm_wrap_flag_fetch() so that the next fpm_wrap / fpm_wrap_check sees the result
int *fpm_wrap_flag_device(void)
{
    if (ensure_init()) return NULL;
    if (!d_wrap_bad) {
        if (cudaMalloc(&d_wrap_bad, sizeof(int)) != cudaSuccess) return NULL;
        if (cudaHostAlloc(&h_wrap_bad, sizeof(int), cudaHostAllocDefault) != cudaSuccess) return NULL;
        *h_wrap_bad = 0;
        cudaMemsetAsync(d_wrap_bad, 0, sizeof(int), g_stream);
    }
    return d_wrap_bad;
}
int fpm_wrap_flag_fetch(void)
{
    FPM_CUDA_OK(cudaMemcpyAsync(h_wrap_bad, d_wrap_bad, sizeof(int), cudaMemcpyDeviceToHost, g_stream));
    return 0;
}

// fastpm_store_wrap + fastpm_paint_local in one pass over the positions (see cic_paint_kernel<.., WRAP>)
int fpm_wrap_paint(const fpm_mesh *m, float *canvas, double *x, int64_t np, double M0, const float *mass, const float *field, int field_stride)
{
    LAZY1(canvas);
    if (ensure_init()) return -1;
    if (!d_wrap_bad) {
        FPM_CUDA_OK(cudaMalloc(&d_wrap_bad, sizeof(int)));
        FPM_CUDA_OK(cudaHostAlloc(&h_wrap_bad, sizeof(int), cudaHostAllocDefault));
        *h_wrap_bad = 0;
        FPM_CUDA_OK(cudaMemsetAsync(d_wrap_bad, 0, sizeof(int), g_stream));
    }
    if (*h_wrap_bad) {
        FPM_CUDA_OK(cudaStreamSynchronize(g_stream));
        *h_wrap_bad = 0;
        FPM_CUDA_OK(cudaMemsetAsync(d_wrap_bad, 0, sizeof(int), g_stream));
        fpm_set_error("A particle is too far from the bounds. Wrapping failed. (store.c:460-471)");
        return -1;
    }
    if (fpm_paint_launch(m, canvas, x, mass, M0, field, field_stride, np, d_wrap_bad, g_stream)) return -1;
    FPM_CUDA_OK(cudaMemcpyAsync(h_wrap_bad, d_wrap_bad, sizeof(int), cudaMemcpyDeviceToHost, g_stream));
    return 0;
}

// performance hint: stores of exactly nc^3 particles are in fastpm_store_fill order (store.c:756-793); 0 clears it
int fpm_particle_grid_hint(int nc) { fpm_set_lagrangian_hint(nc); return 0; }
int fpm_particle_grid_hint_get(void) { return fpm_get_lagrangian_hint(); }
int fpm_tile_stats(uint64_t *out4) { unsigned long long t[4]; if (fpm_tile_stats_fetch(t)) return -1; for (int i = 0; i < 4; i++) out4[i] = t[i]; return 0; }

int fpm_summary(const void *column, int dtype, int ncomp, int64_t np, double *host_out)
{
    return fpm_summary_launch(column, dtype, ncomp, np, host_out, g_stream);
}

int fpm_fill_grid(double *x, uint64_t *id, float *v, int nc, int i0, int64_t np, double boxsize, double shift)
{
    return fpm_fill_grid_launch(x, (unsigned long long *) id, v, nc, i0, np, boxsize / nc, shift, g_stream);
}

int fpm_lpt_evolve(double *x, float *v, const float *dx1, const float *dx2, int64_t np, double D1, double D2, double Dv1, double Dv2)
{
    return fpm_lpt_evolve_launch(x, v, dx1, dx2, D1, D2, Dv1, Dv2, np, g_stream);
}

}   // extern "C"
