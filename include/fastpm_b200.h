/* fastpm_b200 -- C ABI of the B200 (sm_100a) particle-mesh force step.
 *
 * This is the drop-in boundary for the hot path of fastpm/fastpm: plain pointers and sizes, no C++
 * or torch types.  Every entry point names the reference function (file:line under libfastpm/ or
 * api/fastpm/ of the reference tree) whose work it replaces.  The C host layer in
 * fastpm_b200/csrc/host/ (FastPMSolver / fastpm_solver_evolve mirror, include/fastpm_b200_solver.h)
 * and the Python bindings (fastpm_b200/_lib.py) are both built on exactly these calls.
 *
 * Conventions
 *   - all `*_dev` / mesh / particle pointers are DEVICE pointers from fpm_malloc() unless a name ends in _host;
 *   - functions return 0 on success, -1 on failure; fpm_last_error() describes the failure;
 *   - work is issued on the library's stream for the current device; fpm_sync() waits for it;
 *   - one process drives one GPU (one rank of an x-slab decomposition over `nranks` GPUs).
 *
 * Device layouts (FpmGeom in csrc/common.cuh)
 *   real mesh   float  [nx_local (+1 halo plane when nranks > 1)][N][pitch_r],  pitch_r = 2*pitch_c
 *   k mesh      float2 [ny_local][N (kx)][pitch_c],  kz = 0..N/2 used, pitch_c = roundup(N/2+1, 16)
 *               (ky-slab "transposed out" order, cf. PFFT_TRANSPOSED_OUT in pmpfft.c:198-203,281-291)
 *   particles   the reference's column layout (api/fastpm/store.h:101-134): x double[np][3],
 *               v/acc/dx1/dx2 float[np][3], id uint64[np]
 */
#ifndef FASTPM_B200_H
#define FASTPM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct FpmMesh fpm_mesh;

/* ---- runtime ------------------------------------------------------------------------------- */
const char *fpm_last_error(void);
const char *fpm_version(void);
int fpm_device_init(int device);                       /* libfastpm_init, libfastpm.c:10 */
int fpm_device_count(void);
int fpm_device_mem_info(size_t *free_bytes, size_t *total_bytes);
/* diagnostic: out = { current device of this thread, the library's device, device owning ptr, memory type of ptr } */
int fpm_debug_state(const void *ptr, int out[4]);
void *fpm_malloc(size_t bytes);                        /* pm_alloc / store columns: memory.c:182 */
void fpm_free(void *ptr);                              /* memory.c:260 */
void *fpm_host_alloc_pinned(size_t bytes);
void fpm_host_free_pinned(void *ptr);
int fpm_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes);
int fpm_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes);
/* the same on a second stream (pinned host memory), for callers that overlap the PCIe traffic of a run with its first / last force
 * evaluation: a copy is ordered after everything the library has queued when it is issued and nothing queued later waits for it
 * until fpm_copy_fence() (device-side wait of the library stream); fpm_copy_wait() blocks the host until the copies are done */
int fpm_memcpy_h2d_async(void *dst_dev, const void *src_host, size_t bytes);
int fpm_memcpy_d2h_async(void *dst_host, const void *src_dev, size_t bytes);
int fpm_copy_fence(void);
int fpm_copy_wait(void);
/* the fence, placed in front of the next kick / drift / fused particle update the library launches: uploads of v, dx1, dx2, id may
 * travel while a force evaluation (which reads the positions only) runs; no handler, hence no flush of the queued updates, is needed */
int fpm_copy_fence_before_update(void);
int fpm_memcpy_d2d(void *dst_dev, const void *src_dev, size_t bytes);    /* pm_assign, pmapi.c:24 */
int fpm_memset(void *dst_dev, int value, size_t bytes);                  /* pm_clear, pmapi.c:30 */
int fpm_sync(void);
/* CUDA-event timers on the library stream (the reference's CLOCK/ENTER/LEAVE, api/fastpm/prof.h:24-28) */
int fpm_timer_create(void **timer);
int fpm_timer_start(void *timer);
int fpm_timer_stop(void *timer);
int fpm_timer_elapsed_ms(void *timer, double *ms);     /* synchronises on the stop event */
void fpm_timer_destroy(void *timer);
uint64_t fpm_kernel_launch_count(void);                /* kernels launched by this library so far */
/* launches so far by code path, for tests that must know WHICH kernels a run went through; out[i], i < n, in the order:
 * TMA tile pass (one GPU), generic tile pass, row (z) pass with bulk copies, generic z pass, TMA tile pass with several
 * destinations (slab transpose), paint with the brick walk, readout with the brick walk, P(k) accumulated inside the forward
 * x-pass (unused: the row-streaming P(k) kernel is counted here), staged slab transposes, paint through shared-memory tiles,
 * readout through shared-memory tiles, the three force components gathered in one pass (cic_readout3_kernel).  Returns the number
 * of counters the library keeps. */
int fpm_path_counts(uint64_t *out, int n);
/* bytes this rank moved to / from its peers over NVLink so far, counted where the transfers are issued: out4 = { slab-transpose
 * pushes by the copy engines, halo planes, migrating particles, rows the transposing FFT pass stored straight into peers } */
int fpm_comm_byte_counts(uint64_t *out4);
/* optional per-kernel-class timing with CUDA events on the launching stream (off by default); classes in order:
 * paint, readout, fft_tile, fft_z, kick, drift, kspace, pk, summary, other, memset, barrier (cross-GPU, includes the wait for
 * the slowest rank), halo, migrate, push (exposed tail of the copy-engine pushes of a staged slab transpose) */
int fpm_prof_enable(int on);
int fpm_prof_reset(void);
int fpm_prof_get(int64_t *counts, double *total_ms, int ncls);
/* the same, launch by launch in issue order (class index, milliseconds); returns the number of launches recorded */
int fpm_prof_get_launches(int32_t *cls, double *ms, int max);

/* ---- mesh object: struct PM, pm_init / pm_destroy, pmpfft.c:108-342 -------------------------- */
fpm_mesh *fpm_mesh_create(int nmesh, double boxsize, int nranks, int rank);
void fpm_mesh_destroy(fpm_mesh *m);
/* info: [0] Nmesh [1] floats per mesh buffer (pm_allocsize) [2] pitch_r [3] pitch_c [4] nx_local [5] x0
 *       [6] ny_local [7] y0 [8] nranks [9] rank [10] halo planes */
int fpm_mesh_info(const fpm_mesh *m, int64_t info[16]);
/* the five float32 per-axis tables of pm_create_k_factors (pmapi.c:235-275), copied to host_out[5][N]
 * in the order k, kk, k_finite, kk_finite, kk_finite2 */
int fpm_mesh_ktables_host(const fpm_mesh *m, float *host_out);

/* ---- K1 CIC paint: fastpm_paint_local + cic_paint_tuned, painter.c:320 / painter-cic.c:34 ---- */
/* canvas += CIC(x) * (M0 + mass[i]) * field[i*field_stride];  mass and field may be NULL.  The
 * canvas is NOT cleared (call fpm_memset first, like pm_clear in gravity.c:310). */
int fpm_paint(const fpm_mesh *m, float *canvas, const double *x, int64_t np,
              double M0, const float *mass, const float *field, int field_stride);
/* Performance hint for fpm_paint / fpm_readout: stores are laid out as i-planes of nc x nc particles in fastpm_store_fill
 * order (index (i*nc + j)*nc + k, store.c:756-793; on several GPUs the local slab of it) and are walked in 4x8x8 Lagrangian
 * bricks for L2 locality when the mesh is large.  nc = 0 clears the hint.  Results do not depend on it (any traversal order
 * is valid, every particle is visited exactly once). */
int fpm_particle_grid_hint(int nc);
/* the hint in force (0: none).  Set by fastpm_store_fill, by fastpm_sort_snapshot on one rank and by a restart from a catalog in id
 * order (csrc/host/io.c); FASTPM_B200_NO_BRICKS=1 keeps it at 0. */
int fpm_particle_grid_hint_get(void);
/* With the hint set and Nmesh / nc <= 2.5 the deposit and the gather run through shared-memory tiles: a CTA takes an 8 x 8 x 8
 * brick of the particle grid, accumulates / stages the box of mesh cells it touches in shared memory and moves that box to / from
 * the mesh in aligned 16-byte groups; particles far from their brick's box (and stores that are not in grid order) use global
 * reductions / loads, so results never depend on the order.  Diagnostic counters, kept when FASTPM_B200_TILE_STATS is set in the
 * environment: out4 = { deposit: particles that took the global path, CTAs without a tile; gather: the same two }. */
int fpm_tile_stats(uint64_t *out4);
/* ---- K5 CIC readout: fastpm_readout_local + cic_readout_tuned, painter.c:358 / painter-cic.c:113 */
/* out[i*out_stride] = (float) sum_8 (float)(canvas * prescale) * w   (prescale 1.0 = none) */
int fpm_readout(const fpm_mesh *m, const float *canvas, const double *x, int64_t np,
                float *out, int out_stride, double prescale);

/* The last of three component readouts into a column of 3-float rows: out3[i] = { comp0[i], comp1[i], readout(canvas)[i] }, written as
 * whole rows.  comp0 / comp1 are the first two components read out into planar scratch arrays (fpm_readout with out_stride 1): the three
 * passes then move 28 bytes per particle of results instead of the 72 that three strided 4-byte stores into 12-byte rows cost in DRAM
 * (every pass dirties and re-fills whole 32-byte sectors).  Values are those of three fpm_readout calls, bit for bit. */
int fpm_readout_pack3(const fpm_mesh *m, const float *canvas, const double *x, int64_t np, const float *comp0, const float *comp1, float *out3);

/* three readouts in one pass over the particles: out3[i][d] = fpm_readout(canvas_d) for d = 0, 1, 2 (the ACC column of the
 * store, gravity.c:359-396), bit-identical to three separate calls; needs the three fields resident at the same time */
int fpm_readout3(const fpm_mesh *m, const float *canvas0, const float *canvas1, const float *canvas2, const double *x, int64_t np, float *out3);

/* the windows other than CIC, _generic_paint / _generic_readout painter.c:217-317: window = FastPMPainterType (1 linear, support 2;
 * 2 quadratic, support 3; 3 Lanczos with the given support <= 8, including the reference's 1e-3 table quantisation); one GPU: several GPUs go through the _ex forms below with a halo block */
int fpm_paint_window(const fpm_mesh *m, int window, int support, float *canvas, const double *x, int64_t np, double M0, const float *mass,
                     const float *field, int field_stride);
int fpm_readout_window(const fpm_mesh *m, int window, int support, const float *canvas, const double *x, int64_t np, float *out, int out_stride);
/* the same with (a) a derivative direction, fastpm_painter_init_diff painter.c:178-205: along axis diffdir (0..2, -1 none) the window
 * is replaced by its derivative / cellsize -- window 0 (CIC) allowed here: cic_paint_tuned / cic_readout_tuned with diffdir,
 * painter-cic.c:57-60,137-140 -- and (b) on several GPUs the block of halo planes that stands in for the reference's ghost particles
 * (pmghosts.c:45-78): fpm_window_halo_planes() says how many planes lie below (left) and above (right) a slab, `halo` holds
 * left + right planes in that order, zeroed before a deposit; NULL on one GPU */
int fpm_paint_window_ex(const fpm_mesh *m, int window, int support, int diffdir, float *canvas, float *halo, const double *x, int64_t np, double M0,
                        const float *mass, const float *field, int field_stride);
int fpm_readout_window_ex(const fpm_mesh *m, int window, int support, int diffdir, const float *canvas, const float *halo, const double *x, int64_t np,
                          float *out, int out_stride);
int fpm_window_halo_planes(int window, int support, int *left, int *right);

/* ---- K2 / K4 FFT: pm_r2c, pm_c2r, pmpfft.c:370-399 ------------------------------------------ */
/* r2c: cplx = DFT(real) * scale.  `real` is destroyed (as with PFFT_DESTROY_INPUT, pmpfft.c:290).
 * The reference's 1/Norm (pmpfft.c:382-385) is passed as scale = 1/N^3 by the caller. */
int fpm_r2c(fpm_mesh *m, float *real, float *cplx, double scale);
/* same, keeping `real` intact: the z- and y-pass intermediate goes to `work` (work != cplx; work == real allowed) */
int fpm_r2c_ws(fpm_mesh *m, const float *real, float *work, float *cplx, double scale);

/* 1: force the generic shared-memory FFT passes (any Nmesh = 2^a 3^b 5^c); 0 (default): the TMA + register passes for
 * Nmesh in {512, 768, 1024, 1536, 2048, 4096}.  Used by the tests to cross-check one against the other. */
int fpm_fft_set_generic(int on);

/* k-space kernel description: see FpmTransferSpec in csrc/mesh.cuh.  Fused into the first pass of c2r. */
typedef struct {
    int32_t active;
    int32_t potorder;        /* -1 none, 0 kk, 1 kk_finite, 2 kk_finite2: fastpm_apply_laplace_transfer, transfer.c:154 */
    int32_t negate;          /* apply_pot_transfer multiplies by -1, gravity.c:14-18 */
    int32_t ngrad;           /* number of i*k gradients: apply_grad_transfer gravity.c:21 / fastpm_apply_diff_transfer transfer.c:116 */
    int32_t graddir[2];
    int32_t gradorder;       /* 0 k, 1 k_finite */
    int32_t zero_selfconj;   /* 1 on the force path (gravity.c:48-56) and for the in-place IC calls (transfer.c:133-148, see device.py) */
    double scale;
} fpm_transfer;

/* c2r: real = IDFT(kernel(cplx)), unnormalised; cplx is preserved; kernel may be NULL (plain pm_c2r).
 * Replaces gravity_apply_kernel_transfer (gravity.c:174-242) + pm_c2r for one field component. */
int fpm_c2r(fpm_mesh *m, const float *cplx, float *real, const fpm_transfer *kernel);
/* same with an explicit work buffer for the x- and y-pass (work != cplx); `real` may then alias `cplx`,
 * which gives the reference's in-place pm_c2r(pm, inplace) */
int fpm_c2r_ws(fpm_mesh *m, const float *cplx, float *work, float *real, const fpm_transfer *kernel);
/* fills an fpm_transfer for FastPMKernelType `kernel_type` (fastpm_kernel_type_get_orders, gravity.c:111-171):
 * attr 0 = ACC component `memb`, attr 1 = POTENTIAL */
int fpm_transfer_for_kernel(int kernel_type, int attr, int memb, fpm_transfer *out);

/* ---- K3 stand-alone k-space sweeps ----------------------------------------------------------- */
int fpm_apply_transfer(const fpm_mesh *m, const float *from, float *to, const fpm_transfer *kernel);
int fpm_apply_decic(const fpm_mesh *m, const float *from, float *to);      /* transfer.c:78 */
/* in-place fpm_apply_decic(m, cplx, cplx), deferred: fpm_powerspectrum* folds it into its read, any other entry point of this
 * library that is handed the buffer applies it first, fpm_decic_cancel drops it (solver.c:471 before the FORCE/after event) */
int fpm_decic_defer(const fpm_mesh *m, float *cplx);
int fpm_decic_cancel(const float *cplx);
/* Deferred work is applied by every entry point of this library that is handed the buffer, but it is invisible to code that
 * reads a device pointer directly (own CUDA kernels, torch on the pointer): call this first.  Applies the pending deconvolution
 * of fpm_decic_defer, if any, and waits for the library stream. */
int fpm_sync_deferred(void);
/* PGD potential, apply_pgdpot_transfer pgdcorrection.c:28-59: to = alpha exp(-kl^2/k^2 - k^4/ks^4) / k^2 * from */
int fpm_apply_pgd_transfer(const fpm_mesh *m, const float *from, float *to, double alpha, double kl, double ks);
/* force softening, gravity.c:244-270.  Radial: mode 0 = low pass, 1 where k^2 < param else 0 (fastpm_apply_lowpass_transfer,
 * transfer.c:43); mode 1 = exp(-36 (k/param)^36) (gaussian36, gravity.c:104).  Separable: to = from * f[ix] * f[iy] * f[iz] with
 * a host table of n doubles (apply_gaussian_softening gravity.c:66-102, fastpm_apply_smoothing_transfer transfer.c:8-41) */
int fpm_apply_radial(const fpm_mesh *m, const float *from, float *to, int mode, double param);
int fpm_apply_axis_factors(const fpm_mesh *m, const float *from, float *to, const double *factors_host);
int fpm_scale(const float *from, float *to, size_t nfloats, double value); /* transfer.c:213 */
int fpm_divide(const float *from, float *to, size_t nfloats, double value); /* solver.c:738-742 */
int fpm_muladd(float *source, const float *a, const float *b, size_t nfloats, int sign); /* pm2lpt.c:103,118 */
int fpm_set_mode(const fpm_mesh *m, float *cplx, int ix, int iy, int iz, float re, float im); /* transfer.c:306 */
/* delta_k *= sqrt(P(k)/V), P log-log interpolated from the table: initialcondition.c:56-64 */
int fpm_induce_correlation(const fpm_mesh *m, float *cplx, const double *k_host, const double *p_host, int size);
/* unit amplitude, phase kept: fastpm_ic_remove_variance, initialcondition.c:66-99 */
int fpm_remove_variance(const fpm_mesh *m, float *cplx);
/* Gaussian white noise in k-space with the Gadget / N-GenIC seeding scheme and RANLUX (gsl_rng_ranlxd1), the reference's
 * default generator: pmic_fill_gaussian_gadget, initialcondition.c:145-273 (fastpm_ic_fill_gaussiank, FASTPM_DELTAK_GADGET).
 * Same seed -> same field as the reference, up to the last bit of the device's double sin / cos / log. */
int fpm_fill_gaussian_gadget(const fpm_mesh *m, float *cplx, int seed);
/* unit-variance real white noise from a counter-based generator (benchmark-size synthetic ICs only) */
int fpm_fill_whitenoise(const fpm_mesh *m, float *real, uint64_t seed);

/* ---- K9 P(k): fastpm_powerspectrum_init_from_delta, powerspectrum.c:35-124 -------------------- */
/* host outputs, each [N/2]: k (mode-weighted mean |k|), p (<|delta|^2> V), nmodes.  decic != 0 applies
 * fastpm_apply_decic_transfer on the fly (solver.c:471) without writing the mesh. */
int fpm_powerspectrum(const fpm_mesh *m, const float *cplx, int decic, double *k_host, double *p_host, double *nmodes_host);
/* raw sums for multi-GPU callers: sums_host[3*(N/2) + 1] = sum w, sum w |delta|^2, sum w |k| per shell, then the sum of
 * w |delta|^2 over ALL modes (pm_compute_variance, pmapi.c:277-295) */
int fpm_powerspectrum_sums(const fpm_mesh *m, const float *cplx, int decic, double *sums_host);
/* the same for a cross spectrum: sum of w * Re(d1 conj d2) per shell (delta1_k != delta2_k in fastpm_powerspectrum_init_from_delta,
 * powerspectrum.c:87-105); a deconvolution pending on either field is applied first */
int fpm_cross_powerspectrum_sums(const fpm_mesh *m, const float *cplx1, const float *cplx2, double *sums_host);

/* ---- K6 / K7 kick, drift: fastpm_kick_store / fastpm_drift_store, factors.c:176,374 ----------- */
/* factors are the already interpolated differences (fastpm_kick_one, factors.c:148-171) */
int fpm_kick(float *v_out, const float *v_in, const float *acc, const float *dx1, const float *dx2, int64_t np,
             int forcemode, double dda, double q1, double q2, double Dv1, double Dv2);
int fpm_drift(double *x_out, const double *x_in, const float *v, const float *dx1, const float *dx2, int64_t np,
              int forcemode, double dyyy, double da1, double da2, double Dv1, double Dv2);
/* A run of consecutive IN-PLACE kicks and drifts of one store in one pass over the particles (the K K D D between two force
 * evaluations of fastpm_solver_evolve, solver.c:283-356): same operations, order and roundings as the separate calls, v and x
 * stay in registers in between.  ops[nops][7] = { kind (0 kick, 1 drift), mode (kick: 1 = COLA; drift: FastPMForceType),
 * then the five factors in the argument order of fpm_kick (dda q1 q2 Dv1 Dv2) / fpm_drift (dyyy da1 da2 Dv1 Dv2) }; nops <= 8 */
int fpm_update_fused(double *x, float *v, const float *acc, const float *dx1, const float *dx2, int64_t np, int nops, const double *ops);
/* the PGD term of fastpm_drift_one (factors.c:108-113), added to already drifted positions:
 * x += 0.5 * pgdc * dyyy / dyyy_last, with dyyy_last = drift->dyyy[nsamples - 1] != 0 */
int fpm_pgd_shift(double *x, const float *pgdc, int64_t np, double dyyy, double dyyy_last);
/* ---- K8 wrap: fastpm_store_wrap, store.c:447 -------------------------------------------------- */
/* the reference aborts when a particle is > 10000 boxes away (store.c:460-471): that flag is reported by the NEXT
 * fpm_wrap call or by fpm_wrap_check() (which waits for the stream), so the integrator itself never blocks on it */
int fpm_wrap(double *x, int64_t np, double boxsize);
/* fpm_wrap followed by fpm_paint in ONE pass over the positions (same operations on the same values; x is written back only
 * where wrapping changed it).  store.c:447 + painter.c:320 */
int fpm_wrap_paint(const fpm_mesh *m, float *canvas, double *x, int64_t np, double M0, const float *mass, const float *field, int field_stride);
int fpm_wrap_check(void);
/* x[i][d] += s_d in place: the (de-)shift around the 2LPT readouts of cell-centred ICs (USE_SHIFT, pm2lpt.c:30-34,141-145) */
int fpm_shift_positions(double *x, int64_t np, double s0, double s1, double s2);
/* the q column of a freshly filled store: dst[i] = (float) src[i] (store.c:784-789) */
int fpm_cast_f64_to_f32(float *dst, const double *src, int64_t n);
/* ---- fastpm_sort_snapshot by particle id (libfastpmio/io.c:860-960, FastPMSnapshotSortByID) when the ids are dense: the sorted
 * position of a row is id - id0, so one scatter per column replaces the reference's distributed radix sort (mpsort).
 * fpm_id_order_counts: host_counts[0] = rows whose id lies outside [id0, id0 + n), host_counts[1] = rows with id[i] != id0 + i
 * (0: already in id order -- for the Lagrangian ids of fastpm_store_fill, store.c:676-692, the order paint / readout walk fastest).
 * fpm_permute_by_id: dst[id[i] - id0] = src[i] for n rows of elsize bytes, out of place; needs host_counts[0] == 0. */
int fpm_id_order_counts(const uint64_t *id, int64_t n, uint64_t id0, uint64_t *host_counts);
int fpm_permute_by_id(void *dst, const void *src, const uint64_t *id, int64_t n, uint64_t id0, int elsize);
/* ---- sub-sampling and whole-row moves of a store (all pointers device memory)
 * fpm_subsample_mask: mask[i] = f >= 1 || rand[i] <= f with f = fraction, or fraction_each_dev[i] when that is not NULL
 *   (fastpm_store_fill_subsample_mask / _from_array, store.c:967-997).
 * fpm_mask_scan: *host_total = non-zero entries of mask; dest (may be NULL) receives dest[i] = non-zero entries before i, the row a
 *   kept particle moves to in a stable compaction (fastpm_store_subsample, store.c:1004-1034; fastpm_store_get_mask_sum, store.c:289).
 * fpm_compact_rows: dst[dest[i]] = src[i] for the rows with a non-zero mask, rows of elsize bytes, out of place.
 * fpm_gather_rows: dst[i] = src[ind[i]] (fastpm_store_permute, store.c:380-399), out of place. */
int fpm_subsample_mask(const float *rand_dev, const double *fraction_each_dev, double fraction, int64_t n, uint8_t *mask);
int fpm_mask_scan(const uint8_t *mask, int64_t n, int64_t *dest, int64_t *host_total);
int fpm_compact_rows(void *dst, const void *src, const uint8_t *mask, const int64_t *dest, int64_t n, int elsize);
int fpm_gather_rows(void *dst, const void *src, const int32_t *ind, int64_t n, int elsize);
/* the rand column, _fastpm_store_fill_rand store.c:694-720: n deviates of this rank's serial RANLUX stream (drawn on the host, copied up) */
int fpm_fill_rand(float *rand_dev, int64_t n, int rank);
/* ---- K10 summary: fastpm_store_summary, store.c:808.  dtype 4 = float32, 8 = float64;
 * host_out[ncomp][4] = min, max, sum, sum of squares */
int fpm_summary(const void *column, int dtype, int ncomp, int64_t np, double *host_out);
/* ---- IC helpers: fastpm_store_fill store.c:723, pm_2lpt_evolve pm2lpt.c:168 ------------------- */
int fpm_fill_grid(double *x, uint64_t *id, float *v, int nc, int i0, int64_t np, double boxsize, double shift);
int fpm_lpt_evolve(double *x, float *v, const float *dx1, const float *dx2, int64_t np,
                   double D1, double D2, double Dv1, double Dv2);

#ifdef __cplusplus
}
#endif
#endif
