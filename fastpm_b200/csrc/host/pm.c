/* fastpm_b200 host layer -- the PM mesh object and whole-mesh operations behind the reference's
 * pmapi.h / transfer.h entry points (reference: libfastpm/pmpfft.c, pmapi.c, transfer.c).
 * All buffers handed out by pm_alloc are DEVICE memory; every operation is a call into the CUDA
 * library (include/fastpm_b200.h). */
#include "internal.h"

PM *pm_new(int nmesh, double boxsize, MPI_Comm comm)
{
    PM *pm = calloc(1, sizeof(PM));
    pm->NTask = fpm_comm_size(comm);
    pm->ThisTask = fpm_comm_rank(comm);
    pm->comm = comm;
    if (nmesh % 2 != 0) fastpm_raise(-1, "Nmesh must be even, but %d is odd.\n", nmesh);
    pm->mesh = fpm_mesh_create(nmesh, boxsize, pm->NTask, pm->ThisTask);
    if (!pm->mesh) fastpm_raise(-1, "fpm_mesh_create(%d): %s\n", nmesh, fpm_last_error());
    int64_t info[16];
    fpm_mesh_info(pm->mesh, info);
    pm->allocsize = info[1];
    pm->pitch_r = (int) info[2]; pm->pitch_c = (int) info[3];
    pm->nxl = (int) info[4]; pm->x0 = (int) info[5]; pm->nyl = (int) info[6]; pm->y0 = (int) info[7];
    pm->halo = (int) info[10];
    pm->Nproc[0] = pm->NTask; pm->Nproc[1] = 1;          /* x-slabs: the reference's NprocY = 1 (pmpfft.c:117-141) */
    pm->Norm = 1.0; pm->Volume = 1.0;
    for (int d = 0; d < 3; d++) {
        pm->Nmesh[d] = nmesh; pm->BoxSize[d] = boxsize;
        pm->CellSize[d] = pm->BoxSize[d] / pm->Nmesh[d];
        pm->InvCellSize[d] = 1.0 / pm->CellSize[d];
        pm->Norm *= pm->Nmesh[d]; pm->Volume *= pm->BoxSize[d];
    }
    /* regions, in the reference's convention (pmpfft.c:181-211): strides in reals (I) / complex numbers (O) */
    pm->IRegion.start[0] = pm->x0; pm->IRegion.start[1] = 0; pm->IRegion.start[2] = 0;
    pm->IRegion.size[0] = pm->nxl; pm->IRegion.size[1] = nmesh; pm->IRegion.size[2] = nmesh;
    pm->IRegion.strides[2] = 1; pm->IRegion.strides[1] = pm->pitch_r; pm->IRegion.strides[0] = (ptrdiff_t) nmesh * pm->pitch_r;
    pm->IRegion.total = pm->IRegion.size[0] * pm->IRegion.strides[0];
    pm->ORegion.start[0] = 0; pm->ORegion.start[1] = pm->y0; pm->ORegion.start[2] = 0;
    pm->ORegion.size[0] = nmesh; pm->ORegion.size[1] = pm->nyl; pm->ORegion.size[2] = nmesh / 2 + 1;
    pm->ORegion.strides[2] = 1; pm->ORegion.strides[0] = pm->pitch_c; pm->ORegion.strides[1] = (ptrdiff_t) nmesh * pm->pitch_c;
    pm->ORegion.total = pm->ORegion.size[1] * pm->ORegion.strides[1];
    pm->mem = _libfastpm_get_gmem();
    pm->transposed = 1;
    return pm;
}

void pm_delete(PM *pm)
{
    if (!pm) return;
    if (pm->scratch) fpm_free(pm->scratch);
    if (pm->stage2) fastpm_memory_free(pm->mem, pm->stage2);
    if (pm->stage) fastpm_memory_free(pm->mem, pm->stage);
    fpm_mesh_destroy(pm->mesh);
    free(pm);
}

PM *fastpm_create_pm(int Ngrid, int NprocY, int transposed, double BoxSize, MPI_Comm comm)
{
    (void) NprocY; (void) transposed;        /* one decomposition (x-slabs) and one k-space order in this build */
    return pm_new(Ngrid, BoxSize, comm);
}
void fastpm_free_pm(PM *pm) { pm_delete(pm); }

FastPMFloat *pm_alloc_noclear(PM *pm, const char *file, int line)
{ return fastpm_memory_alloc_details(pm->mem, "PMAlloc", sizeof(FastPMFloat) * pm->allocsize, FASTPM_MEMORY_HEAP, file, line); }

FastPMFloat *pm_alloc_details(PM *pm, const char *file, const int line)
{
    FastPMFloat *p = pm_alloc_noclear(pm, file, line);
    FPM_MUST(fpm_memset(p, 0, sizeof(FastPMFloat) * pm->allocsize));       /* pmapi.c:14 */
    return p;
}
void pm_free(PM *pm, FastPMFloat *buf) { fastpm_memory_free(pm->mem, buf); }
void pm_assign(PM *pm, FastPMFloat *from, FastPMFloat *to) { FPM_MUST(fpm_memcpy_d2d(to, from, sizeof(FastPMFloat) * pm->allocsize)); }
void pm_clear(PM *pm, FastPMFloat *buf) { FPM_MUST(fpm_memset(buf, 0, sizeof(FastPMFloat) * pm->allocsize)); }
size_t pm_allocsize(PM *pm) { return pm->allocsize; }
MPI_Comm pm_comm(PM *pm) { return pm->comm; }
double pm_norm(PM *pm) { return pm->Norm; }
ptrdiff_t *pm_nmesh(PM *pm) { return pm->Nmesh; }
int *pm_nproc(PM *pm) { return pm->Nproc; }
double *pm_boxsize(PM *pm) { return pm->BoxSize; }
double pm_volume(PM *pm) { return pm->Volume; }
int pm_unbalanced(PM *pm) { return pm->Nmesh[0] % pm->Nproc[0] != 0; }
PMRegion *pm_i_region(PM *pm) { return &pm->IRegion; }
PMRegion *pm_o_region(PM *pm) { return &pm->ORegion; }

int pm_pos_to_rank(PM *pm, double pos[3])
{
    /* pm_ipos_to_rank, pmpfft.c:353-368, for x-slabs */
    int ipos = (int) floor(pos[0] * pm->InvCellSize[0]);
    int n = (int) pm->Nmesh[0];
    ipos %= n; if (ipos < 0) ipos += n;
    return ipos / pm->nxl;
}

static FastPMFloat *scratch_of(PM *pm)
{
    if (!pm->scratch) {
        pm->scratch = fpm_malloc(sizeof(FastPMFloat) * pm->allocsize);
        if (!pm->scratch) fastpm_raise(-1, "pm scratch mesh: %s\n", fpm_last_error());
    }
    return pm->scratch;
}

/* pm_r2c, pmpfft.c:370-388: out of place, carries 1/Norm.  Unlike PFFT with PFFT_DESTROY_INPUT the input is
 * overwritten with the z/y-pass intermediate only when from != to; from == to goes through the scratch mesh. */
void pm_r2c(PM *pm, FastPMFloat *from, FastPMFloat *to)
{
    if (pm->NTask != 1) {
        /* x-slabs on several GPUs: the distributed transform is out of place and uses its input as work space (PFFT with
         * PFFT_DESTROY_INPUT does the same, pmpfft.c:277-279); an in-place call goes through a copy */
        if (from == to) {
            FastPMFloat *tmp = pm_alloc_noclear(pm, __FILE__, __LINE__);
            pm_assign(pm, from, tmp);
            fpm_mesh_r2c(pm, tmp, to, 1.0 / pm->Norm);
            pm_free(pm, tmp);
        } else {
            fpm_mesh_r2c(pm, from, to, 1.0 / pm->Norm);
        }
        return;
    }
    if (from == to) FPM_MUST(fpm_r2c_ws(pm->mesh, from, scratch_of(pm), to, 1.0 / pm->Norm));
    else FPM_MUST(fpm_r2c(pm->mesh, from, to, 1.0 / pm->Norm));
}

/* pm_c2r, pmpfft.c:390-399: in place, unnormalised */
void pm_c2r(PM *pm, FastPMFloat *inplace)
{
    if (pm->NTask != 1) {
        /* out of place underneath, then back; the plane above the slab is fetched from the x-neighbour right away, so that a CIC
         * fastpm_readout_local on the result is complete (the reference reads the ghosts of its particles instead) */
        FastPMFloat *tmp = pm_alloc_noclear(pm, __FILE__, __LINE__);
        fpm_mesh_c2r(pm, inplace, tmp, NULL);
        pm_assign(pm, tmp, inplace);
        pm_free(pm, tmp);
        fpm_halo_fetch(pm, inplace);
        return;
    }
    FPM_MUST(fpm_c2r_ws(pm->mesh, inplace, scratch_of(pm), inplace, NULL));
}

/* ------------------------------------------------------------------ transfers (transfer.c) */
static void simple_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to, int potorder, int ngrad, int dir, int gradorder)
{
    fpm_transfer t;
    memset(&t, 0, sizeof(t));
    t.active = 1; t.potorder = potorder; t.negate = 0; t.ngrad = ngrad; t.graddir[0] = dir; t.gradorder = gradorder;
    t.zero_selfconj = 1; t.scale = 1.0;
    FPM_MUST(fpm_apply_transfer(pm->mesh, from, to, &t));
}
void fastpm_apply_laplace_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to, int order) { simple_transfer(pm, from, to, order, 0, 0, 0); }
/* transfer.c:116-151.  The reference zeroes the self-conjugate modes of `to` and then overwrites them from
 * `from`; only in-place calls (all its callers) actually end with zeros there.  This build always zeroes. */
void fastpm_apply_diff_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to, int dir, int order) { simple_transfer(pm, from, to, -1, 1, dir, order); }
void fastpm_apply_decic_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to) { FPM_MUST(fpm_apply_decic(pm->mesh, from, to)); }
/* transfer.c:8-41: exp(-k_d^2 sml^2 / 2) per axis, tabulated in double on the host from the float k^2 table like the reference's */
void fastpm_apply_smoothing_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to, double sml)
{
    const int n = (int) pm->Nmesh[0];
    float *tab = malloc(sizeof(float) * 5 * n);
    double *f = malloc(sizeof(double) * n);
    FPM_MUST(fpm_mesh_ktables_host(pm->mesh, tab));
    for (int i = 0; i < n; i++) { double kk = tab[n + i]; f[i] = exp(-0.5 * kk * sml * sml); }
    FPM_MUST(fpm_apply_axis_factors(pm->mesh, from, to, f));
    free(f); free(tab);
}
/* transfer.c:43-66 */
void fastpm_apply_lowpass_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to, double kth)
{ FPM_MUST(fpm_apply_radial(pm->mesh, from, to, 0, kth * kth)); }
void fastpm_apply_multiply_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to, double value)
{ FPM_MUST(fpm_scale(from, to, pm->allocsize, value)); }

void fastpm_apply_modify_mode_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to, ptrdiff_t *mode, double value)
{
    /* transfer.c:290-337 with method 0: set component mode[3] of the mode and of its conjugate */
    if (from != to) pm_assign(pm, from, to);
    ptrdiff_t n = pm->Nmesh[0];
    int self = (mode[0] == (n - mode[0]) % n) && (mode[1] == (n - mode[1]) % n) && (mode[2] == (n - mode[2]) % n);
    if (!self || mode[2] > n / 2) fastpm_raise(-1, "fastpm_b200: modify_mode is implemented for self-conjugate modes only (the DC mode use, src/fastpm.c:541-544)\n");
    /* imaginary part of a self-conjugate mode is pinned to zero (transfer.c:297-303) */
    float re = 0, im = 0;
    float cur[2];
    ptrdiff_t iyl = mode[1] - pm->y0;
    if (iyl >= 0 && iyl < pm->nyl) {
        size_t off = (((size_t) iyl * n + mode[0]) * pm->pitch_c + mode[2]) * 2;
        FPM_MUST(fpm_memcpy_d2h(cur, to + off, sizeof(cur)));
        re = cur[0]; im = cur[1];
        if (mode[3] == 0) re = (float) value; else im = 0;
        FPM_MUST(fpm_set_mode(pm->mesh, to, (int) mode[0], (int) mode[1], (int) mode[2], re, im));
    }
}

static struct { fastpm_fkfunc f; void *d; } induce_cb;
/* initialcondition.c:10-27: the Gadget scheme (the default of the CLI, src/fastpm.c:478) runs on the device */
void fastpm_ic_fill_gaussiank(PM *pm, FastPMFloat *delta_k, int seed, enum FastPMFillDeltaKScheme scheme)
{
    if (scheme != FASTPM_DELTAK_GADGET)
        fastpm_raise(-1, "fastpm_b200: only the FASTPM_DELTAK_GADGET scheme of fastpm_ic_fill_gaussiank is implemented\n");
    FPM_MUST(fpm_fill_gaussian_gadget(pm->mesh, delta_k, seed));
}

void fastpm_ic_remove_variance(PM *pm, FastPMFloat *delta_k) { FPM_MUST(fpm_remove_variance(pm->mesh, delta_k)); }   /* initialcondition.c:66-99 */

void fastpm_ic_induce_correlation(PM *pm, FastPMFloat *delta_k, fastpm_fkfunc pkfunc, void *data)
{
    /* initialcondition.c:56-64 multiplies by sqrt(P(k)/V) through a host callback per mode.  A host callback
     * cannot run in a kernel: the callback is sampled at every distinct |k| the float k^2 tables can produce
     * along a fine log grid and the device interpolates it exactly like fastpm_funck_eval does between
     * table points.  When the callback IS fastpm_powerspectrum_eval2 / fastpm_funck_eval2 the table itself
     * is passed, which reproduces the reference bit for bit. */
    if (pkfunc == (fastpm_fkfunc) fastpm_powerspectrum_eval2 || pkfunc == (fastpm_fkfunc) fastpm_funck_eval2) {
        FastPMFuncK *fk = data;
        FPM_MUST(fpm_induce_correlation(pm->mesh, delta_k, fk->k, fk->f, (int) fk->size));
        return;
    }
    const int ns = 4096;
    double kmin = 2 * M_PI / pm->BoxSize[0] * 0.5, kmax = 2 * M_PI / pm->BoxSize[0] * pm->Nmesh[0] * 2.0;
    double *tk = malloc(sizeof(double) * ns), *tp = malloc(sizeof(double) * ns);
    for (int i = 0; i < ns; i++) {
        tk[i] = kmin * pow(kmax / kmin, i / (double) (ns - 1));
        tp[i] = pkfunc(tk[i], data);
    }
    (void) induce_cb;
    FPM_MUST(fpm_induce_correlation(pm->mesh, delta_k, tk, tp, ns));
    free(tk); free(tp);
}

/* pmapi.c:277-295 */
double pm_compute_variance(PM *pm, FastPMFloat *complx)
{
    int nb = (int) pm->Nmesh[0] / 2;
    double *sums = malloc(sizeof(double) * (3 * nb + 1));
    FPM_MUST(fpm_powerspectrum_sums(pm->mesh, complx, 0, sums));
    double v = sums[3 * nb];                 /* every mode, DC and corners included */
    free(sums);
    fpm_comm_allreduce_double(pm->comm, &v, 1, 0);
    return v / pm->Norm;
}

void pm_check_values(PM *pm, FastPMFloat *field, const char *fmt, ...)
{
    /* pmapi.c:336-356 scans for NaN / |v| > 1e15 and only logs; a min/max summary does the same on the device */
    double out[4];
    FPM_MUST(fpm_summary(field, 4, 1, (int64_t) pm->allocsize, out));
    if (!(out[0] >= -1e15) || !(out[1] <= 1e15) || out[3] != out[3]) {
        char buf[256];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
        fastpm_ilog(INFO, "%s: Task %d has field values that are out of bounds\n", buf, pm->ThisTask);
    }
}

/* ------------------------------------------------------------------ host mirrors in the reference's layouts */
size_t fastpm_b200_mesh_host_size(PM *pm) { return (size_t) pm->Nmesh[0] * pm->Nmesh[1] * (pm->Nmesh[2] + 2); }

int fastpm_b200_mesh_get_real(PM *pm, const FastPMFloat *dev, float *host_dst)
{
    const size_t n = pm->Nmesh[0], rows = (size_t) pm->nxl * n, hp = n + 2;
    float *tmp = malloc(sizeof(float) * rows * pm->pitch_r);
    if (fpm_memcpy_d2h(tmp, dev, sizeof(float) * rows * pm->pitch_r)) { free(tmp); return -1; }
    for (size_t r = 0; r < rows; r++) {
        memcpy(host_dst + r * hp, tmp + r * pm->pitch_r, sizeof(float) * n);
        host_dst[r * hp + n] = 0; host_dst[r * hp + n + 1] = 0;
    }
    free(tmp);
    return 0;
}
int fastpm_b200_mesh_set_real(PM *pm, FastPMFloat *dev, const float *host_src)
{
    const size_t n = pm->Nmesh[0], rows = (size_t) pm->nxl * n, hp = n + 2;
    float *tmp = calloc(rows * pm->pitch_r, sizeof(float));
    for (size_t r = 0; r < rows; r++) memcpy(tmp + r * pm->pitch_r, host_src + r * hp, sizeof(float) * n);
    int rc = fpm_memcpy_h2d(dev, tmp, sizeof(float) * rows * pm->pitch_r);
    free(tmp);
    return rc;
}
/* k-space: device [ky_local][kx][pitch_c]  <->  host [kx][ky][N/2+1] (the full untransposed array; on several ranks
 * every rank passes / receives the full host array and touches only its own ky planes) */
int fastpm_b200_mesh_get_complex(PM *pm, const FastPMFloat *dev, float *host_dst)
{
    const size_t n = pm->Nmesh[0], hc = n / 2 + 1, pc = pm->pitch_c, nyl = pm->nyl, y0 = pm->y0;
    float *tmp = malloc(sizeof(float) * 2 * nyl * n * pc);
    if (fpm_memcpy_d2h(tmp, dev, sizeof(float) * 2 * nyl * n * pc)) { free(tmp); return -1; }
    for (size_t kyl = 0; kyl < nyl; kyl++)
        for (size_t kx = 0; kx < n; kx++)
            memcpy(host_dst + 2 * ((kx * n + kyl + y0) * hc), tmp + 2 * ((kyl * n + kx) * pc), sizeof(float) * 2 * hc);
    free(tmp);
    return 0;
}
int fastpm_b200_mesh_set_complex(PM *pm, FastPMFloat *dev, const float *host_src)
{
    const size_t n = pm->Nmesh[0], hc = n / 2 + 1, pc = pm->pitch_c, nyl = pm->nyl, y0 = pm->y0;
    float *tmp = calloc(2 * nyl * n * pc, sizeof(float));
    for (size_t kyl = 0; kyl < nyl; kyl++)
        for (size_t kx = 0; kx < n; kx++)
            memcpy(tmp + 2 * ((kyl * n + kx) * pc), host_src + 2 * ((kx * n + kyl + y0) * hc), sizeof(float) * 2 * hc);
    int rc = fpm_memcpy_h2d(dev, tmp, sizeof(float) * 2 * nyl * n * pc);
    free(tmp);
    return rc;
}

/* ------------------------------------------------------------------ variable mesh list (vpm.c) */
VPM *vpm_create(VPMInit *vpminit, int base_nmesh, double boxsize, MPI_Comm comm)
{
    int size = 0;
    while (vpminit[size].pm_nc_factor > 0) size++;
    VPM *vpm = malloc(sizeof(VPM) * (size + 1));
    for (int i = 0; i < size; i++) {
        vpm[i].end = 0;
        vpm[i].pm_nc_factor = vpminit[i].pm_nc_factor;
        vpm[i].a_start = vpminit[i].a_start;
        vpm[i].pm = pm_new((int) (base_nmesh * vpm[i].pm_nc_factor), boxsize, comm);
        if (pm_unbalanced(vpm[i].pm)) fastpm_raise(-1, "PM mesh is not divided by the process mesh.\n");
    }
    vpm[size].end = 1; vpm[size].pm = NULL;
    return vpm;
}
VPM *vpm_find(VPM *vpm, double a)
{
    int i;
    for (i = 0; !vpm[i].end; i++) if (vpm[i].a_start > a) break;
    if (i == 0) i = 1;
    return &vpm[i - 1];
}
void vpm_free(VPM *vpm)
{
    for (int i = 0; !vpm[i].end; i++) pm_delete(vpm[i].pm);
    free(vpm);
}
