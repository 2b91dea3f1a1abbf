/* fastpm_b200 host layer -- the small remaining entry points of the headers the force step lives in (solver.h, store.h,
 * transfer.h, string.h, io.h), so that a caller written against libfastpm links without stubs.  None of them launches a kernel
 * of its own: they work on meta data, on single k-space modes (8-byte copies) or call the sweeps that exist. */
#define _GNU_SOURCE
#include "internal.h"
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <sys/stat.h>

/* ------------------------------------------------------------------ solver.h: every species at once, solver.c:604-646 */
void fastpm_set_snapshot(FastPMSolver *fastpm, FastPMSolver *snapshot, FastPMDriftFactor *drift, FastPMKickFactor *kick, double aout)
{
    for (int si = 0; si < FASTPM_SOLVER_NSPECIES; si++) {
        FastPMStore *p = fastpm_solver_get_species(fastpm, si), *po = fastpm_solver_get_species(snapshot, si);
        if (!p || !po) continue;
        fastpm_set_species_snapshot(fastpm, p, drift, kick, po, aout);
    }
}

void fastpm_unset_snapshot(FastPMSolver *fastpm, FastPMSolver *snapshot, FastPMDriftFactor *drift, FastPMKickFactor *kick, double aout)
{
    for (int si = 0; si < FASTPM_SOLVER_NSPECIES; si++) {
        FastPMStore *p = fastpm_solver_get_species(fastpm, si), *po = fastpm_solver_get_species(snapshot, si);
        if (!p || !po) continue;
        fastpm_unset_species_snapshot(fastpm, p, drift, kick, po, aout);
    }
}

/* ------------------------------------------------------------------ store.h: meta data only */
void fastpm_store_set_name(FastPMStore *p, const char *name) { strncpy(p->name, name, 31); p->name[31] = 0; }      /* store.c:258 */

int fastpm_store_has_q(FastPMStore *p) { return p->meta._q_size != 0; }                                          /* store.c:659 */

void fastpm_store_get_q_from_id(FastPMStore *p, uint64_t id, double q[3])                                         /* store.c:664-680 */
{
    ptrdiff_t pabs[3];
    id = id % p->meta._q_size;
    for (int d = 0; d < 3; d++) { pabs[d] = id / p->meta._q_strides[d]; id -= pabs[d] * p->meta._q_strides[d]; }
    for (int d = 0; d < 3; d++) { q[d] = pabs[d] * p->meta._q_scale[d]; q[d] += p->meta._q_shift[d]; }
}

void fastpm_store_get_iq_from_id(FastPMStore *p, uint64_t id, ptrdiff_t pabs[3])                                  /* store.c:682-692 */
{
    for (int d = 0; d < 3; d++) { pabs[d] = id / p->meta._q_strides[d]; id -= pabs[d] * p->meta._q_strides[d]; }
}

void fastpm_store_steal(FastPMStore *p, FastPMStore *po, FastPMColumnTags attributes)                             /* store.c:911-921 */
{
    fpm_store_flush(NULL);
    for (int c = 0; c < 32; c++) {
        if (!(p->_column_info[c].attribute & attributes)) continue;
        po->columns[c] = p->columns[c];
    }
    po->np = p->np;
    po->meta = p->meta;
}

/* ------------------------------------------------------------------ transfer.h: single modes
 * k-space is [ky_local][kx][pitch_c] complex on the device; a mode is stored when kz <= N/2 and ky is in this rank's slab. */
static ptrdiff_t mode_offset(PM *pm, ptrdiff_t ix, ptrdiff_t iy, ptrdiff_t iz)
{
    const ptrdiff_t n = pm->Nmesh[0];
    if (ix < 0 || ix >= n || iy < 0 || iy >= n || iz < 0 || iz > n / 2) return -1;
    if (iy < pm->y0 || iy >= pm->y0 + pm->nyl) return -1;
    return 2 * (((iy - pm->y0) * n + ix) * (ptrdiff_t) pm->pitch_c + iz);
}

/* transfer.c:340-366 */
double fastpm_apply_get_mode_transfer(PM *pm, FastPMFloat *from, ptrdiff_t *mode)
{
    double result = 0.0;
    const ptrdiff_t off = mode_offset(pm, mode[0], mode[1], mode[2]);
    if (off >= 0) {
        float v = 0;
        FPM_MUST(fpm_memcpy_d2h(&v, from + off + mode[3], sizeof(float)));
        result = v;
    }
    fpm_comm_allreduce_double(pm->comm, &result, 1, 0);
    return result;
}

/* transfer.c:290-337: component mode[3] of the mode and of its conjugate, method 0 = set, else add */
void fastpm_apply_set_mode_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to, ptrdiff_t *mode, double value, int method)
{
    const ptrdiff_t n = pm->Nmesh[0];
    const ptrdiff_t cj[3] = { (n - mode[0]) % n, (n - mode[1]) % n, (n - mode[2]) % n };
    if (mode[0] == cj[0] && mode[1] == cj[1] && mode[2] == cj[2] && mode[3] == 1) { method = 0; value = 0; }   /* purely real modes */
    if (from != to) pm_assign(pm, from, to);
    const ptrdiff_t off[2] = { mode_offset(pm, mode[0], mode[1], mode[2]), mode_offset(pm, cj[0], cj[1], cj[2]) };
    const double val[2] = { value, value * ((mode[3] == 0) ? 1 : -1) };
    for (int i = 0; i < 2; i++) {
        if (off[i] < 0) continue;
        float v = 0;
        if (method != 0) FPM_MUST(fpm_memcpy_d2h(&v, to + off[i] + mode[3], sizeof(float)));
        v = method == 0 ? (float) val[i] : (float) (v + val[i]);          /* float = double / float += double, like the reference's */
        FPM_MUST(fpm_memcpy_h2d(to + off[i] + mode[3], &v, sizeof(float)));
    }
}

/* transfer.c:223-247: divide by the real part of the DC mode */
void fastpm_apply_normalize_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to)
{
    ptrdiff_t dc[4] = { 0, 0, 0, 0 };
    const double Norm = fastpm_apply_get_mode_transfer(pm, from, dc);
    if (Norm == 0) fastpm_raise(-1, "It makes no sense to normalize a field with a mean of zero.");
    fastpm_apply_multiply_transfer(pm, from, to, 1 / Norm);
}

/* transfer.c:249-277: weight 2 for every stored mode except the (at most 8) self-conjugate ones, which keep weight 1 */
void fastpm_apply_c2r_weight_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to)
{
    const ptrdiff_t n = pm->Nmesh[0], h = n / 2;
    float keep[8][2];
    ptrdiff_t off[8];
    int cnt = 0;
    for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) for (int c = 0; c < 2; c++) {
        const ptrdiff_t o = mode_offset(pm, a ? h : 0, b ? h : 0, c ? h : 0);
        if (o < 0 || (n % 2 != 0 && (a || b || c))) continue;
        int dup = 0;
        for (int i = 0; i < cnt; i++) if (off[i] == o) dup = 1;
        if (dup) continue;
        FPM_MUST(fpm_memcpy_d2h(keep[cnt], from + o, 2 * sizeof(float)));
        off[cnt++] = o;
    }
    fastpm_apply_multiply_transfer(pm, from, to, 2.0);
    for (int i = 0; i < cnt; i++) FPM_MUST(fpm_memcpy_h2d(to + off[i], keep[i], 2 * sizeof(float)));
}

/* ------------------------------------------------------------------ string.h */
char *fastpm_file_get_content(const char *filename)
{
    FILE *fp = fopen(filename, "r");
    if (!fp) return NULL;
    fseek(fp, 0, SEEK_END);
    const long len = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    char *buf = malloc((size_t) len + 1);
    if (len < 0 || fread(buf, 1, (size_t) len, fp) != (size_t) len) { fclose(fp); free(buf); return NULL; }
    fclose(fp);
    buf[len] = 0;
    return buf;
}

/* one allocation: the NULL-terminated pointer array, then the characters (free() the result) */
char **fastpm_strsplit(const char *str, const char *split)
{
    size_t nparts = 1;
    for (const char *p = str; *p; p++) if (strchr(split, *p)) nparts++;
    char **out = malloc((nparts + 1) * sizeof(char *) + strlen(str) + 1);
    char *text = (char *) (out + nparts + 1);
    strcpy(text, str);
    size_t i = 0;
    out[i++] = text;
    for (char *p = text; *p; p++) if (strchr(split, *p)) { *p = 0; out[i++] = p + 1; }
    out[i] = NULL;
    return out;
}

char *fastpm_strdup(const char *str) { char *d = malloc(strlen(str) + 1); strcpy(d, str); return d; }

char *fastpm_strdup_vprintf(const char *fmt, va_list va)
{
    va_list va2;
    va_copy(va2, va);
    const int n = vsnprintf(NULL, 0, fmt, va);
    char *buf = malloc((size_t) (n < 0 ? 0 : n) + 1);
    vsnprintf(buf, (size_t) (n < 0 ? 0 : n) + 1, fmt, va2);
    va_end(va2);
    return buf;
}

char *fastpm_strdup_printf(const char *fmt, ...)
{
    va_list va;
    va_start(va, fmt);
    char *buf = fastpm_strdup_vprintf(fmt, va);
    va_end(va);
    return buf;
}

/* creates the directories leading to `path` (everything before its last '/') */
void fastpm_path_ensure_dirname(const char *path)
{
    char *dup = fastpm_strdup(path);
    char *slash = strrchr(dup, '/');
    if (slash) {
        *slash = 0;
        for (char *p = dup + 1; *p; p++) if (*p == '/') { *p = 0; mkdir(dup, 0777); *p = '/'; }
        if (*dup) mkdir(dup, 0777);
    }
    free(dup);
}

/* ------------------------------------------------------------------ io.h: a P(k)-like text table, read on every rank */
int read_funck(FastPMFuncK *fk, const char filename[], MPI_Comm comm)
{
    (void) comm;                   /* one node, shared file system: every rank reads the file itself */
    char *content = fastpm_file_get_content(filename);
    if (!content) fastpm_raise(-1, "Failed to read file %s\n", filename);
    if (0 != fastpm_funck_init_from_string(fk, content)) fastpm_raise(-1, "Failed to parse file %s\n", filename);
    free(content);
    return 0;
}

/* ------------------------------------------------------------------ the public mesh calls in one chain (bindings, tests)
 * fastpm_paint (CIC) of the CDM store -> pm_r2c -> fastpm_powerspectrum_init_from_delta -> pm_c2r -> fastpm_readout_local into
 * ACC[:, 0]: what a user's own density / P(k) code does with libfastpm, on one GPU or on the slabs of several.  k, p, nmodes hold
 * Nmesh / 2 bins; dens_host receives the painted density read back at this rank's particles (np values).  Returns np. */
int64_t fastpm_b200_public_mesh_probe(FastPMSolver *fastpm, double a, double *k, double *p, double *nmodes, float *dens_host)
{
    PM *pm = fastpm_find_pm(fastpm, a);
    FastPMStore *cdm = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM);
    FastPMPainter painter[1];
    fastpm_painter_init(painter, pm, FASTPM_PAINTER_CIC, 2);
    FastPMFloat *canvas = pm_alloc(pm), *delta_k = pm_alloc(pm);
    FastPMFieldDescr none = { 0, 0 }, acc0 = { COLUMN_ACC, 0 };
    fastpm_paint(painter, canvas, cdm, none);
    pm_r2c(pm, canvas, delta_k);
    FastPMPowerSpectrum ps;
    fastpm_powerspectrum_init_from_delta(&ps, pm, delta_k, delta_k);
    for (size_t i = 0; i < ps.base.size; i++) { k[i] = ps.base.k[i]; p[i] = ps.base.f[i]; nmodes[i] = ps.Nmodes[i]; }
    fastpm_powerspectrum_destroy(&ps);
    pm_c2r(pm, delta_k);
    fastpm_readout_local(painter, delta_k, cdm, cdm->np, acc0);
    float *tmp = malloc(sizeof(float) * 3 * (cdm->np ? cdm->np : 1));
    FPM_MUST(fpm_memcpy_d2h(tmp, cdm->acc, sizeof(float) * 3 * cdm->np));
    for (size_t i = 0; i < cdm->np; i++) dens_host[i] = tmp[3 * i];
    free(tmp);
    pm_free(pm, delta_k);
    pm_free(pm, canvas);
    return (int64_t) cdm->np;
}

/* ------------------------------------------------------------------ analytic spectra for fastpm_ic_induce_correlation (utils.h:3-14)
 * fastpm_utils_powerspec_eh, utils.c:118-149: P(k) = Norm k T(k)^2 with the zero-baryon-wiggle Eisenstein & Hu (1998) transfer
 * function in the form the reference took from Martin White -- what tests/testpm.c:67-74 hands to fastpm_ic_induce_correlation. */
double fastpm_utils_powerspec_eh(double k, struct fastpm_powerspec_eh_params *param)
{
    const double h = param->hubble_param, om_h2 = param->omegam * h * h, ob_h2 = param->omegab * h * h;
    const double theta_cmb = 2.728 / 2.7, fb = ob_h2 / om_h2;
    const double sound = 44.5 * log(9.83 / om_h2) / sqrt(1. + 10. * exp(0.75 * log(ob_h2))) * h;
    const double alpha = 1. - 0.328 * log(431. * om_h2) * fb + 0.380 * log(22.3 * om_h2) * fb * fb;
    double shape = alpha + (1. - alpha) / (1. + exp(4 * log(0.43 * k * sound)));
    shape *= param->omegam * h;
    const double q = k * theta_cmb * theta_cmb / shape;
    const double L0 = log(2. * exp(1.) + 1.8 * q), C0 = 14.2 + 731. / (1. + 62.5 * q);
    const double tk = L0 / (L0 + C0 * q * q);
    return param->Norm * k * pow(tk, 2);
}

/* utils.c:151-155 */
double fastpm_utils_powerspec_white(double k, double *amplitude) { (void) k; return *amplitude; }

