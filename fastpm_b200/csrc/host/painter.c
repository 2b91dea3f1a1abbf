/* fastpm_b200 host layer -- painter object (reference: libfastpm/painter.c:128-374, painter-cic.c).
 * The CIC window (the default, lua-runtime-fastpm.lua:132-138, and the one BASELINE.json names) has its own kernels; the linear,
 * quadratic and Lanczos windows and the derivative painters of fastpm_painter_init_diff share the generic ones (csrc/paint.cu).
 * The struct keeps the reference's public layout. */
#include "internal.h"
#include "../window.h"

static void no_host_paint(FastPMPainter *painter, FastPMFloat *canvas, double pos[3], double weight, int diffdir)
{ (void) painter; (void) canvas; (void) pos; (void) weight; (void) diffdir;
  fastpm_raise(-1, "painter->paint on a single particle: the canvas is device memory; use fastpm_paint_local\n"); }
static double no_host_readout(FastPMPainter *painter, FastPMFloat *canvas, double pos[3], int diffdir)
{ (void) painter; (void) canvas; (void) pos; (void) diffdir;
  fastpm_raise(-1, "painter->readout on a single particle: the canvas is device memory; use fastpm_readout_local\n"); return 0; }

/* the windows as host functions too (painter->kernel is public, painter.h:27); their addresses tell the painters apart */
static double linear_kernel(double x, double invh) { return fpm_window_linear(x, invh); }
static double quad_kernel(double x, double invh) { return fpm_window_quad(x, invh); }
static double lanczos_kernel(double x, double invh) { return fpm_window_lanczos(x, invh); }
static double linear_diff(double x, double invh) { return fpm_window_linear_diff(x, invh); }
static double quad_diff(double x, double invh) { return fpm_window_quad_diff(x, invh); }
static double lanczos_diff(double x, double invh) { return fpm_window_lanczos_diff(x, invh); }

int fpm_painter_window(const FastPMPainter *painter)
{
    if (painter->kernel == NULL) return FPM_WINDOW_CIC;
    if (painter->kernel == linear_kernel) return FPM_WINDOW_LINEAR;
    if (painter->kernel == quad_kernel) return FPM_WINDOW_QUAD;
    if (painter->kernel == lanczos_kernel) return FPM_WINDOW_LANCZOS;
    fastpm_raise(-1, "fastpm_b200: custom painter kernels cannot run on the device\n");
    return -1;
}

void fastpm_painter_init(FastPMPainter *painter, PM *pm, FastPMPainterType type, int support)
{
    /* painter.c:128-174 */
    painter->pm = pm;
    painter->paint = no_host_paint; painter->readout = no_host_readout;
    painter->kernel = NULL; painter->diff = NULL;
    switch (type) {
        case FASTPM_PAINTER_CIC: support = 2; break;
        case FASTPM_PAINTER_LINEAR: painter->kernel = linear_kernel; painter->diff = linear_diff; support = 2; break;
        case FASTPM_PAINTER_QUAD: painter->kernel = quad_kernel; painter->diff = quad_diff; support = 3; break;
        case FASTPM_PAINTER_LANCZOS: painter->kernel = lanczos_kernel; painter->diff = lanczos_diff; break;
        default: fastpm_raise(-1, "fastpm_b200: painter type %d\n", (int) type);
    }
    if (support < 1 || support > FPM_WINDOW_MAX_SUPPORT) fastpm_raise(-1, "fastpm_b200: painter support %d (1..%d on the device)\n", support, FPM_WINDOW_MAX_SUPPORT);
    painter->diffdir = -1;
    painter->support = support;
    painter->hsupport = 0.5 * support; painter->invh = 1 / (0.5 * support);
    painter->left = (support - 1) / 2;
    painter->Npoints = support * support * support;
    painter->shift = support % 2 == 0 ? 0 : 0.5;
}

/* painter.c:178-182 */
void fastpm_painter_init_diff(FastPMPainter *painter, FastPMPainter *base, int diffdir)
{
    *painter = *base;
    painter->diffdir = diffdir;
}

/* the generic kernels serve every window but plain CIC */
static int generic_path(const FastPMPainter *painter, int window) { return window != FPM_WINDOW_CIC || painter->diffdir >= 0; }

void fastpm_paint_local(FastPMPainter *painter, FastPMFloat *canvas, FastPMStore *p, size_t size, FastPMFieldDescr field)
{
    fpm_store_flush(p);
    const float *fcol = NULL; int fstride = 1;
    if (field.attribute) {
        int ci = fastpm_store_find_column_id(p, field.attribute);
        if (ci < 0 || !p->columns[ci] || p->_column_info[ci].membsize != 4) fastpm_raise(-1, "paint: field column must be an allocated float column\n");
        fcol = (const float *) p->columns[ci] + field.memb;
        fstride = (int) p->_column_info[ci].nmemb;
    }
    const int window = fpm_painter_window(painter);
    if (generic_path(painter, window)) {
        if (fpm_pending_wrap == p) { fpm_pending_wrap = NULL; fastpm_store_wrap(p, painter->pm->BoxSize); }
        FPM_MUST(fpm_paint_window_ex(painter->pm->mesh, window, painter->support, painter->diffdir, canvas, painter->pm->whalo, (const double *) p->x,
                                     (int64_t) size, p->meta.M0, p->mass, fcol, fstride));
        return;
    }
    if (fpm_pending_wrap == p && size == p->np) {
        fpm_pending_wrap = NULL;
        if (fpm_wrap_paint(painter->pm->mesh, canvas, (double *) p->x, (int64_t) size, p->meta.M0, p->mass, fcol, fstride) != 0)
            fastpm_raise(-1, "%s\n", fpm_last_error());
        return;
    }
    FPM_MUST(fpm_paint(painter->pm->mesh, canvas, (const double *) p->x, (int64_t) size, p->meta.M0, p->mass, fcol, fstride));
}

void fastpm_readout_local(FastPMPainter *painter, FastPMFloat *canvas, FastPMStore *p, size_t size, FastPMFieldDescr field)
{
    fpm_store_flush(p);
    int ci = fastpm_store_find_column_id(p, field.attribute);
    if (ci < 0 || !p->columns[ci] || p->_column_info[ci].from_double == NULL) fastpm_raise(-1, "readout: target column is not an allocated float column\n");
    float *out = (float *) p->columns[ci] + field.memb;
    const int window = fpm_painter_window(painter);
    if (generic_path(painter, window)) {
        FPM_MUST(fpm_readout_window_ex(painter->pm->mesh, window, painter->support, painter->diffdir, canvas, painter->pm->whalo, (const double *) p->x,
                                       (int64_t) size, out, (int) p->_column_info[ci].nmemb));
        return;
    }
    FPM_MUST(fpm_readout(painter->pm->mesh, canvas, (const double *) p->x, (int64_t) size, out, (int) p->_column_info[ci].nmemb, 1.0));
}

/* painter.c:342-356: clear + paint (the ghost exchange of the reference is the mesh-plane halo here) */
void fastpm_paint(FastPMPainter *painter, FastPMFloat *canvas, FastPMStore *p, FastPMFieldDescr field)
{
    PM *pm = painter->pm;
    pm_clear(pm, canvas);
    if (pm->NTask == 1) { fastpm_paint_local(painter, canvas, p, p->np, field); return; }
    /* several GPUs: what the particles of this slab deposit outside it goes to the neighbours -- the plane above the slab for CIC,
     * a block of planes on both sides for the wider windows (host/gravity.c does the same inside the force step) */
    const int generic = generic_path(painter, fpm_painter_window(painter));
    FastPMFloat *halo = NULL, *saved = pm->whalo;
    if (generic) {
        int whl = 0, whr = 0;
        FPM_MUST(fpm_window_halo_planes(fpm_painter_window(painter), painter->support, &whl, &whr));
        const size_t hbytes = sizeof(FastPMFloat) * (size_t) (whl + whr) * pm->Nmesh[1] * pm->pitch_r;
        halo = fastpm_memory_alloc(pm->mem, "window halo planes", hbytes, FASTPM_MEMORY_HEAP);
        FPM_MUST(fpm_memset(halo, 0, hbytes));
        pm->whalo = halo; pm->whl = whl; pm->whr = whr;
    }
    fastpm_paint_local(painter, canvas, p, p->np, field);
    fpm_halo_add(pm, canvas);
    if (halo) { pm->whalo = saved; fastpm_memory_free(pm->mem, halo); }
}
