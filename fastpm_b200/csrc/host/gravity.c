/* fastpm_b200 host layer -- the PM force pipeline (reference: libfastpm/gravity.c:273-529).
 *
 *   paint -> [halo add] -> r2c (mean-density normalisation and 1/Norm folded into the first FFT pass)
 *   for each of ACC_x, ACC_y, ACC_z [, POTENTIAL]:
 *       c2r with the Green's function and i*k_d gradient fused into its first pass -> [halo fetch] -> readout
 *
 * against the reference's: ghosts, clear, paint(p)+paint(ghosts), multiply sweep, 2 check sweeps, r2c,
 * 1/Norm sweep, then per component 3 transfer sweeps + 2 check sweeps + c2r + 2 readouts, ghost reduce.
 * The particle ghosts of pmghosts.c are replaced by one mesh plane (support 2 needs planes [x0, x0+nxl]).
 */
#include "internal.h"
#include <math.h>

void fastpm_kernel_type_get_orders(FastPMKernelType type, int *potorder, int *gradorder, int *difforder, int *deconvolveorder)
{
    fpm_transfer t;
    if (fpm_transfer_for_kernel((int) type, 0, 0, &t) != 0) fastpm_raise(-1, "Wrong kernel type\n");
    *potorder = t.potorder;
    *gradorder = t.gradorder;
    *difforder = (type == FASTPM_KERNEL_1_4_DIFF0) ? 0 : 1;
    *deconvolveorder = (type == FASTPM_KERNEL_EASTWOOD || type == FASTPM_KERNEL_GADGET) ? 2 : 0;
}

/* stand-alone version of the fused kernel: canvas = kernel(delta_k) for one field component (gravity.c:172-241) */
void gravity_apply_kernel_transfer(FastPMKernelType type, PM *pm, FastPMFloat *delta_k, FastPMFloat *canvas, FastPMFieldDescr field)
{
    fpm_transfer t;
    if (field.attribute == COLUMN_DENSITY) {               /* fastpm_apply_multiply_transfer(pm, delta_k, canvas, 1.0) */
        fastpm_apply_multiply_transfer(pm, delta_k, canvas, 1.0);
        return;
    }
    if (field.attribute == COLUMN_TIDAL) {
        /* potential, then the gradient along d1 and along d2: memb 0..5 = xx, yy, zz, xy, yz, zx (gravity.c:194-231) */
        static const int D1[6] = { 0, 1, 2, 0, 1, 2 }, D2[6] = { 0, 1, 2, 1, 2, 0 };
        if (field.memb < 0 || field.memb > 5) fastpm_raise(-1, "tidal component %d\n", (int) field.memb);
        if (fpm_transfer_for_kernel((int) type, 0, D1[field.memb], &t) != 0) fastpm_raise(-1, "%s\n", fpm_last_error());
        t.ngrad = 2;
        t.graddir[1] = D2[field.memb];
        FPM_MUST(fpm_apply_transfer(pm->mesh, delta_k, canvas, &t));
        return;
    }
    int attr;
    if (field.attribute == COLUMN_ACC) attr = 0;
    else if (field.attribute == COLUMN_POTENTIAL) attr = 1;
    else { fastpm_raise(-1, "Unknown type for gravity attribute\n"); return; }
    if (fpm_transfer_for_kernel((int) type, attr, field.memb, &t) != 0) fastpm_raise(-1, "%s\n", fpm_last_error());
    FPM_MUST(fpm_apply_transfer(pm->mesh, delta_k, canvas, &t));
}


size_t fastpm_b200_arena_size(void);             /* host/support.c */
size_t fastpm_b200_arena_largest_free(void);
int fastpm_b200_device_room(size_t need, size_t slack);

/* gravity.c:66-102: exp(-(k_d r0)^2 / 2) per axis with r0 = N cells, tabulated in double from the float k table */
static void apply_gaussian_softening(PM *pm, FastPMFloat *from, FastPMFloat *to, double N)
{
    const int n = (int) pm->Nmesh[0];
    const double r0 = N * pm->BoxSize[0] / pm->Nmesh[0];
    float *tab = malloc(sizeof(float) * 5 * n);
    double *f = malloc(sizeof(double) * n);
    FPM_MUST(fpm_mesh_ktables_host(pm->mesh, tab));
    for (int i = 0; i < n; i++) f[i] = exp(-0.5 * pow(tab[i] * r0, 2));
    /* the reference multiplies `to` (gravity.c:93-94); its only caller passes from == to */
    FPM_MUST(fpm_apply_axis_factors(pm->mesh, to, to, f));
    (void) from;
    free(f); free(tab);
}

/* gravity.c:244-270 */
static void apply_softening_transfer(FastPMSofteningType type, PM *pm, FastPMFloat *from, FastPMFloat *to)
{
    const double k_nq = M_PI / pm->BoxSize[0] * pm->Nmesh[0];
    switch (type) {
        case FASTPM_SOFTENING_TWO_THIRD: fastpm_apply_lowpass_transfer(pm, from, to, 2.0 / 3 * k_nq); break;
        case FASTPM_SOFTENING_GAUSSIAN: apply_gaussian_softening(pm, from, to, 1.0); break;
        case FASTPM_SOFTENING_GADGET_LONG_RANGE: apply_gaussian_softening(pm, from, to, pow(2, 0.5) * 1.25); break;
        case FASTPM_SOFTENING_GAUSSIAN36: FPM_MUST(fpm_apply_radial(pm->mesh, from, to, 1, k_nq)); break;
        case FASTPM_SOFTENING_NONE: break;
        default: fastpm_raise(-1, "wrong softening kernel type");
    }
}

/* the pipelined inverse transforms of several GPUs followed by one gather for the three components: default when CDM is the only
 * species, the window is CIC and two more canvases fit (every rank sees the same arena state: the same answer everywhere);
 * FASTPM_B200_FUSED_READOUT=0 keeps the component-by-component gather */
static int pipelined_fused_ok(FastPMSolver *fastpm, PM *pm, FastPMPainter *painter)
{
    static int off = -1;
    if (off < 0) { const char *e = getenv("FASTPM_B200_FUSED_READOUT"); off = (e && atoi(e) == 0) ? 1 : 0; }
    if (off || painter->kernel != NULL || painter->diffdir >= 0) return 0;
    FastPMStore *cdm = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM);
    if (!cdm || !cdm->acc) return 0;                /* nothing rank-local in here: the transforms are collective */
    for (int si = 0; si < FASTPM_SOLVER_NSPECIES; si++)
        if (si != FASTPM_SPECIES_CDM && fastpm_solver_get_species(fastpm, si)) return 0;
    const size_t need = 2 * sizeof(FastPMFloat) * pm->allocsize;
    if (pm->mem->used_bytes + need > pm->mem->total_bytes) return 0;
    return fastpm_b200_arena_largest_free() >= need + (need >> 3);
}

/* opt-in; only when CDM is the only species and two more meshes fit beside what is allocated now */
static int fused_readout_wanted(FastPMSolver *fastpm, PM *pm)
{
    static int want = -1;
    if (want < 0) { const char *e = getenv("FASTPM_B200_FUSED_READOUT"); want = (e && atoi(e) > 0) ? 1 : 0; }
    if (!want) return 0;
    FastPMStore *cdm = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM);
    if (!cdm || !cdm->acc) return 0;
    for (int si = 0; si < FASTPM_SOLVER_NSPECIES; si++)
        if (si != FASTPM_SPECIES_CDM && fastpm_solver_get_species(fastpm, si)) return 0;
    const size_t need = 2 * sizeof(FastPMFloat) * pm->allocsize;
    if (pm->mem->used_bytes + need > pm->mem->total_bytes) return 0;
    int ok;
    if (fastpm_b200_arena_size() > 0) {
        ok = fastpm_b200_arena_largest_free() >= need + (need >> 3);
    } else {
        size_t fr = 0, tot = 0;
        ok = fpm_device_mem_info(&fr, &tot) == 0 && fr >= need + (need >> 3);
    }
    /* every rank must take the same branch: the transforms are collective */
    double all = ok ? 1.0 : 0.0;
    if (pm->NTask > 1) fpm_comm_allreduce_double(fastpm->comm, &all, 1, 1);
    return all > 0.5;
}

/* Readout of force component d of every species.  With the CIC window and a scratch block `planes` (2 x np_upper floats, CDM only)
 * components 0 and 1 go to planar arrays and component 2 writes the ACC rows whole (fpm_readout_pack3): same values, 28 instead
 * of 72 bytes of DRAM traffic per particle for the results. */
static void readout_component(FastPMSolver *fastpm, FastPMPainter *painter, FastPMFloat *canvas, int d, float *planes)
{
    for (int si = 0; si < FASTPM_SOLVER_NSPECIES; si++) {
        FastPMStore *p = fastpm_solver_get_species(fastpm, si);
        if (!p) continue;
        FastPMFieldDescr f = { d < 3 ? COLUMN_ACC : COLUMN_POTENTIAL, d < 3 ? d : 0 };
        if (d == 3 && !p->potential) continue;
        if (planes && si == FASTPM_SPECIES_CDM && d < 3) {
            fpm_store_flush(p);
            float *t0 = planes, *t1 = planes + p->np_upper;
            if (d < 2) FPM_MUST(fpm_readout(painter->pm->mesh, canvas, (const double *) p->x, (int64_t) p->np, d == 0 ? t0 : t1, 1, 1.0));
            else FPM_MUST(fpm_readout_pack3(painter->pm->mesh, canvas, (const double *) p->x, (int64_t) p->np, t0, t1, (float *) p->acc));
            continue;
        }
        fastpm_readout_local(painter, canvas, p, p->np, f);
    }
}

/* the scratch block for readout_component, or NULL: CIC window, CDM with an ACC column, and room for it (every rank decides alike:
 * all ranks see the same arena state; FASTPM_B200_NO_PACKED_READOUT=1 switches it off) */
static float *acc_planes_alloc(FastPMSolver *fastpm, PM *pm, FastPMPainter *painter, size_t reserve)
{
    static int off = -1;
    if (off < 0) off = getenv("FASTPM_B200_NO_PACKED_READOUT") ? 1 : 0;
    FastPMStore *cdm = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM);
    if (off || painter->kernel != NULL || painter->diffdir >= 0 || !cdm || !cdm->acc || cdm->np == 0) return NULL;
    const size_t need = 2 * sizeof(float) * cdm->np_upper;
    (void) pm;
    if (cdm->mem->used_bytes + need > cdm->mem->total_bytes) return NULL;
    /* `reserve`: what is still to be taken after this block (the second canvas of the pipelined transforms) */
    if (!fastpm_b200_device_room(need, reserve + (need >> 2) + ((size_t) 256 << 20))) return NULL;
    return fastpm_memory_alloc(cdm->mem, "ACC component planes", need, FASTPM_MEMORY_STACK);
}

void fastpm_solver_compute_force(FastPMSolver *fastpm, PM *pm, FastPMPainter *painter, FastPMSofteningType dealias,
                                 FastPMKernelType kernel, FastPMFloat *delta_k, double Time)
{
    (void) Time;
    fpm_store_flush(NULL);
    if (fastpm->cosmology->ncdm_linearresponse) fastpm_raise(-1, "fastpm_b200: ncdm linear response is out of scope\n");
    CLOCK(paint);
    LEAVE(paint);
    CLOCK(r2c);
    LEAVE(r2c);
    CLOCK(c2r);
    LEAVE(c2r);
    CLOCK(readout);
    LEAVE(readout);

    FastPMFloat *canvas = pm_alloc_noclear(pm, __FILE__, __LINE__);
    /* Several GPUs and a window that reaches further than CIC's one plane above the slab: the planes outside the slab are kept
     * in a halo block that fpm_halo_add / fpm_halo_fetch exchange with the neighbours (the reference: ghost particles,
     * pmghosts.c:45-78, which handle any support). */
    int whl = 0, whr = 0;
    FPM_MUST(fpm_window_halo_planes(fpm_painter_window(painter), painter->support, &whl, &whr));
    if (pm->NTask > 1 && (painter->kernel != NULL || painter->diffdir >= 0)) {
        const size_t hbytes = sizeof(FastPMFloat) * (size_t) (whl + whr) * pm->Nmesh[1] * pm->pitch_r;
        pm->whalo = fastpm_memory_alloc(pm->mem, "window halo planes", hbytes, FASTPM_MEMORY_HEAP);
        pm->whl = whl; pm->whr = whr;
        FPM_MUST(fpm_memset(pm->whalo, 0, hbytes));
    }

    /* ---- density: gravity.c:305-353 */
    ENTER(paint);
    pm_clear(pm, canvas);
    double total_mass = 0;
    for (int si = 0; si < FASTPM_SOLVER_NSPECIES; si++) {
        FastPMStore *p = fastpm_solver_get_species(fastpm, si);
        if (!p) continue;
        if (p->mass) {
            double s[4];
            FPM_MUST(fpm_summary(p->mass, 4, 1, (int64_t) p->np, s));
            total_mass += p->meta.M0 * p->np + s[2];
        } else {
            total_mass += p->meta.M0 * p->np;         /* sum of fastpm_store_get_mass, gravity.c:330-334 */
        }
        FastPMFieldDescr none = { 0, 0 };
        fastpm_paint_local(painter, canvas, p, p->np, none);
    }
    if (pm->NTask > 1) fpm_halo_add(pm, canvas);
    LEAVE(paint);
    fpm_comm_allreduce_double(fastpm->comm, &total_mass, 1, 0);
    const double mean_mass_per_cell = total_mass / pm->Norm;

    ENTER(r2c);
    /* canvas * (1/mean) (gravity.c:345) and the 1/Norm of pm_r2c (pmpfft.c:382) as one factor on the FFT input */
    const double scale = (1.0 / mean_mass_per_cell) * (1.0 / pm->Norm);
    fpm_mesh_r2c(pm, canvas, delta_k, scale);
    LEAVE(r2c);

    /* ---- dealiasing of the source, in place: gravity.c:476 (the FORCE/after event sees the softened field, like the reference's) */
    if (dealias != FASTPM_SOFTENING_NONE) apply_softening_transfer(dealias, pm, delta_k, delta_k);

    /* ---- force components: gravity.c:359-396 */
    FastPMStore *cdm = fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM);
    const int nacc = (cdm && cdm->potential) ? 4 : 3;
    int d0 = 0;
    const int pipeline = pm->NTask > 1 && fpm_dist_pipeline_ready(pm);
    if (!pipeline && painter->kernel == NULL && painter->diffdir < 0 /* CIC */ && fused_readout_wanted(fastpm, pm)) {
        /* FASTPM_B200_FUSED_READOUT=1: the three inverse transforms into three meshes, then ONE pass over the particles
         * (positions read once instead of three times, ACC written as whole elements). Same values bit for bit. */
        FastPMFloat *cv[3] = { canvas, pm_alloc_noclear(pm, __FILE__, __LINE__), pm_alloc_noclear(pm, __FILE__, __LINE__) };
        ENTER(c2r);
        for (int d = 0; d < 3; d++) {
            fpm_transfer t;
            if (fpm_transfer_for_kernel((int) kernel, 0, d, &t) != 0) fastpm_raise(-1, "%s\n", fpm_last_error());
            fpm_mesh_c2r(pm, delta_k, cv[d], &t);
            if (pm->NTask > 1) fpm_halo_fetch(pm, cv[d]);
        }
        LEAVE(c2r);
        ENTER(readout);
        fpm_store_flush(cdm);
        FPM_MUST(fpm_readout3(pm->mesh, cv[0], cv[1], cv[2], (const double *) cdm->x, (int64_t) cdm->np, (float *) cdm->acc));
        LEAVE(readout);
        pm_free(pm, cv[2]);
        pm_free(pm, cv[1]);
        d0 = 3;
    }
    const size_t canvas_bytes = sizeof(FastPMFloat) * pm->allocsize;
    const int fuse3 = pipeline && nacc == 3 && pipelined_fused_ok(fastpm, pm, painter);
    /* freed last (LIFO): allocated before the second canvas, for which room is kept */
    float *planes = (d0 == 0 && !fuse3) ? acc_planes_alloc(fastpm, pm, painter, pipeline ? canvas_bytes + (canvas_bytes >> 3) : 0) : NULL;
    if (fuse3) {
        /* Several GPUs with room for three canvases: the pipelined transforms below, and then ONE pass over the particles for
         * the three components (cic_readout3_kernel: positions read once, ACC rows written whole; the same values bit for bit).
         * Measured on 2 B200s at nc = 1024: the gather takes 24.7 ms per step instead of 3 x 11.1. */
        FastPMFloat *cv[3] = { canvas, pm_alloc_noclear(pm, __FILE__, __LINE__), pm_alloc_noclear(pm, __FILE__, __LINE__) };
        fpm_transfer t;
        ENTER(c2r);
        if (fpm_transfer_for_kernel((int) kernel, 0, 0, &t) != 0) fastpm_raise(-1, "%s\n", fpm_last_error());
        fpm_dist_c2r_begin(pm, delta_k, cv[0], &t, 0);
        for (int d = 0; d < 3; d++) {
            if (d + 1 < 3) {
                if (fpm_transfer_for_kernel((int) kernel, 0, d + 1, &t) != 0) fastpm_raise(-1, "%s\n", fpm_last_error());
                fpm_dist_c2r_begin(pm, delta_k, cv[d + 1], &t, (d + 1) & 1);
            }
            fpm_dist_c2r_finish(pm, cv[d], d & 1);
            fpm_halo_fetch(pm, cv[d]);
        }
        LEAVE(c2r);
        ENTER(readout);
        fpm_store_flush(cdm);
        FPM_MUST(fpm_readout3(pm->mesh, cv[0], cv[1], cv[2], (const double *) cdm->x, (int64_t) cdm->np, (float *) cdm->acc));
        LEAVE(readout);
        pm_free(pm, cv[2]);
        pm_free(pm, cv[1]);
        d0 = nacc;
    } else if (pipeline && d0 == 0) {
        /* Several GPUs: the slab transpose of component d + 1 travels (copy engines, NVLink) while component d is finished
         * (y- and z-pass) and read out -- two canvases and two staging meshes, used alternately.  The arithmetic of every
         * component is that of fpm_mesh_c2r + readout; only the order in which the work is queued differs. */
        FastPMFloat *cv[2] = { canvas, pm_alloc_noclear(pm, __FILE__, __LINE__) };
        fpm_transfer t;
        ENTER(c2r);
        if (fpm_transfer_for_kernel((int) kernel, 0, 0, &t) != 0) fastpm_raise(-1, "%s\n", fpm_last_error());
        fpm_dist_c2r_begin(pm, delta_k, cv[0], &t, 0);
        LEAVE(c2r);
        for (int d = 0; d < nacc; d++) {
            ENTER(c2r);
            if (d + 1 < nacc) {
                const int e = d + 1;
                if (fpm_transfer_for_kernel((int) kernel, e < 3 ? 0 : 1, e < 3 ? e : 0, &t) != 0) fastpm_raise(-1, "%s\n", fpm_last_error());
                fpm_dist_c2r_begin(pm, delta_k, cv[e & 1], &t, e & 1);
            }
            fpm_dist_c2r_finish(pm, cv[d & 1], d & 1);
            fpm_halo_fetch(pm, cv[d & 1]);
            LEAVE(c2r);
            ENTER(readout);
            readout_component(fastpm, painter, cv[d & 1], d, planes);
            LEAVE(readout);
        }
        pm_free(pm, cv[1]);
        d0 = nacc;
    }
    for (int d = d0; d < nacc; d++) {
        fpm_transfer t;
        if (fpm_transfer_for_kernel((int) kernel, d < 3 ? 0 : 1, d < 3 ? d : 0, &t) != 0) fastpm_raise(-1, "%s\n", fpm_last_error());
        ENTER(c2r);
        fpm_mesh_c2r(pm, delta_k, canvas, &t);
        if (pm->NTask > 1) fpm_halo_fetch(pm, canvas);
        LEAVE(c2r);
        ENTER(readout);
        readout_component(fastpm, painter, canvas, d, planes);
        LEAVE(readout);
    }
    if (planes) fastpm_memory_free(fastpm_solver_get_species(fastpm, FASTPM_SPECIES_CDM)->mem, planes);
    if (pm->whalo) { fastpm_memory_free(pm->mem, pm->whalo); pm->whalo = NULL; }
    pm_free(pm, canvas);
}
