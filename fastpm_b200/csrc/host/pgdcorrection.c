/* fastpm_b200 host layer -- the PGD ("potential gradient descent") correction, reference: libfastpm/pgdcorrection.c.
 *
 * After the force of a step, the particles get one more displacement field, the gradient of a band-pass filtered potential:
 *     pgdc_d = readout( c2r( i k_d[finite] * alpha(a) exp(-kl^2/k^2 - k^4/ks^4) / k^2 * delta_k ) )
 * which every later drift adds to the positions (factors.c:108-113, host/factors.c here).
 *
 * The reference sweeps the filter three times (once per component) and differentiates in place; here the filtered potential
 * is computed ONCE into a third mesh and the three gradients are fused into the first pass of the three inverse transforms,
 * like the force components (host/gravity.c).  The values are the same: each step rounds to float where the reference does.
 * The reference's particle ghosts are the one mesh halo plane of this build.
 */
#include "internal.h"
#include <math.h>

double fastpm_pgdc_get_alpha(FastPMPGDCorrection *pgdc, double a) { return pgdc->alpha0 * pow(10, pgdc->A * a * a - pgdc->B * a); }
double fastpm_pgdc_get_ks(FastPMPGDCorrection *pgdc, double a) { (void) a; return pgdc->ks; }
double fastpm_pgdc_get_kl(FastPMPGDCorrection *pgdc, double a) { (void) a; return pgdc->kl; }

void fastpm_pgdc_calculate(FastPMPGDCorrection *pgdc, PM *pm, FastPMStore *p, FastPMFloat *delta_k, double a, double fac)
{
    if (!p->pgdc) fastpm_raise(-1, "fastpm_pgdc_calculate: the store has no pgdc column (COLUMN_PGDC)\n");
    if (pgdc->PainterType != FASTPM_PAINTER_CIC || (pgdc->PainterSupport != 2 && pgdc->PainterSupport != 0))
        fastpm_raise(-1, "fastpm_b200: only the CIC window (support 2) is implemented for the PGD readout\n");
    fpm_store_flush(p);
    CLOCK(transfer);
    LEAVE(transfer);
    CLOCK(c2r);
    LEAVE(c2r);
    CLOCK(readout);
    LEAVE(readout);

    const double kl = fastpm_pgdc_get_kl(pgdc, a), ks = fastpm_pgdc_get_ks(pgdc, a);
    const double alpha = fastpm_pgdc_get_alpha(pgdc, a) * fac;

    FastPMFloat *pot_k = pm_alloc_noclear(pm, __FILE__, __LINE__);
    FastPMFloat *canvas = pm_alloc_noclear(pm, __FILE__, __LINE__);
    ENTER(transfer);
    FPM_MUST(fpm_apply_pgd_transfer(pm->mesh, delta_k, pot_k, alpha, kl, ks));       /* pgdcorrection.c:28-59 */
    LEAVE(transfer);
    for (int d = 0; d < 3; d++) {
        /* fastpm_apply_diff_transfer(pm, canvas, canvas, d, 1), pgdcorrection.c:108: i * k_finite[d], zero at self-conjugate modes */
        fpm_transfer t;
        memset(&t, 0, sizeof(t));
        t.active = 1; t.potorder = -1; t.negate = 0; t.ngrad = 1; t.graddir[0] = d; t.gradorder = 1; t.zero_selfconj = 1; t.scale = 1.0;
        ENTER(c2r);
        fpm_mesh_c2r(pm, pot_k, canvas, &t);
        if (pm->NTask > 1) fpm_halo_fetch(pm, canvas);
        LEAVE(c2r);
        ENTER(readout);
        FPM_MUST(fpm_readout(pm->mesh, canvas, (const double *) p->x, (int64_t) p->np, (float *) p->pgdc + d, 3, 1.0));
        LEAVE(readout);
    }
    pm_free(pm, canvas);
    pm_free(pm, pot_k);
}
