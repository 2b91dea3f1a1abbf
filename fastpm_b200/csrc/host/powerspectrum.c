/* fastpm_b200 host layer -- P(k) objects (reference: libfastpm/powerspectrum.c).  The shell sums come from
 * the device (fpm_powerspectrum_sums); table parsing, interpolation, sigma(R) and the text writer are host code. */
#include "internal.h"
#include <math.h>

void fastpm_funck_init(FastPMFuncK *fk, const size_t size)
{ fk->size = size; fk->k = malloc(sizeof(double) * size); fk->f = malloc(sizeof(double) * size); }
void fastpm_funck_destroy(FastPMFuncK *fk) { free(fk->f); free(fk->k); }

/* "k<TAB>f" rows; lines that do not parse are skipped (powerspectrum.c:343-378) */
int fastpm_funck_init_from_string(FastPMFuncK *fk, const char *string)
{
    for (int pass = 0; pass < 2; pass++) {
        size_t n = 0;
        const char *s = string;
        while (*s) {
            const char *e = strchr(s, '\n');
            size_t len = e ? (size_t) (e - s) : strlen(s);
            char line[512];
            if (len < sizeof(line)) {
                memcpy(line, s, len); line[len] = 0;
                double k, f;
                if (2 == sscanf(line, "%lg\t%lg", &k, &f)) { if (pass) { fk->k[n] = k; fk->f[n] = f; } n++; }
            }
            if (!e) break;
            s = e + 1;
        }
        if (!pass) fastpm_funck_init(fk, n);
    }
    return fk->size == 0 ? -1 : 0;
}

double fastpm_funck_eval(FastPMFuncK *fk, double k)
{
    if (k == 0) return 1;
    int l = 0, r = (int) fk->size - 1;
    while (r - l > 1) { int m = (r + l) / 2; if (k < fk->k[m]) r = m; else l = m; }
    double k2 = fk->k[r], k1 = fk->k[l], f2 = fk->f[r], f1 = fk->f[l];
    if (l == r) return fk->f[l];
    if (f1 <= 0 || f2 <= 0 || k1 == 0 || k2 == 0) return ((k - k1) * f2 + (k2 - k) * f1) / (k2 - k1);
    k = log(k); f1 = log(f1); f2 = log(f2); k1 = log(k1); k2 = log(k2);
    return exp(((k - k1) * f2 + (k2 - k) * f1) / (k2 - k1));
}
double fastpm_funck_eval2(double k, FastPMFuncK *fk) { return fastpm_funck_eval(fk, k); }

void fastpm_powerspectrum_init(FastPMPowerSpectrum *ps, const size_t size)
{
    fastpm_funck_init(&ps->base, size);
    ps->pm = NULL;
    ps->edges = malloc(sizeof(double) * (size + 1));
    ps->Nmodes = malloc(sizeof(double) * size);
}
int fastpm_powerspectrum_init_from_string(FastPMPowerSpectrum *ps, const char *string)
{
    int r = fastpm_funck_init_from_string(&ps->base, string);
    ps->edges = malloc(sizeof(double) * (ps->base.size + 1));
    ps->Nmodes = malloc(sizeof(double) * ps->base.size);
    return r;
}
void fastpm_powerspectrum_destroy(FastPMPowerSpectrum *ps) { free(ps->edges); free(ps->Nmodes); fastpm_funck_destroy(&ps->base); }

void fastpm_powerspectrum_init_from_delta(FastPMPowerSpectrum *ps, PM *pm, const FastPMFloat *delta1_k, const FastPMFloat *delta2_k)
{
    const int nb = (int) pm->Nmesh[0] / 2;
    fastpm_powerspectrum_init(ps, nb);
    ps->pm = pm;
    ps->Volume = pm->Volume;
    ps->k0 = 2 * M_PI / pm->BoxSize[0];
    for (int i = 0; i <= nb; i++) ps->edges[i] = i * ps->k0;
    double *sums = malloc(sizeof(double) * (3 * nb + 1));
    if (delta1_k == delta2_k) FPM_MUST(fpm_powerspectrum_sums(pm->mesh, delta1_k, 0, sums));
    else FPM_MUST(fpm_cross_powerspectrum_sums(pm->mesh, delta1_k, delta2_k, sums));      /* powerspectrum.c:87-91 */
    fpm_comm_allreduce_double(pm->comm, sums, 3 * nb, 0);          /* powerspectrum.c:113-115 */
    for (int i = 0; i < nb; i++) {
        ps->Nmodes[i] = sums[i];
        ps->base.f[i] = sums[nb + i];
        ps->base.k[i] = sums[2 * nb + i];
        if (ps->Nmodes[i] == 0) continue;
        ps->base.k[i] /= ps->Nmodes[i];
        ps->base.f[i] /= ps->Nmodes[i];
        ps->base.f[i] *= ps->Volume;
    }
    free(sums);
}

void fastpm_powerspectrum_write(FastPMPowerSpectrum *ps, char *filename, double N)
{
    FILE *fp = fopen(filename, "w");
    if (!fp) fastpm_raise(-1, "cannot open %s for writing\n", filename);
    fprintf(fp, "# k p N \n");
    for (size_t i = 0; i < ps->base.size; i++) fprintf(fp, "%g %g %g\n", ps->base.k[i], ps->base.f[i], ps->Nmodes[i]);
    double *L = pm_boxsize(ps->pm);
    fprintf(fp, "# metadata 7\n");
    fprintf(fp, "# volume %g float64\n", ps->Volume);
    fprintf(fp, "# shotnoise %g float64\n", ps->Volume / N);
    fprintf(fp, "# N1 %g int\n", N);
    fprintf(fp, "# N2 %g int\n", N);
    fprintf(fp, "# Lz %g float64\n", L[2]);
    fprintf(fp, "# Lx %g float64\n", L[0]);
    fprintf(fp, "# Ly %g float64\n", L[1]);
    fclose(fp);
}

double fastpm_powerspectrum_large_scale(FastPMPowerSpectrum *ps, int Nmax)
{
    double kmax = Nmax * ps->k0, P = 0, Nm = 0;
    for (size_t i = 0; (i == 0) || (i < ps->base.size && ps->base.k[i] <= kmax); i++) { P += ps->base.f[i] * ps->Nmodes[i]; Nm += ps->Nmodes[i]; }
    return P / Nm;
}
double fastpm_powerspectrum_eval(FastPMPowerSpectrum *ps, double k) { return fastpm_funck_eval(&ps->base, k); }
double fastpm_powerspectrum_eval2(double k, FastPMPowerSpectrum *ps) { return fastpm_funck_eval(&ps->base, k); }

typedef struct { FastPMPowerSpectrum *ps; double R; } SigmaArg;
static double sigma2_integrand(double k, void *param)
{
    SigmaArg *s = param;
    double kr = s->R * k, kr2 = kr * kr, kr3 = kr2 * kr;
    if (kr < 1e-8) return 0;
    double w = 3 * (sin(kr) / kr3 - cos(kr) / kr2);
    return 4 * M_PI * k * k * w * w * fastpm_powerspectrum_eval(s->ps, k) / pow(2 * M_PI, 3);
}
/* powerspectrum.c:250-279: top-hat sigma(R), relative tolerance 1e-4 */
double fastpm_powerspectrum_sigma(FastPMPowerSpectrum *ps, double R)
{
    SigmaArg s = { ps, R };
    return sqrt(fpm_integrate(sigma2_integrand, &s, 0, 500.0 * 1 / R, 0, 1e-4, 20));
}
/* powerspectrum.c:25-33 */
void fastpm_powerspectrum_init_from(FastPMPowerSpectrum *ps, const FastPMPowerSpectrum *other)
{
    fastpm_powerspectrum_init(ps, other->base.size);
    memcpy(ps->base.k, other->base.k, sizeof(double) * ps->base.size);
    memcpy(ps->base.f, other->base.f, sizeof(double) * ps->base.size);
    memcpy(ps->edges, other->edges, sizeof(double) * (ps->base.size + 1));
    memcpy(ps->Nmodes, other->Nmodes, sizeof(double) * ps->base.size);
}

/* powerspectrum.c:127-141: sqrt(P_dest / P_src) per shell */
void fastpm_transferfunction_init(FastPMPowerSpectrum *ps, PM *pm, FastPMFloat *src_k, FastPMFloat *dest_k)
{
    FastPMPowerSpectrum ps2[1];
    fastpm_powerspectrum_init_from_delta(ps, pm, src_k, src_k);
    fastpm_powerspectrum_init_from_delta(ps2, pm, dest_k, dest_k);
    for (size_t i = 0; i < ps->base.size; i++) ps->base.f[i] = sqrt(ps2->base.f[i] / ps->base.f[i]);
    fastpm_powerspectrum_destroy(ps2);
}

/* powerspectrum.c:186-226: callbacks with the inverted signature, and the value of the shell that holds k */
double fastpm_powerspectrum_get(FastPMPowerSpectrum *ps, double k)
{
    if (k == 0) return 1;
    int l = 0, r = (int) ps->base.size;
    while (r - l > 1) {
        const int m = (r + l) / 2;
        if (k <= ps->edges[m]) r = m; else l = m;
    }
    return ps->base.f[l];
}
double fastpm_powerspectrum_get2(double k, FastPMPowerSpectrum *ps) { return fastpm_powerspectrum_get(ps, k); }

/* powerspectrum.c:292-331: mode-weighted merge of consecutive shells */
void fastpm_powerspectrum_rebin(FastPMPowerSpectrum *ps, size_t newsize)
{
    double *k1 = malloc(newsize * sizeof(double)), *p1 = malloc(newsize * sizeof(double));
    double *Nmodes1 = malloc(newsize * sizeof(double)), *edges1 = malloc((newsize + 1) * sizeof(double));
    for (size_t i = 0; i < newsize; i++) {
        const size_t j1 = i * ps->base.size / newsize, j2 = (i + 1) * ps->base.size / newsize;
        k1[i] = 0; p1[i] = 0; Nmodes1[i] = 0;
        edges1[i] = ps->edges[j1]; edges1[i + 1] = ps->edges[j2];
        for (size_t j = j1; j < j2; j++) {
            k1[i] += ps->base.k[j] * ps->Nmodes[j];
            p1[i] += ps->base.f[j] * ps->Nmodes[j];
            Nmodes1[i] += ps->Nmodes[j];
        }
        if (Nmodes1[i] > 0) { k1[i] /= Nmodes1[i]; p1[i] /= Nmodes1[i]; }
    }
    free(ps->base.k); free(ps->base.f); free(ps->Nmodes); free(ps->edges);
    ps->base.k = k1; ps->base.f = p1; ps->Nmodes = Nmodes1; ps->edges = edges1; ps->base.size = newsize;
}

void fastpm_powerspectrum_scale(FastPMPowerSpectrum *ps, double factor)
{ for (size_t i = 1; i < ps->base.size; i++) ps->base.f[i] *= factor; }
