/* fastpm_b200 host layer -- private definitions shared by the C files in this directory.
 * The public contract is include/fastpm_b200_api.h (the libfastpm mirror) on top of
 * include/fastpm_b200.h (the device C ABI). */
#ifndef FASTPM_B200_HOST_INTERNAL_H
#define FASTPM_B200_HOST_INTERNAL_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "fastpm_b200.h"
#include "fastpm_b200_api.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846264338328
#endif

/* struct PM of this build (opaque to API users, api/fastpm/libfastpm.h:20; the reference's is pmpfft.h:43-70) */
struct PM {
    fpm_mesh *mesh;
    int NTask, ThisTask;
    int Nproc[2];
    MPI_Comm comm;
    ptrdiff_t Nmesh[3];
    double BoxSize[3];
    ptrdiff_t allocsize;          /* floats per mesh buffer */
    PMRegion IRegion, ORegion;
    double Norm, Volume;
    double CellSize[3], InvCellSize[3];
    FastPMMemory *mem;
    FastPMFloat *scratch;         /* lazily allocated, for the in-place public pm_c2r / pm_r2c */
    FastPMFloat *stage;           /* several GPUs: staging mesh of the slab transposes (lazily allocated) */
    int stage_off;                /* no room for it in the arena: direct peer stores */
    FastPMFloat *stage2;          /* second staging mesh: pipelined inverse transforms of the force components */
    int stage2_off;
    FastPMFloat *whalo;           /* several GPUs, windows wider than CIC: halo planes below / above the slab (host/gravity.c) */
    int whl, whr;
    int transposed;
    int pitch_r, pitch_c, nxl, x0, nyl, y0, halo;
};

struct VPM {
    PM *pm;
    double a_start;
    double pm_nc_factor;
    int end;
};
int fpm_painter_window(const FastPMPainter *painter);      /* FPM_WINDOW_* of csrc/window.h */
VPM *vpm_create(VPMInit *vpminit, int base_nmesh, double boxsize, MPI_Comm comm);
VPM *vpm_find(VPM *vpm, double a);
void vpm_free(VPM *vpm);

/* device call that must succeed: the reference has no error returns, it raises (logging.c:242-251) */
#define FPM_MUST(call) do { if ((call) != 0) fastpm_raise(-1, "%s: %s\n", #call, fpm_last_error()); } while (0)

/* communicator table (comm.c): rank/size + collectives over host scalars; device exchanges live in comm.c too */
int fpm_comm_rank(MPI_Comm comm);
int fpm_comm_size(MPI_Comm comm);
void fpm_comm_allreduce_double(MPI_Comm comm, double *v, int n, int op);   /* op: 0 sum, 1 min, 2 max */
void fpm_comm_allreduce_i64(MPI_Comm comm, int64_t *v, int n, int op);
void fpm_comm_barrier(MPI_Comm comm);
void fpm_comm_release_migration(void);

/* comm.c: mesh exchanges for one rank or many */
void fpm_halo_add(PM *pm, FastPMFloat *canvas);
void fpm_halo_fetch(PM *pm, FastPMFloat *canvas);
void fpm_mesh_r2c(PM *pm, FastPMFloat *real, FastPMFloat *cplx, double scale);
void fpm_mesh_c2r(PM *pm, const FastPMFloat *cplx, FastPMFloat *real, const fpm_transfer *kernel);
/* several GPUs: the two halves of the inverse transform and the check that a second canvas + staging mesh fit (host/comm.c) */
int fpm_dist_pipeline_ready(PM *pm);
void fpm_dist_c2r_begin(PM *pm, const FastPMFloat *cplx, FastPMFloat *real, const fpm_transfer *kernel, int set);
void fpm_dist_c2r_finish(PM *pm, FastPMFloat *real, int set);
void fpm_mesh_readout(PM *pm, FastPMFloat *canvas, const double *x, int64_t np, float *out, int stride, double prescale);

/* solver.c: store whose wrap is folded into the next fastpm_paint_local */
extern FastPMStore *fpm_pending_wrap;

/* factors.c: applies the queued in-place kicks / drifts of p (NULL: of any store); call before touching store columns */
void fpm_store_flush(FastPMStore *p);

/* numerics.c */
typedef double (*fpm_func1)(double x, void *params);
double fpm_integrate(fpm_func1 f, void *params, double a, double b, double epsabs, double epsrel, int order);
typedef void (*fpm_odefunc)(double t, const double *y, double *dydt, void *params);
int fpm_ode_rkf45(fpm_odefunc f, void *params, int dim, double *t, double t1, double *y, double h0, double epsabs, double epsrel);

/* pm.c internals */
FastPMFloat *pm_alloc_noclear(PM *pm, const char *file, int line);
PM *pm_new(int nmesh, double boxsize, MPI_Comm comm);
void pm_delete(PM *pm);

/* factors.c: interpolated factor differences (fastpm_kick_one / fastpm_drift_one, factors.c:73-171) */
void fpm_kick_factors_at(FastPMKickFactor *kick, double a_v, double af, double *dda, double *Dv1, double *Dv2);
void fpm_drift_factors_at(FastPMDriftFactor *drift, double a_x, double af, double *dyyy, double *da1, double *da2);

/* solver.c: 2LPT on the device (pm2lpt.c:14-210) */
void pm_2lpt_solve(PM *pm, FastPMFloat *delta_k, FastPMFuncK *growth_rate_func_k, FastPMStore *p, double shift[3], FastPMKernelType type);
void pm_2lpt_evolve(double aout, FastPMStore *p, FastPMCosmology *c, int zaonly);

#endif
