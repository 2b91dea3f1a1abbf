/* fastpm_b200 -- host-side mirror of the libfastpm C API for the PM force step and the
 * kick/drift integrator (the part of /root/reference/api/fastpm/ that sits on the hot path).
 *
 * Purpose: libfastpm_b200.so exports these symbols with the reference's names, argument meaning,
 * struct layouts and error behaviour, so that a caller written against libfastpm (src/fastpm.c,
 * tests/testpm.c, external C users) links against it in place of libfastpm.a for this path.
 * The struct layouts below are checked field by field against the reference's headers by
 * tests/test_abi_layout.py whenever the reference tree is present.
 *
 * Differences a caller must know (also in INTEGRATION.md):
 *   - mesh buffers (pm_alloc) and FastPMStore columns are DEVICE memory; host code reads them with
 *     fastpm_b200_* mirror helpers (bottom of this file), not by dereferencing;
 *   - `PM` stays opaque (api/fastpm/libfastpm.h:20); its k-space order is [ky][kx][kz] with a padded
 *     row pitch, reported through pm_o_region() strides like any other layout;
 *   - MPI_Comm: without an MPI installation this header supplies a one-word handle type; ranks are
 *     the processes of the x-slab decomposition, one GPU each, joined by fastpm_b200_comm_init().
 *
 * Each block cites the reference header it mirrors as  [api/fastpm/<file>:<lines>].
 */
#ifndef FASTPM_B200_API_H
#define FASTPM_B200_API_H

#include <stddef.h>
#include <stdint.h>
#include <stdarg.h>

#ifdef FASTPM_B200_WITH_MPI
#include <mpi.h>
#else
#ifndef FASTPM_B200_MPI_STANDIN
#define FASTPM_B200_MPI_STANDIN
typedef int MPI_Comm;              /* handle into the library's communicator table */
#define MPI_COMM_WORLD 1
#endif
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ [libfastpm.h:20-58] */
typedef struct PM PM;
typedef struct FastPMStore FastPMStore;
typedef struct FastPMPainter FastPMPainter;
typedef struct FastPMTransition FastPMTransition;
typedef struct FastPMCosmology FastPMCosmology;
typedef float FastPMFloat;                      /* FASTPM_FFT_PRECISION == 32 only */

typedef enum { FASTPM_FORCE_FASTPM = 0, FASTPM_FORCE_PM, FASTPM_FORCE_COLA, FASTPM_FORCE_2LPT, FASTPM_FORCE_ZA } FastPMForceType;
typedef enum { FASTPM_KERNEL_3_4, FASTPM_KERNEL_3_2, FASTPM_KERNEL_5_4, FASTPM_KERNEL_1_4, FASTPM_KERNEL_1_4_DIFF0,
               FASTPM_KERNEL_GADGET, FASTPM_KERNEL_EASTWOOD, FASTPM_KERNEL_NAIVE } FastPMKernelType;
typedef enum { FASTPM_SOFTENING_NONE, FASTPM_SOFTENING_GAUSSIAN, FASTPM_SOFTENING_GADGET_LONG_RANGE,
               FASTPM_SOFTENING_TWO_THIRD, FASTPM_SOFTENING_GAUSSIAN36 } FastPMSofteningType;

void libfastpm_init(void);
void libfastpm_cleanup(void);
void libfastpm_set_memory_bound(size_t size);
extern const char *LIBFASTPM_VERSION;

typedef double (*fastpm_fkfunc)(double k, void *data);
#define fastpm_pkfunc fastpm_fkfunc

/* ------------------------------------------------------------------ [events.h:1-53] */
typedef struct FastPMEventHandler FastPMEventHandler;
enum FastPMEventStage { FASTPM_EVENT_STAGE_BEFORE, FASTPM_EVENT_STAGE_AFTER };
typedef struct { char type[32]; enum FastPMEventStage stage; } FastPMEvent;
typedef int (*FastPMEventHandlerFunction)(void *context, FastPMEvent *event, void *userdata);
struct FastPMEventHandler {
    char type[32];
    enum FastPMEventStage stage;
    FastPMEventHandlerFunction function;
    void *userdata;
    struct FastPMEventHandler *next;
    void (*free)(void *);
};
void fastpm_add_event_handler(FastPMEventHandler **handlers, const char *type, enum FastPMEventStage stage,
                              FastPMEventHandlerFunction function, void *userdata);
void fastpm_add_event_handler_free(FastPMEventHandler **handlers, const char *where, enum FastPMEventStage stage,
                                   FastPMEventHandlerFunction function, void *userdata, void (*free)(void *ptr));
void fastpm_remove_event_handler(FastPMEventHandler **handlers, const char *type, enum FastPMEventStage stage,
                                 FastPMEventHandlerFunction function, void *userdata);
void fastpm_emit_event(FastPMEventHandler *handlers, const char *type, enum FastPMEventStage stage,
                       FastPMEvent *event, void *context);
void fastpm_destroy_event_handlers(FastPMEventHandler **handlers);

/* ------------------------------------------------------------------ [memory.h:16-67] */
typedef struct MemoryBlock MemoryBlock;
typedef struct FastPMMemory FastPMMemory;
typedef void (*fastpm_memory_func)(FastPMMemory *m, void *userdata);
struct FastPMMemory {
    size_t alignment;
    size_t total_bytes;
    MemoryBlock *pools[8];
    size_t peak_bytes;
    size_t used_bytes;
    size_t free_bytes;
    char *base0;
    char *top;
    char *base;
    void *userdata;
    fastpm_memory_func abortfunc;
    fastpm_memory_func peakfunc;
};
enum FastPMMemoryLocation { FASTPM_MEMORY_HEAP, FASTPM_MEMORY_STACK, FASTPM_MEMORY_FLOATING, FASTPM_MEMORY_MAX };
void fastpm_memory_init(FastPMMemory *m, size_t total_bytes);
void fastpm_memory_set_handlers(FastPMMemory *m, fastpm_memory_func abortfunc, fastpm_memory_func peakfunc, void *userdata);
void fastpm_memory_destroy(FastPMMemory *m);
void fastpm_memory_free(FastPMMemory *m, void *p);
void *fastpm_memory_alloc_details(FastPMMemory *m, const char *name, size_t s, enum FastPMMemoryLocation loc, const char *file, const int line);
void fastpm_memory_dump_status_str(FastPMMemory *m, char *buf, int n);
#define fastpm_memory_alloc(m, name, s, loc) fastpm_memory_alloc_details(m, name, s, loc, __FILE__, __LINE__)
FastPMMemory *_libfastpm_get_gmem(void);
/* freed mesh buffers are kept by size so that no cudaMalloc happens inside a step (csrc/host/support.c): this returns them to the device */
void fastpm_b200_memory_trim(void);

/* ------------------------------------------------------------------ [logging.h:16-79] */
enum FastPMLogLevel { ERROR = 100, INFO = 1 };
enum FastPMLogType { COLLECTIVE = 0, INDIVIDUAL = 1 };
typedef void (*fastpm_msg_handler)(const enum FastPMLogLevel level, const enum FastPMLogType type, const int errcode,
                                   const char *message, MPI_Comm comm, void *userdata);
void fastpm_set_msg_handler(fastpm_msg_handler handler, MPI_Comm comm, void *userdata);
void fastpm_push_msg_handler(fastpm_msg_handler handler, MPI_Comm comm, void *userdata);
void fastpm_pop_msg_handler(void);
void fastpm_default_msg_handler(const enum FastPMLogLevel level, const enum FastPMLogType type, const int errcode,
                                const char *message, MPI_Comm comm, void *userdata);
void fastpm_void_msg_handler(const enum FastPMLogLevel level, const enum FastPMLogType type, const int errcode,
                             const char *message, MPI_Comm comm, void *userdata);
#define fastpm_info(...) fastpm_info_(__FILE__, __LINE__, ## __VA_ARGS__)
#define fastpm_raise(...) fastpm_raise_(__FILE__, __LINE__, ## __VA_ARGS__)
#define fastpm_ilog(...) fastpm_ilog_(__FILE__, __LINE__, ## __VA_ARGS__)
void fastpm_info_(const char *file, int line, const char *fmt, ...);
void fastpm_raise_(const char *file, int line, const int code, const char *fmt, ...);
void fastpm_ilog_(const char *file, int line, const enum FastPMLogLevel level, const char *fmt, ...);

/* ------------------------------------------------------------------ [prof.h:16-33] */
typedef struct FastPMClock FastPMClock;
void fastpm_clock_in(FastPMClock *clock);
void fastpm_clock_out(FastPMClock *clock);
FastPMClock *fastpm_clock_find(const char *file, const char *func, const char *name);
#define CLOCK(name) FastPMClock * CLK ## name = fastpm_clock_find(__FILE__, __func__, # name); fastpm_clock_in(CLK ## name);
#define ENTER(name) fastpm_clock_in(CLK ## name)
#define LEAVE(name) fastpm_clock_out(CLK ## name)
void fastpm_clock_stat(MPI_Comm comm);

/* ------------------------------------------------------------------ [pmapi.h:3-109] */
typedef struct {
    ptrdiff_t start[3];
    ptrdiff_t size[3];
    ptrdiff_t strides[3];      /* in units of real numbers for IRegion, complex numbers for ORegion */
    ptrdiff_t total;
} PMRegion;
FastPMFloat *pm_alloc_details(PM *pm, const char *file, const int line);
#define pm_alloc(pm) pm_alloc_details(pm, __FILE__, __LINE__)
void pm_free(PM *pm, FastPMFloat *buf);
void pm_assign(PM *pm, FastPMFloat *from, FastPMFloat *to);
void pm_clear(PM *pm, FastPMFloat *buf);
size_t pm_allocsize(PM *pm);
MPI_Comm pm_comm(PM *pm);
double pm_norm(PM *pm);
ptrdiff_t *pm_nmesh(PM *pm);
int *pm_nproc(PM *pm);
double *pm_boxsize(PM *pm);
double pm_volume(PM *pm);
int pm_unbalanced(PM *pm);
PMRegion *pm_i_region(PM *pm);
PMRegion *pm_o_region(PM *pm);
double pm_compute_variance(PM *pm, FastPMFloat *complx);
void pm_check_values(PM *pm, FastPMFloat *field, const char *fmt, ...);
int pm_pos_to_rank(PM *pm, double pos[3]);
void pm_r2c(PM *pm, FastPMFloat *from, FastPMFloat *to);     /* out of place */
void pm_c2r(PM *pm, FastPMFloat *inplace);                   /* in place */
PM *fastpm_create_pm(int Ngrid, int NprocY, int transposed, double BoxSize, MPI_Comm comm);
void fastpm_free_pm(PM *pm);

/* ------------------------------------------------------------------ [store.h:9-272] */
typedef uint8_t FastPMParticleMaskType;
enum FastPMSpecies { FASTPM_SPECIES_BARYON = 0, FASTPM_SPECIES_CDM = 1, FASTPM_SPECIES_NCDM = 2 };
typedef enum FastPMColumnTags {
    COLUMN_MASK = 1L << 0, COLUMN_POS = 1L << 1, COLUMN_Q = 1L << 2, COLUMN_VEL = 1L << 3,
    COLUMN_DX1 = 1L << 4, COLUMN_DX2 = 1L << 5, COLUMN_DV1 = 1L << 6, COLUMN_ACC = 1L << 7,
    COLUMN_ID = 1L << 8, COLUMN_AEMIT = 1L << 9, COLUMN_DENSITY = 1L << 10, COLUMN_POTENTIAL = 1L << 11,
    COLUMN_TIDAL = 1L << 12, COLUMN_PGDC = 1L << 13,
    COLUMN_MINID = 1L << 14, COLUMN_TASK = 1L << 15, COLUMN_LENGTH = 1L << 16, COLUMN_RDISP = 1L << 17,
    COLUMN_VDISP = 1L << 18, COLUMN_RVDISP = 1L << 19,
    COLUMN_MASS = 1L << 20, COLUMN_RAND = 1L << 21, COLUMN_RMOM = 1L << 22,
} FastPMColumnTags;

struct FastPMStore {
    FastPMMemory *mem;
    char name[32];
    FastPMColumnTags attributes;
    void *_base;
    size_t np;
    size_t np_upper;
    struct FastPMColumnInfo {
        void (*pack)(FastPMStore *p, ptrdiff_t index, int ci, void *packed);
        void (*unpack)(FastPMStore *p, ptrdiff_t index, int ci, void *packed);
        double (*to_double)(FastPMStore *p, ptrdiff_t index, int ci, int memb);
        void (*from_double)(FastPMStore *p, ptrdiff_t index, int ci, int memb, const double value);
        char name[32];
        char dtype[8];
        size_t elsize;
        size_t membsize;
        size_t nmemb;
        FastPMColumnTags attribute;
    } _column_info[32];
    struct {
        double a_x;
        double a_v;
        double M0;
        double _q_shift[3];
        double _q_scale[3];
        ptrdiff_t _q_strides[3];
        ptrdiff_t _q_size;
    } meta;
    union {
        char *columns[32];
        struct {
            FastPMParticleMaskType *mask;
            double (*x)[3];
            float (*q)[3];
            float (*v)[3];
            float (*dx1)[3];
            float (*dx2)[3];
            float (*dv1)[3];
            float (*acc)[3];
            uint64_t *id;
            float *aemit;
            float *rho;
            float *potential;
            float (*tidal)[6];
            float (*pgdc)[3];
            uint64_t *minid;
            int32_t *task;
            int32_t *length;
            float (*rdisp)[6];
            float (*vdisp)[6];
            float (*rvdisp)[9];
            float *mass;
            float *rand;
            float *rmom;
        };
    };
};
#define FASTPM_STORE_COLUMN_INDEX(column) (((char*) &(((FastPMStore *) NULL)->column) - (char*) &(((FastPMStore *)NULL)->columns[0])) \
                        / sizeof(((FastPMStore *) NULL)->columns[0]))

typedef struct { FastPMColumnTags attribute; int memb; } FastPMFieldDescr;

const char *fastpm_species_get_name(enum FastPMSpecies species);
double fastpm_store_get_mass(FastPMStore *p, ptrdiff_t index);
void fastpm_store_init_details(FastPMStore *p, const char *name, size_t np_upper, FastPMColumnTags attributes,
                               enum FastPMMemoryLocation loc, const char *file, const int line);
#define fastpm_store_init(p, name, np_upper, attributes, loc) fastpm_store_init_details(p, name, np_upper, attributes, loc, __FILE__, __LINE__)
size_t fastpm_store_init_evenly_details(FastPMStore *p, const char *name, size_t np_total, FastPMColumnTags attributes,
                                        double alloc_factor, MPI_Comm comm, const char *file, const int line);
#define fastpm_store_init_evenly(p, name, np_total, attributes, alloc_factor, comm) \
        fastpm_store_init_evenly_details(p, name, np_total, attributes, alloc_factor, comm, __FILE__, __LINE__)
void fastpm_store_fill(FastPMStore *p, PM *pm, double *shift, ptrdiff_t *Nc);
int fastpm_store_find_column_id(FastPMStore *p, FastPMColumnTags attribute);
void fastpm_store_destroy(FastPMStore *p);
void fastpm_store_summary(FastPMStore *p, FastPMColumnTags attribute, MPI_Comm comm, const char *fmt, ...);
void fastpm_store_wrap(FastPMStore *p, double BoxSize[3]);
typedef int (*fastpm_store_target_func)(FastPMStore *p, ptrdiff_t index, void *data);
int fastpm_store_decompose(FastPMStore *p, fastpm_store_target_func target_func, void *data, MPI_Comm comm);
size_t fastpm_store_get_np_total(FastPMStore *p, MPI_Comm comm);
int FastPMTargetPM(FastPMStore *p, ptrdiff_t i, PM *pm);

/* ------------------------------------------------------------------ [painter.h:3-35] */
typedef double (*fastpm_kernelfunc)(double x, double hsupport);
typedef enum { FASTPM_PAINTER_CIC, FASTPM_PAINTER_LINEAR, FASTPM_PAINTER_QUAD, FASTPM_PAINTER_LANCZOS } FastPMPainterType;
struct FastPMPainter {
    PM *pm;
    void (*paint)(FastPMPainter *painter, FastPMFloat *canvas, double pos[3], double weight, int diffdir);
    double (*readout)(FastPMPainter *painter, FastPMFloat *canvas, double pos[3], int diffdir);
    fastpm_kernelfunc kernel;
    fastpm_kernelfunc diff;
    int diffdir;
    int support;
    double hsupport;
    double invh;
    int left;
    int Npoints;
    double shift;
};
void fastpm_painter_init(FastPMPainter *painter, PM *pm, FastPMPainterType type, int support);
void fastpm_painter_init_diff(FastPMPainter *painter, FastPMPainter *base, int diffdir);      /* painter.c:178 */
void fastpm_paint_local(FastPMPainter *painter, FastPMFloat *canvas, FastPMStore *p, size_t size, FastPMFieldDescr field);
void fastpm_readout_local(FastPMPainter *painter, FastPMFloat *canvas, FastPMStore *p, size_t size, FastPMFieldDescr field);
void fastpm_paint(FastPMPainter *painter, FastPMFloat *canvas, FastPMStore *p, FastPMFieldDescr field);

/* ------------------------------------------------------------------ [cosmology.h:3-84] */
extern double HubbleConstant;
extern double HubbleDistance;
typedef enum { FASTPM_GROWTH_MODE_LCDM = 0, FASTPM_GROWTH_MODE_ODE = 1 } FastPMGrowthMode;
typedef struct FastPMFDInterp FastPMFDInterp;
struct FastPMCosmology {
    double h;
    double Omega_m;
    double Omega_cdm;
    double Omega_ncdm;
    double Omega_k;
    double Omega_Lambda;
    double w0;
    double wa;
    double T_cmb;
    double N_eff;
    int N_nu;
    double m_ncdm[3];
    int N_ncdm;
    int ncdm_freestreaming;
    int ncdm_matterlike;
    int ncdm_linearresponse;
    FastPMGrowthMode growth_mode;
    FastPMFDInterp *FDinterp;
};
typedef struct FastPMGrowthInfo { double a; FastPMCosmology *c; double D1, D2, f1, f2; } FastPMGrowthInfo;
double Omega_g(FastPMCosmology *c);
double Gamma_nu(FastPMCosmology *c);
double Omega_ur(FastPMCosmology *c);
double Omega_r(FastPMCosmology *c);
double Omega_DE_TimesHubbleEaSq(double a, FastPMCosmology *c);
double DOmega_DE_TimesHubbleEaSqDa(double a, FastPMCosmology *c);
double D2Omega_DE_TimesHubbleEaSqDa2(double a, FastPMCosmology *c);
double HubbleEa(double a, FastPMCosmology *c);
double Omega_cdm_a(double a, FastPMCosmology *c);
double Omega_m(double a, FastPMCosmology *c);
double Omega_source(double a, FastPMCosmology *c);
double DHubbleEaDa(double a, FastPMCosmology *c);
double D2HubbleEaDa2(double a, FastPMCosmology *c);
void fastpm_cosmology_init(FastPMCosmology *c);
void fastpm_cosmology_destroy(FastPMCosmology *c);
void fastpm_growth_info_init(FastPMGrowthInfo *growth_info, double a, FastPMCosmology *c);
double DGrowthFactorDa(FastPMGrowthInfo *growth_info);
double D2GrowthFactorDa2(FastPMGrowthInfo *growth_info);
double ComovingDistance(double a, FastPMCosmology *c);

/* ------------------------------------------------------------------ [transfer.h:3-37] */
void fastpm_apply_decic_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to);
void fastpm_apply_diff_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to, int dir, int order);
void fastpm_apply_multiply_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to, double value);
void fastpm_apply_smoothing_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to, double sml);    /* transfer.h, transfer.c:8 */
void fastpm_apply_lowpass_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to, double kth);       /* transfer.c:43 */
void fastpm_apply_laplace_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to, int order);
void fastpm_apply_modify_mode_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to, ptrdiff_t *mode, double value);

/* ------------------------------------------------------------------ [powerspectrum.h:3-78] */
typedef struct FastPMFuncK { size_t size; double *k; double *f; } FastPMFuncK;
typedef struct { FastPMFuncK base; double *edges; PM *pm; double k0; double Volume; double *Nmodes; } FastPMPowerSpectrum;
void fastpm_funck_init(FastPMFuncK *fk, const size_t size);
int fastpm_funck_init_from_string(FastPMFuncK *fk, const char *string);
double fastpm_funck_eval(FastPMFuncK *fk, double k);
double fastpm_funck_eval2(double k, FastPMFuncK *fk);
void fastpm_funck_destroy(FastPMFuncK *fk);
void fastpm_powerspectrum_init(FastPMPowerSpectrum *ps, size_t size);
void fastpm_powerspectrum_init_from_delta(FastPMPowerSpectrum *ps, PM *pm, const FastPMFloat *delta1_k, const FastPMFloat *delta2_k);
int fastpm_powerspectrum_init_from_string(FastPMPowerSpectrum *ps, const char *string);
void fastpm_powerspectrum_destroy(FastPMPowerSpectrum *ps);
void fastpm_powerspectrum_write(FastPMPowerSpectrum *ps, char *filename, double N);
double fastpm_powerspectrum_large_scale(FastPMPowerSpectrum *ps, int Nmax);
double fastpm_powerspectrum_eval(FastPMPowerSpectrum *ps, double k);
double fastpm_powerspectrum_eval2(double k, FastPMPowerSpectrum *ps);
double fastpm_powerspectrum_sigma(FastPMPowerSpectrum *ps, double R);
void fastpm_powerspectrum_scale(FastPMPowerSpectrum *ps, double factor);
void fastpm_powerspectrum_init_from(FastPMPowerSpectrum *ps, const FastPMPowerSpectrum *other);                      /* powerspectrum.c:25 */
void fastpm_transferfunction_init(FastPMPowerSpectrum *ps, PM *pm, FastPMFloat *src_k, FastPMFloat *dest_k);        /* :127 */
double fastpm_powerspectrum_get(FastPMPowerSpectrum *ps, double k);                                                  /* :200 */
double fastpm_powerspectrum_get2(double k, FastPMPowerSpectrum *ps);                                                 /* :222 */
void fastpm_powerspectrum_rebin(FastPMPowerSpectrum *ps, size_t newsize);                                            /* :292 */

/* ------------------------------------------------------------------ [initialcondition.h:3-17] */
enum FastPMFillDeltaKScheme { FASTPM_DELTAK_GADGET, FASTPM_DELTAK_FAST, FASTPM_DELTAK_SLOW };     /* [initialcondition.h:3-7] */
/* only the default scheme, FASTPM_DELTAK_GADGET, is implemented (on the device); the others raise */
void fastpm_ic_fill_gaussiank(PM *pm, FastPMFloat *delta_k, int seed, enum FastPMFillDeltaKScheme scheme);
void fastpm_ic_induce_correlation(PM *pm, FastPMFloat *delta_k, fastpm_fkfunc pk, void *pkdata);
void fastpm_ic_remove_variance(PM *pm, FastPMFloat *delta_k);                                /* initialcondition.h, initialcondition.c:66 */

/* ------------------------------------------------------------------ [pgdcorrection.h:3-11] */
typedef struct { FastPMPainterType PainterType; int PainterSupport; double alpha0, A, B, kl, ks; } FastPMPGDCorrection;
double fastpm_pgdc_get_alpha(FastPMPGDCorrection *pgdc, double a);      /* pgdcorrection.h:14-22 */
double fastpm_pgdc_get_kl(FastPMPGDCorrection *pgdc, double a);
double fastpm_pgdc_get_ks(FastPMPGDCorrection *pgdc, double a);
/* fills p->pgdc (device column) from delta_k, which is left unchanged; pgdcorrection.h:24-27 */
void fastpm_pgdc_calculate(FastPMPGDCorrection *pgdc, PM *pm, FastPMStore *p, FastPMFloat *delta_k, double a, double fac);

/* ------------------------------------------------------------------ [timemachine.h:5-47] */
enum FastPMAction { FASTPM_ACTION_FORCE, FASTPM_ACTION_KICK, FASTPM_ACTION_DRIFT };
typedef struct { int force, x, v; } FastPMState;
typedef struct { FastPMState *table; int cycle_len; int cycles; double *timesteps; } FastPMStates;
struct FastPMTransition {
    FastPMStates *states;
    int istart;
    int iend;
    FastPMState *start;
    FastPMState *end;
    enum FastPMAction action;
    struct { double i, f, r; } a;
    struct { int i, f, r; } i;
};
FastPMStates *fastpm_tevo_generate_states(FastPMStates *states, int cycles, FastPMState *templ, double *ts);
void fastpm_tevo_destroy_states(FastPMStates *states);
void fastpm_tevo_transition_init(FastPMTransition *transition, FastPMStates *states, int istart, int iend);
int fastpm_tevo_transition_find_dual(FastPMTransition *transition, FastPMTransition *dual);
int fastpm_tevo_transition_find_next(FastPMTransition *transition, FastPMTransition *next);

/* ------------------------------------------------------------------ [solver.h:1-208] */
#define FASTPM_EVENT_FORCE "FORCE"
#define FASTPM_EVENT_LPT "LPT"
#define FASTPM_EVENT_TRANSITION "TRANSITION"
#define FASTPM_EVENT_INTERPOLATION "INTERPOLATION"
typedef struct VPM VPM;
typedef struct VPMInit { double a_start; double pm_nc_factor; } VPMInit;
typedef struct FastPMDriftFactor FastPMDriftFactor;
typedef struct FastPMKickFactor FastPMKickFactor;
enum { TIMESTEP_START, TIMESTEP_CUR, TIMESTEP_END };
typedef struct { FastPMEvent base; FastPMDriftFactor *drift; FastPMKickFactor *kick; double a1; double a2; int whence; } FastPMInterpolationEvent;
typedef struct { FastPMEvent base; FastPMTransition *transition; } FastPMTransitionEvent;
typedef struct { FastPMEvent base; PM *pm; FastPMFloat *delta_k; FastPMStore *p; } FastPMLPTEvent;
typedef struct {
    FastPMEvent base;
    FastPMKernelType kernel;
    FastPMPainter *painter;
    PM *pm;
    FastPMFloat *delta_k;
    double N;
    double a_f;
    double a_n;
} FastPMForceEvent;
typedef struct {
    size_t nc;
    double boxsize;
    double alloc_factor;
    double lpt_nc_factor;
    FastPMCosmology *cosmology;
    VPMInit *vpminit;
    int USE_DX1_ONLY;
    int USE_SHIFT;
    FastPMColumnTags ExtraAttributes;
    double nLPT;
    FastPMPainterType PAINTER_TYPE;
    int painter_support;
    FastPMForceType FORCE_TYPE;
    FastPMKernelType KERNEL_TYPE;
    FastPMSofteningType SOFTENING_TYPE;
    int NprocY;
    int UseFFTW;
    int pgdc;
    double pgdc_alpha0;
    double pgdc_A;
    double pgdc_B;
    double pgdc_kl;
    double pgdc_ks;
} FastPMConfig;
#define FASTPM_SOLVER_NSPECIES 6
typedef struct {
    FastPMStore *species[FASTPM_SOLVER_NSPECIES];
    char has_species[FASTPM_SOLVER_NSPECIES];
    FastPMStore cdm[1];
    MPI_Comm comm;
    int NTask;
    int ThisTask;
    FastPMConfig config[1];
    FastPMPGDCorrection pgdc[1];
    FastPMCosmology cosmology[1];
    FastPMEventHandler *event_handlers;
    VPM *vpm_list;
    PM *basepm;
    PM *lptpm;
} FastPMSolver;
struct FastPMDriftFactor {
    FastPMForceType forcemode;
    double ai, ac, af;
    int nsamples;
    double Dv1, Dv2;
    double dyyy[32], da1[32], da2[32];
};
struct FastPMKickFactor {
    FastPMForceType forcemode;
    double ai, ac, af;
    int nsamples;
    double q1, q2;
    double dda[32], Dv1[32], Dv2[32];
};
void fastpm_solver_init(FastPMSolver *fastpm, FastPMConfig *config, MPI_Comm comm);
void fastpm_solver_destroy(FastPMSolver *fastpm);
FastPMStore *fastpm_solver_get_species(FastPMSolver *fastpm, enum FastPMSpecies species);
void fastpm_solver_add_species(FastPMSolver *fastpm, enum FastPMSpecies species, FastPMStore *store);
void fastpm_solver_setup_lpt(FastPMSolver *fastpm, enum FastPMSpecies species, FastPMFloat *delta_k_ic,
                             FastPMFuncK *growth_rate_func_k_ic, double a0);
PM *fastpm_find_pm(FastPMSolver *fastpm, double a);
void fastpm_solver_evolve(FastPMSolver *fastpm, double *time_step, int nstep);
void fastpm_drift_init(FastPMDriftFactor *drift, FastPMSolver *fastpm, double ai, double ac, double af);
void fastpm_kick_init(FastPMKickFactor *kick, FastPMSolver *fastpm, double ai, double ac, double af);
void fastpm_kick_store(FastPMKickFactor *kick, FastPMStore *pi, FastPMStore *po, double af);
void fastpm_drift_store(FastPMDriftFactor *drift, FastPMStore *pi, FastPMStore *po, double af);
void fastpm_set_species_snapshot(FastPMSolver *fastpm, FastPMStore *p, FastPMDriftFactor *drift, FastPMKickFactor *kick,
                                 FastPMStore *po, double aout);
void fastpm_unset_species_snapshot(FastPMSolver *fastpm, FastPMStore *p, FastPMDriftFactor *drift, FastPMKickFactor *kick,
                                   FastPMStore *po, double aout);

/* ------------------------------------------------------------------ [gravity.h:3-22] */
#define FASTPM_CRITICAL_DENSITY 27.7455
void fastpm_kernel_type_get_orders(FastPMKernelType type, int *potorder, int *gradorder, int *difforder, int *deconvolveorder);
void fastpm_solver_compute_force(FastPMSolver *fastpm, PM *pm, FastPMPainter *painter, FastPMSofteningType dealias,
                                 FastPMKernelType kernel, FastPMFloat *delta_k, double Time);
void gravity_apply_kernel_transfer(FastPMKernelType kernel, PM *pm, FastPMFloat *delta_k, FastPMFloat *canvas, FastPMFieldDescr field);

/* ------------------------------------------------------------------ ranks (one process per GPU, x-slabs)
 * The launcher supplies two host-buffer collectives (MPI_Allreduce / MPI_Allgather in an MPI build, torch.distributed in
 * fastpm_b200/multigpu.py): allreduce(buf, count, is_int64 (else double), op (0 sum, 1 min, 2 max)), allgather(send, nbytes, recv).
 * fastpm_b200_comm_init also maps the peers' device arenas; _init_host sets the collectives only (host-side tools: file IO). */
typedef void (*fpm_host_allreduce_fn)(void *buf, int count, int is_int64, int op, void *userdata);
typedef void (*fpm_host_allgather_fn)(const void *send, int nbytes, void *recv, void *userdata);
void fastpm_b200_comm_init(int rank, int size, fpm_host_allreduce_fn allreduce, fpm_host_allgather_fn allgather, void *userdata);
void fastpm_b200_comm_init_host(int rank, int size, fpm_host_allreduce_fn allreduce, fpm_host_allgather_fn allgather, void *userdata);

/* ------------------------------------------------------------------ fastpm_b200 extensions
 * Host mirrors of device-resident data, for callers (snapshot writers, FOF, lightcone, custom event
 * handlers) that the reference lets dereference mesh buffers and store columns directly. */
/* copy `count` elements of a store column to / from host memory (elements are whole rows, e.g. double[3]) */
/* Deferred work and direct access to device pointers.  Between two force evaluations the library queues the in-place kicks and drifts of a
 * store and applies them in ONE pass (host/factors.c), and do_force records the CIC deconvolution of event->delta_k instead of sweeping
 * the mesh when the FORCE/after handler is the P(k) measurement (host/solver.c).  Every entry point of this library applies what is
 * pending on the buffers it is handed, so code that goes through the API never sees the difference.  A handler that reads p->x, p->v,
 * event->delta_k, ... with its OWN device code (a CUDA kernel, torch on the pointer) must call this first: it applies everything pending
 * and waits for the device. */
void fastpm_b200_sync_state(void);
/* Promise that an event handler never reads the particles (it prints a progress line, say): events are delivered to it without first
 * applying the queued kicks and drifts, which keeps the fused particle update alive.  Ordinary handlers see a fully updated store. */
void fastpm_b200_mark_handler_passive(FastPMEventHandlerFunction function);
/* 1 when the host-scalar collectives of a multi-rank run go through the shared-memory segment of host/shmcoll.c (all ranks on one
 * host), 0 when they go through the launcher's callbacks (FASTPM_B200_HOST_COLL=callbacks, several hosts) or there is one rank */
int fastpm_b200_host_collectives_shared(void);
/* The ranks of a one-node run without a launcher (fastpm_b200_run -n N forks them): the parent makes the segment before forking --
 * no CUDA call may precede the fork --, every rank attaches by name; no callbacks are involved.  The parent unlinks it at the end. */
int fastpm_b200_local_segment_create(char *name_out, size_t cap);
int fastpm_b200_local_segment_unlink(const char *name);
void fastpm_b200_comm_init_local(int rank, int size, const char *segment);
/* before libfastpm_cleanup in a multi-rank program: waits for every rank, returns the communicator's own device block */
void fastpm_b200_comm_finalize(void);
/* the largest number of exchange rounds a fastpm_store_decompose of this process has needed so far (1 unless a pack buffer overflowed; 0 before the first) */
int fastpm_b200_migrate_rounds_max(void);
int fastpm_b200_store_set_np(FastPMStore *p, int64_t np);
/* a scratch store with q and rand columns filled by fastpm_store_fill on pm's grid, mirrored to the host (bindings, tests); returns np */
int64_t fastpm_b200_fill_probe(PM *pm, int64_t np_upper, float *q_host, float *rand_host);
/* fastpm_paint (CIC) of the CDM store -> pm_r2c -> fastpm_powerspectrum_init_from_delta -> pm_c2r -> fastpm_readout_local into ACC[:, 0],
 * on one GPU or on the slabs of several: P(k) (Nmesh / 2 bins) and the density read back at this rank's particles; returns np */
int64_t fastpm_b200_public_mesh_probe(FastPMSolver *fastpm, double a, double *k, double *p, double *nmodes, float *dens_host);
int fastpm_b200_store_get_column(FastPMStore *p, FastPMColumnTags attribute, void *host_dst, size_t first, size_t count);
int fastpm_b200_store_set_column(FastPMStore *p, FastPMColumnTags attribute, const void *host_src, size_t first, size_t count);
/* mesh buffers: copy to / from a host array in the REFERENCE's layouts -- real [x][y][N+2] floats
 * (pmpfft.c:181-187) and, for k-space, complex [x][y][N/2+1] ("untransposed", ORegion of basepm/lptpm,
 * solver.c:103-109).  allocsize of such a host array is N*N*(N+2) floats. */
int fastpm_b200_mesh_get_real(PM *pm, const FastPMFloat *dev, float *host_dst);
int fastpm_b200_mesh_set_real(PM *pm, FastPMFloat *dev, const float *host_src);
int fastpm_b200_mesh_get_complex(PM *pm, const FastPMFloat *dev, float *host_dst);
int fastpm_b200_mesh_set_complex(PM *pm, FastPMFloat *dev, const float *host_src);
/* number of floats in a host array in the reference layout: N*N*(N+2) */
size_t fastpm_b200_mesh_host_size(PM *pm);
/* ------------------------------------------------------------------ small remaining entry points of the same headers (csrc/host/extras.c) */
void fastpm_set_snapshot(FastPMSolver *fastpm, FastPMSolver *snapshot, FastPMDriftFactor *drift, FastPMKickFactor *kick, double aout);   /* solver.h:188 */
void fastpm_unset_snapshot(FastPMSolver *fastpm, FastPMSolver *snapshot, FastPMDriftFactor *drift, FastPMKickFactor *kick, double aout); /* solver.h:199 */
void fastpm_store_set_name(FastPMStore *p, const char *name);                                             /* store.h:166 */
int fastpm_store_has_q(FastPMStore *p);
void fastpm_store_get_q_from_id(FastPMStore *p, uint64_t id, double q[3]);
void fastpm_store_get_iq_from_id(FastPMStore *p, uint64_t id, ptrdiff_t pabs[3]);
void fastpm_store_steal(FastPMStore *in, FastPMStore *out, FastPMColumnTags attributes);                  /* store.h:260 */
/* store.h:224-268: whole particles between / inside stores, sub-sampling, local sort (csrc/host/store_ops.c).  `mask` and a per-particle
 * `fraction` array are device memory (fastpm_memory_alloc hands out device memory); fastpm_store_permute takes a HOST index array
 * as in the reference; fastpm_store_sort accepts the comparator libfastpm itself uses, FastPMLocalSortByID. */
void fastpm_store_copy(FastPMStore *p, FastPMStore *po);
void fastpm_store_take(FastPMStore *p, ptrdiff_t i, FastPMStore *po, ptrdiff_t j);
void fastpm_store_extend(FastPMStore *p, FastPMStore *extra);
void fastpm_store_get_position(FastPMStore *p, ptrdiff_t index, double pos[3]);
void fastpm_store_get_lagrangian_position(FastPMStore *p, ptrdiff_t index, double pos[3]);
size_t fastpm_store_get_mask_sum(FastPMStore *p, MPI_Comm comm);
void fastpm_store_fill_subsample_mask(FastPMStore *p, double fraction, FastPMParticleMaskType *mask);
void fastpm_store_fill_subsample_mask_from_array(FastPMStore *p, double *fraction, FastPMParticleMaskType *mask);
size_t fastpm_store_subsample(FastPMStore *p, FastPMParticleMaskType *mask, FastPMStore *po);
void fastpm_store_permute(FastPMStore *p, int *ind);
int FastPMLocalSortByID(const int i1, const int i2, FastPMStore *p);
void fastpm_store_sort(FastPMStore *p, int (*cmp_func)(const int i1, const int i2, FastPMStore *p));
/* utils.h:3-14: analytic spectra to pass to fastpm_ic_induce_correlation (tests/testpm.c:67-74) */
struct fastpm_powerspec_eh_params { double hubble_param; double omegam; double omegab; double Norm; };
double fastpm_utils_powerspec_eh(double k, struct fastpm_powerspec_eh_params *param);   /* Eisenstein & Hu, no wiggles */
double fastpm_utils_powerspec_white(double k, double *amplitude);
/* bindings / tests: fill a scratch store on pm's grid, sub-sample it at `fraction` (into a second store or in place), optionally reverse
 * it with fastpm_store_permute and sort it back with fastpm_store_sort, mirror the kept
 * ids and positions; returns the number kept */
int64_t fastpm_b200_subsample_probe(PM *pm, int64_t np_upper, double fraction, int in_place, int sort_back, uint64_t *id_host, double *x_host, int64_t *mask_sum);
double fastpm_apply_get_mode_transfer(PM *pm, FastPMFloat *from, ptrdiff_t *mode);                        /* transfer.c:340 */
void fastpm_apply_set_mode_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to, ptrdiff_t *mode, double value, int method);   /* :290 */
void fastpm_apply_normalize_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to);                         /* :223 */
void fastpm_apply_c2r_weight_transfer(PM *pm, FastPMFloat *from, FastPMFloat *to);                        /* :249 */
char *fastpm_file_get_content(const char *filename);                                                      /* string.h */
char **fastpm_strsplit(const char *str, const char *split);
char *fastpm_strdup(const char *str);
char *fastpm_strdup_printf(const char *fmt, ...);
void fastpm_path_ensure_dirname(const char *path);
int read_funck(FastPMFuncK *fk, const char filename[], MPI_Comm comm);                                    /* io.h */

/* ------------------------------------------------------------------ [io.h] snapshot / mesh files (bigfile directories)
 * Same names, arguments and on-disk result as libfastpmio/io.c; the healpix / light-cone writers are not implemented; on several ranks the sort by id
 * keeps the particles in place and writes every row at the file position its (dense) id gives. */
typedef void (*FastPMSnapshotSorter)(const void *ptr, void *radix, void *arg);
void FastPMSnapshotSortByID(const void *ptr, void *radix, void *arg);
void fastpm_sort_snapshot(FastPMStore *p, MPI_Comm comm, FastPMSnapshotSorter sorter, int redistribute);     /* io.h:19 */
int fastpm_store_write(FastPMStore *p, const char *filebase, const char *mode, int Nwriters, MPI_Comm comm);  /* io.h:22; mode "w" or "r" */
int fastpm_store_read(FastPMStore *p, const char *filebase, int Nwriters, MPI_Comm comm);                     /* io.h:30 */
void write_snapshot_header(FastPMSolver *fastpm, const char *filebase, MPI_Comm comm);                        /* io.h:37 */
void read_snapshot_header(FastPMSolver *fastpm, const char *filebase, double *aout, MPI_Comm comm);           /* io.h:41 */
void write_snapshot_attr(const char *filebase, const char *dataset, const char *attrname, void *buf, const char *dtype, size_t nmemb, MPI_Comm comm); /* io.h:45 */
int write_complex(PM *pm, FastPMFloat *data, const char *filename, const char *blockname, int Nwriters);      /* io.h:56 */
int read_complex(PM *pm, FastPMFloat *data, const char *filename, const char *blockname, int Nwriters);       /* io.h:59 */
/* the same writers on plain arrays (host or device), for bindings and tests */
typedef struct { const char *name; const char *dtype_out; const char *dtype; int nmemb; void *data; int on_device; } FpmIoColumn;
typedef struct { int64_t q_strides[3]; double q_scale[3], q_shift[3]; int64_t q_size; double a_x, a_v, M0; } FpmIoMeta;
typedef struct {
    int64_t NC; double BoxSize, ScalingFactor, GrowthFactor, GrowthRate, HubbleE, RSDFactor, Omega_cdm, OmegaM, OmegaLambda, HubbleParam;
    const char *version; uint64_t TotNumPart[6]; double MassTable[6];
} FpmIoHeader;
int fastpm_b200_io_write_columns(const char *filebase, const char *dataset, const FpmIoColumn *cols, int ncols, int64_t np_local,
                                 const FpmIoMeta *meta, MPI_Comm comm);
int fastpm_b200_io_write_columns_at(const char *filebase, const char *dataset, const FpmIoColumn *cols, int ncols, int64_t np_local,
                                    const FpmIoMeta *meta, const uint64_t *positions, MPI_Comm comm); /* the append mode of fastpm_store_write (io.c:522-537): every column block grows (big_block_mpi_grow_simple) and the ranks' items are
 * written behind its old end; the dataset's own attributes are left alone */
int fastpm_b200_io_append_columns(const char *filebase, const char *dataset, const FpmIoColumn *cols, int ncols, int64_t np_local, MPI_Comm comm);
    /* rows at given file positions */
int fastpm_b200_io_read_columns(const char *filebase, const char *dataset, const FpmIoColumn *cols, int ncols, int64_t *np_local,
                                FpmIoMeta *meta, MPI_Comm comm);
int fastpm_b200_io_write_header(const char *filebase, const FpmIoHeader *h, MPI_Comm comm);
void fastpm_b200_io_header_values(FastPMSolver *fastpm, double aout, double M0_cdm, uint64_t np_total_cdm, FpmIoHeader *h);
int fastpm_b200_io_write_complex_rows(const char *filename, const char *blockname, int nmesh, double boxsize, int y0, int nyl,
                                      const float *rows, size_t pitch_c, int nfile, MPI_Comm comm);
int fastpm_b200_io_read_complex_rows(const char *filename, const char *blockname, int nmesh, int y0, int nyl, float *rows, size_t pitch_c);
void fastpm_b200_io_argsort_u64(const uint64_t *key, size_t n, uint64_t *perm);
/* one snapshot of the CDM species at its current time as src/fastpm.c:1190-1200,1473-1486 writes it (unit conversion, optional
 * sort by id, Header, catalog, conversion reverted), and the restart read of src/fastpm.c:618-635; returns the header's ScalingFactor */
void fastpm_b200_write_snapshot(FastPMSolver *fastpm, const char *filebase, int sort_by_id);
double fastpm_b200_read_snapshot(FastPMSolver *fastpm, const char *filebase);
/* snapshots at the scale factors aout[nout] while fastpm_solver_evolve runs, written to "<base>_%0.04f": the CLI's check_snapshots
 * (src/fastpm.c:1130-1208) as a ready-made INTERPOLATION handler */
void fastpm_b200_add_snapshot_handler(FastPMSolver *fastpm, const char *base, const double *aout, int nout, int sort_by_id);

/* a FastPMConfig/FastPMSolver pair built from scalars, for bindings that cannot lay out the structs */
FastPMSolver *fastpm_b200_solver_new(int64_t nc, double boxsize, const double *pm_nc_factor_pairs, int npairs,
                                     double alloc_factor, double lpt_nc_factor, int force_mode, int kernel_type,
                                     int growth_mode, int compute_potential, double nLPT,
                                     double Omega_m, double h, double T_cmb, double N_eff, int N_nu);
/* same, with the PGD correction of src/fastpm.c:204-217: pgdc = NULL (off) or {alpha0, A, B, kl, ks} (adds COLUMN_PGDC), and the
 * force softening (FastPMSofteningType, gravity.c:244-270) and the painter of the force step (FastPMPainterType + support) */
FastPMSolver *fastpm_b200_solver_new_ex(int64_t nc, double boxsize, const double *pm_nc_factor_pairs, int npairs,
                                        double alloc_factor, double lpt_nc_factor, int force_mode, int kernel_type,
                                        int growth_mode, int compute_potential, double nLPT,
                                        double Omega_m, double h, double T_cmb, double N_eff, int N_nu, const double *pgdc, int softening_type,
                                        int painter_type, int painter_support);
/* FastPMConfig.USE_SHIFT (ICs at cell centres, solver.c:142-150) and USE_DX1_ONLY for the next solver these constructors make */
void fastpm_b200_solver_next_options(int use_shift, int use_dx1_only);
void fastpm_b200_solver_free(FastPMSolver *solver);
/* initial conditions as src/fastpm.c:415-545 makes them from a seed and a linear P(k) table (k, p: `size` doubles each), all on the
 * device: Gadget-scheme white noise, optional remove_variance, colouring, DC mode = 1, 2LPT at a0 */
void fastpm_b200_setup_gadget_ic(FastPMSolver *fastpm, int seed, int remove_variance, const double *k, const double *p, int size, double a0);

#ifdef __cplusplus
}
#endif
#endif
