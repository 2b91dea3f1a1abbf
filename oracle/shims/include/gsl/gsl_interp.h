/* ORACLE / TEST INFRASTRUCTURE ONLY -- mini-GSL subset: linear interpolation */
#ifndef MINIGSL_INTERP_H
#define MINIGSL_INTERP_H
#include <stddef.h>
typedef struct { const char *name; } gsl_interp_type;
typedef struct { const gsl_interp_type *type; size_t size; } gsl_interp;
typedef struct { size_t cache; } gsl_interp_accel;
extern const gsl_interp_type *gsl_interp_linear;
gsl_interp *gsl_interp_alloc(const gsl_interp_type *T, size_t n);
int gsl_interp_init(gsl_interp *obj, const double xa[], const double ya[], size_t size);
double gsl_interp_eval(const gsl_interp *obj, const double xa[], const double ya[], double x, gsl_interp_accel *a);
void gsl_interp_free(gsl_interp *interp);
gsl_interp_accel *gsl_interp_accel_alloc(void);
void gsl_interp_accel_free(gsl_interp_accel *a);
#endif
