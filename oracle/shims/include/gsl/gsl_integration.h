/* ORACLE / TEST INFRASTRUCTURE ONLY -- mini-GSL subset.
 * gsl_integration_qag here is a globally adaptive bisection scheme like
 * QUADPACK QAG, but it uses a Gauss-Legendre rule pair (n, 2n+1 points,
 * nodes found by Newton iteration at start-up) instead of the tabulated
 * Gauss-Kronrod pairs; answers agree with GSL to the requested tolerance,
 * not bisection-by-bisection. */
#ifndef MINIGSL_INTEGRATION_H
#define MINIGSL_INTEGRATION_H
#include <gsl/gsl_math.h>
typedef struct { size_t limit; } gsl_integration_workspace;
enum { GSL_INTEG_GAUSS15 = 1, GSL_INTEG_GAUSS21, GSL_INTEG_GAUSS31,
       GSL_INTEG_GAUSS41, GSL_INTEG_GAUSS51, GSL_INTEG_GAUSS61 };
gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n);
void gsl_integration_workspace_free(gsl_integration_workspace *w);
int gsl_integration_qag(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        size_t limit, int key, gsl_integration_workspace *w, double *result, double *abserr);
#endif
