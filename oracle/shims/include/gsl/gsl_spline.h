#include <gsl/gsl_interp.h>
