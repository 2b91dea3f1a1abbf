/* ORACLE / TEST INFRASTRUCTURE ONLY -- mini-GSL subset */
#ifndef MINIGSL_ERRNO_H
#define MINIGSL_ERRNO_H
enum { GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_EMAXITER = 11, GSL_EROUND = 18 };
typedef void gsl_error_handler_t(const char *reason, const char *file, int line, int gsl_errno);
gsl_error_handler_t *gsl_set_error_handler(gsl_error_handler_t *h);
gsl_error_handler_t *gsl_set_error_handler_off(void);
#endif
