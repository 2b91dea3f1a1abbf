/* ORACLE / TEST INFRASTRUCTURE ONLY -- mini-GSL subset, see ../../src/minigsl.c */
#ifndef MINIGSL_MATH_H
#define MINIGSL_MATH_H
#include <math.h>
#include <stddef.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846264338328
#endif
typedef struct { double (*function)(double x, void *params); void *params; } gsl_function;
#define GSL_FN_EVAL(F, x) (*((F)->function))(x, (F)->params)
#endif
