#ifndef IS3D_GSL_SHIM_SPLINE_H
#define IS3D_GSL_SHIM_SPLINE_H
#include "gsl_interp.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { size_t size; double *x; double *y; double *c; } gsl_spline;
gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size);
int gsl_spline_init(gsl_spline *s, const double xa[], const double ya[], size_t size);
double gsl_spline_eval(const gsl_spline *s, double x, gsl_interp_accel *a);
void gsl_spline_free(gsl_spline *s);
#ifdef __cplusplus
}
#endif
#endif
