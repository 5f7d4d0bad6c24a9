#ifndef IS3D_GSL_SHIM_LINALG_H
#define IS3D_GSL_SHIM_LINALG_H
#include <stddef.h>
#include "gsl_errno.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { size_t size1, size2, tda; double *data; void *block; int owner; } gsl_matrix;
typedef struct { size_t size, stride; double *data; void *block; int owner; } gsl_vector;
typedef struct { gsl_matrix matrix; } gsl_matrix_view;
typedef struct { gsl_vector vector; } gsl_vector_view;
typedef struct { size_t size; size_t *data; } gsl_permutation;
gsl_matrix_view gsl_matrix_view_array(double *base, size_t n1, size_t n2);
gsl_vector_view gsl_vector_view_array(double *base, size_t n);
gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2);
void gsl_matrix_free(gsl_matrix *m);
double gsl_matrix_get(const gsl_matrix *m, size_t i, size_t j);
gsl_vector *gsl_vector_alloc(size_t n);
void gsl_vector_free(gsl_vector *v);
double gsl_vector_get(const gsl_vector *v, size_t i);
gsl_permutation *gsl_permutation_alloc(size_t n);
gsl_permutation *gsl_permutation_calloc(size_t n);
void gsl_permutation_free(gsl_permutation *p);
int gsl_linalg_LU_decomp(gsl_matrix *A, gsl_permutation *p, int *signum);
int gsl_linalg_LU_solve(const gsl_matrix *LU, const gsl_permutation *p, const gsl_vector *b, gsl_vector *x);
int gsl_linalg_LU_invert(const gsl_matrix *LU, const gsl_permutation *p, gsl_matrix *inverse);
#ifdef __cplusplus
}
#endif
#endif
