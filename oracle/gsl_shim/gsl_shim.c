/* Minimal GSL-compatible shim (TEST INFRASTRUCTURE ONLY) -- see gsl/gsl_errno.h.
 *
 * Algorithms restated from the GSL 2.x documentation / published sources:
 *   - natural cubic spline (interpolation/cspline.c): c[0] = c[n-1] = 0, interior c from the
 *     symmetric tridiagonal system  h_{i}c_{i} + 2(h_i+h_{i+1})c_{i+1} + h_{i+1}c_{i+2}
 *       = 3(dy_{i+1}/h_{i+1} - dy_i/h_i), solved by the LDL^T recurrence of
 *     linalg/tridiag.c (solve_tridiag);  eval: b = dy/dx - dx(c_{i+1}+2c_i)/3,
 *     d = (c_{i+1}-c_i)/(3dx), y = y_i + delta(b + delta(c_i + delta d));
 *     x outside [x_0, x_{n-1}] is a domain error (default handler aborts).
 *   - LU with partial pivoting (linalg/lu.c): right-looking elimination, row swap on the
 *     largest |a_ij| in the column (first maximum wins), unit-lower L stored in place.
 *   - LU_solve / LU_invert: permute b, forward substitution with unit L, back substitution with U. */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "gsl/gsl_errno.h"
#include "gsl/gsl_spline.h"
#include "gsl/gsl_linalg.h"

static int handler_off = 0;
gsl_error_handler_t *gsl_set_error_handler_off(void) { handler_off = 1; return NULL; }
void gsl_shim_error(const char *reason, const char *file, int line, int gsl_errno)
{
  if (handler_off) return;
  fprintf(stderr, "gsl: %s:%d: ERROR: %s\nDefault GSL error handler invoked.\n", file, line, reason);
  fflush(stderr);
  (void)gsl_errno;
  abort();
}

static const gsl_interp_type cspline_type = { "cspline", 3 };
const gsl_interp_type *gsl_interp_cspline = &cspline_type;

gsl_interp_accel *gsl_interp_accel_alloc(void) { return (gsl_interp_accel *)calloc(1, sizeof(gsl_interp_accel)); }
void gsl_interp_accel_free(gsl_interp_accel *a) { free(a); }

gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size)
{
  (void)T;
  gsl_spline *s = (gsl_spline *)malloc(sizeof(gsl_spline));
  s->size = size;
  s->x = (double *)malloc(size * sizeof(double));
  s->y = (double *)malloc(size * sizeof(double));
  s->c = (double *)calloc(size, sizeof(double));
  return s;
}

int gsl_spline_init(gsl_spline *s, const double xa[], const double ya[], size_t size)
{
  size_t i;
  for (i = 0; i < size; i++) { s->x[i] = xa[i]; s->y[i] = ya[i]; }
  size_t max_index = size - 1;
  size_t sys = max_index - 1;          /* interior unknowns c[1..n-2] */
  s->c[0] = 0.0; s->c[max_index] = 0.0;
  if (sys == 0) return GSL_SUCCESS;
  double *g = (double *)malloc(sys * sizeof(double));
  double *diag = (double *)malloc(sys * sizeof(double));
  double *off = (double *)malloc(sys * sizeof(double));
  for (i = 0; i < sys; i++) {
    double h_i = xa[i + 1] - xa[i], h_ip1 = xa[i + 2] - xa[i + 1];
    double yd_i = ya[i + 1] - ya[i], yd_ip1 = ya[i + 2] - ya[i + 1];
    double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0, g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    off[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    g[i] = 3.0 * (yd_ip1 * g_ip1 - yd_i * g_i);
  }
  if (sys == 1) { s->c[1] = g[0] / diag[0]; }
  else {
    double *gamma = (double *)malloc(sys * sizeof(double));
    double *alpha = (double *)malloc(sys * sizeof(double));
    double *cc = (double *)malloc(sys * sizeof(double));
    double *z = (double *)malloc(sys * sizeof(double));
    alpha[0] = diag[0]; gamma[0] = off[0] / alpha[0];
    for (i = 1; i < sys - 1; i++) { alpha[i] = diag[i] - off[i - 1] * gamma[i - 1]; gamma[i] = off[i] / alpha[i]; }
    alpha[sys - 1] = diag[sys - 1] - off[sys - 2] * gamma[sys - 2];
    z[0] = g[0];
    for (i = 1; i < sys; i++) z[i] = g[i] - gamma[i - 1] * z[i - 1];
    for (i = 0; i < sys; i++) cc[i] = z[i] / alpha[i];
    s->c[sys] = cc[sys - 1];
    for (i = sys - 1; i-- > 0;) s->c[i + 1] = cc[i] - gamma[i] * s->c[i + 2];
    free(gamma); free(alpha); free(cc); free(z);
  }
  free(g); free(diag); free(off);
  return GSL_SUCCESS;
}

double gsl_spline_eval(const gsl_spline *s, double x, gsl_interp_accel *a)
{
  (void)a;
  size_t n = s->size;
  if (!(x >= s->x[0] && x <= s->x[n - 1])) {
    gsl_shim_error("interpolation error", __FILE__, __LINE__, GSL_EDOM);
    return NAN;
  }
  /* bsearch: index with x[index] <= x < x[index+1]; last interval for x == x[n-1] */
  size_t lo = 0, hi = n - 1;
  while (hi > lo + 1) { size_t i = (hi + lo) / 2; if (s->x[i] > x) hi = i; else lo = i; }
  double x_lo = s->x[lo], x_hi = s->x[lo + 1], dx = x_hi - x_lo;
  double y_lo = s->y[lo], y_hi = s->y[lo + 1], dy = y_hi - y_lo;
  double c_i = s->c[lo], c_ip1 = s->c[lo + 1];
  double b = (dy / dx) - dx * (c_ip1 + 2.0 * c_i) / 3.0;
  double d = (c_ip1 - c_i) / (3.0 * dx);
  double delta = x - x_lo;
  return y_lo + delta * (b + delta * (c_i + delta * d));
}

void gsl_spline_free(gsl_spline *s) { if (!s) return; free(s->x); free(s->y); free(s->c); free(s); }

gsl_matrix_view gsl_matrix_view_array(double *base, size_t n1, size_t n2)
{ gsl_matrix_view v; v.matrix.size1 = n1; v.matrix.size2 = n2; v.matrix.tda = n2; v.matrix.data = base; v.matrix.block = NULL; v.matrix.owner = 0; return v; }
gsl_vector_view gsl_vector_view_array(double *base, size_t n)
{ gsl_vector_view v; v.vector.size = n; v.vector.stride = 1; v.vector.data = base; v.vector.block = NULL; v.vector.owner = 0; return v; }
gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2)
{ gsl_matrix *m = (gsl_matrix *)malloc(sizeof(gsl_matrix)); m->size1 = n1; m->size2 = n2; m->tda = n2; m->data = (double *)calloc(n1 * n2, sizeof(double)); m->block = NULL; m->owner = 1; return m; }
void gsl_matrix_free(gsl_matrix *m) { if (!m) return; if (m->owner) free(m->data); free(m); }
double gsl_matrix_get(const gsl_matrix *m, size_t i, size_t j) { return m->data[i * m->tda + j]; }
gsl_vector *gsl_vector_alloc(size_t n)
{ gsl_vector *v = (gsl_vector *)malloc(sizeof(gsl_vector)); v->size = n; v->stride = 1; v->data = (double *)calloc(n, sizeof(double)); v->block = NULL; v->owner = 1; return v; }
void gsl_vector_free(gsl_vector *v) { if (!v) return; if (v->owner) free(v->data); free(v); }
double gsl_vector_get(const gsl_vector *v, size_t i) { return v->data[i * v->stride]; }
gsl_permutation *gsl_permutation_calloc(size_t n)
{ gsl_permutation *p = (gsl_permutation *)malloc(sizeof(gsl_permutation)); p->size = n; p->data = (size_t *)malloc(n * sizeof(size_t)); for (size_t i = 0; i < n; i++) p->data[i] = i; return p; }
gsl_permutation *gsl_permutation_alloc(size_t n) { return gsl_permutation_calloc(n); }
void gsl_permutation_free(gsl_permutation *p) { if (!p) return; free(p->data); free(p); }

#define M(A, i, j) ((A)->data[(i) * (A)->tda + (j)])

int gsl_linalg_LU_decomp(gsl_matrix *A, gsl_permutation *p, int *signum)
{
  size_t N = A->size1, i, j, k;
  *signum = 1;
  for (i = 0; i < N; i++) p->data[i] = i;
  for (j = 0; j + 1 < N; j++) {
    double max = fabs(M(A, j, j)); size_t i_pivot = j;
    for (i = j + 1; i < N; i++) { double aij = fabs(M(A, i, j)); if (aij > max) { max = aij; i_pivot = i; } }
    if (i_pivot != j) {
      for (k = 0; k < N; k++) { double t = M(A, j, k); M(A, j, k) = M(A, i_pivot, k); M(A, i_pivot, k) = t; }
      size_t t = p->data[j]; p->data[j] = p->data[i_pivot]; p->data[i_pivot] = t;
      *signum = -(*signum);
    }
    double ajj = M(A, j, j);
    if (ajj != 0.0) {
      for (i = j + 1; i < N; i++) {
        double aij = M(A, i, j) / ajj;
        M(A, i, j) = aij;
        for (k = j + 1; k < N; k++) M(A, i, k) -= aij * M(A, j, k);
      }
    }
  }
  return GSL_SUCCESS;
}

static int lu_singular(const gsl_matrix *LU)
{ for (size_t i = 0; i < LU->size1; i++) if (M(LU, i, i) == 0.0) return 1; return 0; }

static void lu_svx(const gsl_matrix *LU, const gsl_permutation *p, double *x, size_t stride)
{
  size_t N = LU->size1, i, j;
  double tmp[16];
  for (i = 0; i < N; i++) tmp[i] = x[p->data[i] * stride];
  for (i = 0; i < N; i++) { double s = tmp[i]; for (j = 0; j < i; j++) s -= M(LU, i, j) * tmp[j]; tmp[i] = s; }
  for (i = N; i-- > 0;) { double s = tmp[i]; for (j = i + 1; j < N; j++) s -= M(LU, i, j) * tmp[j]; tmp[i] = s / M(LU, i, i); }
  for (i = 0; i < N; i++) x[i * stride] = tmp[i];
}

int gsl_linalg_LU_solve(const gsl_matrix *LU, const gsl_permutation *p, const gsl_vector *b, gsl_vector *x)
{
  if (lu_singular(LU)) { gsl_shim_error("matrix is singular", __FILE__, __LINE__, GSL_EDOM); return GSL_EDOM; }
  for (size_t i = 0; i < b->size; i++) x->data[i * x->stride] = b->data[i * b->stride];
  lu_svx(LU, p, x->data, x->stride);
  return GSL_SUCCESS;
}

int gsl_linalg_LU_invert(const gsl_matrix *LU, const gsl_permutation *p, gsl_matrix *inverse)
{
  size_t N = LU->size1;
  if (lu_singular(LU)) { gsl_shim_error("matrix is singular", __FILE__, __LINE__, GSL_EDOM); return GSL_EDOM; }
  for (size_t i = 0; i < N; i++) for (size_t j = 0; j < N; j++) M(inverse, i, j) = (i == j) ? 1.0 : 0.0;
  for (size_t j = 0; j < N; j++) lu_svx(LU, p, inverse->data + j, inverse->tda);
  return GSL_SUCCESS;
}
