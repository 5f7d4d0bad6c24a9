// df coefficient tables on the device and their per-cell evaluation.
// Mirrors Deltaf_Data::evaluate_df_coefficients (reference src/cpp/DeltafData.cpp:324-519):
//   include_baryon = 0 -> natural cubic spline in T over the muB = 0 row (cubic_spline, :324-402)
//   include_baryon = 1 -> bilinear interpolation on the uniform (T, muB) grid (bilinear_interpolation, :419-499)
#pragma once

#include <vector>

#include "common.cuh"

namespace is3d {

// natural cubic spline: nodes x[n], values y[n], second-derivative coefficients c[n] (c[0] = c[n-1] = 0)
struct Spline {
  const double *x = nullptr;
  const double *y = nullptr;
  const double *c = nullptr;
  int n = 0;
};

// Same evaluation formula as GSL's cspline (interpolation/cspline.c, cspline_eval): interval by bisection with
// x[i] <= v < x[i+1] (last interval closed), b = dy/dx - dx (c1 + 2 c0)/3, d = (c1 - c0)/(3 dx).
// Returns false when v is outside the table (GSL: domain error -> abort in the reference).
IS3D_HD bool spline_eval(const Spline &s, double v, double *out)
{
  if (!(v >= s.x[0] && v <= s.x[s.n - 1])) { *out = 0.0; return false; }
  int lo = 0, hi = s.n - 1;
  while (hi > lo + 1) { int i = (hi + lo) >> 1; if (s.x[i] > v) hi = i; else lo = i; }
  double x_lo = s.x[lo], dx = s.x[lo + 1] - x_lo;
  double y_lo = s.y[lo], dy = s.y[lo + 1] - y_lo;
  double c_i = s.c[lo], c_ip1 = s.c[lo + 1];
  double b = (dy / dx) - dx * (c_ip1 + 2.0 * c_i) / 3.0;
  double d = (c_ip1 - c_i) / (3.0 * dx);
  double delta = v - x_lo;
  *out = y_lo + delta * (b + delta * (c_i + delta * d));
  return true;
}

// Natural cubic spline: c[0] = c[n-1] = 0, interior from the symmetric tridiagonal system
//   h_i c_i + 2(h_i + h_{i+1}) c_{i+1} + h_{i+1} c_{i+2} = 3 (dy_{i+1}/h_{i+1} - dy_i/h_i)
// solved by the LDL^T recurrence (same formulation as GSL's cspline_init + solve_tridiag, which the reference
// calls through gsl_spline_init, DeltafData.cpp:313-320).
inline void natural_cspline_coefficients(const double *x, const double *y, int n, double *c)
{
  for (int i = 0; i < n; i++) c[i] = 0.0;
  int sys = n - 2;
  if (sys <= 0) return;
  std::vector<double> g(sys), diag(sys), off(sys);
  for (int i = 0; i < sys; i++) {
    double h_i = x[i + 1] - x[i], h_ip1 = x[i + 2] - x[i + 1];
    double yd_i = y[i + 1] - y[i], yd_ip1 = y[i + 2] - y[i + 1];
    double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0, g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    off[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    g[i] = 3.0 * (yd_ip1 * g_ip1 - yd_i * g_i);
  }
  if (sys == 1) { c[1] = g[0] / diag[0]; return; }
  std::vector<double> gamma(sys), alpha(sys), cc(sys), z(sys);
  alpha[0] = diag[0];
  gamma[0] = off[0] / alpha[0];
  for (int i = 1; i < sys - 1; i++) { alpha[i] = diag[i] - off[i - 1] * gamma[i - 1]; gamma[i] = off[i] / alpha[i]; }
  alpha[sys - 1] = diag[sys - 1] - off[sys - 2] * gamma[sys - 2];
  z[0] = g[0];
  for (int i = 1; i < sys; i++) z[i] = g[i] - gamma[i - 1] * z[i - 1];
  for (int i = 0; i < sys; i++) cc[i] = z[i] / alpha[i];
  c[sys] = cc[sys - 1];
  for (int i = sys - 2; i >= 0; i--) c[i + 1] = cc[i] - gamma[i] * c[i + 2];
}


struct DfTables {
  int n_T = 0, n_muB = 0;
  double T_min = 0, muB_min = 0, dT = 0, dmuB = 0;
  const double *T = nullptr;      // [n_T]
  const double *muB = nullptr;    // [n_muB]
  // raw tables, [n_muB][n_T] row-major: c0 c1 c2 c3 c4 F G betabulk betaV betapi
  const double *tab[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // splines in T of the muB = 0 row (include_baryon = 0): c0, c2, F, betabulk, betapi
  Spline sp_c0, sp_c2, sp_F, sp_betabulk, sp_betapi;
  // PTB: lambda^2(Pi/P), z(Pi/P)
  Spline sp_lambda2, sp_z;
  double bulkPi_over_P_max = 0;
};
enum { TAB_C0 = 0, TAB_C1, TAB_C2, TAB_C3, TAB_C4, TAB_F, TAB_G, TAB_BETABULK, TAB_BETAV, TAB_BETAPI };

// reference readindata.h:93-119
struct DfCoeff {
  double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, shear14_coeff = 0;
  double F = 0, G = 0, betabulk = 0, betaV = 0, betapi = 0;
  double lambda = 0, z = 0, delta_lambda = 0, delta_z = 0;
};

// calculate_bilinear, DeltafData.cpp:404-417
IS3D_HD double bilinear(const double *f, int n_T, double T, double muB, double TL, double TR, double muBL, double muBR,
                        int iTL, int iTR, int iBL, int iBR, double dT, double dmuB)
{
  double f_LL = f[iBL * n_T + iTL];
  double f_LR = f[iBR * n_T + iTL];
  double f_RL = f[iBL * n_T + iTR];
  double f_RR = f[iBR * n_T + iTR];
  return ((f_LL * (TR - T) + f_RL * (T - TL)) * (muBR - muB) + (f_LR * (TR - T) + f_RR * (T - TL)) * (muB - muBL)) / (dT * dmuB);
}

// Returns false if the cell is outside the tables (the reference aborts there).
IS3D_HD bool evaluate_df_coefficients(const DfTables &tb, int df_mode, int include_baryon, double T, double muB,
                                      double E, double P, double bulkPi, DfCoeff *out)
{
  DfCoeff df;
  bool ok = true;
  if (!include_baryon) {
    double T4 = T * T * T * T;
    if (df_mode == 1) {                    // DeltafData.cpp:333-346
      double v;
      ok &= spline_eval(tb.sp_c0, T, &v); df.c0 = v / T4;
      ok &= spline_eval(tb.sp_c2, T, &v); df.c2 = v / T4;
      df.c1 = 0.0; df.c3 = 0.0; df.c4 = 0.0;
      df.shear14_coeff = 2.0 * T * T * (E + P);
    } else if (df_mode == 2 || df_mode == 3 || df_mode == 5) {   // :347-362, :385-390
      double v;
      ok &= spline_eval(tb.sp_F, T, &v); df.F = v * T;
      df.G = 0.0;
      ok &= spline_eval(tb.sp_betabulk, T, &v); df.betabulk = v * T4;
      df.betaV = 1.0;
      ok &= spline_eval(tb.sp_betapi, T, &v); df.betapi = v * T4;
    } else {                               // df_mode 4, :363-384
      double l2, v;
      ok &= spline_eval(tb.sp_lambda2, bulkPi / P, &l2);
      // the reference leaves lambda uninitialised for bulkPi == 0 exactly; lambda^2(0) = 0 there, so use 0
      df.lambda = (bulkPi < 0.0) ? -sqrt(l2) : ((bulkPi > 0.0) ? sqrt(l2) : 0.0);
      ok &= spline_eval(tb.sp_z, bulkPi / P, &df.z);
      ok &= spline_eval(tb.sp_betapi, T, &v); df.betapi = v * T4;
      df.delta_lambda = bulkPi / (5.0 * df.betapi - 3.0 * P * (E + P) / E);
      df.delta_z = -3.0 * df.delta_lambda * P / E;
    }
  } else {                                 // bilinear_interpolation, :419-499
    int iTL = (int)floor((T - tb.T_min) / tb.dT);
    int iTR = iTL + 1;
    int iBL = (int)floor((muB - tb.muB_min) / tb.dmuB);
    int iBR = iBL + 1;
    if (!(iTL >= 0 && iTR < tb.n_T) || !(iBL >= 0 && iBR < tb.n_muB)) { *out = df; return false; }
    double TL = tb.T[iTL], TR = tb.T[iTR], muBL = tb.muB[iBL], muBR = tb.muB[iBR];
    double T3 = T * T * T, T4 = T3 * T, T5 = T4 * T;
#define IS3D_BIL(k) bilinear(tb.tab[k], tb.n_T, T, muB, TL, TR, muBL, muBR, iTL, iTR, iBL, iBR, tb.dT, tb.dmuB)
    if (df_mode == 1) {
      df.c0 = IS3D_BIL(TAB_C0) / T4;
      df.c1 = IS3D_BIL(TAB_C1) / T3;
      df.c2 = IS3D_BIL(TAB_C2) / T4;
      df.c3 = IS3D_BIL(TAB_C3) / T4;
      df.c4 = IS3D_BIL(TAB_C4) / T5;
      df.shear14_coeff = 2.0 * T * T * (E + P);
    } else if (df_mode == 2 || df_mode == 3 || df_mode == 5) {
      df.F = IS3D_BIL(TAB_F) * T;
      df.G = IS3D_BIL(TAB_G);
      df.betabulk = IS3D_BIL(TAB_BETABULK) * T4;
      df.betaV = IS3D_BIL(TAB_BETAV) * T3;
      df.betapi = IS3D_BIL(TAB_BETAPI) * T4;
    } else {
      ok = false;                          // PTB has no muB != 0 tables (:480-484); rejected at create time
    }
#undef IS3D_BIL
  }
  *out = df;
  return ok;
}

}  // namespace is3d
