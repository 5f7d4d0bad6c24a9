// TEST INFRASTRUCTURE ONLY -- see cf_oracle.h.  CPU restatement of the reference's Cooper-Frye hot path.
// Scalar, single-threaded, one loop nest per reference function, same loop order (cell -> species -> pT -> phi ->
// y -> eta) and the same association of the floating-point expressions wherever that matters at 1e-10.
#include "cf_oracle.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

const double hbarC = 0.197327053;                               // reference iS3D.h:14
const double two_pi2_hbarC3 = 2.0 * pow(M_PI, 2) * pow(hbarC, 3);   // iS3D.h:16

// ---------------------------------------------------------------------------------------------------------------
// GSL pieces (third party, GSL 2.x): natural cubic spline (interpolation/cspline.c) and 3x3 LU (linalg/lu.c)
// ---------------------------------------------------------------------------------------------------------------
struct Spline {
  std::vector<double> x, y, c;
  void init(const double *xa, const double *ya, int n)
  {
    x.assign(xa, xa + n); y.assign(ya, ya + n); c.assign(n, 0.0);
    int sys = n - 2;
    if (sys <= 0) return;
    std::vector<double> g(sys), diag(sys), off(sys);
    for (int i = 0; i < sys; i++) {
      double h_i = xa[i + 1] - xa[i], h_ip1 = xa[i + 2] - xa[i + 1];
      double yd_i = ya[i + 1] - ya[i], yd_ip1 = ya[i + 2] - ya[i + 1];
      off[i] = h_ip1;
      diag[i] = 2.0 * (h_ip1 + h_i);
      g[i] = 3.0 * (yd_ip1 * (1.0 / h_ip1) - yd_i * (1.0 / h_i));
    }
    if (sys == 1) { c[1] = g[0] / diag[0]; return; }
    std::vector<double> gamma(sys), alpha(sys), z(sys), cc(sys);
    alpha[0] = diag[0]; gamma[0] = off[0] / alpha[0];
    for (int i = 1; i < sys - 1; i++) { alpha[i] = diag[i] - off[i - 1] * gamma[i - 1]; gamma[i] = off[i] / alpha[i]; }
    alpha[sys - 1] = diag[sys - 1] - off[sys - 2] * gamma[sys - 2];
    z[0] = g[0];
    for (int i = 1; i < sys; i++) z[i] = g[i] - gamma[i - 1] * z[i - 1];
    for (int i = 0; i < sys; i++) cc[i] = z[i] / alpha[i];
    c[sys] = cc[sys - 1];
    for (int i = sys - 2; i >= 0; i--) c[i + 1] = cc[i] - gamma[i] * c[i + 2];
  }
  // false = domain error (the reference aborts in gsl_spline_eval)
  bool eval(double v, double *out) const
  {
    int n = (int)x.size();
    if (!(v >= x[0] && v <= x[n - 1])) return false;
    int lo = 0, hi = n - 1;
    while (hi > lo + 1) { int i = (hi + lo) / 2; if (x[i] > v) hi = i; else lo = i; }
    double dx = x[lo + 1] - x[lo], dy = y[lo + 1] - y[lo];
    double b = (dy / dx) - dx * (c[lo + 1] + 2.0 * c[lo]) / 3.0;
    double d = (c[lo + 1] - c[lo]) / (3.0 * dx);
    double delta = v - x[lo];
    *out = y[lo] + delta * (b + delta * (c[lo] + delta * d));
    return true;
  }
};

// LU decomposition with partial pivoting + inverse, as gsl_linalg_LU_decomp / LU_invert do it
void lu_invert3(const double Ain[9], double inv[9])
{
  double A[3][3];
  int perm[3] = {0, 1, 2};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) A[i][j] = Ain[3 * i + j];
  for (int j = 0; j < 2; j++) {
    double max = fabs(A[j][j]); int ip = j;
    for (int i = j + 1; i < 3; i++) if (fabs(A[i][j]) > max) { max = fabs(A[i][j]); ip = i; }
    if (ip != j) { for (int k = 0; k < 3; k++) { double t = A[j][k]; A[j][k] = A[ip][k]; A[ip][k] = t; } int t = perm[j]; perm[j] = perm[ip]; perm[ip] = t; }
    if (A[j][j] != 0.0)
      for (int i = j + 1; i < 3; i++) {
        double aij = A[i][j] / A[j][j];
        A[i][j] = aij;
        for (int k = j + 1; k < 3; k++) A[i][k] -= aij * A[j][k];
      }
  }
  for (int col = 0; col < 3; col++) {
    double b[3];
    for (int i = 0; i < 3; i++) b[i] = (perm[i] == col) ? 1.0 : 0.0;
    for (int i = 0; i < 3; i++) for (int j = 0; j < i; j++) b[i] -= A[i][j] * b[j];
    for (int i = 2; i >= 0; i--) { for (int j = i + 1; j < 3; j++) b[i] -= A[i][j] * b[j]; b[i] /= A[i][i]; }
    for (int i = 0; i < 3; i++) inv[3 * i + col] = b[i];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// df coefficients: Deltaf_Data::evaluate_df_coefficients (DeltafData.cpp:324-519)
// ---------------------------------------------------------------------------------------------------------------
struct DfCoeff {
  double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, shear14_coeff = 0;
  double F = 0, G = 0, betabulk = 0, betaV = 0, betapi = 0;
  double lambda = 0, z = 0, delta_lambda = 0, delta_z = 0;
};

struct DfData {
  const cf_inputs *in;
  int df_mode, include_baryon;
  Spline c0, c2, F, betabulk, betapi, lambda2, z;
  double T_min, muB_min, dT, dmuB;
  DfData(const cf_params *p, const cf_inputs *in_) : in(in_), df_mode(p->df_mode), include_baryon(p->include_baryon)
  {
    T_min = in->T_arr[0]; muB_min = in->muB_arr ? in->muB_arr[0] : 0.0;
    dT = fabs(in->T_arr[1] - in->T_arr[0]);
    dmuB = (in->n_muB > 1) ? fabs(in->muB_arr[1] - in->muB_arr[0]) : 0.0;
    if (!include_baryon) {     // construct_cubic_splines, DeltafData.cpp:298-321 (muB = 0 row)
      c0.init(in->T_arr, in->c0, in->n_T); c2.init(in->T_arr, in->c2, in->n_T);
      F.init(in->T_arr, in->F, in->n_T); betabulk.init(in->T_arr, in->betabulk, in->n_T); betapi.init(in->T_arr, in->betapi, in->n_T);
      if (in->n_ptb > 0) { lambda2.init(in->ptb_x, in->ptb_lambda2, in->n_ptb); z.init(in->ptb_x, in->ptb_z, in->n_ptb); }
    }
  }
  double bil(const double *f, double T, double muB, int iTL, int iBL) const   // calculate_bilinear, :404-417
  {
    int iTR = iTL + 1, iBR = iBL + 1, nT = in->n_T;
    double TL = in->T_arr[iTL], TR = in->T_arr[iTR], muBL = in->muB_arr[iBL], muBR = in->muB_arr[iBR];
    double f_LL = f[iBL * nT + iTL], f_LR = f[iBR * nT + iTL], f_RL = f[iBL * nT + iTR], f_RR = f[iBR * nT + iTR];
    return ((f_LL * (TR - T) + f_RL * (T - TL)) * (muBR - muB) + (f_LR * (TR - T) + f_RR * (T - TL)) * (muB - muBL)) / (dT * dmuB);
  }
  bool evaluate(double T, double muB, double E, double P, double bulkPi, DfCoeff *out) const
  {
    DfCoeff df;
    bool ok = true;
    if (!include_baryon) {                      // cubic_spline, :324-402
      double T4 = T * T * T * T, v;
      switch (df_mode) {
        case 1:
          ok &= c0.eval(T, &v); df.c0 = v / T4;
          ok &= c2.eval(T, &v); df.c2 = v / T4;
          df.shear14_coeff = 2.0 * T * T * (E + P);
          break;
        case 2: case 3: case 5:
          ok &= F.eval(T, &v); df.F = v * T;
          ok &= betabulk.eval(T, &v); df.betabulk = v * T4;
          df.betaV = 1.0;
          ok &= betapi.eval(T, &v); df.betapi = v * T4;
          break;
        case 4: {
          double l2 = 0;
          ok &= lambda2.eval(bulkPi / P, &l2);
          if (bulkPi < 0.0) df.lambda = -sqrt(l2); else if (bulkPi > 0.0) df.lambda = sqrt(l2);   // 0 left as 0
          ok &= z.eval(bulkPi / P, &df.z);
          ok &= betapi.eval(T, &v); df.betapi = v * T4;
          df.delta_lambda = bulkPi / (5.0 * df.betapi - 3.0 * P * (E + P) / E);
          df.delta_z = -3.0 * df.delta_lambda * P / E;
          break;
        }
      }
    } else {                                    // bilinear_interpolation, :419-499
      int iTL = (int)floor((T - T_min) / dT), iBL = (int)floor((muB - muB_min) / dmuB);
      if (!(iTL >= 0 && iTL + 1 < in->n_T) || !(iBL >= 0 && iBL + 1 < in->n_muB)) { *out = df; return false; }
      double T3 = T * T * T, T4 = T3 * T, T5 = T4 * T;
      if (df_mode == 1) {
        df.c0 = bil(in->c0, T, muB, iTL, iBL) / T4; df.c1 = bil(in->c1, T, muB, iTL, iBL) / T3;
        df.c2 = bil(in->c2, T, muB, iTL, iBL) / T4; df.c3 = bil(in->c3, T, muB, iTL, iBL) / T4;
        df.c4 = bil(in->c4, T, muB, iTL, iBL) / T5;
        df.shear14_coeff = 2.0 * T * T * (E + P);
      } else if (df_mode == 2 || df_mode == 3 || df_mode == 5) {
        df.F = bil(in->F, T, muB, iTL, iBL) * T; df.G = bil(in->G, T, muB, iTL, iBL);
        df.betabulk = bil(in->betabulk, T, muB, iTL, iBL) * T4; df.betaV = bil(in->betaV, T, muB, iTL, iBL) * T3;
        df.betapi = bil(in->betapi, T, muB, iTL, iBL) * T4;
      } else ok = false;
    }
    *out = df;
    return ok;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// GaussThermal.cpp:7-85
// ---------------------------------------------------------------------------------------------------------------
typedef double (*thermal_fn)(double, double, double, double, double);
double neq_int(double pbar, double mbar, double alphaB, double baryon, double sign)
{ double Ebar = sqrt(pbar * pbar + mbar * mbar); return pbar * exp(pbar) / (exp(Ebar - baryon * alphaB) + sign); }
double J10_int(double pbar, double mbar, double alphaB, double baryon, double sign)
{ double Ebar = sqrt(pbar * pbar + mbar * mbar), q = exp(Ebar - baryon * alphaB) + sign; return pbar * exp(pbar + Ebar - baryon * alphaB) / (q * q); }
double J20_int(double pbar, double mbar, double alphaB, double baryon, double sign)
{ double Ebar = sqrt(pbar * pbar + mbar * mbar), q = exp(Ebar - baryon * alphaB) + sign; return Ebar * exp(pbar + Ebar - baryon * alphaB) / (q * q); }
double GaussThermal(thermal_fn f, const double *root, const double *weight, int pts, double mbar, double alphaB, double baryon, double sign)
{ double s = 0.0; for (int k = 0; k < pts; k++) s += weight[k] * f(root[k], mbar, alphaB, baryon, sign); return s; }

// ---------------------------------------------------------------------------------------------------------------
// per-cell pieces shared by all paths
// ---------------------------------------------------------------------------------------------------------------
struct CellState {
  double tau, tau2, eta, x, y, dat, dax, day, dan, ux, uy, un, ut, ut2, ux2, uy2, uperp, utperp, T, P, E;
  double pitt, pitx, pity, pitn, pixx, pixy, pixn, piyy, piyn, pinn, bulkPi;
  double muB, alphaB, nB, Vt, Vx, Vy, Vn, baryon_enthalpy_ratio;
};

// MomentumSpectra.cpp:109-187 (and the identical prologues at :516-599, :1159-1225)
// df_variant: the df_mode 1,2 prologue forms u^tau and pi^{eta eta} through utperp and tau2_un (:124-160); the
// feqmod / famod prologues write the same quantities slightly differently (:531, :564-568)
bool load_cell(const cf_params *p, const cf_inputs *in, long i, bool always_shear_bulk, bool baryon_needs_diff, CellState *c,
               bool df_variant = false)
{
  c->tau = in->col[0][i]; c->tau2 = c->tau * c->tau;
  c->x = in->col[1][i]; c->y = in->col[2][i]; c->eta = in->col[3][i];
  c->dat = in->col[4][i]; c->dax = in->col[5][i]; c->day = in->col[6][i]; c->dan = in->col[7][i];
  c->ux = in->col[8][i]; c->uy = in->col[9][i]; c->un = in->col[10][i];
  c->ux2 = c->ux * c->ux; c->uy2 = c->uy * c->uy;
  c->uperp = sqrt(c->ux * c->ux + c->uy * c->uy);
  c->utperp = sqrt(1.0 + c->ux * c->ux + c->uy * c->uy);
  c->ut = sqrt(1.0 + c->ux * c->ux + c->uy * c->uy + c->tau2 * c->un * c->un);
  if (df_variant) { double tau2_un = c->tau2 * c->un; c->ut = sqrt(c->utperp * c->utperp + tau2_un * c->un); }
  c->ut2 = c->ut * c->ut;
  if (c->ut * c->dat + c->ux * c->dax + c->uy * c->day + c->un * c->dan <= 0.0) return false;
  c->E = in->col[11][i]; c->T = in->col[12][i]; c->P = in->col[13][i];
  c->pitt = c->pitx = c->pity = c->pitn = c->pixx = c->pixy = c->pixn = c->piyy = c->piyn = c->pinn = 0.0;
  if (p->include_shear_deltaf || always_shear_bulk) {
    c->pixx = in->col[14][i]; c->pixy = in->col[15][i]; c->pixn = in->col[16][i]; c->piyy = in->col[17][i]; c->piyn = in->col[18][i];
    double tau2 = c->tau2, un = c->un, ux = c->ux, uy = c->uy, ut = c->ut;
    if (df_variant) {
      double tau2_un = tau2 * un;
      c->pinn = (c->pixx * (c->ux2 - c->ut2) + c->piyy * (c->uy2 - c->ut2) + 2.0 * (c->pixy * ux * uy + tau2_un * (c->pixn * ux + c->piyn * uy))) / (tau2 * c->utperp * c->utperp);
      c->pitn = (c->pixn * ux + c->piyn * uy + tau2_un * c->pinn) / ut;
      c->pity = (c->pixy * ux + c->piyy * uy + tau2_un * c->piyn) / ut;
      c->pitx = (c->pixx * ux + c->pixy * uy + tau2_un * c->pixn) / ut;
      c->pitt = (c->pitx * ux + c->pity * uy + tau2_un * c->pitn) / ut;
    } else {
      c->pinn = (c->pixx * (c->ux2 - c->ut2) + c->piyy * (c->uy2 - c->ut2) + 2.0 * (c->pixy * ux * uy + tau2 * un * (c->pixn * ux + c->piyn * uy))) / (tau2 * c->utperp * c->utperp);
      c->pitn = (c->pixn * ux + c->piyn * uy + tau2 * c->pinn * un) / ut;
      c->pity = (c->pixy * ux + c->piyy * uy + tau2 * c->piyn * un) / ut;
      c->pitx = (c->pixx * ux + c->pixy * uy + tau2 * c->pixn * un) / ut;
      c->pitt = (c->pitx * ux + c->pity * uy + tau2 * c->pitn * un) / ut;
    }
  }
  c->bulkPi = (p->include_bulk_deltaf || always_shear_bulk) ? in->col[19][i] : 0.0;
  c->muB = c->alphaB = c->nB = c->Vt = c->Vx = c->Vy = c->Vn = c->baryon_enthalpy_ratio = 0.0;
  if (baryon_needs_diff) {
    if (p->include_baryon && p->include_baryondiff_deltaf) {
      c->muB = in->col[20][i]; c->nB = in->col[21][i]; c->Vx = in->col[22][i]; c->Vy = in->col[23][i]; c->Vn = in->col[24][i];
      c->Vt = df_variant ? (c->Vx * c->ux + c->Vy * c->uy + c->Vn * (c->tau2 * c->un)) / c->ut
                         : (c->Vx * c->ux + c->Vy * c->uy + c->tau2 * c->Vn * c->un) / c->ut;
      c->alphaB = c->muB / c->T;
      c->baryon_enthalpy_ratio = c->nB / (c->E + c->P);
    }
  } else if (p->include_baryon) {              // famod: muB alone (MomentumSpectra.cpp:1212-1225)
    c->muB = in->col[20][i];
    if (p->include_baryondiff_deltaf) {
      c->Vx = in->col[22][i]; c->Vy = in->col[23][i]; c->Vn = in->col[24][i];
      c->Vt = (c->Vx * c->ux + c->Vy * c->uy + c->tau2 * c->Vn * c->un) / c->ut;
    }
    c->alphaB = c->muB / c->T;
  }
  return true;
}

// Milne_Basis (LocalRestFrame.cpp:12-41) and boosts (:133-154, :173-185)
struct Basis { double Xt, Xx, Xy, Xn, Yx, Yy, Zt, Zn; };
Basis milne_basis(const CellState &c)
{
  Basis b;
  double sinhL = c.tau * c.un / c.utperp, coshL = c.ut / c.utperp;
  b.Xt = c.uperp * coshL; b.Xx = 1; b.Xy = 0; b.Xn = c.uperp * sinhL / c.tau;
  b.Yx = 0; b.Yy = 1; b.Zt = sinhL; b.Zn = coshL / c.tau;
  if (c.uperp > 1.e-5) { b.Xx = c.utperp * c.ux / c.uperp; b.Xy = c.utperp * c.uy / c.uperp; b.Yx = -c.uy / c.uperp; b.Yy = c.ux / c.uperp; }
  return b;
}
struct PiLRF { double xx, xy, xz, yy, yz, zz; };
PiLRF boost_pi(const CellState &c, const Basis &b)
{
  PiLRF l;
  double tau2 = c.tau2, Xt = b.Xt, Xx = b.Xx, Xy = b.Xy, Xn = b.Xn, Yx = b.Yx, Yy = b.Yy, Zt = b.Zt, Zn = b.Zn;
  l.xx = c.pitt * Xt * Xt + c.pixx * Xx * Xx + c.piyy * Xy * Xy + tau2 * tau2 * c.pinn * Xn * Xn
       + 2.0 * (-Xt * (c.pitx * Xx + c.pity * Xy) + c.pixy * Xx * Xy + tau2 * Xn * (c.pixn * Xx + c.piyn * Xy - c.pitn * Xt));
  l.xy = Yx * (-c.pitx * Xt + c.pixx * Xx + c.pixy * Xy + tau2 * c.pixn * Xn) + Yy * (-c.pity * Xt + c.pixy * Xx + c.piyy * Xy + tau2 * c.piyn * Xn);
  l.xz = Zt * (c.pitt * Xt - c.pitx * Xx - c.pity * Xy - tau2 * c.pitn * Xn) - tau2 * Zn * (c.pitn * Xt - c.pixn * Xx - c.piyn * Xy - tau2 * c.pinn * Xn);
  l.yy = c.pixx * Yx * Yx + 2.0 * c.pixy * Yx * Yy + c.piyy * Yy * Yy;
  l.yz = -Zt * (c.pitx * Yx + c.pity * Yy) + tau2 * Zn * (c.pixn * Yx + c.piyn * Yy);
  l.zz = -(l.xx + l.yy);
  return l;
}

struct Grids {
  std::vector<double> cosphi, sinphi, phiw, pT, pTw, y, eta, etaw;
  int ny, neta;
  Grids(const cf_params *p, const cf_inputs *in)
  {
    for (int i = 0; i < in->n_phi; i++) { cosphi.push_back(cos(in->phi[i])); sinphi.push_back(sin(in->phi[i])); phiw.push_back(in->phi_w ? in->phi_w[i] : 1.0); }
    for (int i = 0; i < in->n_pT; i++) { pT.push_back(in->pT[i]); pTw.push_back(in->pT_w ? in->pT_w[i] : 1.0); }
    if (p->dimension == 2) {                   // MomentumSpectra.cpp:73-82
      y.assign(1, 0.0); ny = 1; neta = in->n_eta;
      eta.assign(in->eta, in->eta + in->n_eta); etaw.assign(in->eta_w, in->eta_w + in->n_eta);
    } else {                                   // :83-91 (eta value taken from the cell)
      y.assign(in->y, in->y + in->n_y); ny = in->n_y; neta = 1;
      eta.assign(1, 0.0); etaw.assign(1, 1.0);
    }
  }
};

// does_feqmod_breakdown with fast = 0 (EmissionFunction.cpp:65-109)
bool does_feqmod_breakdown(const cf_inputs *in, double mass_pion0, double T, double F, double bulkPi, double betabulk, double detA,
                           double detA_min, double z, int df_mode)
{
  if (df_mode == 3) {
    const double *r1 = in->gla_root + 1 * in->n_gla, *w1 = in->gla_weight + 1 * in->n_gla;
    const double *r2 = in->gla_root + 2 * in->n_gla, *w2 = in->gla_weight + 2 * in->n_gla;
    double mbar_pion0 = mass_pion0 / T;
    double neq_fact = T * T * T / two_pi2_hbarC3, J20_fact = T * neq_fact;
    double neq_pion0 = neq_fact * GaussThermal(neq_int, r1, w1, in->n_gla, mbar_pion0, 0., 0., -1.);
    double J20_pion0 = J20_fact * GaussThermal(J20_int, r2, w2, in->n_gla, mbar_pion0, 0., 0., -1.);
    double dn_pion0 = bulkPi * (neq_pion0 + J20_pion0 * F / T / T) / betabulk;     // is_linear_pion0_density_negative, :52-63
    bool negative = (neq_pion0 + dn_pion0 < 0.0);
    if (detA <= detA_min || negative) return true;
  } else if (df_mode == 4) {
    if (detA <= detA_min || z < 0.0) return true;
  }
  return false;
}

// ---------------------------------------------------------------------------------------------------------------
// calculate_dN_pTdpTdphidy, df_mode 1,2 (MomentumSpectra.cpp:32-415)
// ---------------------------------------------------------------------------------------------------------------
int spectra_df(const cf_params *p, const cf_inputs *in, double *out, cf_stats *st)
{
  const double prefactor = pow(2.0 * M_PI * hbarC, -3);
  Grids g(p, in);
  DfData dfd(p, in);
  const int npart = in->n_species, npT = in->n_pT, nphi = in->n_phi, ny = g.ny;
  for (long icell = 0; icell < in->n_cells; icell++) {
    CellState c;
    if (!load_cell(p, in, icell, false, true, &c, true)) { st->cells_skipped++; continue; }
    if (p->dimension == 3) g.eta[0] = c.eta;
    double tau2_un = c.tau2 * c.un;
    double tau2_pitn = c.tau2 * c.pitn, tau2_pixn = c.tau2 * c.pixn, tau2_piyn = c.tau2 * c.piyn;
    double tau4_pinn = c.tau2 * c.tau2 * c.pinn, tau2_Vn = c.tau2 * c.Vn;
    DfCoeff df;
    if (!dfd.evaluate(c.T, c.muB, c.E, c.P, c.bulkPi, &df)) { st->cells_out_of_table++; return 3; }
    double shear_coeff = 0, bulk0 = 0, bulk1 = 0, bulk2 = 0, diff0 = 0, diff1 = 0;
    if (p->df_mode == 1) {
      shear_coeff = 1.0 / df.shear14_coeff;
      bulk0 = (df.c0 - df.c2) * c.bulkPi; bulk1 = df.c1 * c.bulkPi; bulk2 = (4. * df.c2 - df.c0) * c.bulkPi;
      diff0 = df.c3; diff1 = df.c4;
    } else {
      shear_coeff = 0.5 / (df.betapi * c.T);
      bulk0 = df.F / (c.T * c.T * df.betabulk) * c.bulkPi; bulk1 = df.G / df.betabulk * c.bulkPi;
      bulk2 = c.bulkPi / (3.0 * c.T * df.betabulk);
      diff0 = c.baryon_enthalpy_ratio / df.betaV; diff1 = 1.0 / df.betaV;
    }
    for (int ipart = 0; ipart < npart; ipart++) {
      double mass = in->mass[ipart], mass_squared = mass * mass, sign = in->sign[ipart], degeneracy = in->degeneracy[ipart];
      double baryon = in->baryon[ipart], chem = baryon * c.alphaB;
      for (int ipT = 0; ipT < npT; ipT++) {
        double pT = g.pT[ipT], mT = sqrt(mass_squared + pT * pT), mT_over_tau = mT / c.tau;
        for (int iphip = 0; iphip < nphi; iphip++) {
          double px = pT * g.cosphi[iphip], py = pT * g.sinphi[iphip];
          for (int iy = 0; iy < ny; iy++) {
            double y = g.y[iy], eta_integral = 0.0;
            for (int ieta = 0; ieta < g.neta; ieta++) {
              double eta = g.eta[ieta], eta_weight = g.etaw[ieta];
              double sinhyeta = sinh(y - eta), coshyeta = sqrt(1.0 + sinhyeta * sinhyeta);
              double pt = mT * coshyeta, pn = mT_over_tau * sinhyeta;
              double pdotdsigma = pt * c.dat + px * c.dax + py * c.day + pn * c.dan;
              if (p->outflow && pdotdsigma <= 0.0) continue;
              double E = pt * c.ut - px * c.ux - py * c.uy - pn * tau2_un;
              double feq = 1.0 / (exp(E / c.T - chem) + sign);
              double feqbar = 1.0 - sign * feq;
              double pimunu_pmu_pnu = c.pitt * pt * pt + c.pixx * px * px + c.piyy * py * py + tau4_pinn * pn * pn
                  + 2.0 * (-(c.pitx * px + c.pity * py) * pt + c.pixy * px * py + pn * (tau2_pixn * px + tau2_piyn * py - tau2_pitn * pt));
              double Vmu_pmu = c.Vt * pt - c.Vx * px - c.Vy * py - tau2_Vn * pn;
              double dfv;
              if (p->df_mode == 1) {
                double df_shear = shear_coeff * pimunu_pmu_pnu;
                double df_bulk = bulk0 * mass_squared + (bulk1 * baryon + bulk2 * E) * E;
                double df_diff = (diff0 * baryon + diff1 * E) * Vmu_pmu;
                dfv = feqbar * (df_shear + df_bulk + df_diff);
              } else {
                double df_shear = shear_coeff * pimunu_pmu_pnu / E;
                double df_bulk = bulk0 * E + bulk1 * baryon + bulk2 * (E - mass_squared / E);
                double df_diff = (diff0 - diff1 * baryon / E) * Vmu_pmu;
                dfv = feqbar * (df_shear + df_bulk + df_diff);
              }
              if (p->regulate_deltaf) dfv = fmax(-1.0, fmin(dfv, 1.0));
              eta_integral += eta_weight * pdotdsigma * feq * (1.0 + dfv);
            }
            out[iy + (long)ny * (iphip + (long)nphi * (ipT + (long)npT * ipart))] += prefactor * degeneracy * eta_integral;
          }
        }
      }
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// calculate_dN_pTdpTdphidy_feqmod, df_mode 3,4 (MomentumSpectra.cpp:419-1044)
// ---------------------------------------------------------------------------------------------------------------
int spectra_feqmod(const cf_params *p, const cf_inputs *in, double *out, cf_stats *st)
{
  const double prefactor = pow(2.0 * M_PI * hbarC, -3);
  Grids g(p, in);
  DfData dfd(p, in);
  const int npart = in->n_species, npT = in->n_pT, nphi = in->n_phi, ny = g.ny, pts = in->n_gla;
  const double *r1 = in->gla_root + 1 * pts, *w1 = in->gla_weight + 1 * pts, *r2 = in->gla_root + 2 * pts, *w2 = in->gla_weight + 2 * pts;
  const double detA_min = p->deta_min;
  for (long icell = 0; icell < in->n_cells; icell++) {
    CellState c;
    if (!load_cell(p, in, icell, false, true, &c)) { st->cells_skipped++; continue; }
    if (p->dimension == 3) g.eta[0] = c.eta;
    double T = c.T, P = c.P, E = c.E, bulkPi = c.bulkPi, tau2 = c.tau2;
    if (p->df_mode == 4) {                       // :603-615
      if (bulkPi < -P) bulkPi = -(1.0 - 1.e-5) * P;
      else if (bulkPi / P > in->ptb_x_max) bulkPi = P * (in->ptb_x_max - 1.e-5);
    }
    double zt = c.tau * c.un / c.utperp, zn = c.ut / (c.tau * c.utperp);
    double pl = P + bulkPi + zt * zt * c.pitt + tau2 * tau2 * zn * zn * c.pinn + 2. * tau2 * zt * zn * c.pitn;
    if (pl < 0) st->cells_pl_negative++;
    DfCoeff df;
    if (!dfd.evaluate(T, c.muB, E, P, bulkPi, &df)) { st->cells_out_of_table++; return 3; }
    Basis b = milne_basis(c);
    PiLRF pl_ = boost_pi(c, b);
    double T_mod = T, alphaB_mod = c.alphaB;
    if (p->df_mode == 3) { T_mod = T + bulkPi * df.F / df.betabulk; alphaB_mod = c.alphaB + bulkPi * df.G / df.betabulk; }
    double shear_coeff = 0.5 / (df.betapi * T), bulk0 = df.F / (T * T * df.betabulk), bulk1 = df.G / df.betabulk, bulk2 = 1.0 / (3.0 * T * df.betabulk);
    double shear_mod = 0.5 / df.betapi, bulk_mod = bulkPi / (3.0 * df.betabulk);
    if (p->df_mode == 4) bulk_mod = df.lambda;
    double Axx = 1.0 + pl_.xx * shear_mod + bulk_mod, Axy = pl_.xy * shear_mod, Axz = pl_.xz * shear_mod;
    double Ayy = 1.0 + pl_.yy * shear_mod + bulk_mod, Ayz = pl_.yz * shear_mod, Azz = 1.0 + pl_.zz * shear_mod + bulk_mod;
    double detA = Axx * (Ayy * Azz - Ayz * Ayz) - Axy * (Axy * Azz - Ayz * Axz) + Axz * (Axy * Ayz - Ayy * Axz);
    double detA_bulk_two_thirds = pow(1.0 + bulk_mod, 2);
    double A[9] = {Axx, Axy, Axz, Axy, Ayy, Ayz, Axz, Ayz, Azz}, A_inv[9];
    lu_invert3(A, A_inv);
    double neq_fact = T * T * T / two_pi2_hbarC3, dn_fact = bulkPi / df.betabulk, J20_fact = T * neq_fact, N10_fact = neq_fact;
    double nmod_fact = T_mod * T_mod * T_mod / two_pi2_hbarC3;
    bool breaks = does_feqmod_breakdown(in, p->mass_pion0, T, df.F, bulkPi, df.betabulk, detA, detA_min, df.z, p->df_mode);
    if (breaks) st->cells_breakdown++;
    double eta_scale = 1.0;
    if (detA > detA_min && p->dimension == 2) eta_scale = detA / detA_bulk_two_thirds;
    for (int ipart = 0; ipart < npart; ipart++) {
      double mass = in->mass[ipart], mass2 = mass * mass, sign = in->sign[ipart], degeneracy = in->degeneracy[ipart], baryon = in->baryon[ipart];
      double chem = baryon * c.alphaB, chem_mod = baryon * alphaB_mod;
      double renorm = 1.0;
      if (p->include_bulk_deltaf) {
        if (p->df_mode == 3) {                   // :795-811
          double mbar = mass / T, mbar_mod = mass / T_mod;
          double neq = neq_fact * degeneracy * GaussThermal(neq_int, r1, w1, pts, mbar, c.alphaB, baryon, sign);
          double N10 = baryon * N10_fact * degeneracy * GaussThermal(J10_int, r1, w1, pts, mbar, c.alphaB, baryon, sign);
          double J20 = J20_fact * degeneracy * GaussThermal(J20_int, r2, w2, pts, mbar, c.alphaB, baryon, sign);
          double n_linear = neq + dn_fact * (neq + N10 * df.G + J20 * df.F / T / T);
          double n_mod = nmod_fact * degeneracy * GaussThermal(neq_int, r1, w1, pts, mbar_mod, alphaB_mod, baryon, sign);
          renorm = n_linear / n_mod;
        } else renorm = df.z;
      }
      if (p->dimension == 2) renorm /= detA_bulk_two_thirds; else renorm /= detA;
      if (std::isnan(renorm) || std::isinf(renorm)) continue;
      for (int ipT = 0; ipT < npT; ipT++) {
        double pT = g.pT[ipT], mT = sqrt(mass2 + pT * pT), mT_over_tau = mT / c.tau;
        for (int iphip = 0; iphip < nphi; iphip++) {
          double px = pT * g.cosphi[iphip], py = pT * g.sinphi[iphip];
          for (int iy = 0; iy < ny; iy++) {
            double y = g.y[iy], eta_integral = 0.0;
            for (int ieta = 0; ieta < g.neta; ieta++) {
              double eta = g.eta[ieta], eta_weight = g.etaw[ieta];
              bool narrow = false;
              if (p->dimension == 3 && !breaks) if (detA < 0.01 && fabs(y - eta) < detA) narrow = true;
              double pdotdsigma, f;
              if (breaks || narrow) {
                double pt = mT * cosh(y - eta), pn = mT_over_tau * sinh(y - eta), tau2_pn = tau2 * pn;
                pdotdsigma = eta_weight * (pt * c.dat + px * c.dax + py * c.day) + pn * c.dan;
                if (p->outflow && pdotdsigma <= 0.0) continue;
                double pdotu = pt * c.ut - px * c.ux - py * c.uy - tau2_pn * c.un;
                double pimunu_pmu_pnu = c.pitt * pt * pt + c.pixx * px * px + c.piyy * py * py + c.pinn * tau2_pn * tau2_pn
                    + 2.0 * (-(c.pitx * px + c.pity * py) * pt + c.pixy * px * py + tau2_pn * (c.pixn * px + c.piyn * py - c.pitn * pt));
                double dfv;
                if (p->df_mode == 3) {
                  double feq = 1.0 / (exp(pdotu / T - chem) + sign), feqbar = 1.0 - sign * feq;
                  double Vmu_pmu = c.Vt * pt - c.Vx * px - c.Vy * py - c.Vn * tau2_pn;
                  double df_shear = shear_coeff * pimunu_pmu_pnu / pdotu;
                  double df_bulk = (bulk0 * pdotu + bulk1 * baryon + bulk2 * (pdotu - mass2 / pdotu)) * bulkPi;
                  double df_diff = (c.baryon_enthalpy_ratio - baryon / pdotu) * Vmu_pmu / df.betaV;
                  dfv = feqbar * (df_shear + df_bulk + df_diff);
                  if (p->regulate_deltaf) dfv = fmax(-1.0, fmin(dfv, 1.0));
                  f = feq * (1.0 + dfv);
                } else {
                  double feq = 1.0 / (exp(pdotu / T) + sign), feqbar = 1.0 - sign * feq;
                  double df_shear = feqbar * shear_coeff * pimunu_pmu_pnu / pdotu;
                  double df_bulk = df.delta_z - 3.0 * df.delta_lambda + feqbar * df.delta_lambda * (pdotu - mass2 / pdotu) / T;
                  dfv = df_shear + df_bulk;
                  if (p->regulate_deltaf) dfv = fmax(-1.0, fmin(dfv, 1.0));
                  f = feq * (1.0 + dfv);
                }
              } else {
                double pt = mT * cosh(y - eta_scale * eta), pn = mT_over_tau * sinh(y - eta_scale * eta), tau2_pn = tau2 * pn;
                pdotdsigma = eta_weight * (pt * c.dat + px * c.dax + py * c.day) + pn * c.dan;
                if (p->outflow && pdotdsigma <= 0.0) continue;
                double pLRF[3] = {-b.Xt * pt + b.Xx * px + b.Xy * py + b.Xn * tau2_pn, b.Yx * px + b.Yy * py, -b.Zt * pt + b.Zn * tau2_pn};
                double pmod[3];
                for (int i = 0; i < 3; i++) pmod[i] = A_inv[3 * i] * pLRF[0] + A_inv[3 * i + 1] * pLRF[1] + A_inv[3 * i + 2] * pLRF[2];
                for (int it = 0; it < 5; it++) {   // iterative refinement, :959-971
                  double prev[3] = {pmod[0], pmod[1], pmod[2]}, back[3], dp[3];
                  for (int i = 0; i < 3; i++) back[i] = A[3 * i] * prev[0] + A[3 * i + 1] * prev[1] + A[3 * i + 2] * prev[2];
                  for (int i = 0; i < 3; i++) dp[i] = pLRF[i] - back[i];
                  if (sqrt(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]) <= 1.e-16) break;
                  for (int i = 0; i < 3; i++) pmod[i] = prev[i] + (A_inv[3 * i] * dp[0] + A_inv[3 * i + 1] * dp[1] + A_inv[3 * i + 2] * dp[2]);
                }
                double E_mod = sqrt(mass2 + pmod[0] * pmod[0] + pmod[1] * pmod[1] + pmod[2] * pmod[2]);
                f = fabs(renorm) / (exp(E_mod / T_mod - chem_mod) + sign);
              }
              eta_integral += pdotdsigma * f;
            }
            out[iy + (long)ny * (iphip + (long)nphi * (ipT + (long)npT * ipart))] += prefactor * degeneracy * eta_integral;
          }
        }
      }
    }
  }
  return 0;
}

int spectra_famod(const cf_params *p, const cf_inputs *in, double *out, cf_stats *st);   // defined below

}  // namespace

extern "C" int cf_oracle_spectra(const cf_params *p, const cf_inputs *in, double *out, cf_stats *st)
{
  cf_stats local;
  if (!st) st = &local;
  memset(st, 0, sizeof(*st));
  const int ny = (p->dimension == 3) ? in->n_y : 1;
  memset(out, 0, sizeof(double) * (size_t)in->n_species * in->n_pT * in->n_phi * ny);
  switch (p->df_mode) {
    case 1: case 2: return spectra_df(p, in, out, st);
    case 3: case 4: return spectra_feqmod(p, in, out, st);
    case 5: return spectra_famod(p, in, out, st);
    default: return 4;
  }
}


// ---------------------------------------------------------------------------------------------------------------
// calculate_dN_dX (df_mode 1,2; SpacetimeDistribution.cpp:31-517) and calculate_dN_dX_feqmod (df_mode 3,4; :520-1246)
// Species-outer loop as in the reference.  The histograms returned are the CLEAN per-species sums; the reference's
// partial memset (:166-168) is a property of its accumulator reuse and is emulated by the caller when needed.
// ---------------------------------------------------------------------------------------------------------------
extern "C" int cf_oracle_dndx(const cf_params *p, const cf_inputs *in, double *tau_hist, double *r_hist, double *phi_hist, cf_stats *st)
{
  cf_stats local;
  if (!st) st = &local;
  memset(st, 0, sizeof(*st));
  if (p->df_mode < 1 || p->df_mode > 4) return 4;
  const bool feqmod = p->df_mode >= 3;
  const double prefactor = pow(2.0 * M_PI * hbarC, -3);
  Grids g(p, in);
  DfData dfd(p, in);
  const int npart = in->n_species, npT = in->n_pT, nphi = in->n_phi, ny = g.ny, pts = in->n_gla;
  const double *r1 = in->gla_root + 1 * pts, *w1 = in->gla_weight + 1 * pts, *r2 = in->gla_root + 2 * pts, *w2 = in->gla_weight + 2 * pts;
  const double TAU_WIDTH = (p->tau_max - p->tau_min) / (double)p->tau_bins, R_WIDTH = (p->r_max - p->r_min) / (double)p->r_bins;
  const double PHIP_WIDTH = 2.0 * M_PI / (double)p->phip_bins;
  memset(tau_hist, 0, sizeof(double) * (size_t)npart * p->tau_bins);
  memset(r_hist, 0, sizeof(double) * (size_t)npart * p->r_bins);
  memset(phi_hist, 0, sizeof(double) * (size_t)npart * p->phip_bins);
  for (int ipart = 0; ipart < npart; ipart++) {
    double mass = in->mass[ipart], mass2 = mass * mass, sign = in->sign[ipart], degeneracy = in->degeneracy[ipart], baryon = in->baryon[ipart];
    for (long icell = 0; icell < in->n_cells; icell++) {
      CellState c;
      if (!load_cell(p, in, icell, false, true, &c)) { if (ipart == 0) st->cells_skipped++; continue; }
      if (p->dimension == 3) g.eta[0] = c.eta;
      double T = c.T, P = c.P, E = c.E, bulkPi = c.bulkPi, tau2 = c.tau2;
      if (p->df_mode == 4) {                     // :784-785 (inclusive comparisons here)
        if (bulkPi <= -P) bulkPi = -(1.0 - 1.e-5) * P;
        else if (bulkPi / P >= in->ptb_x_max) bulkPi = P * (in->ptb_x_max - 1.e-5);
      }
      double chem = baryon * c.alphaB;
      DfCoeff df;
      if (!dfd.evaluate(T, c.muB, E, P, bulkPi, &df)) { st->cells_out_of_table++; return 3; }
      double shear_coeff, bulk0, bulk1, bulk2;
      if (p->df_mode == 1) { shear_coeff = 0.5 / (T * T * (E + P)); bulk0 = df.c0 - df.c2; bulk1 = df.c1; bulk2 = 4.0 * df.c2 - df.c0; }
      else { shear_coeff = 0.5 / (df.betapi * T); bulk0 = df.F / (T * T * df.betabulk); bulk1 = df.G / df.betabulk; bulk2 = 1.0 / (3.0 * T * df.betabulk); }
      // modified-distribution set-up (feqmod only)
      Basis b{};
      double A[9], A_inv[9], detA = 1.0, T_mod = T, alphaB_mod = c.alphaB, eta_scale = 1.0, renorm = 1.0;
      bool breaks = false;
      if (feqmod) {
        b = milne_basis(c);
        PiLRF l = boost_pi(c, b);
        if (p->df_mode == 3) { T_mod = T + bulkPi * df.F / df.betabulk; alphaB_mod = c.alphaB + bulkPi * df.G / df.betabulk; }
        double shear_mod = 0.5 / df.betapi, bulk_mod = (p->df_mode == 4) ? df.lambda : bulkPi / (3.0 * df.betabulk);
        double Axx = 1.0 + l.xx * shear_mod + bulk_mod, Axy = l.xy * shear_mod, Axz = l.xz * shear_mod;
        double Ayy = 1.0 + l.yy * shear_mod + bulk_mod, Ayz = l.yz * shear_mod, Azz = 1.0 + l.zz * shear_mod + bulk_mod;
        detA = Axx * (Ayy * Azz - Ayz * Ayz) - Axy * (Axy * Azz - Ayz * Axz) + Axz * (Axy * Ayz - Ayy * Axz);
        double detA_bulk_two_thirds = pow(1.0 + bulk_mod, 2);
        breaks = does_feqmod_breakdown(in, p->mass_pion0, T, df.F, bulkPi, df.betabulk, detA, p->deta_min, df.z, p->df_mode);
        double Am[9] = {Axx, Axy, Axz, Axy, Ayy, Ayz, Axz, Ayz, Azz};
        memcpy(A, Am, sizeof(A));
        lu_invert3(A, A_inv);
        if (detA > p->deta_min && p->dimension == 2) eta_scale = detA / detA_bulk_two_thirds;
        if (p->include_bulk_deltaf) {
          if (p->df_mode == 3) {
            double neq_fact = T * T * T / two_pi2_hbarC3, dn_fact = bulkPi / df.betabulk, J20_fact = T * neq_fact, N10_fact = neq_fact;
            double nmod_fact = T_mod * T_mod * T_mod / two_pi2_hbarC3, mbar = mass / T, mbar_mod = mass / T_mod;
            double neq = neq_fact * degeneracy * GaussThermal(neq_int, r1, w1, pts, mbar, c.alphaB, baryon, sign);
            double N10 = baryon * N10_fact * degeneracy * GaussThermal(J10_int, r1, w1, pts, mbar, c.alphaB, baryon, sign);
            double J20 = J20_fact * degeneracy * GaussThermal(J20_int, r2, w2, pts, mbar, c.alphaB, baryon, sign);
            double n_linear = neq + dn_fact * (neq + N10 * df.G + J20 * df.F / T / T);
            double n_mod = nmod_fact * degeneracy * GaussThermal(neq_int, r1, w1, pts, mbar_mod, alphaB_mod, baryon, sign);
            renorm = n_linear / n_mod;
          } else renorm = df.z;
        }
        if (p->dimension == 2) renorm /= detA_bulk_two_thirds; else renorm /= detA;
        if (std::isnan(renorm) || std::isinf(renorm)) continue;
      }
      double chem_mod = baryon * alphaB_mod;
      double dN_dy_cell = 0.0;
      for (int ipT = 0; ipT < npT; ipT++) {
        double pT = g.pT[ipT], mT = sqrt(mass2 + pT * pT), mT_over_tau = mT / c.tau, pT_weight = g.pTw[ipT];
        for (int iphip = 0; iphip < nphi; iphip++) {
          double px = pT * g.cosphi[iphip], py = pT * g.sinphi[iphip], phi_weight = g.phiw[iphip];
          for (int iy = 0; iy < ny; iy++) {
            double y = g.y[iy], eta_integral = 0.0;
            for (int ieta = 0; ieta < g.neta; ieta++) {
              double eta = g.eta[ieta], eta_weight = g.etaw[ieta];
              bool linear = !feqmod || breaks;
              if (feqmod && p->dimension == 3 && !breaks && detA < 0.01 && fabs(y - eta) < detA) linear = true;
              double f, pdotdsigma;
              if (linear) {
                double pt = mT * cosh(y - eta), pn = mT_over_tau * sinh(y - eta), tau2_pn = tau2 * pn;
                pdotdsigma = eta_weight * (pt * c.dat + px * c.dax + py * c.day + pn * c.dan);
                if (p->outflow && pdotdsigma <= 0.0) continue;
                double pdotu = pt * c.ut - px * c.ux - py * c.uy - tau2_pn * c.un;
                double pimunu_pmu_pnu = c.pitt * pt * pt + c.pixx * px * px + c.piyy * py * py + c.pinn * tau2_pn * tau2_pn
                    + 2.0 * (-(c.pitx * px + c.pity * py) * pt + c.pixy * px * py + tau2_pn * (c.pixn * px + c.piyn * py - c.pitn * pt));
                double Vmu_pmu = c.Vt * pt - c.Vx * px - c.Vy * py - c.Vn * tau2_pn;
                double dfv;
                if (p->df_mode == 4) {
                  double feq = 1.0 / (exp(pdotu / T) + sign), feqbar = 1.0 - sign * feq;
                  double df_shear = feqbar * shear_coeff * pimunu_pmu_pnu / pdotu;
                  double df_bulk = df.delta_z - 3.0 * df.delta_lambda + feqbar * df.delta_lambda * (pdotu - mass2 / pdotu) / T;
                  dfv = df_shear + df_bulk;
                  if (p->regulate_deltaf) dfv = fmax(-1.0, fmin(dfv, 1.0));
                  f = feq * (1.0 + dfv);
                } else {
                  double feq = 1.0 / (exp(pdotu / T - chem) + sign), feqbar = 1.0 - sign * feq;
                  if (p->df_mode == 1) {
                    double df_shear = shear_coeff * pimunu_pmu_pnu;
                    double df_bulk = (bulk0 * mass2 + (bulk1 * baryon + bulk2 * pdotu) * pdotu) * bulkPi;
                    double df_diff = (df.c3 * baryon + df.c4 * pdotu) * Vmu_pmu;
                    dfv = feqbar * (df_shear + df_bulk + df_diff);
                  } else {
                    double df_shear = shear_coeff * pimunu_pmu_pnu / pdotu;
                    double df_bulk = (bulk0 * pdotu + bulk1 * baryon + bulk2 * (pdotu - mass2 / pdotu)) * bulkPi;
                    double df_diff = (c.baryon_enthalpy_ratio - baryon / pdotu) * Vmu_pmu / df.betaV;
                    dfv = feqbar * (df_shear + df_bulk + df_diff);
                  }
                  if (p->regulate_deltaf) dfv = fmax(-1.0, fmin(dfv, 1.0));
                  f = feq * (1.0 + dfv);
                }
              } else {
                double pt = mT * cosh(y - eta_scale * eta), pn = mT_over_tau * sinh(y - eta_scale * eta), tau2_pn = tau2 * pn;
                pdotdsigma = eta_weight * (pt * c.dat + px * c.dax + py * c.day + pn * c.dan);
                if (p->outflow && pdotdsigma <= 0.0) continue;
                double pLRF[3] = {-b.Xt * pt + b.Xx * px + b.Xy * py + b.Xn * tau2_pn, b.Yx * px + b.Yy * py, -b.Zt * pt + b.Zn * tau2_pn};
                double pmod[3];
                for (int i = 0; i < 3; i++) pmod[i] = A_inv[3 * i] * pLRF[0] + A_inv[3 * i + 1] * pLRF[1] + A_inv[3 * i + 2] * pLRF[2];
                for (int it = 0; it < 5; it++) {
                  double prev[3] = {pmod[0], pmod[1], pmod[2]}, back[3], dp[3];
                  for (int i = 0; i < 3; i++) back[i] = A[3 * i] * prev[0] + A[3 * i + 1] * prev[1] + A[3 * i + 2] * prev[2];
                  for (int i = 0; i < 3; i++) dp[i] = pLRF[i] - back[i];
                  if (sqrt(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]) <= 1.e-16) break;
                  for (int i = 0; i < 3; i++) pmod[i] = prev[i] + (A_inv[3 * i] * dp[0] + A_inv[3 * i + 1] * dp[1] + A_inv[3 * i + 2] * dp[2]);
                }
                double E_mod = sqrt(mass2 + pmod[0] * pmod[0] + pmod[1] * pmod[1] + pmod[2] * pmod[2]);
                f = fabs(renorm) / (exp(E_mod / T_mod - chem_mod) + sign);
              }
              eta_integral += pdotdsigma * f;
            }
            dN_dy_cell += pT_weight * phi_weight * prefactor * degeneracy * eta_integral;
          }
        }
      }
      double r = sqrt(c.x * c.x + c.y * c.y), phi = atan2(c.y, c.x);     // binning, :413-440
      if (phi < 0.0) phi += 2.0 * M_PI;
      long itau = (int)floor((c.tau - p->tau_min) / TAU_WIDTH), ir = (int)floor((r - p->r_min) / R_WIDTH), iphi = (int)floor(phi / PHIP_WIDTH);
      if (itau >= 0 && itau < p->tau_bins) tau_hist[(size_t)ipart * p->tau_bins + itau] += dN_dy_cell;
      if (ir >= 0 && ir < p->r_bins) r_hist[(size_t)ipart * p->r_bins + ir] += dN_dy_cell;
      if (iphi >= 0 && iphi < p->phip_bins) phi_hist[(size_t)ipart * p->phip_bins + iphi] += dN_dy_cell;
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Sampler yields.  calculate_total_yield (ParticleSampler.cpp:447-636) and the per-cell dn_list / dn_tot of
// sample_dN_pTdpTdphidy in fast mode (:680-915, fast_max_particle_number :122-161), df_mode 1-4.
// Note: the reference reads Surface_Element_Vector::dsigma_space in calculate_total_yield without ever calling
// compute_dsigma_magnitude() (:606-609), i.e. an uninitialised value multiplies V.dsigma * dn_diff; the oracle (and
// the product) use the intended sqrt(dsx^2 + dsy^2 + dsz^2).  The term is zero without baryon diffusion.
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct YieldCell { bool valid; double ds_time, ds_space, ds_max, Vdsigma, bulkPi, z, delta_z; bool breaks_yield, breaks_sample;
                   double T, alphaB, F, G, betabulk; };

YieldCell yield_cell(const cf_params *p, const cf_inputs *in, const DfData &dfd, long icell, double F_avg, double betabulk_avg, int *err)
{
  YieldCell y{};
  CellState c;
  if (!load_cell(p, in, icell, false, true, &c)) return y;
  double Vdsigma = c.Vt * c.dat + c.Vx * c.dax + c.Vy * c.day + c.Vn * c.dan;
  double bulkPi = c.bulkPi;
  if (p->df_mode == 4) {
    if (bulkPi <= -c.P) bulkPi = -(1.0 - 1.e-5) * c.P;
    else if (bulkPi / c.P >= in->ptb_x_max) bulkPi = c.P * (in->ptb_x_max - 1.e-5);
  }
  DfCoeff df;
  if (!dfd.evaluate(c.T, c.muB, c.E, c.P, bulkPi, &df)) { *err = 3; return y; }
  Basis b = milne_basis(c);
  double dst = c.dat * c.ut + c.dax * c.ux + c.day * c.uy + c.dan * c.un;
  double dsx = -(c.dat * b.Xt + c.dax * b.Xx + c.day * b.Xy + c.dan * b.Xn);
  double dsy = -(c.dax * b.Yx + c.day * b.Yy);
  double dsz = -(c.dat * b.Zt + c.dan * b.Zn);
  double ds_space = sqrt(dsx * dsx + dsy * dsy + dsz * dsz);
  PiLRF l = boost_pi(c, b);
  double shear_mod = 0, bulk_mod = 0;
  if (p->df_mode == 3) { shear_mod = 0.5 / df.betapi; bulk_mod = bulkPi / (3. * df.betabulk); }
  else if (p->df_mode == 4) { shear_mod = 0.5 / df.betapi; bulk_mod = df.lambda; }
  double Axx = 1.0 + l.xx * shear_mod + bulk_mod, Axy = l.xy * shear_mod, Axz = l.xz * shear_mod;
  double Ayy = 1.0 + l.yy * shear_mod + bulk_mod, Ayz = l.yz * shear_mod, Azz = 1.0 + l.zz * shear_mod + bulk_mod;
  double detA = Axx * (Ayy * Azz - Ayz * Ayz) - Axy * (Axy * Azz - Ayz * Axz) + Axz * (Axy * Ayz - Ayy * Axz);
  y.valid = true; y.ds_time = dst; y.ds_space = ds_space; y.ds_max = fabs(dst) + ds_space; y.Vdsigma = Vdsigma; y.bulkPi = bulkPi;
  y.z = df.z; y.delta_z = df.delta_z;
  y.T = c.T; y.alphaB = c.alphaB; y.F = df.F; y.G = df.G; y.betabulk = df.betabulk;
  y.breaks_yield = does_feqmod_breakdown(in, p->mass_pion0, c.T, df.F, bulkPi, df.betabulk, detA, p->deta_min, df.z, p->df_mode);
  y.breaks_sample = y.breaks_yield;
  if (p->df_mode == 3 && p->fast)   // does_feqmod_breakdown(..., FAST, Tavg, F_avg, betabulk_avg), ParticleSampler.cpp:874
    y.breaks_sample = does_feqmod_breakdown(in, p->mass_pion0, in->T_avg, F_avg, bulkPi, betabulk_avg, detA, p->deta_min, df.z, p->df_mode);
  return y;
}
}  // namespace

extern "C" int cf_oracle_total_yield(const cf_params *p, const cf_inputs *in, double *ntotal)
{
  if (p->df_mode < 1 || p->df_mode > 5) return 4;      // df_mode 5 takes the Chapman-Enskog estimate (:105-110)
  DfData dfd(p, in);
  double Ntot = 0;
  int err = 0;
  for (long icell = 0; icell < in->n_cells; icell++) {
    YieldCell y = yield_cell(p, in, dfd, icell, 0.0, 1.0, &err);
    if (err) return err;
    if (!y.valid) continue;
    for (int s = 0; s < in->n_species; s++) {   // estimate_mean_particle_number, :75-119
      double neq = in->equilibrium_density[s], bd = in->bulk_density[s], dd = in->diffusion_density[s];
      if (p->df_mode == 4) Ntot += y.breaks_yield ? y.ds_time * (1.0 + y.delta_z) * neq : y.ds_time * y.z * neq;
      else Ntot += y.ds_time * (neq + y.bulkPi * bd) - y.ds_space * y.Vdsigma * dd;
    }
  }
  if (p->dimension == 2) Ntot *= (2.0 * p->y_cut);
  *ntotal = Ntot;
  return 0;
}

// dn_tot[cell] (after the 2 y_max ds_max volume factor) and dn_list[cell][species] (before it), df_mode 1-4 (df_mode 5:
// cell_yields_famod below):
// fast = 1 fast_max_particle_number (ParticleSampler.cpp:122-161), fast = 0 max_particle_number (:164-239)
namespace {
int cell_yields_famod(const cf_params *p, const cf_inputs *in, double *dn_tot, double *dn_list);   // defined below
}

extern "C" int cf_oracle_cell_yields(const cf_params *p, const cf_inputs *in, double *dn_tot, double *dn_list)
{
  if (p->df_mode == 5) return cell_yields_famod(p, in, dn_tot, dn_list);
  if (p->df_mode < 1 || p->df_mode > 4) return 4;
  DfData dfd(p, in);
  double F_avg = 0.0, betabulk_avg = 1.0;
  if (p->df_mode == 3 && p->fast) {
    DfCoeff d;
    if (!dfd.evaluate(in->T_avg, in->muB_avg, 0.0, 0.0, 0.0, &d)) return 3;
    F_avg = d.F; betabulk_avg = d.betabulk;
  }
  double y_max = (p->dimension == 2) ? p->y_cut : 0.5;
  const int pts = in->n_gla;
  const double *r1 = in->gla_root + 1 * pts, *w1 = in->gla_weight + 1 * pts;
  const double *r2 = in->gla_root + 2 * pts, *w2 = in->gla_weight + 2 * pts;
  int err = 0;
  for (long icell = 0; icell < in->n_cells; icell++) {
    YieldCell y = yield_cell(p, in, dfd, icell, F_avg, betabulk_avg, &err);
    if (err) return err;
    double tot = 0.0;
    for (int s = 0; s < in->n_species; s++) {
      double v = 0.0;
      if (y.valid && p->fast) {                  // fast_max_particle_number
        double neq = in->equilibrium_density[s], bd = in->bulk_density[s];
        if (p->df_mode <= 2 || y.breaks_sample) v = 2.0 * neq;
        else if (p->df_mode == 3) v = neq + y.bulkPi * bd;
        else v = y.z * neq;
      } else if (y.valid) {                      // max_particle_number at the cell's own (T, alphaB)
        double T = y.T, mbar = in->mass[s] / T, degeneracy = in->degeneracy[s], sign = in->sign[s], baryon = in->baryon[s];
        double neq_fact = T * T * T / two_pi2_hbarC3, J20_fact = T * neq_fact;
        if (p->df_mode <= 2 || y.breaks_sample) {
          double equilibrium_density = neq_fact * degeneracy * GaussThermal(neq_int, r1, w1, pts, mbar, y.alphaB, baryon, sign);
          v = 2.0 * equilibrium_density;
        } else if (p->df_mode == 3) {
          double equilibrium_density = neq_fact * degeneracy * GaussThermal(neq_int, r1, w1, pts, mbar, y.alphaB, baryon, sign);
          double J10 = 0.0;
          if (p->include_baryon) J10 = neq_fact * degeneracy * GaussThermal(J10_int, r1, w1, pts, mbar, y.alphaB, baryon, sign);
          double J20 = J20_fact * degeneracy * GaussThermal(J20_int, r2, w2, pts, mbar, y.alphaB, baryon, sign);
          double bulk_density = (equilibrium_density + (baryon * J10 * y.G) + (J20 * y.F / T / T)) / y.betabulk;
          v = equilibrium_density + y.bulkPi * bulk_density;
        } else {
          double equilibrium_density = neq_fact * degeneracy * GaussThermal(neq_int, r1, w1, pts, mbar, 0.0, 0.0, sign);
          v = y.z * equilibrium_density;
        }
      }
      if (dn_list) dn_list[(size_t)icell * in->n_species + s] = v;
      tot += v;
    }
    dn_tot[icell] = (y.valid && tot > 0.0) ? tot * (2.0 * y_max * y.ds_max) : 0.0;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// df_mode 5: PTMA modified anisotropic distribution.
// AnisoVariables.cpp: compute_F :15-132, compute_J :134-300, line_backtrack :302-391, find_anisotropic_variables
// :393-539, compute_famod_coefficient :541-643; calculate_dN_pTdpTdphidy_famod, MomentumSpectra.cpp:1049-1682.
// The 16-point generalized Gauss-Laguerre rules (AnisoVariables.h:15-121) are regenerated here by Newton iteration on
// the Laguerre recurrence instead of being pasted.
// ---------------------------------------------------------------------------------------------------------------
namespace {

const double four_pi2_hbarC3 = 4.0 * pow(M_PI, 2) * pow(hbarC, 3);

struct GenLaguerre16 {
  double root[4][16], weight[4][16];     // index by alpha = 1..3
  static long double L(int n, int a, long double x, long double *Lnm1)
  {
    long double p0 = 1.0L, p1 = 1.0L + a - x;
    if (n == 0) { *Lnm1 = 0; return p0; }
    for (int k = 1; k < n; k++) { long double p2 = ((2 * k + 1 + a - x) * p1 - (k + a) * p0) / (k + 1); p0 = p1; p1 = p2; }
    *Lnm1 = p0;
    return p1;
  }
  GenLaguerre16()
  {
    const int n = 16;
    for (int a = 1; a <= 3; a++) {
      // initial guesses from the asymptotic formula, refined with Newton + deflation-free bracketing by ordering
      long double x[16];
      for (int i = 0; i < n; i++) {
        if (i == 0) x[i] = (1.0L + a) * (3.0L + 0.92L * a) / (1.0L + 2.4L * n + 1.8L * a);
        else if (i == 1) x[i] = x[0] + (15.0L + 6.25L * a) / (1.0L + 0.9L * a + 2.5L * n);
        else { long double ai = i - 1; x[i] = x[i - 1] + ((1.0L + 2.55L * ai) / (1.9L * ai) + 1.26L * ai * a / (1.0L + 3.5L * ai)) * (x[i - 1] - x[i - 2]) / (1.0L + 0.3L * a); }
        for (int it = 0; it < 100; it++) {
          long double lm1, ln = L(n, a, x[i], &lm1);
          long double dl = (n * ln - (n + a) * lm1) / x[i];      // derivative of L_n^a
          long double dx = ln / dl;
          x[i] -= dx;
          if (fabsl(dx) <= 1e-19L * fabsl(x[i])) break;
        }
        long double lm1, ln = L(n, a, x[i], &lm1);
        (void)ln;
        long double dl = -(n + a) * lm1 / x[i];                   // L_n(x_i) = 0
        // w_i = Gamma(n + a + 1) / (n! x_i [L_n'(x_i)]^2)
        long double w = tgammal((long double)(n + a + 1)) / (tgammal((long double)(n + 1)) * x[i] * dl * dl);
        root[a][i] = (double)x[i]; weight[a][i] = (double)w;
      }
    }
  }
};
const GenLaguerre16 &gl16() { static GenLaguerre16 g; return g; }

struct Hadrons { const double *mass, *sign, *deg; int n; };

void t_functions(double z, double *t200, double *t220, double *t201, double *t402, double *t421, double *t440)
{
  const double delta = 0.01;
  *t200 = *t220 = *t201 = *t402 = *t421 = *t440 = 0.0;
  if (z > delta || (z < -delta && z > -1.)) {
    double t = (z > 0) ? atan(sqrt(z)) / sqrt(z) : atanh(sqrt(-z)) / sqrt(-z), z2 = z * z;
    *t200 = 1. + (1. + z) * t; *t220 = (-1. + (1. + z) * t) / z; *t201 = (1. + (z - 1.) * t) / z;
    *t402 = (3. * (z - 1.) + (z * (3. * z - 2.) + 3.) * t) / (4. * z2);
    *t421 = (3. + z + (1. + z) * (z - 3.) * t) / (4. * z2);
    *t440 = (-(3. + 5. * z) + 3. * (z + 1.) * (z + 1.) * t) / (4. * z2);
  } else if (fabs(z) <= delta) {     // Taylor series of the same functions
    double z2 = z * z, z3 = z2 * z, z4 = z3 * z, z5 = z4 * z, z6 = z5 * z;
    *t200 = 2. + 2. / 3. * z - 2. / 15. * z2 + 2. / 35. * z3 - 2. / 63. * z4 + 2. / 99. * z5 - 2. / 143. * z6;
    *t220 = 2. / 3. - 2. / 15. * z + 2. / 35. * z2 - 2. / 63. * z3 + 2. / 99. * z4 - 2. / 143. * z5 + 2. / 195. * z6;
    *t201 = 4. / 3. - 8. / 15. * z + 12. / 35. * z2 - 16. / 63. * z3 + 20. / 99. * z4 - 24. / 143. * z5 + 28. / 195. * z6;
    *t402 = 16. / 15. - 16. / 35. * z + 32. / 105. * z2 - 160. / 693. * z3 + 80. / 429. * z4 - 112. / 715. * z5 + 448. / 3315. * z6;
    *t421 = 4. / 15. - 8. / 105. * z + 4. / 105. * z2 - 16. / 693. * z3 + 20. / 1287. * z4 - 8. / 715. * z5 + 28. / 3315. * z6;
    *t440 = 2. / 5. - 2. / 35. * z + 2. / 105. * z2 - 2. / 231. * z3 + 2. / 429. * z4 - 2. / 715. * z5 + 2. / 1105. * z6;
  }
}

void compute_F(const Hadrons &h, double Ea, double PTa, double PLa, const double *X, double *F)
{
  double lambda = X[0], aT = X[1], aL = X[2], aT2 = aT * aT, aL2 = aL * aL;
  double common_factor = aT2 * aL * lambda * lambda * lambda * lambda / four_pi2_hbarC3;
  double I_200 = 0, I_220 = 0, I_201 = 0;
  for (int n = 0; n < h.n; n++) {
    if (h.mass[n] == 0) continue;
    double mbar = h.mass[n] / lambda, mbar2 = mbar * mbar, a = 0, b = 0, c = 0;
    for (int i = 0; i < 16; i++) {
      double pbar = gl16().root[2][i], weight = gl16().weight[2][i];
      double Ebar = sqrt(pbar * pbar + mbar2), w = sqrt(aL2 + mbar2 / (pbar * pbar)), z = (aT2 - aL2) / (w * w);
      double t200, t220, t201, u1, u2, u3;
      t_functions(z, &t200, &t220, &t201, &u1, &u2, &u3);
      double cw = pbar * weight * exp(pbar) / (exp(Ebar) + h.sign[n]);
      a += cw * t200 * w; b += cw * t220 / w; c += cw * t201 / w;
    }
    I_200 += a * h.deg[n]; I_220 += b * h.deg[n]; I_201 += c * h.deg[n];
  }
  I_200 *= common_factor; I_220 *= common_factor * aL2; I_201 *= common_factor * aT2 / 2.;
  F[0] = I_200 - Ea; F[1] = I_201 - PTa; F[2] = I_220 - PLa;
}

void j_sums(const Hadrons &h, double lambda, double aT2, double aL2, double S[6])
{
  for (int k = 0; k < 6; k++) S[k] = 0;
  for (int n = 0; n < h.n; n++) {
    if (h.mass[n] == 0) continue;
    double mbar = h.mass[n] / lambda, mbar2 = mbar * mbar, s[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 16; i++) {
      double pbar = gl16().root[3][i], weight = gl16().weight[3][i], pbar2 = pbar * pbar;
      double Ebar = sqrt(pbar2 + mbar2), w = sqrt(aL2 + mbar2 / pbar2), z = (aT2 - aL2) / (w * w);
      double t200, t220, t201, t402, t421, t440;
      t_functions(z, &t200, &t220, &t201, &t402, &t421, &t440);
      double q = exp(Ebar) + h.sign[n], cw = weight * exp(pbar + Ebar) / (q * q);
      s[0] += Ebar * cw * t200 * w; s[1] += Ebar * cw * t201 / w; s[2] += Ebar * cw * t220 / w;
      s[3] += pbar2 / Ebar * cw * t402 / w; s[4] += pbar2 / Ebar * cw * t421 / w; s[5] += pbar2 / Ebar * cw * t440 / w;
    }
    for (int k = 0; k < 6; k++) S[k] += s[k] * h.deg[n];
  }
}

void compute_J(const Hadrons &h, double Ea, double PTa, double PLa, const double *X, const double *F, double J[3][3])
{
  double lambda = X[0], aT = X[1], aL = X[2], aT2 = aT * aT, aL2 = aL * aL, lambda2 = lambda * lambda, lambda3 = lambda2 * lambda;
  double lambda_aT3 = lambda * aT2 * aT, lambda_aL3 = lambda * aL2 * aL, cf = aT2 * aL * lambda2 * lambda3 / four_pi2_hbarC3;
  double S[6];
  j_sums(h, lambda, aT2, aL2, S);
  double J_2001 = S[0] * cf, J_2011 = S[1] * cf * aT2 / 2., J_2201 = S[2] * cf * aL2;
  double J_402m1 = S[3] * cf * aT2 * aT2 / 8., J_421m1 = S[4] * cf * aT2 * aL2 / 2., J_440m1 = S[5] * cf * aL2 * aL2;
  double Eai = F[0] + Ea, PTai = F[1] + PTa, PLai = F[2] + PLa;
  J[0][0] = J_2001 / lambda2; J[0][1] = 2. * (Eai + PTai) / aT;    J[0][2] = (Eai + PLai) / aL;
  J[1][0] = J_2011 / lambda2; J[1][1] = 4. * J_402m1 / lambda_aT3; J[1][2] = J_421m1 / lambda_aL3;
  J[2][0] = J_2201 / lambda2; J[2][1] = 2. * J_421m1 / lambda_aT3; J[2][2] = J_440m1 / lambda_aL3;
}

double line_backtrack(const Hadrons &h, double Ea, double PTa, double PLa, const double *Xc, const double *dX, double dX_abs, double g0, double *F)
{
  const double tol_dX = 1.e-4;
  double X[3] = {Xc[0] + dX[0], Xc[1] + dX[1], Xc[2] + dX[2]};
  compute_F(h, Ea, PTa, PLa, X, F);
  double f = (F[0] * F[0] + F[1] * F[1] + F[2] * F[2]) / 2., gprime0 = -2. * g0, l = 1, alpha = 0.0001, lroot = 0, lprev = 0, fprev = 0;
  for (int n = 0; n < 20; n++) {
    if ((l * dX_abs) <= tol_dX) return l;
    else if (f <= (g0 + l * alpha * gprime0)) return l;
    else if (n == 0) lroot = -gprime0 / (2. * (f - g0 - gprime0));
    else {
      double a = ((f - g0 - l * gprime0) / (l * l) - (fprev - g0 - lprev * gprime0) / (lprev * lprev)) / (l - lprev);
      double b = (-lprev * (f - g0 - l * gprime0) / (l * l) + l * (fprev - g0 - lprev * gprime0) / (lprev * lprev)) / (l - lprev);
      if (a == 0) lroot = -gprime0 / (2. * b);
      else {
        double z = b * b - 3. * a * gprime0;
        if (z < 0) lroot = 0.5 * l; else if (b <= 0) lroot = (-b + sqrt(z)) / (3. * a); else lroot = -gprime0 / (b + sqrt(z));
      }
      lroot = fmin(lroot, 0.5 * l);
    }
    lprev = l; fprev = f;
    l = fmax(lroot, 0.5 * l);
    for (int i = 0; i < 3; i++) X[i] = Xc[i] + l * dX[i];
    compute_F(h, Ea, PTa, PLa, X, F);
    f = (F[0] * F[0] + F[1] * F[1] + F[2] * F[2]) / 2.;
  }
  return l;
}

struct Aniso { double lambda, aT, aL; bool failed; int iterations; };

// J dX = -F by LU with partial pivoting (gsl_linalg_LU_decomp / LU_solve)
void lu_solve3(double A[3][3], const double rhs[3], double x[3])
{
  int perm[3] = {0, 1, 2};
  for (int j = 0; j < 2; j++) {
    int ip = j; double mx = fabs(A[j][j]);
    for (int i = j + 1; i < 3; i++) if (fabs(A[i][j]) > mx) { mx = fabs(A[i][j]); ip = i; }
    if (ip != j) { for (int k = 0; k < 3; k++) { double t = A[j][k]; A[j][k] = A[ip][k]; A[ip][k] = t; } int t = perm[j]; perm[j] = perm[ip]; perm[ip] = t; }
    if (A[j][j] != 0.0) for (int i = j + 1; i < 3; i++) { double a = A[i][j] / A[j][j]; A[i][j] = a; for (int k = j + 1; k < 3; k++) A[i][k] -= a * A[j][k]; }
  }
  double b[3] = {rhs[perm[0]], rhs[perm[1]], rhs[perm[2]]};
  for (int i = 0; i < 3; i++) for (int j = 0; j < i; j++) b[i] -= A[i][j] * b[j];
  for (int i = 2; i >= 0; i--) { for (int j = i + 1; j < 3; j++) b[i] -= A[i][j] * b[j]; b[i] /= A[i][i]; }
  for (int i = 0; i < 3; i++) x[i] = b[i];
}

Aniso find_anisotropic_variables(const Hadrons &h, double E, double pl, double pt, double lambda_0, double aT_0, double aL_0)
{
  Aniso fail{lambda_0, aT_0, aL_0, true, 0};
  double Ea = E, PTa = pt, PLa = pl;
  if (Ea < 0 || PTa < 0 || PLa < 0) return fail;
  double X[3] = {lambda_0, aT_0, aL_0}, dX[3], F[3], J[3][3];
  compute_F(h, Ea, PTa, PLa, X, F);
  double stepmax = 100. * fmax(sqrt(X[0] * X[0] + X[1] * X[1] + X[2] * X[2]), 3.);
  for (int n = 0; n < 30; n++) {
    compute_J(h, Ea, PTa, PLa, X, F, J);
    double f = (F[0] * F[0] + F[1] * F[1] + F[2] * F[2]) / 2.;
    for (int i = 0; i < 3; i++) F[i] *= -1.;
    lu_solve3(J, F, dX);
    double dX_abs = sqrt(dX[0] * dX[0] + dX[1] * dX[1] + dX[2] * dX[2]);
    if (dX_abs > stepmax) { for (int i = 0; i < 3; i++) dX[i] *= stepmax / dX_abs; dX_abs = stepmax; }
    double l = line_backtrack(h, Ea, PTa, PLa, X, dX, dX_abs, f, F);
    for (int i = 0; i < 3; i++) X[i] += (l * dX[i]);
    double F_abs = sqrt(F[0] * F[0] + F[1] * F[1] + F[2] * F[2]);
    dX_abs *= l;
    if (X[0] < 0 || X[1] < 0 || X[2] < 0) { fail.iterations = n + 1; return fail; }
    else if (dX_abs <= 1.e-4 && F_abs <= 1.e-4) return Aniso{X[0], X[1], X[2], false, n + 1};
  }
  fail.iterations = 30;
  return fail;
}

int spectra_famod(const cf_params *p, const cf_inputs *in, double *out, cf_stats *st)
{
  const double prefactor = pow(2.0 * M_PI * hbarC, -3);
  Grids g(p, in);
  const int npart = in->n_species, npT = in->n_pT, nphi = in->n_phi, ny = g.ny;
  Hadrons h{in->pdg_mass, in->pdg_sign, in->pdg_degeneracy, (int)fmin(320, in->n_pdg)};
  const double detB_min = p->deta_min;
  double lambda_prev = 0, aT_prev = 0, aL_prev = 0;
  bool previous_success = false;
  for (long icell = 0; icell < in->n_cells; icell++) {
    CellState c;
    if (!load_cell(p, in, icell, true, false, &c)) { st->cells_skipped++; continue; }
    if (p->dimension == 3) g.eta[0] = c.eta;
    double T = c.T, tau2 = c.tau2;
    Basis b = milne_basis(c);
    PiLRF l = boost_pi(c, b);
    double pl = c.P + c.bulkPi + l.zz, pt = c.P + c.bulkPi - l.zz / 2.;
    double piTxx = 0, piTxy = 0, piTyy = 0, WTzx = 0, WTzy = 0;
    if (p->include_shear_deltaf) { piTxx = (l.xx - l.yy) / 2.; piTxy = l.xy; piTyy = -piTxx; WTzx = l.xz; WTzy = l.yz; }
    double lambda = T, aT = 1, aL = 1, upsilonB = c.alphaB;
    bool breaks = false;
    if (pl < 0 || pt < 0) { st->cells_pl_negative++; breaks = true; }
    else {
      const bool prev = p->famod_chain && previous_success;
      if (prev) { lambda = lambda_prev; aT = aT_prev; aL = aL_prev; }
      Aniso X = find_anisotropic_variables(h, c.E, pl, pt, lambda, aT, aL);
      if (X.failed && prev) {
        lambda = T; aT = 1; aL = 1;
        X = find_anisotropic_variables(h, c.E, pl, pt, lambda, aT, aL);
        if (X.failed) { breaks = true; st->reconstruction_failures++; previous_success = false; }
        else { lambda = X.lambda; aT = X.aT; aL = X.aL; lambda_prev = lambda; aT_prev = aT; aL_prev = aL; previous_success = true; }
      } else { lambda = X.lambda; aT = X.aT; aL = X.aL; lambda_prev = lambda; aT_prev = aT; aL_prev = aL; previous_success = true; }
      st->newton_iterations += X.iterations;
    }
    // compute_famod_coefficient
    double S[6], aT2 = aT * aT, aL2 = aL * aL, lambda2 = lambda * lambda, cf = aT2 * aL * lambda * lambda2 * lambda2 / four_pi2_hbarC3;
    j_sums(h, lambda, aT2, aL2, S);
    double betapiperp = (S[3] * cf * aT2 * aT2 / 8.) / (aT2 * lambda), betaWperp = (S[4] * cf * aT2 * aL2 / 2.) / (aT * aL * lambda);
    double shear_coeff = 0.5 / betapiperp, diff_coeff = 1. / betaWperp, detA = aT * aT * aL;
    double Cxx = 1. + shear_coeff * piTxx, Cxy = shear_coeff * piTxy, Cxz = diff_coeff * WTzx * aT / (aT + aL);
    double Cyx = Cxy, Cyy = 1. + shear_coeff * piTyy, Cyz = diff_coeff * WTzy * aT / (aT + aL);
    double Czx = diff_coeff * WTzx * aL / (aT + aL), Czy = diff_coeff * WTzy * aL / (aT + aL), Czz = 1.;
    double detC = Cxx * (Cyy * Czz - Cyz * Czy) - Cxy * (Cyx * Czz - Cyz * Czx) + Cxz * (Cyx * Czy - Cyy * Czx);
    double Bxx = aT + aT * shear_coeff * piTxx, Bxy = aT * shear_coeff * piTxy, Bxz = diff_coeff * WTzx * aT * aL / (aT + aL);
    double Byy = aT + aT * shear_coeff * piTyy, Byz = diff_coeff * WTzy * aT * aL / (aT + aL), Bzz = aL;
    double detB = detC * detA, detB_bulk_two_thirds = (2. * aT + aL) * (2. * aT + aL) / 9.;
    double B[9] = {Bxx, Bxy, Bxz, Bxy, Byy, Byz, Bxz, Byz, Bzz}, B_inv[9];
    lu_invert3(B, B_inv);
    if (detB <= detB_min) breaks = true;
    double eta_scale = 1;
    if (detB > detB_min && p->dimension == 2) eta_scale = detB / detB_bulk_two_thirds;
    double renorm = eta_scale / detC;
    if (std::isnan(renorm) || std::isinf(renorm)) breaks = true;
    if (breaks) st->cells_breakdown++;
    for (int ipart = 0; ipart < npart; ipart++) {
      double mass = in->mass[ipart], mass2 = mass * mass, sign = in->sign[ipart], degeneracy = in->degeneracy[ipart], baryon = in->baryon[ipart];
      double chem = baryon * c.alphaB, chem_effect = baryon * upsilonB;
      for (int ipT = 0; ipT < npT; ipT++) {
        double pT = g.pT[ipT], mT = sqrt(mass2 + pT * pT), mT_over_tau = mT / c.tau;
        for (int iphip = 0; iphip < nphi; iphip++) {
          double px = pT * g.cosphi[iphip], py = pT * g.sinphi[iphip];
          for (int iy = 0; iy < ny; iy++) {
            double y = g.y[iy], eta_integral = 0;
            for (int ieta = 0; ieta < g.neta; ieta++) {
              double eta = g.eta[ieta], eta_weight = g.etaw[ieta];
              bool narrow = (p->dimension == 3 && !breaks && detB < 0.01 && fabs(y - eta) < detB);
              double p_dsigma, f;
              if (breaks || narrow) {
                double ptau = mT * cosh(y - eta), pn = mT_over_tau * sinh(y - eta), tau2_pn = tau2 * pn;
                p_dsigma = ptau * c.dat + px * c.dax + py * c.day + pn * c.dan;
                if (p->outflow && p_dsigma <= 0) continue;
                double u_p = ptau * c.ut - px * c.ux - py * c.uy - tau2_pn * c.un;
                f = 1. / (exp(u_p / T - chem) + sign);
              } else {
                double ptau = mT * cosh(y - eta_scale * eta), pn = mT_over_tau * sinh(y - eta_scale * eta), tau2_pn = tau2 * pn;
                p_dsigma = ptau * c.dat + px * c.dax + py * c.day + pn * c.dan;
                if (p->outflow && p_dsigma <= 0.0) continue;
                double pLRF[3] = {-b.Xt * ptau + b.Xx * px + b.Xy * py + b.Xn * tau2_pn, b.Yx * px + b.Yy * py, -b.Zt * ptau + b.Zn * tau2_pn};
                double pmod[3];
                for (int i = 0; i < 3; i++) pmod[i] = B_inv[3 * i] * pLRF[0] + B_inv[3 * i + 1] * pLRF[1] + B_inv[3 * i + 2] * pLRF[2];
                for (int it = 0; it < 5; it++) {
                  double prev[3] = {pmod[0], pmod[1], pmod[2]}, back[3], dp[3];
                  for (int i = 0; i < 3; i++) back[i] = B[3 * i] * prev[0] + B[3 * i + 1] * prev[1] + B[3 * i + 2] * prev[2];
                  for (int i = 0; i < 3; i++) dp[i] = pLRF[i] - back[i];
                  if (sqrt(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]) <= 1.e-16) break;
                  for (int i = 0; i < 3; i++) pmod[i] = prev[i] + (B_inv[3 * i] * dp[0] + B_inv[3 * i + 1] * dp[1] + B_inv[3 * i + 2] * dp[2]);
                }
                double E_mod = sqrt(mass2 + pmod[0] * pmod[0] + pmod[1] * pmod[1] + pmod[2] * pmod[2]);
                f = fabs(renorm) / (exp(E_mod / lambda - chem_effect) + sign);
              }
              eta_integral += eta_weight * p_dsigma * f;
            }
            out[iy + (long)ny * (iphip + (long)nphi * (ipT + (long)npT * ipart))] += prefactor * degeneracy * eta_integral;
          }
        }
      }
    }
  }
  return 0;
}

// Per-cell mean yields of sample_dN_pTdpTdphidy_famod (ParticleSampler.cpp:1138-1510): anisotropic variables with the
// SAMPLER's failure rules (a failed first attempt without a previous success is a breakdown and leaves (T, 1, 1), :1369-1374),
// densities g Lambda^3 detA / (2 pi^2 hbarc^3) I_100 with the 16-point a = 1 rule and exp(Ebar + chem) as written (:1490).
int cell_yields_famod(const cf_params *p, const cf_inputs *in, double *dn_tot, double *dn_list)
{
  Hadrons h{in->pdg_mass, in->pdg_sign, in->pdg_degeneracy, (int)fmin(320, in->n_pdg)};
  const GenLaguerre16 &gl = gl16();
  const double y_max = (p->dimension == 2) ? p->y_cut : 0.5;
  double lambda_prev = 0, aT_prev = 0, aL_prev = 0;
  bool previous_success = false;
  for (long icell = 0; icell < in->n_cells; icell++) {
    dn_tot[icell] = 0.0;
    if (dn_list) for (int s = 0; s < in->n_species; s++) dn_list[(size_t)icell * in->n_species + s] = 0.0;
    CellState c;
    if (!load_cell(p, in, icell, true, false, &c)) continue;
    Basis b = milne_basis(c);
    double dst = c.dat * c.ut + c.dax * c.ux + c.day * c.uy + c.dan * c.un;
    double dsx = -(c.dat * b.Xt + c.dax * b.Xx + c.day * b.Xy + c.dan * b.Xn);
    double dsy = -(c.dax * b.Yx + c.day * b.Yy);
    double dsz = -(c.dat * b.Zt + c.dan * b.Zn);
    double ds_max = fabs(dst) + sqrt(dsx * dsx + dsy * dsy + dsz * dsz);
    PiLRF l = boost_pi(c, b);
    double pl = c.P + c.bulkPi + l.zz, pt = c.P + c.bulkPi - l.zz / 2.;
    double lambda = c.T, aT = 1, aL = 1, upsilonB = c.alphaB;
    if (!(pl < 0 || pt < 0)) {
      const bool prev = p->famod_chain && previous_success;
      if (prev) { lambda = lambda_prev; aT = aT_prev; aL = aL_prev; }
      Aniso X = find_anisotropic_variables(h, c.E, pl, pt, lambda, aT, aL);
      if (X.failed) {
        if (prev) {
          lambda = c.T; aT = 1; aL = 1;
          X = find_anisotropic_variables(h, c.E, pl, pt, lambda, aT, aL);
          if (X.failed) previous_success = false;
          else { lambda = X.lambda; aT = X.aT; aL = X.aL; lambda_prev = lambda; aT_prev = aT; aL_prev = aL; previous_success = true; }
        } else previous_success = false;
      } else { lambda = X.lambda; aT = X.aT; aL = X.aL; lambda_prev = lambda; aT_prev = aT; aL_prev = aL; previous_success = true; }
    }
    const double detA = aT * aT * aL, na_fact = lambda * lambda * lambda * detA / two_pi2_hbarC3;
    double tot = 0.0;
    for (int s = 0; s < in->n_species; s++) {
      const double mbar = in->mass[s] / lambda, mbar2 = mbar * mbar, chem = in->baryon[s] * upsilonB, sign = in->sign[s];
      double I_100 = 0;
      for (int k = 0; k < 16; k++) {
        const double pbar = gl.root[1][k], weight = gl.weight[1][k];
        const double Ebar = sqrt(pbar * pbar + mbar2);
        I_100 += pbar * weight * exp(pbar) / (exp(Ebar + chem) + sign);
      }
      const double v = in->degeneracy[s] * na_fact * I_100;
      if (dn_list) dn_list[(size_t)icell * in->n_species + s] = v;
      tot += v;
    }
    dn_tot[icell] = (tot > 0.0) ? tot * (2.0 * y_max * ds_max) : 0.0;
  }
  return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// calculate_spin_polzn, Polarization.cpp:25-263.  f0 is taken at the surface-averaged temperature (:76, :186); cells
// with u.dsigma <= 0 are not skipped; the vorticity index is the in-chunk index (:125-130) when chunk_compat is set.
// Output in the spectra layout (the reference's own storage order, species fastest, :201, :226, is a relabelling).
// ---------------------------------------------------------------------------------------------------------------
extern "C" int cf_oracle_polarization(const cf_params *p, const cf_inputs *in, const double *const w[6], int chunk_compat,
                                      double *St, double *Sx, double *Sy, double *Sn, double *Snorm)
{
  const int npart = in->n_species, npT = in->n_pT, nphi = in->n_phi;
  const int y_pts = (p->dimension == 2) ? 1 : in->n_y, eta_pts = (p->dimension == 2) ? in->n_eta : 1;
  const long FO_chunk = 10000;
  const long total = (long)npart * npT * nphi * y_pts;
  for (long i = 0; i < total; i++) St[i] = Sx[i] = Sy[i] = Sn[i] = Snorm[i] = 0.0;
  std::vector<double> cosphi(nphi), sinphi(nphi), yv(y_pts, 0.0), etav(eta_pts, 0.0), etaw(eta_pts, 1.0);
  for (int j = 0; j < nphi; j++) { cosphi[j] = cos(in->phi[j]); sinphi[j] = sin(in->phi[j]); }
  const double delta_eta = (in->n_eta > 1) ? in->eta[1] - in->eta[0] : 0.0;
  if (p->dimension == 2) for (int k = 0; k < eta_pts; k++) { etav[k] = in->eta[k]; etaw[k] = in->eta_w[k] * delta_eta; }
  else for (int iy = 0; iy < y_pts; iy++) yv[iy] = in->y[iy];
  const double T = in->T_avg;
  for (long icell_glb = 0; icell_glb < in->n_cells; icell_glb++) {
    const long icell = icell_glb % FO_chunk;
    const long iw = chunk_compat ? icell : icell_glb;
    const double tau = in->col[0][icell_glb], tau2 = tau * tau;
    if (p->dimension == 3) etav[0] = in->col[3][icell_glb];
    const double dat = in->col[4][icell_glb], dax = in->col[5][icell_glb], day = in->col[6][icell_glb], dan = in->col[7][icell_glb];
    const double ux = in->col[8][icell_glb], uy = in->col[9][icell_glb], un = in->col[10][icell_glb];
    const double ut = sqrt(fabs(1.0 + ux * ux + uy * uy + tau2 * un * un));
    const double wtx = w[0][iw], wty = w[1][iw], wtn = w[2][iw], wxy = w[3][iw], wxn = w[4][iw], wyn = w[5][iw];
    for (int ipart = 0; ipart < npart; ipart++) {
      const double mass = in->mass[ipart], mass2 = mass * mass, sign = in->sign[ipart];
      for (int ipT = 0; ipT < npT; ipT++) {
        const double pT = in->pT[ipT], mT = sqrt(mass2 + pT * pT), mT_over_tau = mT / tau;
        for (int iphip = 0; iphip < nphi; iphip++) {
          const double px = pT * cosphi[iphip], py = pT * sinphi[iphip];
          for (int iy = 0; iy < y_pts; iy++) {
            const double y = yv[iy];
            double st = 0, sx = 0, sy = 0, sn = 0, snorm = 0;
            for (int ieta = 0; ieta < eta_pts; ieta++) {
              const double eta = etav[ieta], wgt = etaw[ieta];
              const double pt = mT * cosh(y - eta), pn = mT_over_tau * sinh(y - eta), tau2_pn = tau2 * pn;
              const double pdotdsigma = pt * dat + px * dax + py * day + pn * dan;
              const double pdotu = pt * ut - px * ux - py * uy - tau2_pn * un;
              const double f0 = 1.0 / (exp(pdotu / T) + sign);
              const double prefactor = -(1.0 / 8.0 / mass) * (1.0 - sign * f0);
              const double spin_t = prefactor * 2.0 * (wxy * pn - wxn * py + wyn * px);
              const double spin_x = prefactor * 2.0 * (wyn * pt - wtn * py + wty * pn);
              const double spin_y = prefactor * 2.0 * (-wxn * pt + wtn * px - wtx * pn);
              const double spin_n = prefactor * 2.0 * (wtx * py + wxy * pt - wty * px);
              st += (wgt * pdotdsigma * f0 * spin_t);
              sx += (wgt * pdotdsigma * f0 * spin_x);
              sy += (wgt * pdotdsigma * f0 * spin_y);
              sn += (wgt * pdotdsigma * f0 * spin_n);
              snorm += (wgt * pdotdsigma * f0);
            }
            const long idx = iy + (long)y_pts * (iphip + (long)nphi * (ipT + (long)npT * ipart));
            St[idx] += st; Sx[idx] += sx; Sy[idx] += sy; Sn[idx] += sn; Snorm[idx] += snorm;
          }
        }
      }
    }
  }
  return 0;
}
