// K1 math: continuous Cooper-Frye spectra with linear df corrections (df_mode 1 Grad 14-moment, df_mode 2 RTA
// Chapman-Enskog).  Reference: EmissionFunctionArray::calculate_dN_pTdpTdphidy, src/cpp/MomentumSpectra.cpp:32-415.
//
// The reference evaluates, per (cell, species, pT, phi, y[, eta]),
//     p^tau = mT cosh(y-eta), p^eta = mT sinh(y-eta)/tau, p^x = pT cos(phi), p^y = pT sin(phi)
//     f = feq (1 + df),  feq = 1/(exp(u.p/T - b muB/T) + sign)
// Every scalar product of p with a cell tensor is a polynomial in (mT, pT) whose coefficients depend only on
// (cell, y, phi).  Those coefficients are formed once per (cell, iy, iphi) ("item constants", 16 doubles in shared
// memory, warp-uniform) and the inner loop per momentum bin is ~15 FMAs + one exp + one (mode 2: two) reciprocals:
//     x_E  = mT aT - pT bT                (= u.p / T)
//     p.ds = mT c1 + pT d1
//     pi^{mu nu} p_mu p_nu (scaled) = mT^2 q1 + mT pT q2 + pT^2 q3
//     V^mu p_mu = mT v1 - pT v2
#pragma once

#include "cellmath.cuh"
#include "dftables.cuh"

namespace is3d {

// per-cell pack written by the setup kernel, SoA in HBM: pack[k * stride + cell]
enum DfPackIdx {
  DP_VALID = 0, DP_ETA, DP_UTT, DP_TUNT, DP_UXT, DP_UYT, DP_ALPHAB,
  DP_DAT, DP_DAX, DP_DAY, DP_DANT,
  DP_PITT, DP_T2PINN, DP_TPITN, DP_PITX, DP_PITY, DP_TPIXN, DP_TPIYN, DP_PIXX, DP_PIYY, DP_PIXY,
  DP_K0, DP_K1, DP_K2, DP_G0, DP_G1, DP_VT, DP_TVN, DP_VX, DP_VY,
  DP_EBP, DP_EBM,              // exp(-alpha_B), exp(+alpha_B): the chemical-potential factor of the uniform-baryon path
  DP_EPETA, DP_EMETA,          // exp(+eta), exp(-eta): sinh(y - eta) = (e^y e^-eta - e^-y e^eta) / 2 without a libm call per (cell, y)
  DP_SIZE
};

struct DfFlags {
  int df_mode;                 // 1 or 2
  int dimension;
  int include_baryon;
  int include_bulk, include_shear, include_baryondiff;
};

// status bits returned by the per-cell setup
enum { CELL_OK = 0, CELL_SKIPPED = 1, CELL_OUT_OF_TABLE = 2 };

// Per-cell prologue of the reference loop (MomentumSpectra.cpp:109-246) folded into the pack.
IS3D_HD int df_setup_cell(const Cell &c, const DfTables &tb, const DfFlags &fl, double pack[DP_SIZE])
{
  for (int k = 0; k < DP_SIZE; k++) pack[k] = 0.0;
  double tau = c.tau, tau2 = tau * tau;
  double ux = c.ux, uy = c.uy, un = c.un;
  double utperp = sqrt(1.0 + ux * ux + uy * uy);
  double tau2_un = tau2 * un;
  double ut = sqrt(utperp * utperp + tau2_un * un);
  // skip cells with u.dsigma <= 0 (:132)
  if (ut * c.dat + ux * c.dax + uy * c.day + un * c.dan <= 0.0) return CELL_SKIPPED;

  double T = c.T, P = c.P, E = c.E;
  Shear pi;
  if (fl.include_shear) pi = complete_shear(c.pixx, c.pixy, c.pixn, c.piyy, c.piyn, ut, ux, uy, un, tau2);
  double bulkPi = fl.include_bulk ? c.bulkPi : 0.0;
  double muB = 0.0, alphaB = 0.0, nB = 0.0, Vt = 0.0, Vx = 0.0, Vy = 0.0, Vn = 0.0, baryon_enthalpy_ratio = 0.0;
  if (fl.include_baryon && fl.include_baryondiff) {      // :176-187
    muB = c.muB; nB = c.nB; Vx = c.Vx; Vy = c.Vy; Vn = c.Vn;
    Vt = (Vx * ux + Vy * uy + Vn * tau2_un) / ut;
    alphaB = muB / T;
    baryon_enthalpy_ratio = nB / (E + P);
  }
  DfCoeff df;
  if (!evaluate_df_coefficients(tb, fl.df_mode, fl.include_baryon, T, muB, E, P, bulkPi, &df)) return CELL_OUT_OF_TABLE;

  double invT = 1.0 / T;
  double sc, K0, K1, K2, G0, G1;
  if (fl.df_mode == 1) {                                 // :222-231
    double shear_coeff = 1.0 / df.shear14_coeff;
    double bulk0 = (df.c0 - df.c2) * bulkPi, bulk1 = df.c1 * bulkPi, bulk2 = (4. * df.c2 - df.c0) * bulkPi;
    sc = shear_coeff;
    K0 = bulk0; K1 = bulk1 * T; K2 = bulk2 * T * T;      // df_bulk = K0 m^2 + (K1 b + K2 xE) xE
    G0 = df.c3; G1 = df.c4 * T;                          // df_diff = (G0 b + G1 xE) V.p
  } else {                                               // :232-241
    double shear_coeff = 0.5 / (df.betapi * T);
    double bulk0 = df.F / (T * T * df.betabulk) * bulkPi, bulk1 = df.G / df.betabulk * bulkPi;
    double bulk2 = bulkPi / (3.0 * T * df.betabulk);
    sc = shear_coeff * invT;                             // df_shear = sc' pipp / xE
    K0 = (bulk0 + bulk2) * T; K1 = bulk1; K2 = bulk2 * invT;   // df_bulk = K0 xE + K1 b - K2 m^2 / xE
    G0 = baryon_enthalpy_ratio / df.betaV; G1 = invT / df.betaV;   // df_diff = (G0 - G1 b / xE) V.p
  }
  pack[DP_VALID] = 1.0;
  pack[DP_ETA] = c.eta;
  pack[DP_UTT] = ut * invT;  pack[DP_TUNT] = tau * un * invT;  pack[DP_UXT] = ux * invT;  pack[DP_UYT] = uy * invT;
  pack[DP_ALPHAB] = alphaB;
  pack[DP_EBP] = exp(-alphaB); pack[DP_EBM] = exp(alphaB);
  pack[DP_EPETA] = exp(c.eta); pack[DP_EMETA] = exp(-c.eta);
  pack[DP_DAT] = c.dat; pack[DP_DAX] = c.dax; pack[DP_DAY] = c.day; pack[DP_DANT] = c.dan / tau;
  pack[DP_PITT] = sc * pi.tt; pack[DP_T2PINN] = sc * tau2 * pi.nn; pack[DP_TPITN] = sc * tau * pi.tn;
  pack[DP_PITX] = sc * pi.tx; pack[DP_PITY] = sc * pi.ty; pack[DP_TPIXN] = sc * tau * pi.xn; pack[DP_TPIYN] = sc * tau * pi.yn;
  pack[DP_PIXX] = sc * pi.xx; pack[DP_PIYY] = sc * pi.yy; pack[DP_PIXY] = sc * pi.xy;
  pack[DP_K0] = K0; pack[DP_K1] = K1; pack[DP_K2] = K2; pack[DP_G0] = G0; pack[DP_G1] = G1;
  pack[DP_VT] = Vt; pack[DP_TVN] = tau * Vn; pack[DP_VX] = Vx; pack[DP_VY] = Vy;
  return CELL_OK;
}

// item constants: one (cell, eta-node) seen from a fixed (y, phi)
// Field order = read order of the momentum loop, so that the slots a kernel variant does not need sit at the end of
// the struct and cost no LDS.128: KX is the bulk coefficient the mode multiplies by xE (mode 1: K2, mode 2: K0); the
// other one is folded into q1 / q3 (see df_make_item); pad is read by the PTB fallback only.
struct alignas(16) DfItem {
  double aT, bT, c1, d1;
  double q1, q2, q3, KX;
  double K1, alphaB, v1, v2;
  double G0, G1, pad, unused_;
};

// pk(k) returns pack entry k of this cell.  sh/ch = sinh, cosh of (y - eta); w = eta quadrature weight (1 in 3+1d).
// The bulk term proportional to m^2 = mT^2 - pT^2 is folded into the shear quadratic form, so the momentum loop
// never needs m^2:   mode 1:  K0 m^2 + pi.p.p  = mT^2 (q1 + K0) + mT pT q2 + pT^2 (q3 - K0)
//                    mode 2:  pi.p.p - K2 m^2  = mT^2 (q1 - K2) + mT pT q2 + pT^2 (q3 + K2)
template <class PackFn>
IS3D_HD DfItem df_make_item(PackFn pk, int mode, double sh, double ch, double cphi, double sphi, double w)
{
  DfItem it;
  it.aT = ch * pk(DP_UTT) - sh * pk(DP_TUNT);
  it.bT = cphi * pk(DP_UXT) + sphi * pk(DP_UYT);
  it.c1 = w * (ch * pk(DP_DAT) + sh * pk(DP_DANT));
  it.d1 = w * (cphi * pk(DP_DAX) + sphi * pk(DP_DAY));
  it.K1 = pk(DP_K1);
  it.KX = (mode == 1) ? pk(DP_K2) : pk(DP_K0);
  const double fold = (mode == 1) ? pk(DP_K0) : -pk(DP_K2);
  it.q1 = ch * ch * pk(DP_PITT) + sh * sh * pk(DP_T2PINN) - 2.0 * ch * sh * pk(DP_TPITN) + fold;
  it.q2 = 2.0 * (sh * (pk(DP_TPIXN) * cphi + pk(DP_TPIYN) * sphi) - ch * (pk(DP_PITX) * cphi + pk(DP_PITY) * sphi));
  it.q3 = pk(DP_PIXX) * cphi * cphi + pk(DP_PIYY) * sphi * sphi + 2.0 * pk(DP_PIXY) * cphi * sphi - fold;
  it.alphaB = pk(DP_ALPHAB);
  it.v1 = pk(DP_VT) * ch - pk(DP_TVN) * sh;
  it.v2 = pk(DP_VX) * cphi + pk(DP_VY) * sphi;
  it.G0 = pk(DP_G0); it.G1 = pk(DP_G1);
  it.pad = 0.0; it.unused_ = 0.0;
  return it;
}

// per-bin registers.  The spectra kernels give one thread R species at ONE pT node, so everything that multiplies
// pT alone (DfShared) is formed once per item and shared by the thread's R evaluations.
struct DfBin {
  double mT, mT2, baryon, sign;
};

struct DfShared {
  double pb, pd, pq2, pq3, pv;          // pT bT, pT d1, pT q2, pT^2 q3, pT v2
};

template <bool BARYON>
IS3D_HD DfShared df_share(const DfItem &it, double pT, double pT2)
{
  DfShared s;
  s.pb = pT * it.bT; s.pd = pT * it.d1; s.pq2 = pT * it.q2; s.pq3 = pT2 * it.q3;
  s.pv = BARYON ? pT * it.v2 : 0.0;
  return s;
}

// The distribution feq (1 + df) at one momentum (MomentumSpectra.cpp:317-359)
// PAD: the item's additive term outside feqbar (PTB fallback only: delta_z - 3 delta_lambda); off = the slot is never read
template <int MODE, bool BARYON, bool REGULATE, bool PAD = false>
IS3D_HD double df_distribution(const DfItem &it, const DfShared &s, const DfBin &b, const double *__restrict__ exptab)
{
  double xE = fma(b.mT, it.aT, -s.pb);
  double x = xE;
  if (BARYON) x = fma(-b.baryon, it.alphaB, xE);
  const double q = fast_exp(x, exptab) + b.sign;                   // e^x + sign
  double feq, rxE = 0.0;
  if (MODE == 1) {
    feq = fast_rcp(q);
  } else {
    // df_mode 2 also needs 1/xE: one reciprocal of the product serves both (e^x <= 2.1e295 by fast_exp's clamp, so the
    // product stays finite for any xE a surface can produce)
    const double y = fast_rcp(q * xE);
    feq = y * xE;
    rxE = y * q;
  }
  double feqbar = fma(-b.sign, feq, 1.0);
  double pipp = fma(b.mT2, it.q1, fma(b.mT, s.pq2, s.pq3));      // shear + the folded bulk m^2 term
  double dfv;
  if (MODE == 1) {
    double lin = it.KX * xE;                         // (K1 b + K2 xE)
    if (BARYON) lin = fma(it.K1, b.baryon, lin);
    dfv = fma(lin, xE, pipp);
    if (BARYON) {
      double Vp = fma(b.mT, it.v1, -s.pv);
      dfv = fma(fma(it.G1, xE, it.G0 * b.baryon), Vp, dfv);
    }
  } else {
    const double r = rxE;
    dfv = fma(pipp, r, it.KX * xE);
    if (BARYON) {
      dfv = fma(it.K1, b.baryon, dfv);
      double Vp = fma(b.mT, it.v1, -s.pv);
      dfv = fma(fma(-it.G1 * b.baryon, r, it.G0), Vp, dfv);
    }
  }
  double df = PAD ? fma(feqbar, dfv, it.pad) : feqbar * dfv;
  if (REGULATE) df = fmax(-1.0, fmin(df, 1.0));
  return fma(feq, df, feq);
}

// One integrand evaluation: returns w * p.dsigma * feq (1 + df)   (MomentumSpectra.cpp:304-361)
template <int MODE, bool BARYON, bool REGULATE, bool OUTFLOW, bool PAD = false>
IS3D_HD double df_eval(const DfItem &it, const DfShared &s, const DfBin &b, const double *__restrict__ exptab)
{
  double pds = fma(b.mT, it.c1, s.pd);
  double contrib = pds * df_distribution<MODE, BARYON, REGULATE, PAD>(it, s, b, exptab);
  if (OUTFLOW) contrib = (pds <= 0.0) ? 0.0 : contrib;
  return contrib;
}

// ---------------------------------------------------------------------------------------------------------------------
// K1 production path: thread groups with ONE baryon number.
//
// df_spectra_kernel gives a thread R species classes at one pT node, and the host orders the classes so that the R
// classes of a thread carry the same baryon number b (run_spectra_df, "slot table").  With b and pT both fixed per
// thread, every term of df that is polynomial in (mT, pT) collapses into coefficients formed ONCE per (item, thread)
// and shared by the thread's R evaluations:
//   df_mode 1 (MomentumSpectra.cpp:325-338): the whole correction is a quadratic form,
//       (K1 b + K2 xE) xE + K0 m^2 + pi.p.p + (G1 xE + G0 b) V.p  =  mT^2 Q1 + mT (pT Q2 + b M1) + (pT^2 Q3 - b pT M2)
//       Q1 = q1 + K2 aT^2 + G1 aT v1,  Q2 = q2 - 2 K2 aT bT - G1 (aT v2 + bT v1),  Q3 = q3 + K2 bT^2 + G1 bT v2,
//       M1 = K1 aT + G0 v1,  M2 = K1 bT + G0 v2                      (q1..q3 already hold the folded K0 m^2 term)
//   df_mode 2 (:339-352): the terms over xE and the regular terms separate,
//       (pi.p.p - K2 m^2 - G1 b V.p) / xE + K0 xE + K1 b + G0 V.p  =  [mT^2 q1 + mT A + B] / xE + (mT L1 + C)
//       A = pT q2 - b (G1 v1),  B = pT^2 q3 + b pT (G1 v2),  L1 = K0 aT + G0 v1,  C = K1 b - pT (K0 bT + G0 v2)
//   both: exp((u.p - b mu_B)/T) = exp(xE) exp(-b alpha_B); the second factor is an item constant selected by b, so the
//       chemical-potential shift and the "+ sign" become one FMA.
// FP64-pipe instructions per evaluation with baryon terms (SASS, R = 4): mode 1 28.7 -> 20.5, mode 2 31.7 -> 26.
// eb[2 + b] = exp(-b alpha_B) for b = -2 .. 2: hadrons and the deuteron, the only nucleus the reference's PDG readers
// produce (readindata.cpp:1098-1214); a thread reads the slot of its group's baryon number with one LDS.64.
//
// df_mode 2 without regulate_deltaf goes one step further (FOLD): with y = 1/((e^x + sign) xE), feq = y xE and 1/xE = y (e^x + sign),
//       feq (1 + df) = y xE + y feqbar [quad + lin xE] = y (xE + feqbar quad'),   quad' = quad + lin xE,
// and quad' is again a quadratic form in mT whose coefficients are formed per (item, thread):
//       q1' = q1 + L1 aT (per item),  A' = A + C aT - L1 pT bT,  B' = B - C pT bT       (lin = mT L1 + C)
// without baryon terms (lin = K0 xE):  q1' = q1 + K0 aT^2,  A' = A - 2 K0 aT pT bT,  B' = B + K0 (pT bT)^2.
// The evaluation then needs neither 1/xE nor df itself: 24 instead of 26 FP64-pipe instructions (regulate_deltaf clamps
// df, so that variant keeps the explicit form).
constexpr int kMaxBaryon = 2;
struct alignas(16) DfItemU {
  double aT, bT, c1, d1;
  double q1, q2, q3, L1;         // mode 1: Q1, Q2, Q3, M1;  mode 2 without baryon terms: L1 = K0
  double L2, K1, Gv1, Gv2;       // mode 1: L2 = M2;  K1, Gv1 = G1 v1, Gv2 = G1 v2: mode 2 with baryon terms only
  double eb[2 * kMaxBaryon + 1];
};

template <int MODE, bool BARYON, bool FOLD = false, class PackFn>
IS3D_HD DfItemU df_make_item_u(PackFn pk, double sh, double ch, double cphi, double sphi, double w)
{
  DfItemU it;
  const double aT = ch * pk(DP_UTT) - sh * pk(DP_TUNT);
  const double bT = cphi * pk(DP_UXT) + sphi * pk(DP_UYT);
  it.aT = aT; it.bT = bT;
  it.c1 = w * (ch * pk(DP_DAT) + sh * pk(DP_DANT));
  it.d1 = w * (cphi * pk(DP_DAX) + sphi * pk(DP_DAY));
  const double fold = (MODE == 1) ? pk(DP_K0) : -pk(DP_K2);
  const double q1 = ch * ch * pk(DP_PITT) + sh * sh * pk(DP_T2PINN) - 2.0 * ch * sh * pk(DP_TPITN) + fold;
  const double q2 = 2.0 * (sh * (pk(DP_TPIXN) * cphi + pk(DP_TPIYN) * sphi) - ch * (pk(DP_PITX) * cphi + pk(DP_PITY) * sphi));
  const double q3 = pk(DP_PIXX) * cphi * cphi + pk(DP_PIYY) * sphi * sphi + 2.0 * pk(DP_PIXY) * cphi * sphi - fold;
  double v1 = 0.0, v2 = 0.0, G0 = 0.0, G1 = 0.0, K1 = 0.0;
  for (int i = 0; i < 2 * kMaxBaryon + 1; i++) it.eb[i] = 1.0;
  if (BARYON) {
    v1 = pk(DP_VT) * ch - pk(DP_TVN) * sh;
    v2 = pk(DP_VX) * cphi + pk(DP_VY) * sphi;
    G0 = pk(DP_G0); G1 = pk(DP_G1); K1 = pk(DP_K1);
    const double ebp = pk(DP_EBP), ebm = pk(DP_EBM);
    it.eb[0] = ebm * ebm; it.eb[1] = ebm; it.eb[3] = ebp; it.eb[4] = ebp * ebp;
  }
  if (MODE == 1) {
    const double K2 = pk(DP_K2);
    it.q1 = q1 + K2 * aT * aT + G1 * aT * v1;
    it.q2 = q2 - 2.0 * K2 * aT * bT - G1 * (aT * v2 + bT * v1);
    it.q3 = q3 + K2 * bT * bT + G1 * bT * v2;
    it.L1 = K1 * aT + G0 * v1;
    it.L2 = K1 * bT + G0 * v2;
  } else {
    const double K0 = pk(DP_K0);
    it.q1 = q1; it.q2 = q2; it.q3 = q3;
    it.L1 = BARYON ? K0 * aT + G0 * v1 : K0;
    it.L2 = K0 * bT + G0 * v2;
    if (FOLD) {
      it.q1 = fma(it.L1, aT, q1);                                   // q1 + L1 aT
      if (!BARYON) { it.q1 = fma(K0 * aT, aT, q1); it.L2 = 2.0 * K0 * aT; }   // L2 slot: 2 K0 aT
    }
  }
  it.K1 = K1; it.Gv1 = G1 * v1; it.Gv2 = G1 * v2;
  return it;
}

// Upper bound of |w p.dsigma feq (1 + df)| over every (class, pT) column a block can hold, for an item whose xE = u.p/T is at
// least xe_lo > 0 on all of them (mT <= mT_hi, pT <= pT_hi, |b| <= kMaxBaryon).  Used for the items df_spectra_kernel drops as
// negligible: the sum of these bounds per block row is compared with the finished spectra (a-posteriori check, spectra_df.cu).
// Term by term from df_share_u / df_eval_u_tail with absolute values; feq <= 1 / (e^xE min_b e^(-b alphaB) - 1).  +inf when the
// bound cannot be formed (the check then fails and the call is repeated without dropping anything).
template <int MODE, bool BARYON, bool REGULATE>
IS3D_HD double df_item_term_bound(const DfItemU &it, double xe_lo, double mT_hi, double pT_hi, const double *__restrict__ exptab)
{
  const double bm = BARYON ? (double)kMaxBaryon : 0.0;
  double ebmin = 1.0;
  if (BARYON) {
    for (int i = 0; i < 2 * kMaxBaryon + 1; i++) ebmin = fmin(ebmin, it.eb[i]);
  }
  const double E = fast_exp(fmin(xe_lo, 680.0), exptab) * ebmin;
  if (!(E > 4.0) || !(xe_lo > 0.0)) return as_double(0x7ff0000000000000ll);
  const double feq_max = 1.0 / (E - 1.0);
  const double P = mT_hi * fabs(it.c1) + pT_hi * fabs(it.d1);
  const double pbmax = pT_hi * fabs(it.bT);
  double df_max;
  if (MODE == 1) {
    const double A = pT_hi * fabs(it.q2) + bm * fabs(it.L1), B = pT_hi * pT_hi * fabs(it.q3) + bm * pT_hi * fabs(it.L2);
    df_max = 2.0 * (mT_hi * mT_hi * fabs(it.q1) + mT_hi * A + B);            // |feqbar| <= 2
    if (REGULATE) df_max = fmin(df_max, 1.0);
  } else if (REGULATE) {
    df_max = 1.0;
  } else {
    double A, B;
    if (BARYON) {
      const double C = fabs(it.K1) * bm + pT_hi * fabs(it.L2);
      A = pT_hi * fabs(it.q2) + bm * fabs(it.Gv1) + C * fabs(it.aT) + fabs(it.L1) * pbmax;
      B = pT_hi * pT_hi * fabs(it.q3) + bm * pT_hi * fabs(it.Gv2) + C * pbmax;
    } else {
      A = pT_hi * fabs(it.q2) + fabs(it.L2) * pbmax;                         // L2 slot: 2 K0 aT
      B = pT_hi * pT_hi * fabs(it.q3) + fabs(it.L1) * pbmax * pbmax;         // L1 = K0
    }
    df_max = 2.0 * (mT_hi * mT_hi * fabs(it.q1) + mT_hi * A + B) / xe_lo;    // feqbar quad' / xE
  }
  const double bound = 1.001 * P * feq_max * (1.0 + df_max);
  return bound == bound ? bound : as_double(0x7ff0000000000000ll);
}

// what reduce_partials_kernel needs to test the dropped items against the finished bins (bsum == nullptr: no test)
struct PruneCheck {
  const double *bsum = nullptr;        // [block row group][Ny * Nphi]: summed bounds of the dropped terms
  const int *bin_row = nullptr;        // [(class, pT) bin] -> block row group of bsum
  int Ny = 1, NyNphi = 1;
  double eps = 0.0;                    // a bin passes when bound <= eps |bin|
  unsigned long long *violations = nullptr;
};

// thread constants of the uniform-baryon path
struct DfThreadU {
  double pT, pT2, b, bpT;        // b = the group's baryon number, bpT = b pT
  int eslot;                     // kMaxBaryon + b: the thread's slot of DfItemU::eb
};

// per (item, thread) coefficients shared by the thread's R evaluations
struct DfSharedU {
  double pb, pd, A, B, C, eb;
};

template <int MODE, bool BARYON, bool FOLD = false>
IS3D_HD DfSharedU df_share_u(const DfItemU &it, const DfThreadU &th)
{
  DfSharedU s;
  s.pb = th.pT * it.bT;
  s.pd = th.pT * it.d1;
  s.eb = 1.0; s.C = 0.0;
  if (MODE == 1) {
    s.A = th.pT * it.q2;
    s.B = th.pT2 * it.q3;
    if (BARYON) { s.A = fma(th.b, it.L1, s.A); s.B = fma(-th.bpT, it.L2, s.B); }
  } else {
    s.A = th.pT * it.q2;
    s.B = th.pT2 * it.q3;
    if (BARYON) {
      s.A = fma(-th.b, it.Gv1, s.A);
      s.B = fma(th.bpT, it.Gv2, s.B);
      s.C = fma(it.K1, th.b, -th.pT * it.L2);
    }
    if (FOLD) {                                  // quad' = quad + lin xE (see above)
      if (BARYON) {
        s.A = fma(-it.L1, s.pb, fma(s.C, it.aT, s.A));
        s.B = fma(-s.C, s.pb, s.B);
      } else {
        s.A = fma(-it.L2, s.pb, s.A);              // L2 slot = 2 K0 aT
        s.B = fma(it.L1 * s.pb, s.pb, s.B);        // L1 = K0
      }
    }
  }
  if (BARYON) s.eb = it.eb[th.eslot];
  return s;
}

// One integrand evaluation w p.dsigma feq (1 + df) on the uniform-baryon path (MomentumSpectra.cpp:304-361), in two steps:
// df_eval_u_x / fast_exp give xE = u.p/T and e^xE, which do NOT depend on the baryon number; df_eval_u_tail does the rest.
// A baryon class and its antibaryon class (same mass, same statistics, b -> -b) therefore share xE and the exponential:
// the pair path of df_spectra_kernel evaluates both tails from one exp (charge-conjugate pairs, spectra_df.cu).
template <int MODE, bool BARYON, bool REGULATE, bool OUTFLOW>
IS3D_HD double df_eval_u_tail(const DfItemU &it, const DfSharedU &s, double mT, double mT2, double sign, double xE, double e,
                              int spare = 0);

// CLAMP = false: the caller knows that no xE of this item can reach 680 (df_item_needs_clamp)
template <bool CLAMP = true>
IS3D_HD double df_eval_u_x(const DfItemU &it, const DfSharedU &s, double mT)
{
  // clamped in place (x <= 680, common.cuh): beyond that feq < 1e-295 and every later use of xE multiplies feq
  const double xE = fma(mT, it.aT, -s.pb);
  return CLAMP ? clamp_hi_word_680(xE) : xE;
}

// true when some (class, pT) bin of the launch could see xE = mT aT - pT bT above 600 for this item (mT_max, pT_max: the
// largest table entries; aT > 0 for a time-like flow velocity).  NaN coefficients also say true.
IS3D_HD bool df_item_needs_clamp(double aT, double bT, double mT_max, double pT_max)
{
  return !(fma(mT_max, aT, pT_max * fabs(bT)) < 600.0);
}

template <int MODE, bool BARYON, bool REGULATE, bool OUTFLOW, bool CLAMP = true>
IS3D_HD double df_eval_u(const DfItemU &it, const DfSharedU &s, double mT, double mT2, double sign,
                         const double *__restrict__ exptab)
{
  const double xE = df_eval_u_x<CLAMP>(it, s, mT);
  int spare;
  const double e = fast_exp_k<false>(xE, exptab, spare);
  return df_eval_u_tail<MODE, BARYON, REGULATE, OUTFLOW>(it, s, mT, mT2, sign, xE, e, spare);
}

// the part of an evaluation behind the reciprocal: rc = 1 / (q xE) (df_mode 2) or 1 / q (df_mode 1), q = e^x + sign
template <int MODE, bool BARYON, bool REGULATE, bool OUTFLOW>
IS3D_HD double df_eval_u_finish(const DfItemU &it, const DfSharedU &s, double mT, double mT2, double sign, double xE, double q, double rc)
{
  const double quad = fma(mT2, it.q1, fma(mT, s.A, s.B));      // Horner form without mT^2: same speed (profiles/r02_k1_variants_horner_pairR.txt)
  if (MODE == 2 && !REGULATE) {
    // folded form: (it.q1, s.A, s.B) hold quad' = quad + lin xE (df_make_item_u / df_share_u with FOLD = true)
    const double y = rc;
    const double feq = y * xE;
    const double feqbar = fma(-sign, feq, 1.0);
    const double pds = fma(mT, it.c1, s.pd);
    double contrib = (pds * y) * fma(feqbar, quad, xE);
    if (OUTFLOW) contrib = (pds <= 0.0) ? 0.0 : contrib;
    return contrib;
  }
  double feq, dfv;
  if (MODE == 1) {
    feq = rc;
    dfv = quad;
  } else {
    const double y = rc;
    feq = y * xE;
    const double r = y * q;
    dfv = fma(quad, r, BARYON ? fma(mT, it.L1, s.C) : it.L1 * xE);
  }
  const double feqbar = fma(-sign, feq, 1.0);
  double df = feqbar * dfv;
  if (REGULATE) df = fmax(-1.0, fmin(df, 1.0));
  const double pds = fma(mT, it.c1, s.pd);
  double contrib = pds * fma(feq, df, feq);
  if (OUTFLOW) contrib = (pds <= 0.0) ? 0.0 : contrib;
  return contrib;
}

template <int MODE, bool BARYON, bool REGULATE, bool OUTFLOW>
IS3D_HD double df_eval_u_tail(const DfItemU &it, const DfSharedU &s, double mT, double mT2, double sign, double xE, double e,
                              int spare)
{
  const double q = BARYON ? fma(e, s.eb, sign) : e + sign;         // e^x + sign, x = xE - b alpha_B
  // df_mode 2: one reciprocal serves 1/(e^x + sign) and 1/xE (e^x <= 2.1e295 by fast_exp's clamp and exp(|b| alpha_B) < 1e4, so
  // the product stays finite for any xE a surface can produce)
  const double rc = fast_rcp(MODE == 1 ? q : q * xE, spare);
  return df_eval_u_finish<MODE, BARYON, REGULATE, OUTFLOW>(it, s, mT, mT2, sign, xE, q, rc);
}

// A charge-conjugate pair (s: the baryon member, sm: its antibaryon partner) from ONE reciprocal: with D = q qm [xE],
// 1 / (q [xE]) = qm / D and 1 / (qm [xE]) = q / D -- one FP64 instruction (df_mode 2) and one MUFU seed fewer than two
// reciprocals.  Only where D cannot overflow: callers use it for items whose every xE is below kXePairShared.
constexpr double kXePairShared = 340.0;
template <int MODE, bool REGULATE, bool OUTFLOW>
IS3D_HD void df_eval_u_pair_shared(const DfItemU &it, const DfSharedU &s, const DfSharedU &sm, double mT, double mT2, double sign,
                                   double xE, double e, int spare, double &acc, double &accm)
{
  const double q = fma(e, s.eb, sign), qm = fma(e, sm.eb, sign);
  const double qx = MODE == 1 ? q : q * xE;
  const double Y = fast_rcp(qx * qm, spare);
  acc += df_eval_u_finish<MODE, true, REGULATE, OUTFLOW>(it, s, mT, mT2, sign, xE, q, Y * qm);
  accm += df_eval_u_finish<MODE, true, REGULATE, OUTFLOW>(it, sm, mT, mT2, sign, xE, qm, Y * q);
}

}  // namespace is3d
