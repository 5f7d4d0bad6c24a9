// K2 math: continuous spectra with the PTM / PTB modified equilibrium distributions (df_mode 3, 4).
// Reference: EmissionFunctionArray::calculate_dN_pTdpTdphidy_feqmod, src/cpp/MomentumSpectra.cpp:419-1044.
//
// Non-breakdown cells evaluate f = |renorm| / (exp(E'/T' - b alphaB') + sign) with p' = A^-1 p_LRF.  p_LRF is linear
// in (mT, pT):  p_LRF = mT (ch a + sh b) + pT (cphi c + sphi d)  with a,b,c,d built from the Milne basis, so
//     E'^2 / T'^2 = m^2/T'^2 + mT^2 h1 + mT pT h2 + pT^2 h3,   h = quadratic forms of A^-1{a,b,c,d}/T'
// per (cell, y, phi): 4 FMAs + one square root per momentum bin instead of a 3x3 solve.  The reference's LU inverse
// plus per-momentum iterative refinement (:954-971) is replaced by a refined cofactor inverse per cell.
// Breakdown cells (and the "narrow" |y - eta| < detA case, :865-871) fall back to the linear df of K1 with the
// PTM / PTB coefficient sets, evaluated by the same df_eval as df_mode 2.
#pragma once

#include "gauss_thermal.cuh"
#include "spectra_df.cuh"

namespace is3d {

enum FeqmodPackIdx {
  // 0 .. DP_SIZE-1: the linear-df pack of spectra_df.cuh (used on breakdown)
  FP_BREAKDOWN = DP_SIZE, FP_DETA, FP_ETA_SCALE, FP_RENORM, FP_IT2, FP_ALPHAB_MOD,
  FP_A1X, FP_A1Y, FP_A1Z, FP_A2X, FP_A2Y, FP_A2Z, FP_A3X, FP_A3Y, FP_A3Z, FP_A4X, FP_A4Y, FP_A4Z,
  FP_ADD,                       // PTB linear df: additive delta_z - 3 delta_lambda (outside feqbar)
  // inputs of the per-(cell, species) PTM renormalisation
  FP_T, FP_TMOD, FP_DNFACT, FP_G, FP_F_T2, FP_RENORM_DIV,
  FP_SIZE
};

struct FeqmodFlags {
  int df_mode;                 // 3 or 4
  int dimension;
  int include_baryon, include_bulk, include_shear, include_baryondiff;
  double deta_min, mass_pion0, bulkPi_over_P_max;
  int clamp_inclusive = 0;     // PTB bulk clamp: spectra use < / > (MomentumSpectra.cpp:607-614), dN/dX and the sampler
                               // use <= / >= (SpacetimeDistribution.cpp:784-785, ParticleSampler.cpp:552-559)
};

enum { CELL_BREAKDOWN = 4, CELL_PL_NEGATIVE = 8 };

// does_feqmod_breakdown with fast = 0 (EmissionFunction.cpp:65-109)
IS3D_HD bool feqmod_breaks_down(const FeqmodFlags &fl, double T, double F, double bulkPi, double betabulk, double detA,
                                double z, const double *gla_root, const double *gla_weight, int gla_pts)
{
  if (fl.df_mode == 3) {
    const double *r1 = gla_root + 1 * gla_pts, *w1 = gla_weight + 1 * gla_pts;
    const double *r2 = gla_root + 2 * gla_pts, *w2 = gla_weight + 2 * gla_pts;
    double mbar = fl.mass_pion0 / T;
    double neq_fact = T * T * T / kTwoPi2HbarC3, J20_fact = T * neq_fact;
    double neq = neq_fact * gauss_thermal<TI_NEQ>(r1, w1, gla_pts, mbar, 0., 0., -1.);
    double J20 = J20_fact * gauss_thermal<TI_J20>(r2, w2, gla_pts, mbar, 0., 0., -1.);
    double dn = bulkPi * (neq + J20 * F / T / T) / betabulk;
    return detA <= fl.deta_min || (neq + dn < 0.0);
  }
  return detA <= fl.deta_min || z < 0.0;
}

// Per-cell prologue (MomentumSpectra.cpp:516-773).  Returns CELL_* bits.
IS3D_HD int feqmod_setup_cell(const Cell &c, const DfTables &tb, const FeqmodFlags &fl, const double *gla_root,
                              const double *gla_weight, int gla_pts, double pack[FP_SIZE])
{
  for (int k = 0; k < FP_SIZE; k++) pack[k] = 0.0;
  double tau = c.tau, tau2 = tau * tau;
  double ux = c.ux, uy = c.uy, un = c.un;
  double ut = sqrt(1.0 + ux * ux + uy * uy + tau2 * un * un);
  if (ut * c.dat + ux * c.dax + uy * c.day + un * c.dan <= 0.0) return CELL_SKIPPED;
  int status = CELL_OK;
  double utperp = sqrt(1.0 + ux * ux + uy * uy);
  double T = c.T, P = c.P, E = c.E;
  Shear pi;
  if (fl.include_shear) pi = complete_shear(c.pixx, c.pixy, c.pixn, c.piyy, c.piyn, ut, ux, uy, un, tau2);
  double bulkPi = fl.include_bulk ? c.bulkPi : 0.0;
  double muB = 0.0, alphaB = 0.0, nB = 0.0, Vt = 0.0, Vx = 0.0, Vy = 0.0, Vn = 0.0, ber = 0.0;
  if (fl.include_baryon && fl.include_baryondiff) {
    muB = c.muB; nB = c.nB; Vx = c.Vx; Vy = c.Vy; Vn = c.Vn;
    Vt = (Vx * ux + Vy * uy + tau2 * Vn * un) / ut;
    alphaB = muB / T;
    ber = nB / (E + P);
  }
  if (fl.df_mode == 4) {        // keep Pi/P inside the PTB table (:603-615)
    if (fl.clamp_inclusive) {
      if (bulkPi <= -P) bulkPi = -(1.0 - 1.e-5) * P;
      else if (bulkPi / P >= fl.bulkPi_over_P_max) bulkPi = P * (fl.bulkPi_over_P_max - 1.e-5);
    } else {
      if (bulkPi < -P) bulkPi = -(1.0 - 1.e-5) * P;
      else if (bulkPi / P > fl.bulkPi_over_P_max) bulkPi = P * (fl.bulkPi_over_P_max - 1.e-5);
    }
  }
  double zt = tau * un / utperp, zn = ut / (tau * utperp);
  double pl = P + bulkPi + zt * zt * pi.tt + tau2 * tau2 * zn * zn * pi.nn + 2. * tau2 * zt * zn * pi.tn;
  if (pl < 0) status |= CELL_PL_NEGATIVE;

  DfCoeff df;
  if (!evaluate_df_coefficients(tb, fl.df_mode, fl.include_baryon, T, muB, E, P, bulkPi, &df)) return CELL_OUT_OF_TABLE;
  Basis b = milne_basis(ut, ux, uy, un, tau);
  ShearLRF l = boost_shear_to_lrf(pi, b, tau2);

  double T_mod = T, alphaB_mod = alphaB;
  if (fl.df_mode == 3) { T_mod = T + bulkPi * df.F / df.betabulk; alphaB_mod = alphaB + bulkPi * df.G / df.betabulk; }
  double shear_mod = 0.5 / df.betapi;
  double bulk_mod = (fl.df_mode == 4) ? df.lambda : bulkPi / (3.0 * df.betabulk);
  double A[9];
  A[0] = 1.0 + l.xx * shear_mod + bulk_mod; A[1] = l.xy * shear_mod; A[2] = l.xz * shear_mod;
  A[3] = A[1]; A[4] = 1.0 + l.yy * shear_mod + bulk_mod; A[5] = l.yz * shear_mod;
  A[6] = A[2]; A[7] = A[5]; A[8] = 1.0 + l.zz * shear_mod + bulk_mod;
  double detA = A[0] * (A[4] * A[8] - A[5] * A[5]) - A[1] * (A[1] * A[8] - A[5] * A[2]) + A[2] * (A[1] * A[5] - A[4] * A[2]);
  double detA_bulk_two_thirds = (1.0 + bulk_mod) * (1.0 + bulk_mod);
  double Ainv[9], det_unused;
  invert3x3(A, Ainv, &det_unused);

  bool breaks = feqmod_breaks_down(fl, T, df.F, bulkPi, df.betabulk, detA, df.z, gla_root, gla_weight, gla_pts);
  if (breaks) status |= CELL_BREAKDOWN;
  double eta_scale = 1.0;
  if (detA > fl.deta_min && fl.dimension == 2) eta_scale = detA / detA_bulk_two_thirds;

  // ---- linear-df pack (breakdown branch, :887-928) ----
  double invT = 1.0 / T;
  double shear_coeff = 0.5 / (df.betapi * T);
  double sc = shear_coeff * invT, K0, K1, K2, G0, G1, add = 0.0;
  if (fl.df_mode == 3) {
    double bulk0 = df.F / (T * T * df.betabulk) * bulkPi, bulk1 = df.G / df.betabulk * bulkPi, bulk2 = bulkPi / (3.0 * T * df.betabulk);
    K0 = (bulk0 + bulk2) * T; K1 = bulk1; K2 = bulk2 * invT;
    G0 = ber / df.betaV; G1 = invT / df.betaV;
  } else {                     // PTB: df = feqbar (shear + dl (xE - m^2/T^2 / xE)) + dz - 3 dl
    K0 = df.delta_lambda; K1 = 0.0; K2 = df.delta_lambda * invT * invT; G0 = 0.0; G1 = 0.0;
    add = df.delta_z - 3.0 * df.delta_lambda;
    alphaB = 0.0;              // the PTB fallback has no chemical-potential term (:913)
  }
  pack[DP_VALID] = 1.0;
  pack[DP_ETA] = c.eta;
  pack[DP_UTT] = ut * invT; pack[DP_TUNT] = tau * un * invT; pack[DP_UXT] = ux * invT; pack[DP_UYT] = uy * invT;
  pack[DP_ALPHAB] = alphaB;
  pack[DP_DAT] = c.dat; pack[DP_DAX] = c.dax; pack[DP_DAY] = c.day; pack[DP_DANT] = c.dan / tau;
  pack[DP_PITT] = sc * pi.tt; pack[DP_T2PINN] = sc * tau2 * pi.nn; pack[DP_TPITN] = sc * tau * pi.tn;
  pack[DP_PITX] = sc * pi.tx; pack[DP_PITY] = sc * pi.ty; pack[DP_TPIXN] = sc * tau * pi.xn; pack[DP_TPIYN] = sc * tau * pi.yn;
  pack[DP_PIXX] = sc * pi.xx; pack[DP_PIYY] = sc * pi.yy; pack[DP_PIXY] = sc * pi.xy;
  pack[DP_K0] = K0; pack[DP_K1] = K1; pack[DP_K2] = K2; pack[DP_G0] = G0; pack[DP_G1] = G1;
  pack[DP_VT] = Vt; pack[DP_TVN] = tau * Vn; pack[DP_VX] = Vx; pack[DP_VY] = Vy;
  pack[FP_ADD] = add;

  // ---- modified-distribution pack ----
  pack[FP_BREAKDOWN] = breaks ? 1.0 : 0.0;
  pack[FP_DETA] = detA;
  pack[FP_ETA_SCALE] = eta_scale;
  double renorm_div = (fl.dimension == 2) ? detA_bulk_two_thirds : detA;
  double renorm = 1.0;
  if (fl.include_bulk && fl.df_mode == 4) renorm = df.z;
  renorm /= renorm_div;
  if (not_finite(renorm)) renorm = 0.0;          // the reference skips the species (:828-832)
  pack[FP_RENORM] = fabs(renorm);
  double iTm = 1.0 / T_mod;
  pack[FP_IT2] = iTm * iTm;
  pack[FP_ALPHAB_MOD] = alphaB_mod;
  // p_LRF = mT (ch a + sh b) + pT (cphi c + sphi d),  a = (-Xt, 0, -Zt), b = tau (Xn, 0, Zn), c = (Xx, Yx, 0), d = (Xy, Yy, 0)
  const double va[3] = {-b.Xt, 0.0, -b.Zt}, vb[3] = {tau * b.Xn, 0.0, tau * b.Zn};
  const double vc[3] = {b.Xx, b.Yx, 0.0}, vd[3] = {b.Xy, b.Yy, 0.0};
  for (int i = 0; i < 3; i++) {
    pack[FP_A1X + i] = iTm * (Ainv[3 * i] * va[0] + Ainv[3 * i + 1] * va[1] + Ainv[3 * i + 2] * va[2]);
    pack[FP_A2X + i] = iTm * (Ainv[3 * i] * vb[0] + Ainv[3 * i + 1] * vb[1] + Ainv[3 * i + 2] * vb[2]);
    pack[FP_A3X + i] = iTm * (Ainv[3 * i] * vc[0] + Ainv[3 * i + 1] * vc[1] + Ainv[3 * i + 2] * vc[2]);
    pack[FP_A4X + i] = iTm * (Ainv[3 * i] * vd[0] + Ainv[3 * i + 1] * vd[1] + Ainv[3 * i + 2] * vd[2]);
  }
  pack[FP_T] = T; pack[FP_TMOD] = T_mod; pack[FP_DNFACT] = bulkPi / df.betabulk; pack[FP_G] = df.G;
  pack[FP_F_T2] = df.F / T / T; pack[FP_RENORM_DIV] = renorm_div;
  return status;
}

// PTM renormalisation n_linear / n_mod of one species in one cell (MomentumSpectra.cpp:795-826); 0 encodes the
// reference's "skip this species" for a NaN / inf factor.
template <class PackFn>
IS3D_HD double feqmod_renorm_ptm(PackFn pk, double mass, double degeneracy, double baryon, double sign,
                                 const double *gla_root, const double *gla_weight, int pts)
{
  const double *r1 = gla_root + 1 * pts, *w1 = gla_weight + 1 * pts, *r2 = gla_root + 2 * pts, *w2 = gla_weight + 2 * pts;
  double T = pk(FP_T), T_mod = pk(FP_TMOD), alphaB = pk(DP_ALPHAB), alphaB_mod = pk(FP_ALPHAB_MOD);
  double neq_fact = T * T * T / kTwoPi2HbarC3, J20_fact = T * neq_fact, N10_fact = neq_fact;
  double nmod_fact = T_mod * T_mod * T_mod / kTwoPi2HbarC3;
  double mbar = mass / T, mbar_mod = mass / T_mod;
  double neq = neq_fact * degeneracy * gauss_thermal<TI_NEQ>(r1, w1, pts, mbar, alphaB, baryon, sign);
  double N10 = baryon * N10_fact * degeneracy * gauss_thermal<TI_J10>(r1, w1, pts, mbar, alphaB, baryon, sign);
  double J20 = J20_fact * degeneracy * gauss_thermal<TI_J20>(r2, w2, pts, mbar, alphaB, baryon, sign);
  double n_linear = neq + pk(FP_DNFACT) * (neq + N10 * pk(FP_G) + J20 * pk(FP_F_T2));
  double n_mod = nmod_fact * degeneracy * gauss_thermal<TI_NEQ>(r1, w1, pts, mbar_mod, alphaB_mod, baryon, sign);
  double renorm = (n_linear / n_mod) / pk(FP_RENORM_DIV);
  if (not_finite(renorm)) return 0.0;
  return fabs(renorm);
}

// item constants of the modified branch
// The mass term of E'^2 / T'^2 = m^2 / T'^2 + mT^2 h1 + mT pT h2 + pT^2 h3 is folded into the quadratic form with
// m^2 = mT^2 - pT^2 (h1 += 1/T'^2, h3 -= 1/T'^2), so the momentum loop never needs m^2.
// eb[kMaxBaryon + b] = exp(-b alphaB') for b = -2 .. 2 (K2's uniform-baryon thread groups, see DfItemU in spectra_df.cuh):
// exp(E'/T' - b alphaB') + sign = exp(E'/T') eb + sign is one FMA.  Filled only when asked for (eb_slots).
struct alignas(16) FeqmodItem {
  double c1, d1, h1, h2;
  double h3, renorm, alphaB_mod, eb[2 * kMaxBaryon + 1];
};

// sh/ch = sinh, cosh of (y - eta_scale eta); eta weight placement of the feqmod spectra path:
// w (p^tau ds_tau + p^x ds_x + p^y ds_y) + p^eta ds_eta  (MomentumSpectra.cpp:883, :936); the dN/dX path weights the
// whole p.dsigma (SpacetimeDistribution.cpp:1022, :1075) -> w_on_dan
template <class PackFn>
IS3D_HD FeqmodItem feqmod_make_item(PackFn pk, double sh, double ch, double cphi, double sphi, double w, bool w_on_dan = false,
                                    bool eb_slots = false, bool fold_renorm = false)
{
  FeqmodItem it;
  it.c1 = w * ch * pk(DP_DAT) + (w_on_dan ? w : 1.0) * sh * pk(DP_DANT);
  it.d1 = w * (cphi * pk(DP_DAX) + sphi * pk(DP_DAY));
  double g1[3], g2[3];
  for (int i = 0; i < 3; i++) {
    g1[i] = ch * pk(FP_A1X + i) + sh * pk(FP_A2X + i);
    g2[i] = cphi * pk(FP_A3X + i) + sphi * pk(FP_A4X + i);
  }
  it.h1 = g1[0] * g1[0] + g1[1] * g1[1] + g1[2] * g1[2];
  it.h2 = 2.0 * (g1[0] * g2[0] + g1[1] * g2[1] + g1[2] * g2[2]);
  it.h3 = g2[0] * g2[0] + g2[1] * g2[1] + g2[2] * g2[2];
  const double iT2 = pk(FP_IT2);
  it.h1 += iT2; it.h3 -= iT2;
  it.alphaB_mod = pk(FP_ALPHAB_MOD);
  it.renorm = pk(FP_RENORM);
  if (fold_renorm) { it.c1 *= it.renorm; it.d1 *= it.renorm; }     // the cell's |renorm| >= 0 rides on p.dsigma
  for (int i = 0; i < 2 * kMaxBaryon + 1; i++) it.eb[i] = 1.0;
  if (eb_slots) {
    const double ebp = exp(-it.alphaB_mod), ebm = exp(it.alphaB_mod);
    it.eb[0] = ebm * ebm; it.eb[1] = ebm; it.eb[3] = ebp; it.eb[4] = ebp * ebp;
  }
  return it;
}

// linear-df item of the breakdown branch: as df_make_item but with the feqmod weight placement
template <class PackFn>
IS3D_HD DfItem feqmod_make_linear_item(PackFn pk, double sh, double ch, double cphi, double sphi, double w, bool w_on_dan = false)
{
  DfItem it = df_make_item(pk, 2, sh, ch, cphi, sphi, 1.0);
  it.c1 = w * ch * pk(DP_DAT) + (w_on_dan ? w : 1.0) * sh * pk(DP_DANT);
  it.d1 = w * (cphi * pk(DP_DAX) + sphi * pk(DP_DAY));
  it.pad = pk(FP_ADD);
  return it;
}

// sqrt(a) for a > 0 in the FMA pipe: hardware rsqrt seed y (sees the upper 32 bits of a: ~20 bits) and ONE third-order
// step, g = a y, e = 1 - g y (<= 2^-18), sqrt(a) = g (1 - e)^(-1/2) = g (1 + e/2 + 3 e^2/8) + O(5/16 e^3 < 2^-56):
// 5 FP64 instructions (the coupled Goldschmidt step + residual fix it replaces took 7)
IS3D_HD double fast_sqrt(double a)
{
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double g = a * y;
  const double e = fma(-g, y, 1.0);
  double p = fma(e, 0.375, 0.5);
  p = p * e;
  return fma(g, p, g);
#else
  return sqrt(a);
#endif
}

// per-item products with the thread's pT (shared by its R species, see DfShared)
struct FeqmodShared {
  double pd, ph2, ph3;                  // pT d1, pT h2, pT^2 h3
};

IS3D_HD FeqmodShared feqmod_share(const FeqmodItem &it, double pT, double pT2)
{
  FeqmodShared s;
  s.pd = pT * it.d1; s.ph2 = pT * it.h2; s.ph3 = pT2 * it.h3;
  return s;
}

// the modified distribution |renorm| / (exp(E'/T' - b alphaB') + sign) at one momentum (MomentumSpectra.cpp:941-979);
// renorm_sp = |renorm| of this (cell, species)
template <bool BARYON>
IS3D_HD double feqmod_distribution(const FeqmodItem &it, const FeqmodShared &s, const DfBin &b, double renorm_sp,
                                   const double *__restrict__ exptab)
{
  double e2 = fma(b.mT2, it.h1, fma(b.mT, s.ph2, s.ph3));
  double x = fast_sqrt(e2);
  if (BARYON) x = fma(-b.baryon, it.alphaB_mod, x);
  return renorm_sp * fast_rcp(fast_exp(x, exptab) + b.sign);
}

// K2 production path (thread groups with one baryon number, see DfItemU): eb = it.eb[kMaxBaryon + b] read once per
// (item, thread); with FOLDED the cell's renormalisation already multiplies c1 / d1 (feqmod_make_item fold_renorm),
// otherwise renorm_sp = |renorm| of this (cell, class).  FP64-pipe instructions per evaluation: 3 (E'^2) + 5 (sqrt) +
// 7 (exp) + 1 + 3 (rcp) + 2 (+ 1 unfolded) = 21 (22); the first version took 28 (29).
// Accumulates into acc with the final FMA inside the branch (returning the contribution lets the compiler merge the
// "+=" of the modified and the linear branch behind their join, which costs a DMUL + DADD instead of one DFMA).
// CLAMP = false: the caller knows E'/T' < kXePairShared for every column of its block (feqmod_item_range)
template <bool BARYON, bool OUTFLOW, bool FOLDED, bool CLAMP = true>
IS3D_HD void feqmod_accum_u(double &acc, const FeqmodItem &it, const FeqmodShared &s, double eb, double mT, double mT2, double sign,
                            double renorm_sp, const double *__restrict__ exptab)
{
  const double e2 = fma(mT2, it.h1, fma(mT, s.ph2, s.ph3));
  int spare;
  const double e = fast_exp_k<CLAMP>(fast_sqrt(e2), exptab, spare);
  const double f = fast_rcp(BARYON ? fma(e, eb, sign) : e + sign, spare);
  const double pds = fma(mT, it.c1, s.pd);
  const double sum = FOLDED ? fma(pds, f, acc) : fma(pds * f, renorm_sp, acc);
  acc = (OUTFLOW && pds <= 0.0) ? acc : sum;
}

// Charge-conjugate pair of classes (baryon class + its antibaryon class: same mass and statistics, b -> -b): E' and its
// exponential do not depend on b, so both members are evaluated from ONE sqrt + exp (16 shared FP64 instructions + 5-6 per
// member instead of 21-22 each).  eb / ebm = exp(-+|b| alphaB'), rn / rnm = the members' PTM renormalisations.
// CLAMP = false (exponents below kXePairShared): also ONE reciprocal, 1 / q = qm / (q qm), 1 / qm = q / (q qm).
template <bool OUTFLOW, bool FOLDED, bool CLAMP = true>
IS3D_HD void feqmod_accum_pair_u(double &acc, double &accm, const FeqmodItem &it, const FeqmodShared &s, double eb, double ebm, double mT,
                                 double mT2, double sign, double rn, double rnm, const double *__restrict__ exptab)
{
  const double e2 = fma(mT2, it.h1, fma(mT, s.ph2, s.ph3));
  int spare;
  const double e = fast_exp_k<CLAMP>(fast_sqrt(e2), exptab, spare);
  const double q = fma(e, eb, sign), qm = fma(e, ebm, sign);
  double f, fm;
  if (CLAMP) { f = fast_rcp(q, spare); fm = fast_rcp(qm, spare); }
  else { const double Y = fast_rcp(q * qm, spare); f = Y * qm; fm = Y * q; }
  const double pds = fma(mT, it.c1, s.pd);
  const double sum = FOLDED ? fma(pds, f, acc) : fma(pds * f, rn, acc);
  const double summ = FOLDED ? fma(pds, fm, accm) : fma(pds * fm, rnm, accm);
  acc = (OUTFLOW && pds <= 0.0) ? acc : sum;
  accm = (OUTFLOW && pds <= 0.0) ? accm : summ;
}

// Range of the exponent x = E'/T' - b alphaB' an item can produce on the columns of a block (mT in [mT_lo, mT_hi], pT <= pT_hi,
// pT <= mT, mass^2 >= m2_lo).  E'^2/T'^2 = |mT g1 + pT g2|^2 + m^2/T'^2 with |g1|^2 = h1 - 1/T'^2, |g2|^2 = h3 + 1/T'^2
// (feqmod_make_item), hence  (mT_lo max(|g1| - |g2|, 0))^2 + m2_lo/T'^2  <=  E'^2/T'^2  <=  (mT_hi |g1| + pT_hi |g2|)^2 + mT_hi^2/T'^2.
struct FeqmodRange { double lo, hi; };
IS3D_HD FeqmodRange feqmod_item_range(const FeqmodItem &it, double iT2, double mT_lo, double mT_hi, double pT_hi, double m2_lo, double shift)
{
  const double g1 = sqrt(fmax(it.h1 - iT2, 0.0)), g2 = sqrt(fmax(it.h3 + iT2, 0.0));
  const double a = mT_lo * fmax(g1 - g2, 0.0), b = fma(mT_hi, g1, pT_hi * g2);
  FeqmodRange r;
  r.lo = sqrt(fma(a, a, m2_lo * iT2)) - shift;
  r.hi = sqrt(fma(b, b, mT_hi * mT_hi * iT2)) + shift;
  return r;
}

// Upper bound of |w p.dsigma renorm / (e^x + sign)| over the columns of a block for an item whose exponent is at least x_lo
// (the chemical-potential shift already taken off); rn_max = the largest per-class renormalisation of the cell (1 when the
// cell's renormalisation is folded into c1 / d1).  +inf when it cannot be formed.
IS3D_HD double feqmod_item_term_bound(const FeqmodItem &it, double x_lo, double mT_hi, double pT_hi, double rn_max,
                                      const double *__restrict__ exptab)
{
  const double E = fast_exp(fmin(x_lo, 680.0), exptab);
  if (!(E > 4.0)) return as_double(0x7ff0000000000000ll);
  const double bound = 1.001 * (mT_hi * fabs(it.c1) + pT_hi * fabs(it.d1)) * fabs(rn_max) / (E - 1.0);
  return bound == bound ? bound : as_double(0x7ff0000000000000ll);
}

#if defined(__CUDACC__)
// ---- fused PTM renormalisation (device) ---------------------------------------------------------------------------
// The four 32-point Gauss-Laguerre sums of feqmod_renorm_ptm in one pass over the nodes: neq and J10 share their
// exp(Ebar - b alphaB); node constants w p e^p (alpha = 1) and w e^p (alpha = 2) are staged in shared memory by
// RenormNodes::load; exp / sqrt / reciprocals are the FP64-pipe versions.
constexpr int kRenormMaxPts = 64;

struct RenormNodes {
  double p1sq[kRenormMaxPts], c1[kRenormMaxPts], p2sq[kRenormMaxPts], c2[kRenormMaxPts];
  __device__ void load(const double *__restrict__ gla_root, const double *__restrict__ gla_weight, int gla_pts)
  {
    for (int k = threadIdx.x; k < gla_pts; k += blockDim.x) {
      const double r1 = gla_root[1 * gla_pts + k], w1 = gla_weight[1 * gla_pts + k];
      const double r2 = gla_root[2 * gla_pts + k], w2 = gla_weight[2 * gla_pts + k];
      p1sq[k] = r1 * r1; c1[k] = w1 * (r1 * exp(r1));
      p2sq[k] = r2 * r2; c2[k] = w2 * exp(r2);
    }
  }
};

template <class PackFn>
__device__ __forceinline__ double feqmod_renorm_ptm_fused(PackFn pk, double m, double g, double b, double sg, const RenormNodes &nd,
                                                          int gla_pts, const double *__restrict__ exptab)
{
  const double T = pk(FP_T), T_mod = pk(FP_TMOD), alphaB = pk(DP_ALPHAB), alphaB_mod = pk(FP_ALPHAB_MOD);
  const double mbar = m / T, mbar_mod = m / T_mod, mb2 = mbar * mbar, mm2 = mbar_mod * mbar_mod;
  const double chem = b * alphaB, chem_mod = b * alphaB_mod;
  double sneq = 0.0, sJ10 = 0.0, sJ20 = 0.0, smod = 0.0;
  for (int k = 0; k < gla_pts; k++) {
    const double E1 = fast_sqrt(nd.p1sq[k] + mb2);
    const double e1 = fast_exp(E1 - chem, exptab), iq1 = fast_rcp(e1 + sg);
    const double t1 = nd.c1[k] * iq1;
    sneq += t1;                                   // w p e^p / (e^(E - b alphaB) + sign)
    sJ10 += t1 * (e1 * iq1);                      // w p e^(p + E - b alphaB) / q^2
    const double E2 = fast_sqrt(nd.p2sq[k] + mb2);
    const double e2 = fast_exp(E2 - chem, exptab), iq2 = fast_rcp(e2 + sg);
    sJ20 += nd.c2[k] * E2 * (e2 * iq2 * iq2);     // w E e^(p + E - b alphaB) / q^2
    const double Em = fast_sqrt(nd.p1sq[k] + mm2);
    smod += nd.c1[k] * fast_rcp(fast_exp(Em - chem_mod, exptab) + sg);
  }
  const double neq_fact = T * T * T / kTwoPi2HbarC3, J20_fact = T * neq_fact;
  const double nmod_fact = T_mod * T_mod * T_mod / kTwoPi2HbarC3;
  const double neq = neq_fact * g * sneq, N10 = b * neq_fact * g * sJ10, J20 = J20_fact * g * sJ20;
  const double n_linear = neq + pk(FP_DNFACT) * (neq + N10 * pk(FP_G) + J20 * pk(FP_F_T2));
  const double n_mod = nmod_fact * g * smod;
  const double r = (n_linear / n_mod) / pk(FP_RENORM_DIV);
  return not_finite(r) ? 0.0 : fabs(r);
}
#endif

}  // namespace is3d
