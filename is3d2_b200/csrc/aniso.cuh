// K3 math: PTMA modified anisotropic distribution (df_mode 5).
// Reference: src/cpp/AnisoVariables.cpp (compute_F :15-132, compute_J :134-300, line_backtrack :302-391,
// find_anisotropic_variables :393-539, compute_famod_coefficient :541-643) and the per-cell prologue of
// calculate_dN_pTdpTdphidy_famod (src/cpp/MomentumSpectra.cpp:1159-1481).
//
// Per cell three nonlinear equations I_200 = E, I_201 = P_T, I_220 = P_L are solved for (Lambda, alpha_T, alpha_L) by
// Newton iteration with Numerical-Recipes line backtracking.  Every function / Jacobian evaluation is a sum over
// (hadron, quadrature node) pairs -- min(320, N_pdg) x 16 terms -- which is what gets parallelised: the sums go
// through a Reducer policy (serial on the host, lane-strided + warp shuffle on the device) while the scalar Newton
// logic runs redundantly on every lane.
#pragma once

#include "cellmath.cuh"
#include "spectra_feqmod.cuh"

namespace is3d {

#define IS3D_GL16_CONST static const
#include "aniso_gl16.inc"
#undef IS3D_GL16_CONST

// layout of the 6 x 16 table handed to the device: roots / weights for alpha = 1, 2, 3
inline void fill_gl16_table(double out[96])
{
  for (int i = 0; i < 16; i++) {
    out[i] = kGL16RootA1[i]; out[16 + i] = kGL16WeightA1[i];
    out[32 + i] = kGL16RootA2[i]; out[48 + i] = kGL16WeightA2[i];
    out[64 + i] = kGL16RootA3[i]; out[80 + i] = kGL16WeightA3[i];
  }
}

// AnisoVariables.h:5-12
constexpr int kAnisoNmax = 30;
constexpr int kAnisoBacktracks = 20;
constexpr double kAnisoTolDX = 1.e-4, kAnisoTolF = 1.e-4, kAnisoDelta = 0.01;
constexpr int kAnisoPts = 16;
constexpr int kAnisoMaxHadrons = 320;      // MomentumSpectra.cpp:1295

struct AnisoHadrons {
  const double *mass, *sign, *deg;
  int n;                                   // min(320, N_pdg)
  const double *gl16;                      // fill_gl16_table layout
  const double *exptab;                    // 2^(m/1024) table of fast_exp (shared memory on the device)
  int exact;                               // 1: the reference's own formulas with libm atan / exp / sqrt and true divisions
                                           // (parity mode, see aniso_t_functions); 0: FP64-pipe approximations
};

// sqrt(a) and 1/sqrt(a) for a > 0 from one hardware seed y and ONE third-order step (see fast_sqrt): with g = a y and
// e = 1 - g y, (1 - e)^(-1/2) = 1 + p + O(5/16 e^3), p = e (1/2 + 3/8 e); sqrt(a) = g (1 + p), 1/sqrt(a) = y (1 + p).
// 6 FP64 instructions (the coupled Goldschmidt step + residual fix + second half-step it replaces took 10).
IS3D_HD void fast_sqrt_rsqrt(double a, double *root, double *iroot)
{
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double g = a * y;
  const double e = fma(-g, y, 1.0);
  double p = fma(e, 0.375, 0.5);
  p = p * e;
  *root = fma(g, p, g);
  *iroot = fma(y, p, y);
#else
  *root = sqrt(a); *iroot = 1.0 / *root;
#endif
}

// atan(s), s >= 0: three-range reduction to |x| <= 0.66 (selects, no divergent branches) and the degree-4 / degree-5
// rational x + x z P(z)/Q(z) (Cephes atan.c coefficients, public domain; tools/gen_atan.py: max relative error 2e-16)
IS3D_HD double fast_atan(double s)
{
#if defined(__CUDA_ARCH__)
  const bool big = s > 2.414213562373095, mid = s > 0.66;
  const double num = big ? -1.0 : (mid ? s - 1.0 : s), den = big ? s : (mid ? s + 1.0 : 1.0);
  const double y0 = big ? 1.5707963267948966 : (mid ? 0.7853981633974483 : 0.0);
  const double more = big ? 6.123233995736765886130e-17 : (mid ? 3.061616997868382943065e-17 : 0.0);
  const double x = num * fast_rcp(den), z = x * x;
  double pn = fma(-8.750608600031904122785e-1, z, -1.615753718733365076637e1);
  pn = fma(pn, z, -7.500855792314704667340e1); pn = fma(pn, z, -1.228866684490136173410e2); pn = fma(pn, z, -6.485021904942025371773e1);
  double qn = z + 2.485846490142306297962e1;
  qn = fma(qn, z, 1.650270098316988542046e2); qn = fma(qn, z, 4.328810604912902668951e2); qn = fma(qn, z, 4.853903996359136964868e2);
  qn = fma(qn, z, 1.945506571482613964425e2);
  const double r = fma(x * z, pn * fast_rcp(qn), x) + more;
  return y0 + r;
#else
  return atan(s);
#endif
}

// ln(v) for normal v > 0 in the FP64 FMA pipe: v = 2^k m with m in [0.75, 1.5), ln m = 2 atanh(f), f = (m - 1)/(m + 1),
// |f| <= 0.2, odd series to f^23 (truncation f^24/25 < 7e-19 relative to 2f); one reciprocal seed, no libm.
IS3D_HD double fast_log(double v)
{
#if defined(__CUDA_ARCH__)
  int hi = __double2hiint(v);
  const int lo = __double2loint(v);
  int k = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;                   // m in [1, 2)
  if (hi >= 0x3ff80000) { hi -= 0x00100000; k++; }       // m >= 1.5: halve it
  const double m = __hiloint2double(hi, lo);
  const double f = (m - 1.0) * fast_rcp(m + 1.0), f2 = f * f;
  double p = fma(f2, 1.0 / 23.0, 1.0 / 21.0);
  p = fma(p, f2, 1.0 / 19.0); p = fma(p, f2, 1.0 / 17.0); p = fma(p, f2, 1.0 / 15.0); p = fma(p, f2, 1.0 / 13.0);
  p = fma(p, f2, 1.0 / 11.0); p = fma(p, f2, 1.0 / 9.0); p = fma(p, f2, 1.0 / 7.0); p = fma(p, f2, 1.0 / 5.0);
  p = fma(p, f2, 1.0 / 3.0);
  const double lnm = fma(2.0 * f, f2 * p, 2.0 * f);
  // k ln2 in two pieces (hi part exact for |k| < 2^11): ln v = k ln2_hi + (k ln2_lo + ln m)
  const double kd = (double)k;
  return fma(kd, 6.93147180369123816490e-01, fma(kd, 1.90821492927058770002e-10, lnm));
#else
  return log(v);
#endif
}

// atanh(s) / s for 0 < s < 1 given s and 1/s: atanh(s) = ln((1 + s)/(1 - s)) / 2 through log1p(u), u = 2 s / (1 - s), with the
// rounding of 1 + u compensated (ln(1 + u) = ln(v) + (u - (v - 1)) / v, v = fl(1 + u)) so that small s keeps its relative
// accuracy: the closed forms of the angular functions divide by z = -s^2 once or twice (AnisoVariables.cpp:72-89).
IS3D_HD double fast_atanh_over_s(double s, double is)
{
#if defined(__CUDA_ARCH__)
  const double u = 2.0 * s * fast_rcp(1.0 - s);
  const double v = 1.0 + u;
  const double c = u - (v - 1.0);
  const double l = fma(c, fast_rcp(v), fast_log(v));
  return 0.5 * l * is;
#else
  (void)is;
  return atanh(s) / s;
#endif
}

// hypergeometric-type angular functions of z = (aT^2 - aL^2) / w^2 (closed forms for |z| > delta, series inside)
struct AnisoT { double t200, t220, t201, t402, t421, t440; };

// The closed forms cancel catastrophically towards z -> delta (t402, t421, t440 lose a factor ~3/z^2 = 3e4 at z = 0.01),
// so the reference's result carries the rounding of ITS atan to the 1e-11 level; bit-level agreement with it needs the same
// libm call and the same division order (exact = true, used by the chain-faithful parity mode).
IS3D_HD AnisoT aniso_t_functions(double z, bool need_j, bool exact)
{
  AnisoT r;
  r.t200 = r.t220 = r.t201 = r.t402 = r.t421 = r.t440 = 0.0;     // reference leaves them unset when z <= -1
  if (z > kAnisoDelta || (z < -kAnisoDelta && z > -1.)) {
    double t, iz;
    if (z > 0.0 && !exact) { double s, is; fast_sqrt_rsqrt(z, &s, &is); t = fast_atan(s) * is; iz = is * is; }
    else if (z > 0.0) { double s = sqrt(z); t = atan(s) / s; iz = 1.0 / z; }
    else if (!exact) { double s, is; fast_sqrt_rsqrt(-z, &s, &is); t = fast_atanh_over_s(s, is); iz = -(is * is); }
    else { double s = sqrt(-z); t = atanh(s) / s; iz = 1.0 / z; }
    if (exact) {             // AnisoVariables.cpp:72-89, :206-231 verbatim (divisions, not reciprocal multiplies)
      r.t200 = 1. + (1. + z) * t;
      r.t220 = (-1. + (1. + z) * t) / z;
      r.t201 = (1. + (z - 1.) * t) / z;
      if (need_j) {
        double z2 = z * z;
        r.t402 = (3. * (z - 1.) + (z * (3. * z - 2.) + 3.) * t) / (4. * z2);
        r.t421 = (3. + z + (1. + z) * (z - 3.) * t) / (4. * z2);
        r.t440 = (-(3. + 5. * z) + 3. * (z + 1.) * (z + 1.) * t) / (4. * z2);
      }
      return r;
    }
    r.t200 = 1. + (1. + z) * t;
    r.t220 = (-1. + (1. + z) * t) * iz;
    r.t201 = (1. + (z - 1.) * t) * iz;
    if (need_j) {
      double iz2 = 0.25 * iz * iz;
      r.t402 = (3. * (z - 1.) + (z * (3. * z - 2.) + 3.) * t) * iz2;
      r.t421 = (3. + z + (1. + z) * (z - 3.) * t) * iz2;
      r.t440 = (-(3. + 5. * z) + 3. * (z + 1.) * (z + 1.) * t) * iz2;
    }
  } else if (fabs(z) <= kAnisoDelta) {
    double z2 = z * z, z3 = z2 * z, z4 = z3 * z, z5 = z4 * z, z6 = z5 * z;
    r.t200 = 2. + (2. / 3.) * z - (2. / 15.) * z2 + (2. / 35.) * z3 - (2. / 63.) * z4 + (2. / 99.) * z5 - (2. / 143.) * z6;
    r.t220 = (2. / 3.) - (2. / 15.) * z + (2. / 35.) * z2 - (2. / 63.) * z3 + (2. / 99.) * z4 - (2. / 143.) * z5 + (2. / 195.) * z6;
    r.t201 = (4. / 3.) - (8. / 15.) * z + (12. / 35.) * z2 - (16. / 63.) * z3 + (20. / 99.) * z4 - (24. / 143.) * z5 + (28. / 195.) * z6;
    if (need_j) {
      r.t402 = (16. / 15.) - (16. / 35.) * z + (32. / 105.) * z2 - (160. / 693.) * z3 + (80. / 429.) * z4 - (112. / 715.) * z5 + (448. / 3315.) * z6;
      r.t421 = (4. / 15.) - (8. / 105.) * z + (4. / 105.) * z2 - (16. / 693.) * z3 + (20. / 1287.) * z4 - (8. / 715.) * z5 + (28. / 3315.) * z6;
      r.t440 = (2. / 5.) - (2. / 35.) * z + (2. / 105.) * z2 - (2. / 231.) * z3 + (2. / 429.) * z4 - (2. / 715.) * z5 + (2. / 1105.) * z6;
    }
  }
  return r;
}

// per-node constants: everything in a term that depends on the quadrature node alone (a lane of the device reducer
// keeps ONE node for the whole sum: 32 lanes = 2 hadrons x 16 nodes per step)
struct AnisoNode { double pbar, pbar2, ipbar2, cF, wJ; };   // cF = pbar w e^pbar (F terms) or w e^pbar (J terms); wJ = w

IS3D_HD AnisoNode aniso_node_F(const AnisoHadrons &h, int i)
{
  const double pbar = h.gl16[32 + i], weight = h.gl16[48 + i];
  return AnisoNode{pbar, pbar * pbar, 1.0 / (pbar * pbar), pbar * weight * exp(pbar), weight};
}
IS3D_HD AnisoNode aniso_node_J(const AnisoHadrons &h, int i)
{
  const double pbar = h.gl16[64 + i], weight = h.gl16[80 + i];
  return AnisoNode{pbar, pbar * pbar, 1.0 / (pbar * pbar), weight * exp(pbar), weight};
}

// one (hadron, node) term of compute_F (:36-107): out = {I_200, I_220, I_201} before the common factors
IS3D_HD void aniso_F_term(const AnisoHadrons &h, int n, const AnisoNode &nd, double lambda, double ilambda, double aT2, double aL2,
                          double out[3])
{
  const double mass = h.mass[n];
  if (mass == 0) return;                   // photons skipped
  if (h.exact) {                           // reference expression order
    const double mbar = mass / lambda, mbar2 = mbar * mbar;
    const double Ebar = sqrt(nd.pbar2 + mbar2);
    const double w = sqrt(aL2 + mbar2 / nd.pbar2);
    const double z = (aT2 - aL2) / (w * w);
    const AnisoT t = aniso_t_functions(z, false, true);
    const double cw = h.deg[n] * (nd.cF / (exp(Ebar) + h.sign[n]));
    out[0] += cw * t.t200 * w;
    out[1] += cw * t.t220 / w;
    out[2] += cw * t.t201 / w;
    return;
  }
  const double mbar = mass * ilambda, mbar2 = mbar * mbar;
  const double Ebar = fast_sqrt(nd.pbar2 + mbar2);
  double w, iw;
  fast_sqrt_rsqrt(fma(mbar2, nd.ipbar2, aL2), &w, &iw);
  const double z = (aT2 - aL2) * (iw * iw);
  const AnisoT t = aniso_t_functions(z, false, false);
  const double cw = h.deg[n] * (nd.cF * fast_rcp(fast_exp(Ebar, h.exptab) + h.sign[n]));
  out[0] += cw * t.t200 * w;
  out[1] += cw * t.t220 * iw;
  out[2] += cw * t.t201 * iw;
}

// one term of compute_J (:175-258) / compute_famod_coefficient (:573-627): out = {J_2001, J_2011, J_2201, J_402m1, J_421m1, J_440m1}
IS3D_HD void aniso_J_term(const AnisoHadrons &h, int n, const AnisoNode &nd, double lambda, double ilambda, double aT2, double aL2,
                          double out[6])
{
  const double mass = h.mass[n];
  if (mass == 0) return;
  if (h.exact) {
    const double mbar = mass / lambda, mbar2 = mbar * mbar;
    const double Ebar = sqrt(nd.pbar2 + mbar2);
    const double w = sqrt(aL2 + mbar2 / nd.pbar2);
    const double z = (aT2 - aL2) / (w * w);
    const AnisoT t = aniso_t_functions(z, true, true);
    const double q = exp(Ebar) + h.sign[n];
    const double cw = h.deg[n] * (nd.wJ * exp(nd.pbar + Ebar) / (q * q));
    out[0] += Ebar * cw * t.t200 * w;
    out[1] += Ebar * cw * t.t201 / w;
    out[2] += Ebar * cw * t.t220 / w;
    out[3] += nd.pbar2 / Ebar * cw * t.t402 / w;
    out[4] += nd.pbar2 / Ebar * cw * t.t421 / w;
    out[5] += nd.pbar2 / Ebar * cw * t.t440 / w;
    return;
  }
  const double mbar = mass * ilambda, mbar2 = mbar * mbar;
  double Ebar, iEbar, w, iw;
  fast_sqrt_rsqrt(nd.pbar2 + mbar2, &Ebar, &iEbar);
  fast_sqrt_rsqrt(fma(mbar2, nd.ipbar2, aL2), &w, &iw);
  const double z = (aT2 - aL2) * (iw * iw);
  const AnisoT t = aniso_t_functions(z, true, false);
  const double e = fast_exp(Ebar, h.exptab), iq = fast_rcp(e + h.sign[n]);
  const double cw = h.deg[n] * (nd.cF * e * iq * iq);           // e^(pbar + Ebar) / (e^Ebar + sign)^2
  const double ecw = Ebar * cw, pcw = nd.pbar2 * iEbar * cw * iw;
  out[0] += ecw * t.t200 * w;
  out[1] += ecw * t.t201 * iw;
  out[2] += ecw * t.t220 * iw;
  out[3] += pcw * t.t402;
  out[4] += pcw * t.t421;
  out[5] += pcw * t.t440;
}

// serial reducer (host / single thread): hadron-major, node-minor like the reference's loops
struct SerialReducer {
  IS3D_HD void sum_F(const AnisoHadrons &h, double lambda, double aT2, double aL2, double out[3]) const
  {
    for (int k = 0; k < 3; k++) out[k] = 0.0;
    const double il = 1.0 / lambda;
    for (int n = 0; n < h.n; n++)
      for (int i = 0; i < kAnisoPts; i++) aniso_F_term(h, n, aniso_node_F(h, i), lambda, il, aT2, aL2, out);
  }
  IS3D_HD void sum_J(const AnisoHadrons &h, double lambda, double aT2, double aL2, double out[6]) const
  {
    for (int k = 0; k < 6; k++) out[k] = 0.0;
    const double il = 1.0 / lambda;
    for (int n = 0; n < h.n; n++)
      for (int i = 0; i < kAnisoPts; i++) aniso_J_term(h, n, aniso_node_J(h, i), lambda, il, aT2, aL2, out);
  }
};

#if defined(__CUDACC__)
// one warp: lane = (hadron parity, node); partial sums over every second hadron, butterfly reduction -> every lane
// holds the same total.  Not inlined: the Newton / line-search driver calls these from five places, and five inlined
// copies of the term loops overflow the instruction cache (ncu: "no instruction" was the top stall).  The hadron table
// travels as a by-value struct of pointers and the sums come back by value, so the call keeps everything in registers
// (a by-reference struct / output array lives in local memory and cost 5 500 local loads per cell).
#ifndef IS3D_K3_UNROLL
#define IS3D_K3_UNROLL 1
#endif
constexpr int kAnisoUnroll = IS3D_K3_UNROLL;    // hadrons per loop trip of a lane (independent dependency chains)
struct AnisoSum3 { double v[3]; };
struct AnisoSum6 { double v[6]; };

static __device__ __noinline__ AnisoSum3 aniso_warp_sum_F(AnisoHadrons h, double lambda, double aT2, double aL2)
{
  const int lane = threadIdx.x & 31;
  const AnisoNode nd = aniso_node_F(h, lane & 15);
  const double il = 1.0 / lambda;
  double o[3] = {0.0, 0.0, 0.0};
#pragma unroll kAnisoUnroll
  for (int n = lane >> 4; n < h.n; n += 2) aniso_F_term(h, n, nd, lambda, il, aT2, aL2, o);
  AnisoSum3 r;
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) o[k] += __shfl_xor_sync(0xffffffffu, o[k], off);
    r.v[k] = o[k];
  }
  return r;
}
static __device__ __noinline__ AnisoSum6 aniso_warp_sum_J(AnisoHadrons h, double lambda, double aT2, double aL2)
{
  const int lane = threadIdx.x & 31;
  const AnisoNode nd = aniso_node_J(h, lane & 15);
  const double il = 1.0 / lambda;
  double o[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll kAnisoUnroll
  for (int n = lane >> 4; n < h.n; n += 2) aniso_J_term(h, n, nd, lambda, il, aT2, aL2, o);
  AnisoSum6 r;
#pragma unroll
  for (int k = 0; k < 6; k++) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) o[k] += __shfl_xor_sync(0xffffffffu, o[k], off);
    r.v[k] = o[k];
  }
  return r;
}
struct WarpReducer {
  __device__ void sum_F(const AnisoHadrons &h, double lambda, double aT2, double aL2, double out[3]) const
  {
    const AnisoSum3 r = aniso_warp_sum_F(h, lambda, aT2, aL2);
    out[0] = r.v[0]; out[1] = r.v[1]; out[2] = r.v[2];
  }
  __device__ void sum_J(const AnisoHadrons &h, double lambda, double aT2, double aL2, double out[6]) const
  {
    const AnisoSum6 r = aniso_warp_sum_J(h, lambda, aT2, aL2);
#pragma unroll
    for (int k = 0; k < 6; k++) out[k] = r.v[k];
  }
};
#endif

template <class Reducer>
IS3D_HD void aniso_compute_F(const Reducer &red, const AnisoHadrons &h, double Ea, double PTa, double PLa, const double X[3], double F[3])
{
  const double lambda = X[0], aT2 = X[1] * X[1], aL2 = X[2] * X[2], aL = X[2];
  const double common_factor = aT2 * aL * lambda * lambda * lambda * lambda / kFourPi2HbarC3;
  double I[3];
  red.sum_F(h, lambda, aT2, aL2, I);
  F[0] = I[0] * common_factor - Ea;                       // I_200 - E
  F[1] = I[2] * (common_factor * aT2 / 2.) - PTa;          // I_201 - PT
  F[2] = I[1] * (common_factor * aL2) - PLa;               // I_220 - PL
}

template <class Reducer>
IS3D_HD void aniso_compute_J(const Reducer &red, const AnisoHadrons &h, double Ea, double PTa, double PLa, const double X[3],
                             const double F[3], double J[9])
{
  const double lambda = X[0], aT = X[1], aL = X[2], aT2 = aT * aT, aL2 = aL * aL;
  const double lambda2 = lambda * lambda, lambda3 = lambda2 * lambda;
  const double lambda_aT3 = lambda * aT2 * aT, lambda_aL3 = lambda * aL2 * aL;
  const double common_factor = aT2 * aL * lambda2 * lambda3 / kFourPi2HbarC3;
  double S[6];
  red.sum_J(h, lambda, aT2, aL2, S);
  const double J_2001 = S[0] * common_factor, J_2011 = S[1] * (common_factor * aT2 / 2.), J_2201 = S[2] * (common_factor * aL2);
  const double J_402m1 = S[3] * (common_factor * aT2 * aT2 / 8.), J_421m1 = S[4] * (common_factor * aT2 * aL2 / 2.);
  const double J_440m1 = S[5] * (common_factor * aL2 * aL2);
  const double Eai = F[0] + Ea, PTai = F[1] + PTa, PLai = F[2] + PLa;
  J[0] = J_2001 / lambda2; J[1] = 2. * (Eai + PTai) / aT;   J[2] = (Eai + PLai) / aL;
  J[3] = J_2011 / lambda2; J[4] = 4. * J_402m1 / lambda_aT3; J[5] = J_421m1 / lambda_aL3;
  J[6] = J_2201 / lambda2; J[7] = 2. * J_421m1 / lambda_aT3; J[8] = J_440m1 / lambda_aL3;
}

// 3x3 solve J dX = rhs by LU with partial pivoting (gsl_linalg_LU_decomp + LU_solve, AnisoVariables.cpp:457-465)
IS3D_HD void solve3x3_lu(const double Jin[9], const double rhs[3], double x[3])
{
  double A[3][3], b[3] = {rhs[0], rhs[1], rhs[2]};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) A[i][j] = Jin[3 * i + j];
  for (int j = 0; j < 2; j++) {
    int ip = j;
    double mx = fabs(A[j][j]);
    for (int i = j + 1; i < 3; i++) if (fabs(A[i][j]) > mx) { mx = fabs(A[i][j]); ip = i; }
    if (ip != j) {
      for (int k = 0; k < 3; k++) { double t = A[j][k]; A[j][k] = A[ip][k]; A[ip][k] = t; }
      double t = b[j]; b[j] = b[ip]; b[ip] = t;
    }
    if (A[j][j] != 0.0)
      for (int i = j + 1; i < 3; i++) {
        double a = A[i][j] / A[j][j];
        A[i][j] = a;
        for (int k = j + 1; k < 3; k++) A[i][k] -= a * A[j][k];
      }
  }
  // forward substitution with the unit-lower factor (multipliers were applied to A only), then back substitution
  b[1] -= A[1][0] * b[0];
  b[2] -= A[2][0] * b[0] + A[2][1] * b[1];
  x[2] = b[2] / A[2][2];
  x[1] = (b[1] - A[1][2] * x[2]) / A[1][1];
  x[0] = (b[0] - A[0][1] * x[1] - A[0][2] * x[2]) / A[0][0];
}

// line_backtrack (:302-391); on return F = F(Xcurrent + l dX)
template <class Reducer>
IS3D_HD double aniso_line_backtrack(const Reducer &red, const AnisoHadrons &h, double Ea, double PTa, double PLa, const double Xc[3],
                                    const double dX[3], double dX_abs, double g0, double F[3])
{
  double X[3] = {Xc[0] + dX[0], Xc[1] + dX[1], Xc[2] + dX[2]};
  aniso_compute_F(red, h, Ea, PTa, PLa, X, F);
  double f = (F[0] * F[0] + F[1] * F[1] + F[2] * F[2]) / 2.;
  const double gprime0 = -2. * g0, alpha = 0.0001;
  double l = 1, lroot = 0, lprev = 0, fprev = 0;
  for (int n = 0; n < kAnisoBacktracks; n++) {
    if ((l * dX_abs) <= kAnisoTolDX) return l;
    else if (f <= (g0 + l * alpha * gprime0)) return l;
    else if (n == 0) lroot = -gprime0 / (2. * (f - g0 - gprime0));
    else {
      double a = ((f - g0 - l * gprime0) / (l * l) - (fprev - g0 - lprev * gprime0) / (lprev * lprev)) / (l - lprev);
      double b = (-lprev * (f - g0 - l * gprime0) / (l * l) + l * (fprev - g0 - lprev * gprime0) / (lprev * lprev)) / (l - lprev);
      if (a == 0) lroot = -gprime0 / (2. * b);
      else {
        double z = b * b - 3. * a * gprime0;
        if (z < 0) lroot = 0.5 * l;
        else if (b <= 0) lroot = (-b + sqrt(z)) / (3. * a);
        else lroot = -gprime0 / (b + sqrt(z));
      }
      lroot = fmin(lroot, 0.5 * l);
    }
    lprev = l; fprev = f;
    l = fmax(lroot, 0.5 * l);
    for (int i = 0; i < 3; i++) X[i] = Xc[i] + l * dX[i];
    aniso_compute_F(red, h, Ea, PTa, PLa, X, F);
    f = (F[0] * F[0] + F[1] * F[1] + F[2] * F[2]) / 2.;
  }
  return l;
}

struct AnisoSolution {
  double lambda, aT, aL;
  bool failed;
  int iterations;
};

// find_anisotropic_variables (:393-539)
template <class Reducer>
IS3D_HD AnisoSolution aniso_find_variables(const Reducer &red, const AnisoHadrons &h, double E, double pl, double pt, double lambda_0,
                                           double aT_0, double aL_0)
{
  AnisoSolution fail{lambda_0, aT_0, aL_0, true, 0};
  const double Ea = E, PTa = pt, PLa = pl;
  if (Ea < 0 || PTa < 0 || PLa < 0) return fail;
  double X[3] = {lambda_0, aT_0, aL_0}, dX[3], F[3], J[9];
  aniso_compute_F(red, h, Ea, PTa, PLa, X, F);
  const double stepmax = 100. * fmax(sqrt(X[0] * X[0] + X[1] * X[1] + X[2] * X[2]), 3.);
  for (int n = 0; n < kAnisoNmax; n++) {
    aniso_compute_J(red, h, Ea, PTa, PLa, X, F, J);
    double f = (F[0] * F[0] + F[1] * F[1] + F[2] * F[2]) / 2.;
    double mF[3] = {-F[0], -F[1], -F[2]};
    solve3x3_lu(J, mF, dX);
    double dX_abs = sqrt(dX[0] * dX[0] + dX[1] * dX[1] + dX[2] * dX[2]);
    if (dX_abs > stepmax) {
      for (int i = 0; i < 3; i++) dX[i] *= stepmax / dX_abs;
      dX_abs = stepmax;
    }
    double l = aniso_line_backtrack(red, h, Ea, PTa, PLa, X, dX, dX_abs, f, F);
    for (int i = 0; i < 3; i++) X[i] += (l * dX[i]);
    double F_abs = sqrt(F[0] * F[0] + F[1] * F[1] + F[2] * F[2]);
    dX_abs *= l;
    if (X[0] < 0 || X[1] < 0 || X[2] < 0) { fail.iterations = n + 1; return fail; }
    else if (dX_abs <= kAnisoTolDX && F_abs <= kAnisoTolF) return AnisoSolution{X[0], X[1], X[2], false, n + 1};
  }
  fail.iterations = kAnisoNmax;
  return fail;
}

// compute_famod_coefficient (:541-643)
template <class Reducer>
IS3D_HD void aniso_famod_coefficient(const Reducer &red, const AnisoHadrons &h, double lambda, double aT, double aL,
                                     double *betapiperp, double *betaWperp)
{
  const double lambda2 = lambda * lambda, aT2 = aT * aT, aL2 = aL * aL;
  const double common_factor = aT2 * aL * lambda * lambda2 * lambda2 / kFourPi2HbarC3;
  double S[6];
  red.sum_J(h, lambda, aT2, aL2, S);
  const double J_402m1 = S[3] * (common_factor * aT2 * aT2 / 8.), J_421m1 = S[4] * (common_factor * aT2 * aL2 / 2.);
  *betapiperp = J_402m1 / (aT2 * lambda);
  *betaWperp = J_421m1 / (aT * aL * lambda);
}

struct FamodFlags {
  int dimension;
  int include_baryon, include_shear, include_baryondiff;
  double deta_min;
  int sampler = 0;             // 1: failure rules of sample_dN_pTdpTdphidy_famod (ParticleSampler.cpp:1335-1383): a failed first
                               // attempt without a previous success IS a breakdown (the spectra path adopts (T,1,1) instead,
                               // MomentumSpectra.cpp:1353-1364) and there is no renormalisation factor to go non-finite
};

// where the famod stage leaves its solution in the feqmod pack layout (slots the PTMA spectra kernel does not read)
enum { FP_FAMOD_LAMBDA = FP_T, FP_FAMOD_AT = FP_TMOD, FP_FAMOD_AL = FP_DNFACT, FP_FAMOD_BETAPIPERP = FP_G, FP_FAMOD_BETAWPERP = FP_F_T2 };

enum { CELL_RECONSTRUCTION_FAIL = 16 };

// initial-guess chain carried from cell to cell (MomentumSpectra.cpp:1132-1135, :1308-1364)
struct FamodChain {
  double lambda_prev, aT_prev, aL_prev;
  bool previous_success;
};

// Per-cell prologue of calculate_dN_pTdpTdphidy_famod (:1159-1481) into the feqmod pack layout, so that the K2
// spectra kernel evaluates the momentum loop: modified branch f = |renorm| / (exp(E'/Lambda - b upsilonB) + sign)
// with p' = B^-1 p_LRF, fallback f = feq (no df).  `chain` may be NULL (chain-free policy: always start at (T,1,1)).
template <class Reducer>
IS3D_HD int famod_setup_cell(const Reducer &red, const Cell &c, const FamodFlags &fl, const AnisoHadrons &h, FamodChain *chain,
                             double pack[FP_SIZE], int *iterations)
{
  for (int k = 0; k < FP_SIZE; k++) pack[k] = 0.0;
  *iterations = 0;
  double tau = c.tau, tau2 = tau * tau;
  double ux = c.ux, uy = c.uy, un = c.un;
  double ut = sqrt(1. + ux * ux + uy * uy + tau2 * un * un);
  if (ut * c.dat + ux * c.dax + uy * c.day + un * c.dan <= 0) return CELL_SKIPPED;
  int status = CELL_OK;
  double T = c.T, P = c.P, E = c.E;
  // shear and bulk are always read here (:1192-1204)
  Shear pi = complete_shear(c.pixx, c.pixy, c.pixn, c.piyy, c.piyn, ut, ux, uy, un, tau2);
  double bulkPi = c.bulkPi;
  double muB = fl.include_baryon ? c.muB : 0.0;
  double alphaB = muB / T;
  Basis b = milne_basis(ut, ux, uy, un, tau);
  ShearLRF l = boost_shear_to_lrf(pi, b, tau2);
  double pl = P + bulkPi + l.zz, pt = P + bulkPi - l.zz / 2.;
  double piTxx = 0, piTxy = 0, piTyy = 0, WTzx = 0, WTzy = 0;
  if (fl.include_shear) { piTxx = (l.xx - l.yy) / 2.; piTxy = l.xy; piTyy = -piTxx; WTzx = l.xz; WTzy = l.yz; }

  double lambda = T, aT = 1, aL = 1;
  bool breaks = false;
  if (pl < 0 || pt < 0) { status |= CELL_PL_NEGATIVE; breaks = true; }
  else {
    const bool prev = chain && chain->previous_success;
    if (prev) { lambda = chain->lambda_prev; aT = chain->aT_prev; aL = chain->aL_prev; }
    AnisoSolution X = aniso_find_variables(red, h, E, pl, pt, lambda, aT, aL);
    if (X.failed && prev) {
      lambda = T; aT = 1; aL = 1;
      X = aniso_find_variables(red, h, E, pl, pt, lambda, aT, aL);
      if (X.failed) {
        breaks = true;
        status |= CELL_RECONSTRUCTION_FAIL;
        if (chain) chain->previous_success = false;
      } else {
        lambda = X.lambda; aT = X.aT; aL = X.aL;
        if (chain) { chain->lambda_prev = lambda; chain->aT_prev = aT; chain->aL_prev = aL; chain->previous_success = true; }
      }
    } else if (X.failed && fl.sampler) {
      // sampler rule (ParticleSampler.cpp:1369-1374): breakdown, the variables stay at the initial guess (T, 1, 1)
      breaks = true;
      status |= CELL_RECONSTRUCTION_FAIL;
      if (chain) chain->previous_success = false;
    } else {
      // also taken when the FIRST attempt fails without a previous success: the reference then adopts the returned
      // (lambda_0, aT_0, aL_0) = (T, 1, 1) as if it were a solution (:1353-1364)
      lambda = X.lambda; aT = X.aT; aL = X.aL;
      if (chain) { chain->lambda_prev = lambda; chain->aT_prev = aT; chain->aL_prev = aL; chain->previous_success = true; }
    }
    *iterations = X.iterations;
  }
  double betapiperp, betaWperp;
  aniso_famod_coefficient(red, h, lambda, aT, aL, &betapiperp, &betaWperp);
  double shear_coeff = 0.5 / betapiperp, diff_coeff = 1. / betaWperp;
  double detA = aT * aT * aL;
  double Cxx = 1. + shear_coeff * piTxx, Cxy = shear_coeff * piTxy, Cxz = diff_coeff * WTzx * aT / (aT + aL);
  double Cyx = Cxy, Cyy = 1. + shear_coeff * piTyy, Cyz = diff_coeff * WTzy * aT / (aT + aL);
  double Czx = diff_coeff * WTzx * aL / (aT + aL), Czy = diff_coeff * WTzy * aL / (aT + aL), Czz = 1.;
  double detC = Cxx * (Cyy * Czz - Cyz * Czy) - Cxy * (Cyx * Czz - Cyz * Czx) + Cxz * (Cyx * Czy - Cyy * Czx);
  double B[9];
  B[0] = aT + aT * shear_coeff * piTxx; B[1] = aT * shear_coeff * piTxy; B[2] = diff_coeff * WTzx * aT * aL / (aT + aL);
  B[3] = B[1]; B[4] = aT + aT * shear_coeff * piTyy; B[5] = diff_coeff * WTzy * aT * aL / (aT + aL);
  B[6] = B[2]; B[7] = B[5]; B[8] = aL;
  double detB = detC * detA;
  double detB_bulk_two_thirds = (2. * aT + aL) * (2. * aT + aL) / 9.;
  double Binv[9], det_unused;
  invert3x3(B, Binv, &det_unused);
  if (detB <= fl.deta_min) breaks = true;
  double eta_scale = 1;
  if (detB > fl.deta_min && fl.dimension == 2) eta_scale = detB / detB_bulk_two_thirds;
  double renorm = eta_scale / detC;
  if (not_finite(renorm)) { if (!fl.sampler) breaks = true; renorm = 0.0; }
  if (breaks) status |= CELL_BREAKDOWN;

  // fallback: plain equilibrium distribution (no df), eta weight on the whole p.dsigma (:1540-1553, :1617)
  double invT = 1.0 / T;
  pack[DP_VALID] = 1.0;
  pack[DP_ETA] = c.eta;
  pack[DP_UTT] = ut * invT; pack[DP_TUNT] = tau * un * invT; pack[DP_UXT] = ux * invT; pack[DP_UYT] = uy * invT;
  pack[DP_ALPHAB] = alphaB;
  pack[DP_DAT] = c.dat; pack[DP_DAX] = c.dax; pack[DP_DAY] = c.day; pack[DP_DANT] = c.dan / tau;
  // modified branch
  pack[FP_BREAKDOWN] = breaks ? 1.0 : 0.0;
  pack[FP_DETA] = detB;
  pack[FP_ETA_SCALE] = eta_scale;
  pack[FP_RENORM] = fabs(renorm);
  double iL = 1.0 / lambda;
  pack[FP_IT2] = iL * iL;
  pack[FP_ALPHAB_MOD] = alphaB;            // upsilonB = alphaB (:1292)
  const double va[3] = {-b.Xt, 0.0, -b.Zt}, vb[3] = {tau * b.Xn, 0.0, tau * b.Zn};
  const double vc[3] = {b.Xx, b.Yx, 0.0}, vd[3] = {b.Xy, b.Yy, 0.0};
  for (int i = 0; i < 3; i++) {
    pack[FP_A1X + i] = iL * (Binv[3 * i] * va[0] + Binv[3 * i + 1] * va[1] + Binv[3 * i + 2] * va[2]);
    pack[FP_A2X + i] = iL * (Binv[3 * i] * vb[0] + Binv[3 * i + 1] * vb[1] + Binv[3 * i + 2] * vb[2]);
    pack[FP_A3X + i] = iL * (Binv[3 * i] * vc[0] + Binv[3 * i + 1] * vc[1] + Binv[3 * i + 2] * vc[2]);
    pack[FP_A4X + i] = iL * (Binv[3 * i] * vd[0] + Binv[3 * i + 1] * vd[1] + Binv[3 * i + 2] * vd[2]);
  }
  pack[FP_FAMOD_LAMBDA] = lambda; pack[FP_FAMOD_AT] = aT; pack[FP_FAMOD_AL] = aL;     // the solution and its coefficients:
  pack[FP_FAMOD_BETAPIPERP] = betapiperp; pack[FP_FAMOD_BETAWPERP] = betaWperp;        // read by the PTMA sampler stage
  return status;
}

}  // namespace is3d
