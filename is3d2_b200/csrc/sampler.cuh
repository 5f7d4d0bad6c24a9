// K5 / K6 math: per-cell mean yields and Monte-Carlo momentum sampling (operation 2), df_mode 1-5.
// Reference: src/cpp/ParticleSampler.cpp -- estimate_mean_particle_number :75-119, fast_max_particle_number :122-161,
// max_particle_number :164-239, sample_momentum :243-405, rescale_momentum :407-426, calculate_total_yield :447-636,
// sample_dN_pTdpTdphidy :638-1134.
//
// Random numbers: counter-based Philox4x32-10.  key = sampler seed; counter = (global cell index, hadron index within
// the cell, draw block).  A sampled hadron therefore depends only on (seed, cell, hadron index), never on the launch
// geometry or on how cells are sharded over GPUs.  The reference's four std::default_random_engine (minstd_rand0)
// streams (:650-654) are implementation-defined; parity with it is distributional (hard part 5).
#pragma once

#include "cellmath.cuh"
#include "dftables.cuh"
#include "gauss_thermal.cuh"
#include "spectra_feqmod.cuh"

namespace is3d {

// ---------------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11)
// ---------------------------------------------------------------------------------------------------------------
struct Philox {
  uint32_t ctr[4];
  uint32_t key[2];
  uint32_t out[4];
  int have;               // unread 32-bit words left in out[]

  IS3D_HD static void mulhilo(uint32_t a, uint32_t b, uint32_t *hi, uint32_t *lo)
  {
    uint64_t p = (uint64_t)a * (uint64_t)b;
    *hi = (uint32_t)(p >> 32);
    *lo = (uint32_t)p;
  }
  // stream = (seed; global cell, event block, hadron index within (cell, block)); cells < 2^44, event blocks < 2^20
  IS3D_HD void init(uint64_t seed, uint64_t cell, uint32_t hadron, uint32_t block = 0)
  {
    key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
    ctr[0] = (uint32_t)cell; ctr[1] = (uint32_t)(cell >> 32) | (block << 12); ctr[2] = hadron; ctr[3] = 0;
    have = 0;
  }
  // not inlined on the device: ~25 call sites of canonical() would each carry the ten rounds, and the sampler kernel is
  // instruction-cache bound (ncu: 15 "no instruction" stall cycles per issued instruction before this)
#if defined(__CUDACC__)
  __host__ __device__ __noinline__ void generate()
#else
  void generate()
#endif
  {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
#pragma unroll
    for (int r = 0; r < 10; r++) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(0xD2511F53u, c0, &hi0, &lo0);
      mulhilo(0xCD9E8D57u, c2, &hi1, &lo1);
      uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    ctr[3]++;               // next draw block
    have = 4;
  }
  // uniform double in [0, 1) with 53 random bits: generate_canonical<double, 53> of the reference (:25-29)
  IS3D_HD double canonical()
  {
    if (have < 2) generate();
    uint32_t a = out[4 - have], b = out[5 - have];
    have -= 2;
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
  }
};

// Poisson variate of mean lam: product-of-uniforms for lam < 10, Hoermann's transformed rejection (PTRS) above.
IS3D_HD long poisson_sample(Philox &rng, double lam)
{
  if (!(lam > 0.0)) return 0;
  if (lam < 10.0) {
    double L = exp(-lam), p = 1.0;
    long k = 0;
    do { k++; p *= rng.canonical(); } while (p > L);
    return k - 1;
  }
  double slam = sqrt(lam), loglam = log(lam);
  double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
  double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
  for (int iter = 0; iter < 100000; iter++) {      // bounded: a GPU thread must never spin forever
    double U = rng.canonical() - 0.5, V = rng.canonical();
    double us = 0.5 - fabs(U);
    double kf = floor((2.0 * a / us + b) * U + lam + 0.43);
    if (us >= 0.07 && V <= vr) return (long)kf;
    if (kf < 0.0 || (us < 0.013 && V > us)) continue;
    if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + kf * loglam - lgamma(kf + 1.0)) return (long)kf;
  }
  return (long)lam;
}

// ---------------------------------------------------------------------------------------------------------------
// per-cell pack of the sampler
// ---------------------------------------------------------------------------------------------------------------
enum SamplerPackIdx {
  SP_VALID = 0, SP_DNTOT,            // mean hadron number of the cell, volume factor 2 y_max ds_max included (:915)
  SP_WA, SP_WB,                      // species weights w_s = WA neq_s + WB dn_bulk_s (fast mode)
  SP_BREAKDOWN,
  SP_TAU, SP_X, SP_Y, SP_ETA, SP_UT, SP_UX, SP_UY, SP_UN,
  SP_XT, SP_XX, SP_XY, SP_XN, SP_YX, SP_YY, SP_ZT, SP_ZN,
  SP_DST, SP_DSX, SP_DSY, SP_DSZ, SP_DSMAX,
  SP_PIXX, SP_PIXY, SP_PIXZ, SP_PIYY, SP_PIYZ, SP_PIZZ, SP_VX, SP_VY, SP_VZ,
  SP_T, SP_TSAMPLE, SP_ALPHAB, SP_ALPHAB_SAMPLE, SP_BULKPI, SP_BER,
  SP_SHEAR_MOD, SP_ISO_SCALE, SP_DIFF_MOD,
  SP_C0, SP_C1, SP_C2, SP_C3, SP_C4, SP_C5, SP_C6,     // df_mode-specific linear-df coefficients (see sampler_setup_cell)
  // yield estimate of calculate_total_yield (:447-636)
  SP_DS_TIME, SP_DS_SPACE_VDSIGMA, SP_Z, SP_DELTA_Z,
  SP_SIZE
};

struct SamplerFlags {
  int df_mode, dimension;
  int include_baryon, include_bulk, include_shear, include_baryondiff;
  int fast;
  double deta_min, mass_pion0, bulkPi_over_P_max, y_cut;
  double T_avg, F_avg, betabulk_avg;     // fast-mode breakdown test (:660-669, :874)
};

// per-cell prologue of sample_dN_pTdpTdphidy (:680-915) and of calculate_total_yield (:452-620).
// sum_neq / sum_bulk: sums over the chosen species of the fast-mode densities (cell independent).
IS3D_HD int sampler_setup_cell(const Cell &c, const DfTables &tb, const SamplerFlags &fl, const double *gla_root,
                               const double *gla_weight, int gla_pts, double sum_neq, double sum_bulk, double pack[SP_SIZE])
{
  for (int k = 0; k < SP_SIZE; k++) pack[k] = 0.0;
  double tau = c.tau, tau2 = tau * tau;
  double ux = c.ux, uy = c.uy, un = c.un;
  double ut = sqrt(1.0 + ux * ux + uy * uy + tau2 * un * un);
  if (ut * c.dat + ux * c.dax + uy * c.day + un * c.dan <= 0.0) return CELL_SKIPPED;
  int status = CELL_OK;
  double T = c.T, P = c.P, E = c.E;
  Shear pi;
  if (fl.include_shear) pi = complete_shear(c.pixx, c.pixy, c.pixn, c.piyy, c.piyn, ut, ux, uy, un, tau2);
  double bulkPi = fl.include_bulk ? c.bulkPi : 0.0;
  double muB = 0.0, alphaB = 0.0, nB = 0.0, Vt = 0.0, Vx = 0.0, Vy = 0.0, Vn = 0.0, Vdsigma = 0.0, ber = 0.0;
  if (fl.include_baryon && fl.include_baryondiff) {
    muB = c.muB; nB = c.nB; Vx = c.Vx; Vy = c.Vy; Vn = c.Vn;
    Vt = (Vx * ux + Vy * uy + tau2 * Vn * un) / ut;
    Vdsigma = Vt * c.dat + Vx * c.dax + Vy * c.day + Vn * c.dan;
    alphaB = muB / T;
    ber = nB / (E + P);
  }
  if (fl.df_mode == 4) {          // inclusive clamp (:552-559, :774-775)
    if (bulkPi <= -P) bulkPi = -(1.0 - 1.e-5) * P;
    else if (bulkPi / P >= fl.bulkPi_over_P_max) bulkPi = P * (fl.bulkPi_over_P_max - 1.e-5);
  }
  DfCoeff df;
  if (!evaluate_df_coefficients(tb, fl.df_mode, fl.include_baryon, T, muB, E, P, bulkPi, &df)) return CELL_OUT_OF_TABLE;

  Basis b = milne_basis(ut, ux, uy, un, tau);
  DsigmaLRF ds = boost_dsigma_to_lrf(c.dat, c.dax, c.day, c.dan, b, ut, ux, uy, un);
  ShearLRF l = boost_shear_to_lrf(pi, b, tau2);
  double Vx_LRF, Vy_LRF, Vz_LRF;
  boost_V_to_lrf(Vt, Vx, Vy, Vn, b, tau2, &Vx_LRF, &Vy_LRF, &Vz_LRF);

  double T_mod = T, alphaB_mod = alphaB, shear_mod = 0.0, bulk_mod = 0.0, diff_mod = 0.0;
  if (fl.df_mode == 3) {
    T_mod = T + bulkPi * df.F / df.betabulk;
    alphaB_mod = alphaB + bulkPi * df.G / df.betabulk;
    shear_mod = 0.5 / df.betapi; bulk_mod = bulkPi / (3.0 * df.betabulk); diff_mod = T / df.betaV;
  } else if (fl.df_mode == 4) {
    shear_mod = 0.5 / df.betapi; bulk_mod = df.lambda; diff_mod = 0.0;
  }
  double Axx = 1.0 + l.xx * shear_mod + bulk_mod, Axy = l.xy * shear_mod, Axz = l.xz * shear_mod;
  double Ayy = 1.0 + l.yy * shear_mod + bulk_mod, Ayz = l.yz * shear_mod, Azz = 1.0 + l.zz * shear_mod + bulk_mod;
  double detA = Axx * (Ayy * Azz - Ayz * Ayz) - Axy * (Axy * Azz - Ayz * Axz) + Axz * (Axy * Ayz - Ayy * Axz);   // compute_detA

  // does_feqmod_breakdown(..., FAST, Tavg, F_avg, betabulk_avg) (:874); total-yield path always uses fast = 0
  FeqmodFlags ff;
  ff.df_mode = fl.df_mode; ff.deta_min = fl.deta_min; ff.mass_pion0 = fl.mass_pion0;
  bool breaks = false, breaks_yield = false;
  if (fl.df_mode == 3 || fl.df_mode == 4) {
    breaks_yield = feqmod_breaks_down(ff, T, df.F, bulkPi, df.betabulk, detA, df.z, gla_root, gla_weight, gla_pts);
    if (fl.df_mode == 3 && fl.fast)
      breaks = feqmod_breaks_down(ff, fl.T_avg, fl.F_avg, bulkPi, fl.betabulk_avg, detA, df.z, gla_root, gla_weight, gla_pts);
    else breaks = breaks_yield;
  }
  if (breaks) status |= CELL_BREAKDOWN;

  // fast_max_particle_number summed over species (:884-894): w_s = WA neq_s + WB dn_bulk_s
  double WA = 2.0, WB = 0.0;
  if (fl.df_mode == 3 && !breaks) { WA = 1.0; WB = bulkPi; }
  if (fl.df_mode == 4 && !breaks) { WA = df.z; WB = 0.0; }
  double dn_tot = WA * sum_neq + WB * sum_bulk;
  double y_max = (fl.dimension == 2) ? fl.y_cut : 0.5;
  pack[SP_VALID] = 1.0;
  pack[SP_DNTOT] = (dn_tot <= 0.0) ? 0.0 : dn_tot * (2.0 * y_max * ds.magnitude);      // :913-915
  pack[SP_WA] = WA; pack[SP_WB] = WB;
  pack[SP_BREAKDOWN] = breaks ? 1.0 : 0.0;
  pack[SP_TAU] = tau; pack[SP_X] = c.x; pack[SP_Y] = c.y; pack[SP_ETA] = c.eta;
  pack[SP_UT] = ut; pack[SP_UX] = ux; pack[SP_UY] = uy; pack[SP_UN] = un;
  pack[SP_XT] = b.Xt; pack[SP_XX] = b.Xx; pack[SP_XY] = b.Xy; pack[SP_XN] = b.Xn;
  pack[SP_YX] = b.Yx; pack[SP_YY] = b.Yy; pack[SP_ZT] = b.Zt; pack[SP_ZN] = b.Zn;
  pack[SP_DST] = ds.t; pack[SP_DSX] = ds.x; pack[SP_DSY] = ds.y; pack[SP_DSZ] = ds.z; pack[SP_DSMAX] = ds.magnitude;
  pack[SP_PIXX] = l.xx; pack[SP_PIXY] = l.xy; pack[SP_PIXZ] = l.xz; pack[SP_PIYY] = l.yy; pack[SP_PIYZ] = l.yz; pack[SP_PIZZ] = l.zz;
  pack[SP_VX] = Vx_LRF; pack[SP_VY] = Vy_LRF; pack[SP_VZ] = Vz_LRF;
  const bool modified = (fl.df_mode == 3 && !breaks);         // PTM samples at (T_mod, alphaB_mod); PTB always at (T, 0)
  pack[SP_T] = T;
  pack[SP_TSAMPLE] = modified ? T_mod : T;
  pack[SP_ALPHAB] = alphaB;
  pack[SP_ALPHAB_SAMPLE] = (fl.df_mode == 4) ? 0.0 : (modified ? alphaB_mod : alphaB);
  pack[SP_BULKPI] = bulkPi; pack[SP_BER] = ber;
  pack[SP_SHEAR_MOD] = shear_mod; pack[SP_ISO_SCALE] = 1.0 + bulk_mod; pack[SP_DIFF_MOD] = diff_mod;
  if (fl.df_mode == 1) {            // :947-968
    pack[SP_C0] = df.shear14_coeff; pack[SP_C1] = df.c0 - df.c2; pack[SP_C2] = df.c1; pack[SP_C3] = 4.0 * df.c2 - df.c0;
    pack[SP_C4] = df.c3; pack[SP_C5] = df.c4;
  } else if (fl.df_mode == 2 || fl.df_mode == 3) {   // :971-994
    pack[SP_C0] = 2.0 * df.betapi * T; pack[SP_C1] = df.G; pack[SP_C2] = df.F / (T * T); pack[SP_C3] = 3.0 * T;
    pack[SP_C4] = bulkPi / df.betabulk; pack[SP_C5] = df.betaV;
  } else {                           // PTB fallback, :1030-1047
    pack[SP_C0] = 2.0 * df.betapi * T; pack[SP_C1] = df.delta_z - 3.0 * df.delta_lambda; pack[SP_C2] = df.delta_lambda / T;
  }
  // calculate_total_yield (:606-619): ds_time, ds_space * V.dsigma, z, delta_z, with its own (fast = 0) breakdown flag
  pack[SP_DS_TIME] = ds.t;
  pack[SP_DS_SPACE_VDSIGMA] = ds.space * Vdsigma;
  pack[SP_Z] = (fl.df_mode == 4) ? (breaks_yield ? (1.0 + df.delta_z) : df.z) : 1.0;
  pack[SP_DELTA_Z] = df.delta_z;
  return status;
}

// mean hadron number of one cell for the event-count estimate: sum over species of estimate_mean_particle_number
// (:75-119) with the species sums factored out
IS3D_HD double cell_mean_yield(const double pack[SP_SIZE], int df_mode, double sum_neq, double sum_bulk, double sum_diff)
{
  if (df_mode == 4) return pack[SP_DS_TIME] * pack[SP_Z] * sum_neq;
  return pack[SP_DS_TIME] * (sum_neq + pack[SP_BULKPI] * sum_bulk) - pack[SP_DS_SPACE_VDSIGMA] * sum_diff;
}

// rational fit of the maximum of the light-boson thermal weight (:41-70)
IS3D_HD double pion_thermal_weight_max(double x)
{
  double x2 = x * x, x3 = x2 * x, x4 = x3 * x;
  double max = (143206.88623164667 - 95956.76008684626 * x - 21341.937407169076 * x2 + 14388.446116867359 * x3 - 6083.775788504437 * x4) /
               (-0.3541350577684533 + 143218.69233952634 * x - 24516.803600065778 * x2 - 115811.59391199696 * x3 + 35814.36403387459 * x4);
  return 1.00001 * max;
}

struct LrfMomentum { double E, px, py, pz, feq; bool ok; };
constexpr int kMaxRejectionIterations = 1000000;   // acceptance is O(0.5); only NaN inputs (e.g. T_mod <= 0) get here

// thermal momentum in the local rest frame by rejection (:243-405); counts proposals in *samples
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
LrfMomentum sample_momentum(Philox &rng, long *samples, double mass, double sign, double T, double chem)
{
  double mbar = mass / T, mbar_squared = mbar * mbar;
  double pbar = 0.0, Ebar = mbar, phi_over_2pi = 0.0, costheta = 1.0, feq = 0.0;
  bool accepted = false;
  if (mbar < 1.008) {
    double weq_max = 1.0;
    if (mbar < 0.8554 && sign == -1.0) weq_max = pion_thermal_weight_max(mbar);
    for (int iter = 0; iter < kMaxRejectionIterations; iter++) {
      (*samples)++;
      double r1 = 1.0 - rng.canonical(), r2 = 1.0 - rng.canonical(), r3 = 1.0 - rng.canonical();
      double l1 = log(r1), l2 = log(r2), l3 = log(r3);
      pbar = -(l1 + l2 + l3);
      Ebar = sqrt(pbar * pbar + mbar_squared);
      feq = 1.0 / (exp(Ebar) + sign);
      double weight = feq / weq_max / (r1 * r2 * r3);
      if (rng.canonical() < weight) {
        phi_over_2pi = (l1 + l2) * (l1 + l2) / (pbar * pbar);
        costheta = (l1 - l2) / (l1 + l2);
        accepted = true;
        break;
      }
    }
  } else {
    // mixture of k^n exp(-k) proposals with weights (mbar^2, 2 mbar, 2) (:316-323)
    const double w0 = mbar_squared, w1 = 2.0 * mbar, wsum = w0 + w1 + 2.0;
    for (int iter = 0; iter < kMaxRejectionIterations; iter++) {
      (*samples)++;
      double pick = rng.canonical() * wsum, kbar;
      if (pick < w0) {
        kbar = -log(1.0 - rng.canonical());
        phi_over_2pi = rng.canonical();
        costheta = 2.0 * rng.canonical() - 1.0;
      } else if (pick < w0 + w1) {
        double l1 = log(1.0 - rng.canonical()), l2 = log(1.0 - rng.canonical());
        kbar = -(l1 + l2);
        phi_over_2pi = -l1 / kbar;
        costheta = 2.0 * rng.canonical() - 1.0;
      } else {
        double l1 = log(1.0 - rng.canonical()), l2 = log(1.0 - rng.canonical()), l3 = log(1.0 - rng.canonical());
        kbar = -(l1 + l2 + l3);
        phi_over_2pi = (l1 + l2) * (l1 + l2) / (kbar * kbar);
        costheta = (l1 - l2) / (l1 + l2);
      }
      Ebar = kbar + mbar;
      pbar = sqrt(Ebar * Ebar - mbar_squared);
      double boltz = exp(Ebar - chem);
      feq = 1.0 / (boltz + sign);
      double weight = pbar / Ebar * boltz * feq;
      if (rng.canonical() < weight) { accepted = true; break; }
    }
  }
  double p = pbar * T, phi = phi_over_2pi * kTwoPi;
  double sintheta = sqrt(1.0 - costheta * costheta);
  LrfMomentum q;
  q.E = Ebar * T;
  q.px = p * sintheta * cos(phi);
  q.py = p * sintheta * sin(phi);
  q.pz = p * costheta;
  q.feq = feq;
  q.ok = accepted;
  return q;
}

// One proposed hadron of species (mass, sign, baryon) in the cell described by `pk`: momentum, viscous and flux
// weights, acceptance (:929-1059).  Returns true when accepted; pLRF holds the (rescaled) rest-frame momentum.
template <int df_mode, class PackFn>
IS3D_HD bool sample_hadron(Philox &rng, PackFn pk, double mass, double sign, double baryon, long *samples, LrfMomentum *out)
{
  const double mass_squared = mass * mass;
  const bool breakdown = pk(SP_BREAKDOWN) != 0.0;
  const double T = pk(SP_T);
  double w_visc = 1.0;
  LrfMomentum p;
  const bool linear = (df_mode == 1 || df_mode == 2 || (df_mode == 3 && breakdown));
  if (linear) {
    p = sample_momentum(rng, samples, mass, sign, T, baryon * pk(SP_ALPHAB));
    double E = p.E, px = p.px, py = p.py, pz = p.pz, feqbar = 1.0 - sign * p.feq;
    double pipp = px * px * pk(SP_PIXX) + py * py * pk(SP_PIYY) + pz * pz * pk(SP_PIZZ) +
                  2. * (px * py * pk(SP_PIXY) + px * pz * pk(SP_PIXZ) + py * pz * pk(SP_PIYZ));
    double Vp = -(px * pk(SP_VX) + py * pk(SP_VY) + pz * pk(SP_VZ));
    double df_shear, df_bulk, df_diff;
    if (df_mode == 1) {
      df_shear = pipp / pk(SP_C0);
      df_bulk = (pk(SP_C1) * mass_squared + (baryon * pk(SP_C2) + pk(SP_C3) * E) * E) * pk(SP_BULKPI);
      df_diff = (baryon * pk(SP_C4) + pk(SP_C5) * E) * Vp;
    } else {
      df_shear = pipp / (pk(SP_C0) * E);
      df_bulk = (baryon * pk(SP_C1) + pk(SP_C2) * E + (E - mass_squared / E) / pk(SP_C3)) * pk(SP_C4);
      df_diff = (pk(SP_BER) - baryon / E) * Vp / pk(SP_C5);
    }
    double df_reg = fmax(-1.0, fmin(1.0, feqbar * (df_shear + df_bulk + df_diff)));
    w_visc = (1.0 + df_reg) / 2.0;
  } else if (df_mode == 3) {
    p = sample_momentum(rng, samples, mass, sign, pk(SP_TSAMPLE), baryon * pk(SP_ALPHAB_SAMPLE));
    // rescale_momentum (:407-426): p_i = A_ij p'_j + diff_mod (E' ber + b) V_i
    double dm = pk(SP_DIFF_MOD) * (p.E * pk(SP_BER) + baryon), iso = pk(SP_ISO_SCALE), sm = pk(SP_SHEAR_MOD);
    double px = iso * p.px + sm * (pk(SP_PIXX) * p.px + pk(SP_PIXY) * p.py + pk(SP_PIXZ) * p.pz) + dm * pk(SP_VX);
    double py = iso * p.py + sm * (pk(SP_PIXY) * p.px + pk(SP_PIYY) * p.py + pk(SP_PIYZ) * p.pz) + dm * pk(SP_VY);
    double pz = iso * p.pz + sm * (pk(SP_PIXZ) * p.px + pk(SP_PIYZ) * p.py + pk(SP_PIZZ) * p.pz) + dm * pk(SP_VZ);
    p.px = px; p.py = py; p.pz = pz;
    p.E = sqrt(mass_squared + px * px + py * py + pz * pz);
  } else if (df_mode == 5) {
    // PTMA (sample_dN_pTdpTdphidy_famod, ParticleSampler.cpp:1518-1531): thermal momentum at (Lambda, b upsilon_B), then
    // p_i = B_ij p'_j (rescale_momentum_famod :428-445; the pack holds B, or the identity on breakdown); no df weight
    p = sample_momentum(rng, samples, mass, sign, pk(SP_TSAMPLE), baryon * pk(SP_ALPHAB_SAMPLE));
    double px = pk(SP_PIXX) * p.px + pk(SP_PIXY) * p.py + pk(SP_PIXZ) * p.pz;
    double py = pk(SP_PIXY) * p.px + pk(SP_PIYY) * p.py + pk(SP_PIYZ) * p.pz;
    double pz = pk(SP_PIXZ) * p.px + pk(SP_PIYZ) * p.py + pk(SP_PIZZ) * p.pz;
    p.px = px; p.py = py; p.pz = pz;
    p.E = sqrt(mass_squared + px * px + py * py + pz * pz);
  } else {   // df_mode 4
    p = sample_momentum(rng, samples, mass, sign, T, 0.0);
    if (!breakdown) {
      double iso = pk(SP_ISO_SCALE), sm = pk(SP_SHEAR_MOD);
      double px = iso * p.px + sm * (pk(SP_PIXX) * p.px + pk(SP_PIXY) * p.py + pk(SP_PIXZ) * p.pz);
      double py = iso * p.py + sm * (pk(SP_PIXY) * p.px + pk(SP_PIYY) * p.py + pk(SP_PIYZ) * p.pz);
      double pz = iso * p.pz + sm * (pk(SP_PIXZ) * p.px + pk(SP_PIYZ) * p.py + pk(SP_PIZZ) * p.pz);
      p.px = px; p.py = py; p.pz = pz;
      p.E = sqrt(mass_squared + px * px + py * py + pz * pz);
    } else {
      double E = p.E, px = p.px, py = p.py, pz = p.pz, feqbar = 1.0 - sign * p.feq;
      double pipp = px * px * pk(SP_PIXX) + py * py * pk(SP_PIYY) + pz * pz * pk(SP_PIZZ) +
                    2. * (px * py * pk(SP_PIXY) + px * pz * pk(SP_PIXZ) + py * pz * pk(SP_PIYZ));
      double df_shear = feqbar * pipp / (pk(SP_C0) * E);
      double df_bulk = pk(SP_C1) + feqbar * pk(SP_C2) * (E - mass_squared / E);
      double df_reg = fmax(-1.0, fmin(1.0, df_shear + df_bulk));
      w_visc = (1.0 + df_reg) / 2.0;
    }
  }
  double w_flux = fmax(0.0, p.E * pk(SP_DST) - p.px * pk(SP_DSX) - p.py * pk(SP_DSY) - p.pz * pk(SP_DSZ)) / (p.E * pk(SP_DSMAX));
  *out = p;
  if (!p.ok) return false;
  return rng.canonical() < (w_flux * w_visc);
}

// lab-frame momentum and spacetime point of an accepted hadron (:1062-1109)
struct LabParticle { double E, px, py, pz, eta, t, z, rapidity; };

template <class PackFn>
IS3D_HD LabParticle boost_to_lab(Philox &rng, PackFn pk, const LrfMomentum &p, double mass, int dimension, double y_max)
{
  // Lab_Momentum::boost_pLRF_to_lab_frame (Momentum.cpp:14-31)
  double ptau = p.E * pk(SP_UT) + p.px * pk(SP_XT) + p.pz * pk(SP_ZT);
  double px = p.E * pk(SP_UX) + p.px * pk(SP_XX) + p.py * pk(SP_YX);
  double py = p.E * pk(SP_UY) + p.px * pk(SP_XY) + p.py * pk(SP_YY);
  double pn = p.E * pk(SP_UN) + p.px * pk(SP_XN) + p.pz * pk(SP_ZN);
  double tau = pk(SP_TAU);
  LabParticle q;
  q.px = px; q.py = py;
  if (dimension == 2) {
    double rapidity = y_max * (2.0 * rng.canonical() - 1.0);
    double sinhy = sinh(rapidity), coshy = sqrt(1.0 + sinhy * sinhy);
    double tau_pn = tau * pn;
    double mT = sqrt(ptau * ptau - tau_pn * tau_pn);
    double sinheta = (ptau * sinhy - tau_pn * coshy) / mT;
    q.eta = asinh(sinheta);
    double cosheta = sqrt(1.0 + sinheta * sinheta);
    q.pz = mT * sinhy; q.E = mT * coshy; q.rapidity = rapidity;
    q.t = tau * cosheta; q.z = tau * sinheta;
  } else {
    double eta = pk(SP_ETA), sinheta = sinh(eta), cosheta = sqrt(1.0 + sinheta * sinheta);
    q.pz = tau * pn * cosheta + ptau * sinheta;
    q.E = sqrt(mass * mass + px * px + py * py + q.pz * q.pz);
    q.rapidity = 0.5 * log((q.E + q.pz) / (q.E - q.pz));
    q.eta = eta; q.t = tau * cosheta; q.z = tau * sinheta;
  }
  return q;
}

}  // namespace is3d
