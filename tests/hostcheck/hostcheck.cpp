// TEST INFRASTRUCTURE ONLY (development aid) -- never built into, loaded by or called from the product.
//
// Compiles the IS3D_HD per-cell / per-momentum formulas of is3d_b200/csrc/*.cuh with g++ and runs them in plain
// loops over a working directory, so that a formula slip is caught here (no GPU in the build container) before a
// GPU run.  The CUDA kernels call the very same inline functions; what this cannot check is the kernels' launch
// geometry, shared-memory staging and reductions -- the `-m gpu` parity tests do that through the C ABI.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <vector>

#include "../../is3d_b200/csrc/spectra_df.cuh"
#include "../../is3d_b200/csrc/spectra_feqmod.cuh"
#include "../../is3d_b200/host/host_dfview.hpp"
#include "../../is3d_b200/host/is3d_host.hpp"

static const double *exp_table()
{
  static double tab[is3d::kExpTableSize];
  static bool ready = false;
  if (!ready) { is3d::fill_exp_table(tab); ready = true; }
  return tab;
}

using namespace is3dhost;

namespace {
struct Loaded {
  ParameterReader par;
  FO_surface surf;
  std::vector<particle_info> pdg;
  std::vector<int> chosen;
  Table pT, phi, y, eta;
  Deltaf_Data *df = nullptr;
  void load(const char *root)
  {
    set_root(root);
    par.readFromFile("iS3D_parameters.dat");
    FO_data_reader rd(&par, "input");
    rd.get_number_cells();
    rd.read_freezeout_surface(surf);
    PDG_Data p(&par);
    p.read_resonances(pdg);
    Table ch("PDG/chosen_particles.dat");
    for (long m = 0; m < ch.getNumberOfRows(); m++)
      for (size_t n = 0; n < pdg.size(); n++)
        if (pdg[n].mc_id == (long)ch.get(1, m + 1)) { chosen.push_back((int)n); break; }
    df = new Deltaf_Data(&par);
    df->load_df_coefficient_data();
    if (!(int)par.getVal("include_baryon")) df->compute_jonah_coefficients(pdg);
    df->compute_particle_densities(pdg);
    pT.loadTableFromFile("tables/momentum/pT_table.dat");
    phi.loadTableFromFile("tables/momentum/phi_table.dat");
    y.loadTableFromFile("tables/momentum/y_table.dat");
    eta.loadTableFromFile("tables/spacetime_rapidity/eta_table.dat");
  }
};

template <int MODE, bool BARYON>
double eval_dispatch(bool reg, bool outflow, const is3d::DfItem &it, const is3d::DfBin &b)
{
  if (reg && outflow) return is3d::df_eval<MODE, BARYON, true, true>(it, b, exp_table());
  if (reg) return is3d::df_eval<MODE, BARYON, true, false>(it, b, exp_table());
  if (outflow) return is3d::df_eval<MODE, BARYON, false, true>(it, b, exp_table());
  return is3d::df_eval<MODE, BARYON, false, false>(it, b, exp_table());
}
}  // namespace

// out: Ns*NpT*Nphi*Ny doubles, reference indexing.  Returns the number of doubles written (0 on unsupported mode).
extern "C" long hostcheck_spectra_df(const char *root, double *out, long capacity)
{
  Loaded L;
  L.load(root);
  is3d::DfFlags fl;
  fl.df_mode = L.par.getVal("df_mode"); fl.dimension = L.par.getVal("dimension");
  fl.include_baryon = L.par.getVal("include_baryon"); fl.include_bulk = L.par.getVal("include_bulk_deltaf");
  fl.include_shear = L.par.getVal("include_shear_deltaf"); fl.include_baryondiff = L.par.getVal("include_baryondiff_deltaf");
  const bool reg = (int)L.par.getVal("regulate_deltaf"), outflow = (int)L.par.getVal("outflow");
  if (fl.df_mode != 1 && fl.df_mode != 2) return 0;
  HostDfView view(*L.df);
  const int ns = (int)L.chosen.size(), npT = (int)L.pT.getNumberOfRows(), nphi = (int)L.phi.getNumberOfRows();
  const int ny = fl.dimension == 3 ? (int)L.y.getNumberOfRows() : 1;
  const int neta = fl.dimension == 3 ? 1 : (int)L.eta.getNumberOfRows();
  const long total = (long)ns * npT * nphi * ny;
  if (total > capacity) return -total;
  std::vector<double> acc(total, 0.0);
  is3d::SurfaceView sv;
  for (int k = 0; k < 25; k++) sv.col[k] = L.surf.col[k].data();
  sv.n = L.surf.size();
  for (int64_t ic = 0; ic < sv.n; ic++) {
    is3d::Cell c = is3d::load_cell(sv, ic, fl.include_baryon != 0);
    double pack[is3d::DP_SIZE];
    int st = is3d::df_setup_cell(c, view.tb, fl, pack);
    if (st == is3d::CELL_OUT_OF_TABLE) { printf("hostcheck: cell %ld out of table\n", (long)ic); return 0; }
    if (st != is3d::CELL_OK) continue;
    auto pk = [&](int k) { return pack[k]; };
    for (int iy = 0; iy < ny; iy++) {
      double yv = fl.dimension == 3 ? L.y.get(1, iy + 1) : 0.0;
      for (int ie = 0; ie < neta; ie++) {
        double etav = fl.dimension == 3 ? pack[is3d::DP_ETA] : L.eta.get(1, ie + 1);
        double w = fl.dimension == 3 ? 1.0 : L.eta.get(2, ie + 1);
        double sh = sinh(yv - etav), ch = sqrt(1.0 + sh * sh);
        for (int ip = 0; ip < nphi; ip++) {
          double ph = L.phi.get(1, ip + 1);
          is3d::DfItem it = is3d::df_make_item(pk, sh, ch, cos(ph), sin(ph), w);
          for (int s = 0; s < ns; s++) {
            const particle_info &p = L.pdg[L.chosen[s]];
            for (int ipT = 0; ipT < npT; ipT++) {
              double pTv = L.pT.get(1, ipT + 1), m2 = p.mass * p.mass, mT = sqrt(m2 + pTv * pTv);
              is3d::DfBin b{mT, pTv, mT * mT, mT * pTv, pTv * pTv, m2, (double)p.baryon, (double)p.sign};
              double v;
              if (fl.df_mode == 1) v = fl.include_baryon ? eval_dispatch<1, true>(reg, outflow, it, b) : eval_dispatch<1, false>(reg, outflow, it, b);
              else v = fl.include_baryon ? eval_dispatch<2, true>(reg, outflow, it, b) : eval_dispatch<2, false>(reg, outflow, it, b);
              acc[iy + (long)ny * (ip + (long)nphi * (ipT + (long)npT * s))] += v;
            }
          }
        }
      }
    }
  }
  for (int s = 0; s < ns; s++) {
    double g = (double)L.pdg[L.chosen[s]].gspin;
    for (long k = 0; k < (long)npT * nphi * ny; k++) out[(long)s * npT * nphi * ny + k] = is3d::kCooperFryePrefactor * g * acc[(long)s * npT * nphi * ny + k];
  }
  return total;
}

// df_mode 3, 4 (K2).  Same loop emulation: per cell setup pack, PTM renorm per (cell, species), per (y, phi, eta) item.
extern "C" long hostcheck_spectra_feqmod(const char *root, double *out, long capacity, long *stats4)
{
  Loaded L;
  L.load(root);
  is3d::FeqmodFlags fl;
  fl.df_mode = L.par.getVal("df_mode"); fl.dimension = L.par.getVal("dimension");
  fl.include_baryon = L.par.getVal("include_baryon"); fl.include_bulk = L.par.getVal("include_bulk_deltaf");
  fl.include_shear = L.par.getVal("include_shear_deltaf"); fl.include_baryondiff = L.par.getVal("include_baryondiff_deltaf");
  fl.deta_min = L.par.getVal("deta_min"); fl.mass_pion0 = L.par.getVal("mass_pion0");
  fl.bulkPi_over_P_max = L.df->bulkPi_over_Peq_max;
  const bool reg = (int)L.par.getVal("regulate_deltaf"), outflow = (int)L.par.getVal("outflow");
  if (fl.df_mode != 3 && fl.df_mode != 4) return 0;
  const bool species_renorm = fl.df_mode == 3 && fl.include_bulk;
  HostDfView view(*L.df);
  Gauss_Laguerre gla;
  gla.load_roots_and_weights("tables/gauss/gla_roots_weights.txt");
  const int ns = (int)L.chosen.size(), npT = (int)L.pT.getNumberOfRows(), nphi = (int)L.phi.getNumberOfRows();
  const int ny = fl.dimension == 3 ? (int)L.y.getNumberOfRows() : 1;
  const int neta = fl.dimension == 3 ? 1 : (int)L.eta.getNumberOfRows();
  const long total = (long)ns * npT * nphi * ny;
  if (total > capacity) return -total;
  std::vector<double> acc(total, 0.0);
  is3d::SurfaceView sv;
  for (int k = 0; k < 25; k++) sv.col[k] = L.surf.col[k].data();
  sv.n = L.surf.size();
  long nbreak = 0, npl = 0, nskip = 0;
  for (int64_t ic = 0; ic < sv.n; ic++) {
    is3d::Cell c = is3d::load_cell(sv, ic, fl.include_baryon != 0);
    double pack[is3d::FP_SIZE];
    int st = is3d::feqmod_setup_cell(c, view.tb, fl, gla.root.data(), gla.weight.data(), gla.points, pack);
    if (st == is3d::CELL_OUT_OF_TABLE) { printf("hostcheck: cell %ld out of table\n", (long)ic); return 0; }
    if (st == is3d::CELL_SKIPPED) { nskip++; continue; }
    if (st & is3d::CELL_BREAKDOWN) nbreak++;
    if (st & is3d::CELL_PL_NEGATIVE) npl++;
    auto pk = [&](int k) { return pack[k]; };
    std::vector<double> rn(ns, pack[is3d::FP_RENORM]);
    if (species_renorm)
      for (int s = 0; s < ns; s++) {
        const particle_info &p = L.pdg[L.chosen[s]];
        rn[s] = is3d::feqmod_renorm_ptm(pk, p.mass, (double)p.gspin, (double)p.baryon, (double)p.sign, gla.root.data(), gla.weight.data(), gla.points);
      }
    for (int iy = 0; iy < ny; iy++) {
      double yv = fl.dimension == 3 ? L.y.get(1, iy + 1) : 0.0;
      for (int ie = 0; ie < neta; ie++) {
        double etav = fl.dimension == 3 ? pack[is3d::DP_ETA] : L.eta.get(1, ie + 1);
        double w = fl.dimension == 3 ? 1.0 : L.eta.get(2, ie + 1);
        bool linear = pack[is3d::FP_BREAKDOWN] != 0.0;
        if (fl.dimension == 3 && !linear && pack[is3d::FP_DETA] < 0.01 && fabs(yv - etav) < pack[is3d::FP_DETA]) linear = true;
        double d = linear ? (yv - etav) : (yv - pack[is3d::FP_ETA_SCALE] * etav);
        double sh = sinh(d), ch = cosh(d);
        for (int ip = 0; ip < nphi; ip++) {
          double ph = L.phi.get(1, ip + 1);
          is3d::DfItem lin;
          is3d::FeqmodItem mod;
          if (linear) lin = is3d::feqmod_make_linear_item(pk, sh, ch, cos(ph), sin(ph), w);
          else mod = is3d::feqmod_make_item(pk, sh, ch, cos(ph), sin(ph), w);
          for (int s = 0; s < ns; s++) {
            const particle_info &p = L.pdg[L.chosen[s]];
            for (int ipT = 0; ipT < npT; ipT++) {
              double pTv = L.pT.get(1, ipT + 1), m2 = p.mass * p.mass, mT = sqrt(m2 + pTv * pTv);
              is3d::DfBin b{mT, pTv, mT * mT, mT * pTv, pTv * pTv, m2, (double)p.baryon, (double)p.sign};
              double v;
              if (!linear) {
                if (fl.include_baryon) v = outflow ? is3d::feqmod_eval<true, true>(mod, b, rn[s], exp_table()) : is3d::feqmod_eval<true, false>(mod, b, rn[s], exp_table());
                else v = outflow ? is3d::feqmod_eval<false, true>(mod, b, rn[s], exp_table()) : is3d::feqmod_eval<false, false>(mod, b, rn[s], exp_table());
              } else {
                v = fl.include_baryon ? eval_dispatch<2, true>(reg, outflow, lin, b) : eval_dispatch<2, false>(reg, outflow, lin, b);
                if (rn[s] == 0.0) v = 0.0;
              }
              acc[iy + (long)ny * (ip + (long)nphi * (ipT + (long)npT * s))] += v;
            }
          }
        }
      }
    }
  }
  for (int s = 0; s < ns; s++) {
    double g = (double)L.pdg[L.chosen[s]].gspin;
    for (long k = 0; k < (long)npT * nphi * ny; k++) out[(long)s * npT * nphi * ny + k] = is3d::kCooperFryePrefactor * g * acc[(long)s * npT * nphi * ny + k];
  }
  if (stats4) { stats4[0] = nskip; stats4[1] = nbreak; stats4[2] = npl; stats4[3] = 0; }
  return total;
}

// ---- sampler (K5/K6) emulation: same Philox streams, same per-cell / per-hadron functions, serial loops ----------
#include "../../is3d_b200/csrc/sampler.cuh"

// hists: dN_dy [ns][y_bins], dN_deta [ns][eta_bins], dN_pT [ns][pT_bins]; returns accepted hadrons, *yield = mean total
extern "C" long hostcheck_sampler(const char *root, long nevents, double *dN_dy, double *dN_deta, double *dN_pT, double *yield_out,
                                  long *proposed_out)
{
  Loaded L;
  L.load(root);
  is3d::SamplerFlags fl;
  fl.df_mode = L.par.getVal("df_mode"); fl.dimension = L.par.getVal("dimension");
  fl.include_baryon = L.par.getVal("include_baryon"); fl.include_bulk = L.par.getVal("include_bulk_deltaf");
  fl.include_shear = L.par.getVal("include_shear_deltaf"); fl.include_baryondiff = L.par.getVal("include_baryondiff_deltaf");
  fl.fast = L.par.getVal("fast"); fl.deta_min = L.par.getVal("deta_min"); fl.mass_pion0 = L.par.getVal("mass_pion0");
  fl.bulkPi_over_P_max = L.df->bulkPi_over_Peq_max; fl.y_cut = L.par.getVal("y_cut");
  const uint64_t seed = (uint64_t)L.par.getVal("sampler_seed");
  HostDfView view(*L.df);
  Plasma QGP;
  QGP.load_thermodynamic_averages();
  fl.T_avg = QGP.temperature; fl.F_avg = 0.0; fl.betabulk_avg = 1.0;
  if (fl.df_mode == 3 && fl.fast) {
    is3d::DfCoeff d;
    is3d::evaluate_df_coefficients(view.tb, 3, fl.include_baryon, QGP.temperature, QGP.baryon_chemical_potential, 0.0, 0.0, 0.0, &d);
    fl.F_avg = d.F; fl.betabulk_avg = d.betabulk;
  }
  Gauss_Laguerre gla;
  gla.load_roots_and_weights("tables/gauss/gla_roots_weights.txt");
  const int ns = (int)L.chosen.size();
  std::vector<double> cumA(ns), cumB(ns);
  double totA = 0, totB = 0, totD = 0;
  for (int s = 0; s < ns; s++) {
    const particle_info &p = L.pdg[L.chosen[s]];
    totA += p.equilibrium_density; totB += p.bulk_density; totD += p.diff_density;
    cumA[s] = totA; cumB[s] = totB;
  }
  const int y_bins = L.par.getVal("y_bins"), eta_bins = L.par.getVal("eta_bins"), pT_bins = L.par.getVal("pT_bins");
  const double y_cut = fl.y_cut, y_w = 2.0 * y_cut / y_bins, eta_cut = L.par.getVal("eta_cut"), eta_w = 2.0 * eta_cut / eta_bins;
  const double pT_min = L.par.getVal("pT_min"), pT_w = (L.par.getVal("pT_max") - pT_min) / pT_bins;
  const double y_max = fl.dimension == 2 ? y_cut : 0.5;
  is3d::SurfaceView sv;
  for (int k = 0; k < 25; k++) sv.col[k] = L.surf.col[k].data();
  sv.n = L.surf.size();
  long accepted = 0, proposed = 0;
  double yield = 0.0;
  for (int64_t ic = 0; ic < sv.n; ic++) {
    is3d::Cell c = is3d::load_cell(sv, ic, fl.include_baryon != 0);
    double pack[is3d::SP_SIZE];
    int st = is3d::sampler_setup_cell(c, view.tb, fl, gla.root.data(), gla.weight.data(), gla.points, totA, totB, pack);
    if (st == is3d::CELL_SKIPPED || st == is3d::CELL_OUT_OF_TABLE) continue;
    yield += is3d::cell_mean_yield(pack, fl.df_mode, totA, totB, totD);
    if (!(pack[is3d::SP_DNTOT] > 0.0)) continue;
    is3d::Philox rng;
    rng.init(seed, (uint64_t)ic, 0xFFFFFFFFu);
    long n = is3d::poisson_sample(rng, (double)nevents * pack[is3d::SP_DNTOT]);
    auto pk = [&](int k) { return pack[k]; };
    for (long h = 0; h < n; h++) {
      proposed++;
      is3d::Philox r;
      r.init(seed, (uint64_t)ic, (uint32_t)h);
      (void)r.canonical();     // event label
      double target = r.canonical() * (pack[is3d::SP_WA] * totA + pack[is3d::SP_WB] * totB);
      int a = 0, b = ns - 1;
      while (a < b) { int m = (a + b) >> 1; if (pack[is3d::SP_WA] * cumA[m] + pack[is3d::SP_WB] * cumB[m] > target) b = m; else a = m + 1; }
      const particle_info &p = L.pdg[L.chosen[a]];
      long samples = 0;
      is3d::LrfMomentum q;
      if (!is3d::sample_hadron(r, pk, fl.df_mode, p.mass, (double)p.sign, (double)p.baryon, &samples, &q)) continue;
      is3d::LabParticle lab = is3d::boost_to_lab(r, pk, q, p.mass, fl.dimension, y_max);
      accepted++;
      int iy = (int)floor((lab.rapidity + y_cut) / y_w);
      if (iy >= 0 && iy < y_bins) dN_dy[(size_t)a * y_bins + iy] += 1.0;
      int ie = (int)floor((lab.eta + eta_cut) / eta_w);
      if (ie >= 0 && ie < eta_bins) dN_deta[(size_t)a * eta_bins + ie] += 1.0;
      int ip = (int)floor((sqrt(lab.px * lab.px + lab.py * lab.py) - pT_min) / pT_w);
      if (ip >= 0 && ip < pT_bins) dN_pT[(size_t)a * pT_bins + ip] += 1.0;
    }
  }
  if (fl.dimension == 2) yield *= 2.0 * y_cut;
  *yield_out = yield;
  *proposed_out = proposed;
  return accepted;
}

// ---- df_mode 5 (K3): serial-reducer Newton solve per cell with the reference's chain, then the K2 momentum math ----
#include "../../is3d_b200/csrc/aniso.cuh"

extern "C" long hostcheck_spectra_famod(const char *root, int chain_on, double *out, long capacity, long *stats5)
{
  Loaded L;
  L.load(root);
  is3d::FamodFlags fl;
  fl.dimension = L.par.getVal("dimension"); fl.include_baryon = L.par.getVal("include_baryon");
  fl.include_shear = L.par.getVal("include_shear_deltaf"); fl.include_baryondiff = L.par.getVal("include_baryondiff_deltaf");
  fl.deta_min = L.par.getVal("deta_min");
  const bool outflow = (int)L.par.getVal("outflow");
  const int ns = (int)L.chosen.size(), npT = (int)L.pT.getNumberOfRows(), nphi = (int)L.phi.getNumberOfRows();
  const int ny = fl.dimension == 3 ? (int)L.y.getNumberOfRows() : 1;
  const int neta = fl.dimension == 3 ? 1 : (int)L.eta.getNumberOfRows();
  const long total = (long)ns * npT * nphi * ny;
  if (total > capacity) return -total;
  std::vector<double> acc(total, 0.0);
  std::vector<double> pm, ps, pd;
  for (auto &p : L.pdg) { pm.push_back(p.mass); ps.push_back(p.sign); pd.push_back(p.gspin); }
  double gl16[96];
  is3d::fill_gl16_table(gl16);
  is3d::AnisoHadrons h{pm.data(), ps.data(), pd.data(), (int)std::min<size_t>(320, pm.size()), gl16, exp_table(), 1};
  is3d::SurfaceView sv;
  for (int k = 0; k < 25; k++) sv.col[k] = L.surf.col[k].data();
  sv.n = L.surf.size();
  is3d::FamodChain chain{0, 0, 0, false};
  is3d::SerialReducer red;
  long nbreak = 0, npl = 0, nfail = 0, iters = 0;
  for (int64_t ic = 0; ic < sv.n; ic++) {
    is3d::Cell c = is3d::load_cell(sv, ic, fl.include_baryon != 0);
    double pack[is3d::FP_SIZE];
    int it;
    int st = is3d::famod_setup_cell(red, c, fl, h, chain_on ? &chain : nullptr, pack, &it);
    if (st == is3d::CELL_SKIPPED) continue;
    iters += it;
    if (st & is3d::CELL_BREAKDOWN) nbreak++;
    if (st & is3d::CELL_PL_NEGATIVE) npl++;
    if (st & is3d::CELL_RECONSTRUCTION_FAIL) nfail++;
    auto pk = [&](int k) { return pack[k]; };
    for (int iy = 0; iy < ny; iy++) {
      double yv = fl.dimension == 3 ? L.y.get(1, iy + 1) : 0.0;
      for (int ie = 0; ie < neta; ie++) {
        double etav = fl.dimension == 3 ? pack[is3d::DP_ETA] : L.eta.get(1, ie + 1);
        double w = fl.dimension == 3 ? 1.0 : L.eta.get(2, ie + 1);
        bool linear = pack[is3d::FP_BREAKDOWN] != 0.0;
        if (fl.dimension == 3 && !linear && pack[is3d::FP_DETA] < 0.01 && fabs(yv - etav) < pack[is3d::FP_DETA]) linear = true;
        double d = linear ? (yv - etav) : (yv - pack[is3d::FP_ETA_SCALE] * etav);
        double sh = sinh(d), ch = cosh(d);
        for (int ip = 0; ip < nphi; ip++) {
          double ph = L.phi.get(1, ip + 1);
          is3d::DfItem lin;
          is3d::FeqmodItem mod;
          if (linear) lin = is3d::feqmod_make_linear_item(pk, sh, ch, cos(ph), sin(ph), w, true);
          else mod = is3d::feqmod_make_item(pk, sh, ch, cos(ph), sin(ph), w, true);
          for (int s = 0; s < ns; s++) {
            const particle_info &p = L.pdg[L.chosen[s]];
            for (int ipT = 0; ipT < npT; ipT++) {
              double pTv = L.pT.get(1, ipT + 1), m2 = p.mass * p.mass, mT = sqrt(m2 + pTv * pTv);
              is3d::DfBin b{mT, pTv, mT * mT, mT * pTv, pTv * pTv, m2, (double)p.baryon, (double)p.sign};
              double v;
              if (!linear) {
                if (fl.include_baryon) v = outflow ? is3d::feqmod_eval<true, true>(mod, b, mod.renorm, exp_table()) : is3d::feqmod_eval<true, false>(mod, b, mod.renorm, exp_table());
                else v = outflow ? is3d::feqmod_eval<false, true>(mod, b, mod.renorm, exp_table()) : is3d::feqmod_eval<false, false>(mod, b, mod.renorm, exp_table());
              } else {
                v = fl.include_baryon ? eval_dispatch<2, true>(false, outflow, lin, b) : eval_dispatch<2, false>(false, outflow, lin, b);
              }
              acc[iy + (long)ny * (ip + (long)nphi * (ipT + (long)npT * s))] += v;
            }
          }
        }
      }
    }
  }
  for (int s = 0; s < ns; s++) {
    double g = (double)L.pdg[L.chosen[s]].gspin;
    for (long k = 0; k < (long)npT * nphi * ny; k++) out[(long)s * npT * nphi * ny + k] = is3d::kCooperFryePrefactor * g * acc[(long)s * npT * nphi * ny + k];
  }
  if (stats5) { stats5[0] = nbreak; stats5[1] = npl; stats5[2] = nfail; stats5[3] = iters; stats5[4] = 0; }
  return total;
}
