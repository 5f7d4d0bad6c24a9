// Deltaf_Data host side: coefficient-table loading, PTB (lambda, z) tables and the fast-mode species densities.
// Follows reference src/cpp/DeltafData.cpp:21-295 and :555-690.  The per-cell evaluation
// (evaluate_df_coefficients) lives on the GPU (csrc/dftables.cuh); the same header is compiled here for the single
// evaluation at the surface averages that compute_particle_densities needs at start-up.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../csrc/dftables.cuh"
#include "../csrc/gauss_thermal.cuh"
#include "host_parallel.hpp"
#include "is3d_host.hpp"
#include "host_dfview.hpp"

namespace is3dhost {

Deltaf_Data::Deltaf_Data(ParameterReader *paraRdr)
{
  hrg_eos = paraRdr->getVal("hrg_eos");
  mode = paraRdr->getVal("mode");
  df_mode = paraRdr->getVal("df_mode");
  include_baryon = paraRdr->getVal("include_baryon");
  if (hrg_eos == 1) hrg_eos_path = "deltaf_coefficients/vh/urqmd/";
  else if (hrg_eos == 2) hrg_eos_path = "deltaf_coefficients/vh/smash/";
  else if (hrg_eos == 3) hrg_eos_path = "deltaf_coefficients/vh/smash_box/";
  else fatal("Error: please choose hrg_eos = (1,2,3)");
}

// Each file: line 1 = #T points, line 2 = #muB points, line 3 = header, then rows "T muB value" with T fastest.
// For include_baryon = 0 only the muB = 0 block is read (DeltafData.cpp:134).
void Deltaf_Data::load_df_coefficient_data()
{
  static const char *names[10] = {"c0.dat", "c1.dat", "c2.dat", "c3.dat", "c4.dat", "F.dat", "G.dat", "betabulk.dat", "betaV.dat", "betapi.dat"};
  for (int k = 0; k < 10; k++) {
    std::string fn = hrg_eos_path + names[k];
    FILE *f = fopen(path(fn).c_str(), "r");
    if (!f) fatal("Couldn't open coefficient file " + fn);
    int nT = 0, nB = 0;
    if (fscanf(f, "%d\n%d\n", &nT, &nB) != 2) fatal("bad header in " + fn);
    char header[300];
    if (!fgets(header, 100, f)) fatal("bad header in " + fn);
    points_T = nT;
    points_muB = include_baryon ? nB : 1;
    if (k == 0) { T_array.assign(points_T, 0.0); muB_array.assign(points_muB, 0.0); }
    tab[k].assign((size_t)points_T * points_muB, 0.0);
    for (int iB = 0; iB < points_muB; iB++)
      for (int iT = 0; iT < points_T; iT++)
        if (fscanf(f, "%lf\t\t%lf\t\t%lf\n", &T_array[iT], &muB_array[iB], &tab[k][(size_t)iB * points_T + iT]) != 3)
          fatal("bad row in " + fn);
    fclose(f);
  }
}

// lambda^2(Pi/P) and z(Pi/P) on 301 lambda nodes in [-1, 2] at the average temperature (DeltafData.cpp:220-295)
void Deltaf_Data::compute_jonah_coefficients(const std::vector<particle_info> &particle_data)
{
  const double lambda_min = -1.0, lambda_max = 2.0;
  const double delta_lambda = (lambda_max - lambda_min) / ((double)jonah_points - 1.0);
  lambda_squared_array.assign(jonah_points, 0.0);
  z_array.assign(jonah_points, 0.0);
  bulkPi_over_Peq_array.assign(jonah_points, 0.0);
  bulkPi_over_Peq_max = -1.0;
  Plasma QGP;
  QGP.load_thermodynamic_averages();
  const double T = QGP.temperature;
  Gauss_Laguerre gla;
  gla.load_roots_and_weights("tables/gauss/gla_roots_weights.txt");
  const int pts = gla.points;
  const double *root2 = gla.roots(2), *weight2 = gla.weights(2);
  auto gauss1d = [&](double (*fn)(double, double, double, double), double mbar, double lambda, double sign) {
    double sum = 0.0;
    for (int k = 0; k < pts; k++) sum += weight2[k] * fn(root2[k], mbar, lambda, sign);
    return sum;
  };
  // The reference recomputes the lambda = 0 sums E, P inside the node loop (same value every time); here they are taken
  // once, and the 301 nodes -- independent O(N_pdg x 64) Gauss-Laguerre sums each -- are spread over the host threads
  // (the visible serial prefix of a run).  Every node's arithmetic is unchanged.
  double E = 0.0, P = 0.0;
  for (const particle_info &p : particle_data) {
    double degeneracy = (double)p.gspin, mass = p.mass, sign = (double)p.sign;
    double mbar = mass / T;
    if (mass == 0.0) continue;                          // photon skipped (:266)
    E += degeneracy * gauss1d(is3d::E_mod_int, mbar, 0.0, sign);
    P += (1.0 / 3.0) * degeneracy * gauss1d(is3d::P_mod_int, mbar, 0.0, sign);
  }
  const int nt = host_threads((size_t)jonah_points, "IS3D_TABLE_THREADS");
  parallel_for(nt, [&](int t) {
    for (int i = t; i < jonah_points; i += nt) {
      double lambda = lambda_min + (double)i * delta_lambda;
      double E_mod = 0.0, P_mod = 0.0;
      for (const particle_info &p : particle_data) {
        double degeneracy = (double)p.gspin, mass = p.mass, sign = (double)p.sign;
        double mbar = mass / T;
        if (mass == 0.0) continue;
        E_mod += degeneracy * gauss1d(is3d::E_mod_int, mbar, lambda, sign);
        P_mod += (1.0 / 3.0) * degeneracy * gauss1d(is3d::P_mod_int, mbar, lambda, sign);
      }
      double z = E / E_mod;
      lambda_squared_array[i] = lambda * lambda;
      z_array[i] = z;
      bulkPi_over_Peq_array[i] = (P_mod / P) * z - 1.0;
    }
  });
  for (int i = 0; i < jonah_points; i++) bulkPi_over_Peq_max = fmax(bulkPi_over_Peq_max, bulkPi_over_Peq_array[i]);
  have_jonah = true;
}

// host-pointer view of the coefficient tables (+ spline coefficients) in the layout the device code evaluates
HostDfView::HostDfView(const Deltaf_Data &d)
{
  tb.n_T = d.points_T; tb.n_muB = d.points_muB;
  tb.T = d.T_array.data(); tb.muB = d.muB_array.data();
  tb.T_min = d.T_array[0]; tb.muB_min = d.muB_array[0];
  tb.dT = fabs(d.T_array[1] - d.T_array[0]);
  tb.dmuB = d.points_muB > 1 ? fabs(d.muB_array[1] - d.muB_array[0]) : 0.0;
  for (int k = 0; k < 10; k++) tb.tab[k] = d.tab[k].data();
  auto mk = [&](int slot, const std::vector<double> &x, const double *y, int n, is3d::Spline *sp) {
    spc[slot].assign(n, 0.0);
    is3d::natural_cspline_coefficients(x.data(), y, n, spc[slot].data());
    sp->x = x.data(); sp->y = y; sp->c = spc[slot].data(); sp->n = n;
  };
  if (!d.include_baryon) {
    mk(0, d.T_array, d.tab[is3d::TAB_C0].data(), d.points_T, &tb.sp_c0);
    mk(1, d.T_array, d.tab[is3d::TAB_C2].data(), d.points_T, &tb.sp_c2);
    mk(2, d.T_array, d.tab[is3d::TAB_F].data(), d.points_T, &tb.sp_F);
    mk(3, d.T_array, d.tab[is3d::TAB_BETABULK].data(), d.points_T, &tb.sp_betabulk);
    mk(4, d.T_array, d.tab[is3d::TAB_BETAPI].data(), d.points_T, &tb.sp_betapi);
    if (d.have_jonah) {
      mk(5, d.bulkPi_over_Peq_array, d.lambda_squared_array.data(), Deltaf_Data::jonah_points, &tb.sp_lambda2);
      mk(6, d.bulkPi_over_Peq_array, d.z_array.data(), Deltaf_Data::jonah_points, &tb.sp_z);
      tb.bulkPi_over_P_max = d.bulkPi_over_Peq_max;
    }
  }
}

// equilibrium density and linear bulk / diffusion corrections of every PDG entry at (T_avg, muB_avg)
// (DeltafData.cpp:555-690)
void Deltaf_Data::compute_particle_densities(std::vector<particle_info> &particle_data)
{
  Plasma QGP;
  QGP.load_thermodynamic_averages();
  const double T = QGP.temperature, E = QGP.energy_density, P = QGP.pressure;
  const double muB = QGP.baryon_chemical_potential, nB = QGP.net_baryon_density;

  HostDfView view(*this);
  const is3d::DfTables &tb = view.tb;
  is3d::DfCoeff df;
  if (!is3d::evaluate_df_coefficients(tb, df_mode, include_baryon, T, muB, E, P, 0.0, &df)) {
    if (include_baryon) fatal("Error: (T,muB) outside df coefficient table. Exiting...");
    fprintf(stderr, "gsl: interp.c: ERROR: interpolation error\nDefault GSL error handler invoked.\n");
    abort();                                              // the reference dies in gsl_spline_eval here
  }
  double alphaB = muB / T;
  double baryon_enthalpy_ratio = nB / (E + P);

  Gauss_Laguerre gla;
  gla.load_roots_and_weights("tables/gauss/gla_roots_weights.txt");
  const int pts = gla.points;
  const double *r1 = gla.roots(1), *w1 = gla.weights(1), *r2 = gla.roots(2), *w2 = gla.weights(2);
  const double *r3 = gla.roots(3), *w3 = gla.weights(3);
  const double norm = two_pi2_hbarC3();

  for (particle_info &p : particle_data) {
    double mass = p.mass, degeneracy = (double)p.gspin, baryon = (double)p.baryon, sign = (double)p.sign;
    double mbar = mass / T;
    double neq_fact = degeneracy * pow(T, 3) / norm;
    double neq = neq_fact * is3d::gauss_thermal<is3d::TI_NEQ>(r1, w1, pts, mbar, alphaB, baryon, sign);
    double dn_bulk = 0.0, dn_diff = 0.0;
    if (df_mode == 1) {
      double J10_fact = degeneracy * pow(T, 3) / norm;
      double J20_fact = degeneracy * pow(T, 4) / norm;
      double J30_fact = degeneracy * pow(T, 5) / norm;
      double J31_fact = degeneracy * pow(T, 5) / norm / 3.0;
      double J10 = J10_fact * is3d::gauss_thermal<is3d::TI_J10>(r1, w1, pts, mbar, alphaB, baryon, sign);
      double J20 = J20_fact * is3d::gauss_thermal<is3d::TI_J20>(r2, w2, pts, mbar, alphaB, baryon, sign);
      double J30 = J30_fact * is3d::gauss_thermal<is3d::TI_J30>(r3, w3, pts, mbar, alphaB, baryon, sign);
      double J31 = J31_fact * is3d::gauss_thermal<is3d::TI_J31>(r3, w3, pts, mbar, alphaB, baryon, sign);
      dn_bulk = ((df.c0 - df.c2) * mass * mass * J10 + df.c1 * baryon * J20 + (4.0 * df.c2 - df.c0) * J30);
      dn_diff = baryon * df.c3 * neq * T + df.c4 * J31;
    } else if (df_mode == 2 || df_mode == 3 || df_mode == 5) {
      double J10_fact = degeneracy * pow(T, 3) / norm;
      double J11_fact = degeneracy * pow(T, 3) / norm / 3.0;
      double J20_fact = degeneracy * pow(T, 4) / norm;
      double J10 = J10_fact * is3d::gauss_thermal<is3d::TI_J10>(r1, w1, pts, mbar, alphaB, baryon, sign);
      double J11 = J11_fact * is3d::gauss_thermal<is3d::TI_J11>(r1, w1, pts, mbar, alphaB, baryon, sign);
      double J20 = J20_fact * is3d::gauss_thermal<is3d::TI_J20>(r2, w2, pts, mbar, alphaB, baryon, sign);
      dn_bulk = (neq + (baryon * J10 * df.G) + (J20 * df.F / pow(T, 2))) / df.betabulk;
      dn_diff = (neq * T * baryon_enthalpy_ratio - baryon * J11) / df.betaV;
    }
    p.equilibrium_density = neq;
    p.bulk_density = dn_bulk;
    p.diff_density = dn_diff;
  }
}

}  // namespace is3dhost
