// Freezeout-surface readers (modes 1/5/6/7) into a structure-of-arrays surface, plus the volume-weighted
// thermodynamic averages.  Column contracts and unit conversions follow reference src/cpp/readindata.cpp:167-729.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>

#include "is3d_host.hpp"

namespace is3dhost {

void FO_surface::resize(int64_t n, bool with_vorticity)
{
  for (auto &c : col) c.assign((size_t)n, 0.0);
  for (auto &c : vorticity) c.assign(with_vorticity ? (size_t)n : 0, 0.0);
}

// ds_max-weighted averages, identical in all three readers (e.g. readindata.cpp:330-360)
void compute_thermodynamic_averages(const FO_surface &s, double avg[5])
{
  double sums[6];
  compute_thermodynamic_sums(s, sums);
  for (int k = 0; k < 5; k++) avg[k] = sums[k] / sums[5];
}

// the numerators (T, E, P, muB, nB weighted by ds_max) and the denominator (sum of ds_max) of the averages: additive
// over cell blocks, so ranks that each hold a block can all-reduce these six numbers to get the surface averages
void compute_thermodynamic_sums(const FO_surface &s, double sums[6])
{
  double T_avg = 0, E_avg = 0, P_avg = 0, muB_avg = 0, nB_avg = 0, max_volume = 0;
  const int64_t n = s.size();
  for (int64_t i = 0; i < n; i++) {
    double tau = s.col[IS3D_COL_TAU][i], tau2 = tau * tau;
    double ux = s.col[IS3D_COL_UX][i], uy = s.col[IS3D_COL_UY][i], un = s.col[IS3D_COL_UN][i];
    double ut = sqrt(1. + ux * ux + uy * uy + tau2 * un * un);
    double dat = s.col[IS3D_COL_DAT][i], dax = s.col[IS3D_COL_DAX][i], day = s.col[IS3D_COL_DAY][i], dan = s.col[IS3D_COL_DAN][i];
    double uds = ut * dat + ux * dax + uy * day + un * dan;
    double ds_ds = dat * dat - dax * dax - day * day - dan * dan / tau2;
    double ds_max = fabs(uds) + sqrt(fabs(uds * uds - ds_ds));
    max_volume += ds_max;
    E_avg += (s.col[IS3D_COL_E][i] * ds_max);
    T_avg += (s.col[IS3D_COL_T][i] * ds_max);
    P_avg += (s.col[IS3D_COL_P][i] * ds_max);
    muB_avg += (s.col[IS3D_COL_MUB][i] * ds_max);
    nB_avg += (s.col[IS3D_COL_NB][i] * ds_max);
  }
  sums[0] = T_avg; sums[1] = E_avg; sums[2] = P_avg; sums[3] = muB_avg; sums[4] = nB_avg; sums[5] = max_volume;
}

// 15 significant digits, no trailing newline (readindata.cpp:363-366); later stages re-read this file, so the
// rounding to 15 digits is part of the reference's arithmetic
void write_thermodynamic_averages(const double avg[5])
{
  std::ofstream f(path("tables/thermodynamic/average_thermodynamic_quantities.dat").c_str(), std::ios_base::out);
  f << std::setprecision(15) << avg[0] << "\n" << avg[1] << "\n" << avg[2] << "\n" << avg[3] << "\n" << avg[4];
}

FO_data_reader::FO_data_reader(ParameterReader *paraRdr, const std::string &)
{
  mode = paraRdr->getVal("mode");
  dimension = paraRdr->getVal("dimension");
  include_baryon = paraRdr->getVal("include_baryon");
}

// number of cells = number of newline-terminated rows of input/surface.dat (Table rule, readindata.cpp:137-146)
long FO_data_reader::get_number_cells()
{
  FILE *f = fopen(path("input/surface.dat").c_str(), "rb");
  if (!f) fatal("Table::loadTableFromFile error: the data file input/surface.dat cannot be opened.");
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  text_.resize(sz + 1);
  if (sz > 0 && fread(text_.data(), 1, sz, f) != (size_t)sz) fatal("short read of input/surface.dat");
  fclose(f);
  text_[sz] = '\0';
  long rows = 0;
  for (long i = 0; i < sz; i++) rows += (text_[i] == '\n');
  number_of_cells = rows;
  return rows;
}

// the readers consume a flat stream of numbers (ifstream >> double), `columns` per cell
std::vector<double> FO_data_reader::slurp(long columns)
{
  if (text_.empty()) get_number_cells();
  std::vector<double> v((size_t)number_of_cells * columns, 0.0);
  const char *p = text_.data();
  for (size_t k = 0; k < v.size(); k++) {
    char *e = nullptr;
    double x = strtod(p, &e);
    if (e == p) break;                                // stream failure: remaining values stay 0
    v[k] = x;
    p = e;
  }
  return v;
}

void FO_data_reader::read_freezeout_surface(FO_surface &surf)
{
  if (mode == 1 || mode == 5) read_surface_cpu_vh(surf);
  else if (mode == 6) read_surface_music(surf);
  else if (mode == 7) read_surface_hic_eventgen(surf);
}

// mode 1: t x y n ds_t ds_x ds_y ds_n u^x u^y u^n E T P pi^xx pi^xy pi^xn pi^yy pi^yn Pi [muB nB V^x V^y V^n] [6 wbar]
void FO_data_reader::read_surface_cpu_vh(FO_surface &s)
{
  const long ncol = 20 + (include_baryon ? 5 : 0) + (mode == 5 ? 6 : 0);
  std::vector<double> v = slurp(ncol);
  const long n = number_of_cells;
  s.resize(n, mode == 5);
  for (long i = 0; i < n; i++) {
    const double *r = &v[(size_t)i * ncol];
    for (int k = 0; k <= IS3D_COL_UN; k++) s.col[k][i] = r[k];          // tau..un unchanged
    s.col[IS3D_COL_E][i] = r[11] * hbarC;
    s.col[IS3D_COL_T][i] = r[12] * hbarC;
    s.col[IS3D_COL_P][i] = r[13] * hbarC;
    s.col[IS3D_COL_PIXX][i] = r[14] * hbarC;
    s.col[IS3D_COL_PIXY][i] = r[15] * hbarC;
    s.col[IS3D_COL_PIXN][i] = r[16] * hbarC;
    s.col[IS3D_COL_PIYY][i] = r[17] * hbarC;
    s.col[IS3D_COL_PIYN][i] = r[18] * hbarC;
    s.col[IS3D_COL_BULKPI][i] = r[19] * hbarC;
    long k = 20;
    if (include_baryon) {
      s.col[IS3D_COL_MUB][i] = r[20] * hbarC;
      s.col[IS3D_COL_NB][i] = r[21];
      s.col[IS3D_COL_VX][i] = r[22];
      s.col[IS3D_COL_VY][i] = r[23];
      s.col[IS3D_COL_VN][i] = r[24];
      k = 25;
    }
    if (mode == 5) for (int w = 0; w < 6; w++) s.vorticity[w][i] = r[k + w];
    if (dimension == 2 && s.col[IS3D_COL_ETA][i] != 0) s.col[IS3D_COL_ETA][i] = 0;   // readindata.cpp:311-319
  }
  double avg[5];
  compute_thermodynamic_averages(s, avg);
  write_thermodynamic_averages(avg);
}

// mode 6 (MUSIC): t x y n ds_t/t ds_x/t ds_y/t ds_n/t u^t u^x u^y t.u^n E T muB muS muC (E+P)/T pi^tt pi^tx pi^ty
// t.pi^tn pi^xx pi^xy t.pi^xn pi^yy t.pi^yn t2.pi^nn Pi [nB V^t V^x V^y t.V^n]
void FO_data_reader::read_surface_music(FO_surface &s)
{
  const long ncol = 29 + (include_baryon ? 5 : 0);
  std::vector<double> v = slurp(ncol);
  const long n = number_of_cells;
  s.resize(n);
  for (long i = 0; i < n; i++) {
    const double *r = &v[(size_t)i * ncol];
    double tau = r[0];
    s.col[IS3D_COL_TAU][i] = tau; s.col[IS3D_COL_X][i] = r[1]; s.col[IS3D_COL_Y][i] = r[2]; s.col[IS3D_COL_ETA][i] = r[3];
    s.col[IS3D_COL_DAT][i] = r[4] * tau; s.col[IS3D_COL_DAX][i] = r[5] * tau;
    s.col[IS3D_COL_DAY][i] = r[6] * tau; s.col[IS3D_COL_DAN][i] = r[7] * tau;
    s.col[IS3D_COL_UX][i] = r[9]; s.col[IS3D_COL_UY][i] = r[10]; s.col[IS3D_COL_UN][i] = r[11] / tau;
    double E = r[12] * hbarC, T = r[13] * hbarC;
    s.col[IS3D_COL_E][i] = E; s.col[IS3D_COL_T][i] = T;
    s.col[IS3D_COL_MUB][i] = r[14] * hbarC;
    s.col[IS3D_COL_P][i] = r[17] * T - E;
    s.col[IS3D_COL_PIXX][i] = r[22] * hbarC;
    s.col[IS3D_COL_PIXY][i] = r[23] * hbarC;
    s.col[IS3D_COL_PIXN][i] = r[24] * hbarC / tau;
    s.col[IS3D_COL_PIYY][i] = r[25] * hbarC;
    s.col[IS3D_COL_PIYN][i] = r[26] * hbarC / tau;
    s.col[IS3D_COL_BULKPI][i] = r[28] * hbarC;
    if (include_baryon) {
      s.col[IS3D_COL_NB][i] = r[29];
      s.col[IS3D_COL_VX][i] = r[31];
      s.col[IS3D_COL_VY][i] = r[32];
      s.col[IS3D_COL_VN][i] = r[33] / tau;
    }
    if (dimension == 2 && s.col[IS3D_COL_ETA][i] != 0) s.col[IS3D_COL_ETA][i] = 0;
  }
  double avg[5];
  compute_thermodynamic_averages(s, avg);
  write_thermodynamic_averages(avg);
}

// mode 7 (HIC-EventGen, 2+1d, GeV units): t x y n ds_t/t ds_x/t ds_y/t ds_n/t v^x v^y t.v^n pi^tt pi^tx pi^ty t.pi^tn
// pi^xx pi^xy t.pi^xn pi^yy t.pi^yn t2.pi^nn Pi T E P muB
void FO_data_reader::read_surface_hic_eventgen(FO_surface &s)
{
  if (dimension != 2) fatal("read_surface_hic_eventgen error: HIC-EventGen surface must be 2+1d (set dimension = 2)");
  if (include_baryon) fatal("read_surface_hic_eventgen error: HIC-EventGen has no baryon chemical potential (set include_baryon = 0)");
  const long ncol = 26;
  std::vector<double> v = slurp(ncol);
  const long n = number_of_cells;
  s.resize(n);
  for (long i = 0; i < n; i++) {
    const double *r = &v[(size_t)i * ncol];
    double tau = r[0];
    s.col[IS3D_COL_TAU][i] = tau; s.col[IS3D_COL_X][i] = r[1]; s.col[IS3D_COL_Y][i] = r[2]; s.col[IS3D_COL_ETA][i] = 0;
    s.col[IS3D_COL_DAT][i] = r[4] * tau; s.col[IS3D_COL_DAX][i] = r[5] * tau; s.col[IS3D_COL_DAY][i] = r[6] * tau;
    s.col[IS3D_COL_DAN][i] = 0;
    double vx = r[8], vy = r[9];
    double ut = 1. / sqrt(fabs(1. - vx * vx - vy * vy));
    s.col[IS3D_COL_UX][i] = ut * vx; s.col[IS3D_COL_UY][i] = ut * vy; s.col[IS3D_COL_UN][i] = 0;
    s.col[IS3D_COL_PIXX][i] = r[15]; s.col[IS3D_COL_PIXY][i] = r[16]; s.col[IS3D_COL_PIXN][i] = 0;
    s.col[IS3D_COL_PIYY][i] = r[18]; s.col[IS3D_COL_PIYN][i] = 0;
    s.col[IS3D_COL_BULKPI][i] = r[21];
    s.col[IS3D_COL_T][i] = r[22]; s.col[IS3D_COL_E][i] = r[23]; s.col[IS3D_COL_P][i] = r[24];
    s.col[IS3D_COL_MUB][i] = r[25];
  }
  // the reference averages with its local ut (from v) and nB = 0; ut from u^x,u^y is the same number up to rounding
  double avg[5];
  compute_thermodynamic_averages(s, avg);
  avg[4] = 0.0;
  write_thermodynamic_averages(avg);
}

}  // namespace is3dhost
