// IS3D driver (reference src/cpp/iS3D.cpp) and the C entry points of include/is3d_host.h.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>

#include "../../include/is3d_host.h"
#include "is3d_host.hpp"

namespace is3dhost {

// everything IS3D::run_particlization builds between reading the parameter file and calling calculate_spectra
struct Session {
  ParameterReader paraRdr;
  FO_surface surf;
  std::vector<particle_info> particle_data;
  std::unique_ptr<Table> chosen_particles, pT_tab, phi_tab, y_tab, eta_tab;
  std::unique_ptr<Deltaf_Data> df_data;
  std::unique_ptr<EmissionFunctionArray> efa;
  std::vector<std::vector<Sampled_Particle>> events;
  bool have_surface = false, tables_ready = false;
  // contexts created ahead of time (executable path: while surface.dat is parsed)
  std::thread early;
  is3d_group *early_group = nullptr;
  is3d_params early_params{};
  std::string early_error;                  // reported where the contexts are needed, in the order the reference reports errors
  void start_contexts()
  {
    early_params = EmissionFunctionArray::params_from(&paraRdr);
    early = std::thread([this] { early_group = EmissionFunctionArray::create_group(early_params, &early_error); });
  }
  ~Session() { if (early.joinable()) early.join(); if (early_group && !efa) is3d_group_destroy(early_group); }

  void open(const char *const *overrides)
  {
    printf("\n\nReading in parameters...\n\n");
    paraRdr.readFromFile("iS3D_parameters.dat");
    if (overrides)
      for (int i = 0; overrides[i]; i++) {
        char *fake[2] = {nullptr, const_cast<char *>(overrides[i])};
        paraRdr.readFromArguments(2, fake);
      }
  }
  long read_surface()
  {
    printf("\n\nReading in freezeout surface from input/surface.dat\n");
    FO_data_reader reader(&paraRdr, "input");
    long n = reader.get_number_cells();
    reader.read_freezeout_surface(surf);
    have_surface = true;
    printf("Number of freezeout cells = %ld\n\n", n);
    return n;
  }
  // in-memory branch of run_particlization (iS3D.cpp:126-220): averages computed here and written to the side file
  long set_surface(int64_t n, const double *const cols[IS3D_SURFACE_COLUMNS])
  {
    surf.resize(n);
    for (int k = 0; k < IS3D_SURFACE_COLUMNS; k++)
      if (cols[k]) surf.col[k].assign(cols[k], cols[k] + n);
    double avg[5];
    compute_thermodynamic_averages(surf, avg);
    write_thermodynamic_averages(avg);
    have_surface = true;
    return (long)n;
  }
  void prepare_tables()
  {
    int include_baryon = paraRdr.getVal("include_baryon");
    printf("\n\nReading in particle info from PDG/...\n");
    PDG_Data pdg(&paraRdr);
    pdg.read_resonances(particle_data);
    chosen_particles.reset(new Table("PDG/chosen_particles.dat"));
    printf("Number of chosen particles = %ld\n", chosen_particles->getNumberOfRows());
    df_data.reset(new Deltaf_Data(&paraRdr));
    df_data->load_df_coefficient_data();
    if (!include_baryon) df_data->compute_jonah_coefficients(particle_data);
    df_data->compute_particle_densities(particle_data);
    pT_tab.reset(new Table("tables/momentum/pT_table.dat"));
    phi_tab.reset(new Table("tables/momentum/phi_table.dat"));
    y_tab.reset(new Table("tables/momentum/y_table.dat"));
    eta_tab.reset(new Table("tables/spacetime_rapidity/eta_table.dat"));
    tables_ready = true;
  }
  void create_context()
  {
    if (!tables_ready) prepare_tables();
    if (early.joinable()) early.join();
    if (!early_error.empty()) fatal(early_error);
    efa.reset(new EmissionFunctionArray(&paraRdr, chosen_particles.get(), pT_tab.get(), phi_tab.get(), y_tab.get(),
                                        eta_tab.get(), &particle_data, have_surface ? &surf : nullptr, df_data.get(),
                                        early_group, early_group ? &early_params : nullptr));
  }
  void run() { efa->calculate_spectra(events); }
};

void IS3D::read_fo_surf_from_memory(std::vector<double> tau_in, std::vector<double> x_in, std::vector<double> y_in,
                                    std::vector<double> eta_in, std::vector<double> dsigma_tau_in,
                                    std::vector<double> dsigma_x_in, std::vector<double> dsigma_y_in,
                                    std::vector<double> dsigma_eta_in, std::vector<double> E_in, std::vector<double> T_in,
                                    std::vector<double> P_in, std::vector<double> ux_in, std::vector<double> uy_in,
                                    std::vector<double> un_in, std::vector<double> pixx_in, std::vector<double> pixy_in,
                                    std::vector<double> pixn_in, std::vector<double> piyy_in, std::vector<double> piyn_in,
                                    std::vector<double> pinn_in, std::vector<double> Pi_in)
{
  tau = std::move(tau_in); x = std::move(x_in); y = std::move(y_in); eta = std::move(eta_in);
  dsigma_tau = std::move(dsigma_tau_in); dsigma_x = std::move(dsigma_x_in); dsigma_y = std::move(dsigma_y_in);
  dsigma_eta = std::move(dsigma_eta_in);
  E = std::move(E_in); T = std::move(T_in); P = std::move(P_in);
  ux = std::move(ux_in); uy = std::move(uy_in); un = std::move(un_in);
  pixx = std::move(pixx_in); pixy = std::move(pixy_in); pixn = std::move(pixn_in); piyy = std::move(piyy_in);
  piyn = std::move(piyn_in); pinn = std::move(pinn_in);   // pinn is extraneous (recomputed on the device)
  Pi = std::move(Pi_in);
}

void IS3D::run_particlization(int fo_from_file)
{
  printf("\n::::::::::::::::::::::::::::::::::::::::\n");
  printf("::                                    ::\n");
  printf("::    Starting iS3D particlization    ::\n");
  printf("::          (B200-native path)        ::\n");
  printf("::::::::::::::::::::::::::::::::::::::::\n\n");
  // IS3D_TIMING=1: wall time of each phase of the run (no counterpart in the reference, which times calculate_spectra only)
  const bool timing = getenv("IS3D_TIMING") != nullptr;
  auto t_phase = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    auto t1 = std::chrono::steady_clock::now();
    if (timing) printf("[timing] %-44s %9.3f s\n", what, std::chrono::duration<double>(t1 - t_phase).count());
    t_phase = t1;
  };
  Session s;
  s.open(nullptr);
  s.start_contexts();          // CUDA contexts (+ the NCCL communicator of a multi-GPU run) come up while the surface is parsed
  if (fo_from_file == 1) { s.read_surface(); lap("surface.dat -> structure of arrays"); }
  else {
    printf("from memory (please check that you've already undone hbarc = 1 units, tau factors from hydro module)...\n\n");
    const double *cols[IS3D_SURFACE_COLUMNS] = {tau.data(), x.data(), y.data(), eta.data(), dsigma_tau.data(), dsigma_x.data(),
                                                dsigma_y.data(), dsigma_eta.data(), ux.data(), uy.data(), un.data(), E.data(),
                                                T.data(), P.data(), pixx.data(), pixy.data(), pixn.data(), piyy.data(),
                                                piyn.data(), Pi.data(), nullptr, nullptr, nullptr, nullptr, nullptr};
    s.set_surface((int64_t)tau.size(), cols);
    printf("Number of freezeout cells = %ld\n\n", (long)tau.size());
  }
  s.prepare_tables();
  lap("PDG / df tables / densities");
  s.create_context();
  lap("wait for CUDA contexts + tables + surface H2D");
  s.run();
  lap("calculate_spectra (compute + result files)");
  int operation = s.paraRdr.getVal("operation");
  if (operation == 2) {
    printf("\nCopying final particle list to memory (JETSCAPE)\n");
    printf("Event particle list contains %ld events\n", (long)s.events.size());
    final_particles_ = s.events;
  }
}

}  // namespace is3dhost

using namespace is3dhost;

struct is3d_host {
  std::string root;
  Session s;
};

// The reference works with fixed relative paths; here they resolve against ONE process-wide root (io.cpp).  Every entry
// point installs its session's root first, so several live sessions can be used in turn (not concurrently: like the
// reference's EmissionFunctionArray, the host layer is single-caller).
#define ENTER(h) set_root((h)->root)

extern "C" {

is3d_host *is3d_host_open(const char *root, const char *const *overrides)
{
  is3d_host *h = new is3d_host;
  h->root = root ? root : "";
  ENTER(h);
  h->s.open(overrides);
  return h;
}
void is3d_host_close(is3d_host *h) { delete h; }
int64_t is3d_host_read_surface(is3d_host *h) { ENTER(h); return h->s.read_surface(); }
int64_t is3d_host_set_surface(is3d_host *h, int64_t n, const double *const cols[IS3D_SURFACE_COLUMNS]) { ENTER(h); return h->s.set_surface(n, cols); }
void is3d_host_thermo_sums(is3d_host *h, double sums6[6]) { compute_thermodynamic_sums(h->s.surf, sums6); }
void is3d_host_set_thermo_averages(is3d_host *h, const double avg5[5]) { ENTER(h); write_thermodynamic_averages(avg5); }
void is3d_host_prepare_tables(is3d_host *h) { ENTER(h); h->s.prepare_tables(); }
void is3d_host_prepare(is3d_host *h) { ENTER(h); h->s.create_context(); }
is3d_ctx *is3d_host_context(is3d_host *h) { return h->s.efa ? h->s.efa->context() : nullptr; }
is3d_group *is3d_host_group(is3d_host *h) { return h->s.efa ? h->s.efa->group() : nullptr; }
void is3d_host_run(is3d_host *h) { ENTER(h); h->s.run(); }
int64_t is3d_host_spectra(is3d_host *h, const double **data, int64_t dims[4])
{
  EmissionFunctionArray *e = h->s.efa.get();
  if (!e) return 0;
  *data = e->dN_pTdpTdphidy.data();
  dims[0] = e->number_of_chosen_particles; dims[1] = e->pT_tab_length; dims[2] = e->phi_tab_length; dims[3] = e->y_tab_length;
  return (int64_t)e->dN_pTdpTdphidy.size();
}
int64_t is3d_host_dndx(is3d_host *h, const double **tau, const double **r, const double **phi)
{
  EmissionFunctionArray *e = h->s.efa.get();
  if (!e) return 0;
  *tau = e->dN_taudtaudy.data(); *r = e->dN_twopirdrdy.data(); *phi = e->dN_dphisdy.data();
  return e->number_of_chosen_particles;
}
int64_t is3d_host_events(is3d_host *h) { return (int64_t)h->s.events.size(); }
int64_t is3d_host_event_particles(is3d_host *h, int64_t event, double *out)
{
  const auto &ev = h->s.events[event];
  if (out)
    for (size_t i = 0; i < ev.size(); i++) {
      const Sampled_Particle &p = ev[i];
      double rec[13] = {(double)p.chosen_index, (double)p.mcID, p.mass, p.tau, p.x, p.y, p.eta, p.t, p.z, p.E, p.px, p.py, p.pz};
      memcpy(out + 13 * i, rec, sizeof(rec));
    }
  return (int64_t)ev.size();
}
double is3d_host_seconds(is3d_host *h) { return h->s.efa ? h->s.efa->seconds_compute : 0.0; }
void is3d_host_stats(is3d_host *h, is3d_stats *out) { if (h->s.efa) *out = h->s.efa->stats; }
int64_t is3d_host_pdg(is3d_host *h, double *out)
{
  const auto &v = h->s.particle_data;
  if (out)
    for (size_t i = 0; i < v.size(); i++) {
      double rec[8] = {(double)v[i].mc_id, v[i].mass, (double)v[i].gspin, (double)v[i].baryon, (double)v[i].sign,
                       v[i].equilibrium_density, v[i].bulk_density, v[i].diff_density};
      memcpy(out + 8 * i, rec, sizeof(rec));
    }
  return (int64_t)v.size();
}
int64_t is3d_host_ptb(is3d_host *h, double *x, double *l2, double *z, double *xmax)
{
  Deltaf_Data *d = h->s.df_data.get();
  if (!d || !d->have_jonah) return 0;
  const int n = Deltaf_Data::jonah_points;
  if (x) memcpy(x, d->bulkPi_over_Peq_array.data(), n * sizeof(double));
  if (l2) memcpy(l2, d->lambda_squared_array.data(), n * sizeof(double));
  if (z) memcpy(z, d->z_array.data(), n * sizeof(double));
  if (xmax) *xmax = d->bulkPi_over_Peq_max;
  return n;
}
int64_t is3d_host_surface_column(is3d_host *h, int k, const double **data)
{
  *data = h->s.surf.col[k].data();
  return h->s.surf.size();
}
int64_t is3d_host_chosen(is3d_host *h, int *mcid)
{
  EmissionFunctionArray *e = h->s.efa.get();
  if (!e) return 0;
  if (mcid) for (size_t i = 0; i < e->MCID.size(); i++) mcid[i] = e->MCID[i];
  return (int64_t)e->MCID.size();
}

}  // extern "C"
