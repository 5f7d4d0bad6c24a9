// EmissionFunctionArray: owns the run-time switches, marshals species / tables / surface into one CUDA context and
// forwards the compute members to the C ABI; writes the results/ tree.  Mirrors reference
// src/cpp/EmissionFunction.cpp (ctor :114-391, calculate_spectra :981-1386, writers :406-975).
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>

#include <charconv>
#include <string>

#include "host_parallel.hpp"
#include "is3d_host.hpp"

namespace is3dhost {

void EmissionFunctionArray::check(is3d_status st, const char *what)
{
  if (st == IS3D_OK) return;
  // the reference's convention: print and exit(-1) (GSL domain errors abort)
  printf("%s error: %s\n", what, grp && *is3d_group_last_error(grp) ? is3d_group_last_error(grp) : is3d_last_error(ctx));
  fflush(stdout);
  if (st == IS3D_ERR_TABLE_RANGE && !prm.include_baryon) abort();
  exit(-1);
}

is3d_params EmissionFunctionArray::params_from(ParameterReader *paraRdr, int *polzn_file_compat_out)
{
  is3d_params prm;
  is3d_default_params(&prm);
  prm.operation = paraRdr->getVal("operation");
  prm.dimension = paraRdr->getVal("dimension");
  if (prm.dimension != 2 && prm.dimension != 3) fatal("EmissionFunctionArray error: need to set dimension = (2,3)");
  prm.df_mode = paraRdr->getVal("df_mode");
  if (prm.df_mode < 1 || prm.df_mode > 5) fatal("EmissionFunctionArray error: need to set df_mode = (1,2,3,4,5)");
  prm.include_baryon = paraRdr->getVal("include_baryon");
  prm.include_bulk_deltaf = paraRdr->getVal("include_bulk_deltaf");
  prm.include_shear_deltaf = paraRdr->getVal("include_shear_deltaf");
  prm.include_baryondiff_deltaf = paraRdr->getVal("include_baryondiff_deltaf");
  prm.regulate_deltaf = paraRdr->getVal("regulate_deltaf");
  prm.outflow = paraRdr->getVal("outflow");
  prm.deta_min = paraRdr->getVal("deta_min");
  prm.mass_pion0 = paraRdr->getVal("mass_pion0");
  prm.fast = paraRdr->getVal("fast");
  long seed = (long)paraRdr->getVal("sampler_seed");
  if (seed < 0) seed = (long)std::chrono::system_clock::now().time_since_epoch().count();   // "< 0 => clock"
  prm.sampler_seed = seed;
  prm.test_sampler = paraRdr->getVal("test_sampler");
  prm.pT_min = paraRdr->getVal("pT_min"); prm.pT_max = paraRdr->getVal("pT_max"); prm.pT_bins = paraRdr->getVal("pT_bins");
  prm.y_cut = paraRdr->getVal("y_cut"); prm.y_bins = paraRdr->getVal("y_bins");
  prm.phip_bins = paraRdr->getVal("phip_bins");
  prm.eta_cut = paraRdr->getVal("eta_cut"); prm.eta_bins = paraRdr->getVal("eta_bins");
  prm.tau_min = paraRdr->getVal("tau_min"); prm.tau_max = paraRdr->getVal("tau_max"); prm.tau_bins = paraRdr->getVal("tau_bins");
  prm.r_min = paraRdr->getVal("r_min"); prm.r_max = paraRdr->getVal("r_max"); prm.r_bins = paraRdr->getVal("r_bins");
  if (const char *dev = getenv("IS3D_DEVICE")) prm.device = atoi(dev);
  if (const char *v = getenv("IS3D_FAMOD_CHAIN")) prm.famod_chain = atoi(v);
  if (const char *v = getenv("IS3D_DNDX_BUG_COMPAT")) prm.dndx_bug_compat = atoi(v);
  if (const char *v = getenv("IS3D_POLZN_CHUNK_COMPAT")) prm.polzn_chunk_compat = atoi(v);
  if (const char *v = getenv("IS3D_NEGLIGIBLE_MARGIN")) prm.negligible_margin = atof(v);
  if (polzn_file_compat_out)
    if (const char *v = getenv("IS3D_POLZN_FILE_COMPAT")) *polzn_file_compat_out = atoi(v);
  return prm;
}

// one CUDA context per GPU of the run: IS3D_DEVICES = "all" | "0-7" | "0,2,3" (default: the single device IS3D_DEVICE, else 0)
is3d_group *EmissionFunctionArray::create_group(const is3d_params &prm, std::string *error)
{
  std::vector<int> devices = parse_device_list(getenv("IS3D_DEVICES"), prm.device);
  is3d_group *g = nullptr;
  is3d_status st = is3d_group_create(&prm, (int)devices.size(), devices.data(), &g);
  if (st != IS3D_OK) {
    const std::string msg = std::string("is3d_create error: ") + is3d_group_last_error(nullptr);
    if (!error) fatal(msg);
    *error = msg;
    return nullptr;
  }
  return g;
}

EmissionFunctionArray::EmissionFunctionArray(ParameterReader *paraRdr_in, Table *chosen_particles_in, Table *pT_tab_in,
                                             Table *phi_tab_in, Table *y_tab_in, Table *eta_tab_in,
                                             std::vector<particle_info> *particles_in, FO_surface *surf_in,
                                             Deltaf_Data *df_data_in, is3d_group *ready_group, const is3d_params *ready_params)
{
  paraRdr = paraRdr_in;
  pT_tab = pT_tab_in; phi_tab = phi_tab_in; y_tab = y_tab_in; eta_tab = eta_tab_in;
  particles = particles_in; surf = surf_in; df_data = df_data_in;
  pT_tab_length = pT_tab->getNumberOfRows();
  phi_tab_length = phi_tab->getNumberOfRows();
  y_tab_length = y_tab->getNumberOfRows();
  eta_tab_length = eta_tab->getNumberOfRows();

  prm = params_from(paraRdr, &polzn_file_compat);
  if (ready_group && ready_params) prm = *ready_params;       // the switches the ready contexts were created with (clock seeds!)
  OPERATION = prm.operation; DIMENSION = prm.dimension; DF_MODE = prm.df_mode; TEST_SAMPLER = prm.test_sampler;
  MODE = paraRdr->getVal("mode");
  if (DIMENSION == 2) y_tab_length = 1;
  else eta_tab_length = 1;
  const int GROUP_PARTICLES = paraRdr->getVal("group_particles");
  (void)paraRdr->getVal("particle_diff_tolerance");
  (void)paraRdr->getVal("lightest_particle");
  (void)paraRdr->getVal("do_resonance_decays");
  OVERSAMPLE = paraRdr->getVal("oversample");
  MIN_NUM_HADRONS = paraRdr->getVal("min_num_hadrons");
  MAX_NUM_SAMPLES = paraRdr->getVal("max_num_samples");
  if (OPERATION == 2) printf("Sampler seed set to %ld \n", (long)prm.sampler_seed);

  // chosen species in file order, matched by Monte-Carlo id (EmissionFunction.cpp:357-372); the optional mass sort
  // of group_particles only affected the (removed) resonance-decay code
  number_of_chosen_particles = chosen_particles_in->getNumberOfRows();
  for (int m = 0; m < number_of_chosen_particles; m++) {
    long mc_id = (long)chosen_particles_in->get(1, m + 1);
    for (size_t n = 0; n < particles->size(); n++)
      if ((*particles)[n].mc_id == mc_id) { chosen_particles_sampling_table.push_back((int)n); break; }
  }
  if ((int)chosen_particles_sampling_table.size() != number_of_chosen_particles)
    fatal("EmissionFunctionArray error: a chosen particle is not in the PDG table");
  if (GROUP_PARTICLES == 1) {   // adjacent-swap sort by mass (EmissionFunction.cpp:375-390): sets the species order
    auto &tab = chosen_particles_sampling_table;
    for (int m = 0; m < number_of_chosen_particles; m++)
      for (int n = 0; n < number_of_chosen_particles - m - 1; n++)
        if ((*particles)[tab[n]].mass > (*particles)[tab[n + 1]].mass) std::swap(tab[n], tab[n + 1]);
  }

  dN_pTdpTdphidy.assign((size_t)number_of_chosen_particles * pT_tab_length * phi_tab_length * y_tab_length, 0.0);

  // ---- one CUDA context per GPU, each holding every static input; cells are sharded over them ----
  grp = ready_group ? ready_group : create_group(prm);
  if (is3d_group_size(grp) > 1) printf("Sharding the freezeout surface over %d GPUs (one NCCL all-reduce per result)\n", is3d_group_size(grp));
  ctx = is3d_group_ctx(grp, 0);

  const int ns = number_of_chosen_particles;
  std::vector<double> Mass(ns), Sign(ns), Degeneracy(ns), Baryon(ns), Neq(ns), Dnb(ns), Dnd(ns);
  MCID.assign(ns, 0);
  for (int i = 0; i < ns; i++) {
    const particle_info &p = (*particles)[chosen_particles_sampling_table[i]];
    Mass[i] = p.mass; Sign[i] = p.sign; Degeneracy[i] = p.gspin; Baryon[i] = p.baryon; MCID[i] = (int)p.mc_id;
    Neq[i] = p.equilibrium_density; Dnb[i] = p.bulk_density; Dnd[i] = p.diff_density;
  }
  Gauss_Laguerre gla;
  Gauss_Legendre legendre;
  gla.load_roots_and_weights("tables/gauss/gla_roots_weights.txt");
  legendre.load_roots_and_weights("tables/gauss/gauss_legendre.dat");
  Plasma QGP;
  QGP.load_thermodynamic_averages();
  const int np = (int)particles->size();
  std::vector<double> Mp(np), Sp(np), Dp(np), Bp(np);
  for (int i = 0; i < np; i++) { const particle_info &p = (*particles)[i]; Mp[i] = p.mass; Sp[i] = p.sign; Dp[i] = p.gspin; Bp[i] = p.baryon; }
  for (int d = 0; d < is3d_group_size(grp); d++) {
    is3d_ctx *c = ctx = is3d_group_ctx(grp, d);      // `check` reports the failing context's message
    check(is3d_set_species(c, ns, Mass.data(), Sign.data(), Degeneracy.data(), Baryon.data(), MCID.data(), Neq.data(),
                           Dnb.data(), Dnd.data()), "is3d_set_species");
    check(is3d_set_pdg(c, np, Mp.data(), Sp.data(), Dp.data(), Bp.data()), "is3d_set_pdg");
    check(is3d_set_momentum_tables(c, (int)pT_tab->getNumberOfRows(), pT_tab->column(1).data(), pT_tab->column(2).data(),
                                   (int)phi_tab->getNumberOfRows(), phi_tab->column(1).data(), phi_tab->column(2).data(),
                                   (int)y_tab->getNumberOfRows(), y_tab->column(1).data(), y_tab->column(2).data(),
                                   (int)eta_tab->getNumberOfRows(), eta_tab->column(1).data(), eta_tab->column(2).data()),
          "is3d_set_momentum_tables");
    check(is3d_set_gauss_tables(c, gla.alpha, gla.points, gla.root.data(), gla.weight.data(), legendre.points,
                                legendre.root.data(), legendre.weight.data()), "is3d_set_gauss_tables");
    check(is3d_set_thermo_averages(c, QGP.temperature, QGP.energy_density, QGP.pressure, QGP.baryon_chemical_potential,
                                   QGP.net_baryon_density), "is3d_set_thermo_averages");
    check(is3d_set_df_tables(c, df_data->points_T, df_data->points_muB, df_data->T_array.data(), df_data->muB_array.data(),
                             df_data->tab[0].data(), df_data->tab[1].data(), df_data->tab[2].data(), df_data->tab[3].data(),
                             df_data->tab[4].data(), df_data->tab[5].data(), df_data->tab[6].data(), df_data->tab[7].data(),
                             df_data->tab[8].data(), df_data->tab[9].data()), "is3d_set_df_tables");
    if (df_data->have_jonah)
      check(is3d_set_ptb_tables(c, Deltaf_Data::jonah_points, df_data->bulkPi_over_Peq_array.data(),
                                df_data->lambda_squared_array.data(), df_data->z_array.data(), df_data->bulkPi_over_Peq_max),
            "is3d_set_ptb_tables");
  }
  ctx = is3d_group_ctx(grp, 0);
  if (surf) set_surface_on_device();
}

EmissionFunctionArray::~EmissionFunctionArray() { is3d_group_destroy(grp); }

void EmissionFunctionArray::set_surface_on_device()
{
  const double *cols[IS3D_SURFACE_COLUMNS];
  for (int k = 0; k < IS3D_SURFACE_COLUMNS; k++) cols[k] = surf->col[k].data();
  check(is3d_group_set_surface(grp, surf->size(), cols, 0), "is3d_set_surface");
}

void EmissionFunctionArray::calculate_dN_pTdpTdphidy()
{
  check(is3d_group_spectra(grp, dN_pTdpTdphidy.data(), &stats), "calculate_dN_pTdpTdphidy");
  if (DF_MODE == 3 || DF_MODE == 4) {
    printf("\nfeqmod breaks down for %ld / %ld cells until t = %.3f fm/c\n", (long)stats.cells_breakdown, (long)stats.cells_total, stats.tau_breakdown);
    printf("pl went negative for %ld / %ld cells until t = %.3f fm/c\n\n", (long)stats.cells_pl_negative, (long)stats.cells_total, stats.tau_pl_negative);
  } else if (DF_MODE == 5) {
    printf("\nfamod breaks down for %ld / %ld cells until t = %.3f fm/c\n", (long)stats.cells_breakdown, (long)stats.cells_total, stats.tau_breakdown);
    printf("pl went negative for %ld / %ld cells until t = %.3f fm/c\n\n", (long)stats.cells_pl_negative, (long)stats.cells_total, stats.tau_pl_negative);
    printf("Number of reconstruction failures = %ld\n", (long)stats.reconstruction_failures);
    printf("Average number of iterations = %lf\n\n", (double)stats.newton_iterations / (double)stats.cells_total);
  }
}

void EmissionFunctionArray::calculate_dN_dX()
{
  const int ns = number_of_chosen_particles;
  dN_taudtaudy.assign((size_t)ns * prm.tau_bins, 0.0);
  dN_twopirdrdy.assign((size_t)ns * prm.r_bins, 0.0);
  dN_dphisdy.assign((size_t)ns * prm.phip_bins, 0.0);
  check(is3d_group_dndx(grp, dN_taudtaudy.data(), dN_twopirdrdy.data(), dN_dphisdy.data(), &stats), "calculate_dN_dX");
  // dndx_bug_compat (the reference's partial memset) is applied by the library itself (is3d_dndx)
}

double EmissionFunctionArray::calculate_total_yield()
{
  double ntot = 0.0;
  check(is3d_group_total_yield(grp, &ntot, &stats), "calculate_total_yield");
  return ntot;
}

void EmissionFunctionArray::sample_dN_pTdpTdphidy()
{
  // 64-byte wire records over PCIe (the list is transfer-bound); the Sampled_Particle fields that are functions of the
  // others -- mass, mcID, E, t, z -- are restored here while the per-event vectors are filled (one host thread per event range)
  is3d_particle_compact *plist = nullptr;
  int64_t total = 0;
  std::vector<int64_t> counts(Nevents, 0);
  check(is3d_group_sample_compact(grp, Nevents, &plist, &total, counts.data(), &stats), "sample_dN_pTdpTdphidy");
  particle_event_list.assign(Nevents, {});
  std::vector<int64_t> first(Nevents + 1, 0);
  for (long e = 0; e < Nevents; e++) first[e + 1] = first[e] + counts[e];
  const int nt = host_threads((size_t)Nevents, "IS3D_WRITER_THREADS");
  parallel_for(nt, [&](int t) {
    for (int64_t e = (int64_t)Nevents * t / nt; e < (int64_t)Nevents * (t + 1) / nt; e++) {
      auto &ev = particle_event_list[e];
      ev.resize((size_t)counts[e]);
      for (int64_t i = 0; i < counts[e]; i++) {
        const is3d_particle_compact &q = plist[first[e] + i];
        const particle_info &info = (*particles)[chosen_particles_sampling_table[q.chosen_index]];
        Sampled_Particle &p = ev[(size_t)i];
        p.chosen_index = q.chosen_index; p.mcID = (int)info.mc_id; p.mass = info.mass;
        p.tau = q.tau; p.x = q.x; p.y = q.y; p.eta = q.eta;
        const double sh = sinh(q.eta);
        p.t = q.tau * sqrt(1.0 + sh * sh); p.z = q.tau * sh;
        p.px = q.px; p.py = q.py; p.pz = q.pz;
        p.E = sqrt(info.mass * info.mass + q.px * q.px + q.py * q.py + q.pz * q.pz);
      }
    }
  });
  is3d_free_particles(plist);
  if (stats.sampler_proposals > 0)
    printf("\nMomentum sampling efficiency = %f %%\n", 100.0 * (double)stats.sampler_accepted / (double)stats.sampler_proposals);
}

// EmissionFunction.cpp:981-1386
void EmissionFunctionArray::calculate_spectra(std::vector<std::vector<Sampled_Particle>> &particle_event_list_in)
{
  static const char *names[6] = {"", "Grad 14-moment approximation", "RTA Chapman-Enskog expansion",
                                 "PTM modified equilibrium distribution", "PTB modified equilibrium distribution",
                                 "PTM modified anisotropic distribution"};
  printf("\n\nRunning particlization with %s\n\n", names[DF_MODE]);
  auto t0 = std::chrono::steady_clock::now();
  switch (OPERATION) {
    case 0: {
      printf("\nComputing particle spacetime distributions...\n\n");
      if (DF_MODE == 5) fatal("calculate_spectra error: no spacetime distribution routine for famod yet");
      calculate_dN_dX();
      write_dN_dX_toFile();
      break;
    }
    case 1: {
      printf("\nComputing continuous momentum spectra...\n\n");
      calculate_dN_pTdpTdphidy();
      if (getenv("IS3D_TIMING"))
        printf("[timing] %-44s %9.3f s (dominant kernels %.3f s on the slowest GPU)\n", "calculate_dN_pTdpTdphidy (GPU, all-reduce, D2H)",
               std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), stats.kernel_ms * 1e-3);
      write_dN_pTdpTdphidy_toFile();
      write_continuous_vn_toFile();
      write_dN_twopipTdpTdy_toFile();
      write_dN_dphidy_toFile();
      write_dN_dy_toFile();
      break;
    }
    case 2: {
      if (OVERSAMPLE) {
        double Ntotal = calculate_total_yield();
        Nevents = (long)fmin(ceil(MIN_NUM_HADRONS / Ntotal), MAX_NUM_SAMPLES);
        printf("\nSampling %ld particlization events...\n\n", Nevents);
      } else {
        printf("\nSampling 1 particlization event...\n\n");
      }
      sample_dN_pTdpTdphidy();
      if (TEST_SAMPLER) write_sampled_tests_to_file();
      else write_particle_list_OSC();
      particle_event_list_in = particle_event_list;
      break;
    }
    default: fatal("calculate_spectra error: need to set operation = (0, 1, 2)");
  }
  if (MODE == 5) {                       // EmissionFunction.cpp:1304-1310
    printf("\nComputing spin polarization...\n");
    calculate_spin_polzn();
    write_polzn_vector_toFile();
  }
  seconds_compute = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("\nSpectra calculation took %g seconds\n\n", seconds_compute);
}

#define IDX(iy, iphi, ipT, is) ((iy) + y_tab_length * ((iphi) + phi_tab_length * ((ipT) + pT_tab_length * (long)(is))))

static FILE *open_result(const char *fmt, int id, const char *mode);

// calculate_spin_polzn (Polarization.cpp:25-263) through the C ABI; arrays come back in the spectra layout
void EmissionFunctionArray::calculate_spin_polzn()
{
  if (surf->vorticity[0].size() != (size_t)surf->size()) fatal("calculate_spin_polzn error: the surface carries no thermal vorticity (mode 5 file needed)");
  const double *w[6];
  for (int k = 0; k < 6; k++) w[k] = surf->vorticity[k].data();
  check(is3d_group_set_vorticity(grp, surf->size(), w), "is3d_set_vorticity");
  const size_t total = (size_t)is3d_spectra_size(ctx);
  for (auto *v : {&St, &Sx, &Sy, &Sn, &Snorm}) v->assign(total, 0.0);
  is3d_stats pst;
  check(is3d_group_polarization(grp, St.data(), Sx.data(), Sy.data(), Sn.data(), Snorm.data(), &pst), "calculate_spin_polzn");
}

// write_polzn_vector_toFile (EmissionFunction.cpp:561-609): results/{St,Sx,Sy,Sn}.dat, rows `y phip pT S/Snorm`, loop order
// species -> y -> phi -> pT, blank line after each phi block.  The reference ACCUMULATES the arrays species-fastest
// (iS3D = ipart + npart (ipT + NpT (iphip + Nphi iy)), Polarization.cpp:226) but READS them here with the spectra index
// (iy + Ny (iphip + Nphi (ipT + NpT ipart)), :591), so the value printed next to (species, y, phi, pT) belongs to another
// bin.  polzn_file_compat = 1 (default; IS3D_POLZN_FILE_COMPAT=0 switches it off) reproduces the files as the reference
// writes them; 0 prints each bin's own value.
void EmissionFunctionArray::write_polzn_vector_toFile()
{
  printf("Writing polarization vector to file...\n");
  FILE *f[4] = {open_result("results/St.dat", 0, "w"), open_result("results/Sx.dat", 0, "w"), open_result("results/Sy.dat", 0, "w"),
                open_result("results/Sn.dat", 0, "w")};
  const std::vector<double> *S[4] = {&St, &Sx, &Sy, &Sn};
  const long np = number_of_chosen_particles, npT = pT_tab_length, nphi = phi_tab_length, ny = y_tab_length;
  for (long ipart = 0; ipart < np; ipart++)
    for (long iy = 0; iy < ny; iy++) {
      double y = (DIMENSION == 3) ? y_tab->get(1, iy + 1) : 0.0;
      for (long iphip = 0; iphip < nphi; iphip++) {
        double phip = phi_tab->get(1, iphip + 1);
        for (long ipT = 0; ipT < npT; ipT++) {
          long j = IDX(iy, iphip, ipT, ipart);               // the index the reference reads with
          long src = j;
          if (polzn_file_compat) {
            // decode j in the storage order of the accumulation, then address that bin in our (spectra-layout) arrays
            long sp = j % np, r = j / np, p2 = r % npT, r2 = r / npT, ph = r2 % nphi, yy = r2 / nphi;
            src = IDX(yy, ph, p2, sp);
          }
          for (int k = 0; k < 4; k++)
            fprintf(f[k], "%.8e\t%.8e\t%.8e\t%.8e\n", y, phip, pT_tab->get(1, ipT + 1), (*S[k])[src] / Snorm[src]);
        }
        for (int k = 0; k < 4; k++) fprintf(f[k], "\n");
      }
    }
  for (int k = 0; k < 4; k++) fclose(f[k]);
}

// ---- writers ---------------------------------------------------------------------------------------------------
// one result file set per species: the species are independent, so the writers run on all host threads
// (IS3D_WRITER_THREADS overrides); every file's content is what the serial loop writes
template <class Fn>
static void parallel_species(int ns, Fn fn)
{
  const int nt = host_threads((size_t)(ns > 0 ? ns : 1), "IS3D_WRITER_THREADS");
  parallel_for(nt, [&](int t) { for (int is = t; is < ns; is += nt) fn(is); });
}

static FILE *open_result(const char *fmt, int id, const char *mode = "w")
{
  char name[300];
  snprintf(name, sizeof(name), fmt, id);
  FILE *f = fopen(path(name).c_str(), mode);
  if (!f) fatal(std::string("cannot open ") + name + " (the results/ tree must exist, clear_results.sh)");
  return f;
}


void EmissionFunctionArray::write_dN_pTdpTdphidy_toFile()
{
  printf("Writing thermal spectra to file...\n");
  parallel_species(number_of_chosen_particles, [&](int is) {
    FILE *f = open_result("results/continuous/dN_pTdpTdphidy_%d.dat", MCID[is]);
    fprintf(f, "y\tphip\tpT\tdN_pTdpTdphidy\n");
    for (long iy = 0; iy < y_tab_length; iy++) {
      double y = (DIMENSION == 3) ? y_tab->get(1, iy + 1) : 0.0;
      for (long iphi = 0; iphi < phi_tab_length; iphi++) {
        double phip = phi_tab->get(1, iphi + 1);
        for (long ipT = 0; ipT < pT_tab_length; ipT++)
          fprintf(f, "%.8e\t%.8e\t%.8e\t%.8e\n", y, phip, pT_tab->get(1, ipT + 1), dN_pTdpTdphidy[IDX(iy, iphi, ipT, is)]);
        fprintf(f, "\n");
      }
    }
    fclose(f);
  });
}

void EmissionFunctionArray::write_dN_dphidy_toFile()
{
  printf("Writing thermal dN_dphidy to file...\n");
  parallel_species(number_of_chosen_particles, [&](int is) {
    FILE *f = open_result("results/continuous/dN_dphidy_%d.dat", MCID[is]);
    for (long iy = 0; iy < y_tab_length; iy++) {
      double y = (DIMENSION == 3) ? y_tab->get(1, iy + 1) : 0.0;
      for (long iphi = 0; iphi < phi_tab_length; iphi++) {
        double sum = 0.0;
        for (long ipT = 0; ipT < pT_tab_length; ipT++) sum += pT_tab->get(2, ipT + 1) * dN_pTdpTdphidy[IDX(iy, iphi, ipT, is)];
        fprintf(f, "%.8e\t%.8e\t%.8e\n", y, phi_tab->get(1, iphi + 1), sum);
      }
      if (iy < y_tab_length - 1) fprintf(f, "\n");
    }
    fclose(f);
  });
}

void EmissionFunctionArray::write_dN_twopipTdpTdy_toFile()
{
  printf("Writing thermal dN_twopipTdpTdy to file...\n");
  parallel_species(number_of_chosen_particles, [&](int is) {
    FILE *f = open_result("results/continuous/dN_2pipTdpTdy_%d.dat", MCID[is]);
    for (long iy = 0; iy < y_tab_length; iy++) {
      double y = (DIMENSION == 3) ? y_tab->get(1, iy + 1) : 0.0;
      for (long ipT = 0; ipT < pT_tab_length; ipT++) {
        double sum = 0.0;
        for (long iphi = 0; iphi < phi_tab_length; iphi++) sum += phi_tab->get(2, iphi + 1) * dN_pTdpTdphidy[IDX(iy, iphi, ipT, is)] / two_pi;
        fprintf(f, "%.8e\t%.8e\t%.8e\n", y, pT_tab->get(1, ipT + 1), sum);
      }
      if (iy < y_tab_length - 1) fprintf(f, "\n");
    }
    fclose(f);
  });
}

void EmissionFunctionArray::write_dN_dy_toFile()
{
  printf("Writing thermal dN_dy to file...\n");
  parallel_species(number_of_chosen_particles, [&](int is) {
    FILE *f = open_result("results/continuous/dN_dy_%d.dat", MCID[is]);
    for (long iy = 0; iy < y_tab_length; iy++) {
      double y = (DIMENSION == 3) ? y_tab->get(1, iy + 1) : 0.0;
      double sum = 0.0;
      for (long iphi = 0; iphi < phi_tab_length; iphi++)
        for (long ipT = 0; ipT < pT_tab_length; ipT++)
          sum += phi_tab->get(2, iphi + 1) * pT_tab->get(2, ipT + 1) * dN_pTdpTdphidy[IDX(iy, iphi, ipT, is)];
      fprintf(f, "%5.8g\t%.8g\n", y, sum);
    }
    fclose(f);
  });
}

void EmissionFunctionArray::write_continuous_vn_toFile()
{
  printf("Writing continuous vn(pT,y) to file (for testing vn's)...\n");
  const int k_max = 7;
  parallel_species(number_of_chosen_particles, [&](int is) {
    FILE *f = open_result("results/continuous/vn_%d.dat", MCID[is]);
    for (long iy = 0; iy < y_tab_length; iy++) {
      double y = (DIMENSION == 3) ? y_tab->get(1, iy + 1) : 0.0;
      for (long ipT = 0; ipT < pT_tab_length; ipT++) {
        double re[k_max] = {0}, im[k_max] = {0}, den = 0.0;
        for (long iphi = 0; iphi < phi_tab_length; iphi++) {
          double phip = phi_tab->get(1, iphi + 1), w = phi_tab->get(2, iphi + 1);
          double v = dN_pTdpTdphidy[IDX(iy, iphi, ipT, is)];
          for (int k = 0; k < k_max; k++) {
            re[k] += cos(((double)k + 1.0) * phip) * w * v;
            im[k] += sin(((double)k + 1.0) * phip) * w * v;
          }
          den += w * v;
        }
        fprintf(f, "%.8e\t%.8e", y, pT_tab->get(1, ipT + 1));
        for (int k = 0; k < k_max; k++) {
          double vn = std::abs(std::complex<double>(re[k], im[k])) / den;
          if (den < 1.e-15) vn = 0.0;
          fprintf(f, "\t%.8e", vn);
        }
        fprintf(f, "\n");
      }
      fprintf(f, "\n");
    }
    fclose(f);
  });
}

// SpacetimeDistribution.cpp:448-490: bin mid-point and the histogram normalised by tau dtau / 2 pi r dr / dphi,
// files opened in append mode.  With dndx_bug_compat the histograms arrive already accumulated over species.
void EmissionFunctionArray::write_dN_dX_toFile()
{
  const double tau_w = (prm.tau_max - prm.tau_min) / (double)prm.tau_bins;
  const double r_w = (prm.r_max - prm.r_min) / (double)prm.r_bins;
  const double phi_w = two_pi / (double)prm.phip_bins;
  for (int is = 0; is < number_of_chosen_particles; is++) {
    printf("Writing dN_dX of particle %d to file...\n", MCID[is]);
    FILE *ft = open_result("results/continuous/dN_taudtaudy_%d.dat", MCID[is], "a");
    FILE *fr = open_result("results/continuous/dN_2pirdrdy_%d.dat", MCID[is], "a");
    FILE *fp = open_result("results/continuous/dN_dphidy_%d.dat", MCID[is], "a");
    for (int i = 0; i < prm.tau_bins; i++) {
      double mid = prm.tau_min + tau_w * ((double)i + 0.5);
      fprintf(ft, "%.6e\t%.6e\n", mid, dN_taudtaudy[(size_t)is * prm.tau_bins + i] / (mid * tau_w));
    }
    for (int i = 0; i < prm.r_bins; i++) {
      double mid = prm.r_min + r_w * ((double)i + 0.5);
      fprintf(fr, "%.6e\t%.6e\n", mid, dN_twopirdrdy[(size_t)is * prm.r_bins + i] / (two_pi * mid * r_w));
    }
    for (int i = 0; i < prm.phip_bins; i++) {
      double mid = phi_w * ((double)i + 0.5);
      fprintf(fp, "%.6e\t%.6e\n", mid, dN_dphisdy[(size_t)is * prm.phip_bins + i] / phi_w);
    }
    fclose(ft); fclose(fr); fclose(fp);
  }
}

void EmissionFunctionArray::write_particle_list_OSC()
{
  // one file per event (EmissionFunction.cpp:645-678): events are independent, so they are formatted and written by all
  // host threads (IS3D_WRITER_THREADS overrides); each row is built with std::to_chars, which prints the same correctly
  // rounded 17 significant digits as the reference's `scientific << setprecision(16)` stream
  printf("Writing sampled particles list to OSCAR File...\n");
  const int nt = host_threads((size_t)(Nevents > 0 ? Nevents : 1), "IS3D_WRITER_THREADS");
  parallel_for(nt, [&](int t) {
    std::string buf;
    char num[64];
    for (long e = t; e < Nevents; e += nt) {
      const auto &ev = particle_event_list[e];
      buf.clear();
      buf.reserve(ev.size() * 240 + 64);
      buf += "n pid px py pz E m x y z t\n";
      for (size_t i = 0; i < ev.size(); i++) {
        const Sampled_Particle &p = ev[i];
        auto r = std::to_chars(num, num + sizeof(num), (int)i);
        buf.append(num, r.ptr);
        buf += ' ';
        r = std::to_chars(num, num + sizeof(num), p.mcID);
        buf.append(num, r.ptr);
        const double v[9] = {p.px, p.py, p.pz, p.E, p.mass, p.x, p.y, p.z, p.t};
        for (double x : v) {
          buf += ' ';
          r = std::to_chars(num, num + sizeof(num), x, std::chars_format::scientific, 16);
          buf.append(num, r.ptr);
        }
        buf += '\n';
      }
      FILE *f = open_result("results/particle_list_osc_%d.dat", (int)e + 1);
      fwrite(buf.data(), 1, buf.size(), f);
      fclose(f);
    }
  });
}

// EmissionFunction.cpp:685-975 (six writers): normalisations kept, histograms come from the device
void EmissionFunctionArray::write_sampled_tests_to_file()
{
  const int ns = number_of_chosen_particles;
  const int K = 7;
  std::vector<double> dN_dy((size_t)ns * prm.y_bins), dN_deta((size_t)ns * prm.eta_bins), dN_dphip((size_t)ns * prm.phip_bins),
      dN_pT((size_t)ns * prm.pT_bins), pT_count((size_t)ns * prm.pT_bins), vn_re((size_t)K * ns * prm.pT_bins),
      vn_im((size_t)K * ns * prm.pT_bins), dN_tau((size_t)ns * prm.tau_bins), dN_r((size_t)ns * prm.r_bins),
      dN_phis((size_t)ns * prm.phip_bins);
  check(is3d_group_sample_histograms(grp, dN_dy.data(), dN_deta.data(), dN_dphip.data(), dN_pT.data(), pT_count.data(), vn_re.data(),
                               vn_im.data(), dN_tau.data(), dN_r.data(), dN_phis.data()), "is3d_sample_histograms");
  const double Y_CUT = prm.y_cut, nev = (double)Nevents;
  const double Y_WIDTH = 2.0 * Y_CUT / (double)prm.y_bins, ETA_WIDTH = 2.0 * prm.eta_cut / (double)prm.eta_bins;
  const double PT_WIDTH = (prm.pT_max - prm.pT_min) / (double)prm.pT_bins, PHIP_WIDTH = two_pi / (double)prm.phip_bins;
  const double TAU_WIDTH = (prm.tau_max - prm.tau_min) / (double)prm.tau_bins, R_WIDTH = (prm.r_max - prm.r_min) / (double)prm.r_bins;
  printf("Writing event-averaged distributions of each species to file...\n");
  for (int is = 0; is < ns; is++) {
    FILE *f = open_result("results/sampled/dN_dy/dN_dy_%d_test.dat", MCID[is]);
    FILE *fa = open_result("results/sampled/dN_dy/dN_dy_%d_average_test.dat", MCID[is]);
    double average = 0.0;
    for (int i = 0; i < prm.y_bins; i++) {
      double c = dN_dy[(size_t)is * prm.y_bins + i];
      average += c;
      fprintf(f, "%.6g\t%.6g\n", -Y_CUT + Y_WIDTH * ((double)i + 0.5), c / (Y_WIDTH * nev));
    }
    fprintf(fa, "%.6g\n", average / (2.0 * Y_CUT * nev));
    fclose(f); fclose(fa);
    f = open_result("results/sampled/dN_deta/dN_deta_%d_test.dat", MCID[is]);
    for (int i = 0; i < prm.eta_bins; i++)
      fprintf(f, "%.6g\t%.6g\n", -prm.eta_cut + ETA_WIDTH * ((double)i + 0.5), dN_deta[(size_t)is * prm.eta_bins + i] / (ETA_WIDTH * nev));
    fclose(f);
    f = open_result("results/sampled/dN_2pipTdpTdy/dN_2pipTdpTdy_%d_test.dat", MCID[is]);
    for (int i = 0; i < prm.pT_bins; i++) {
      double mid = prm.pT_min + PT_WIDTH * ((double)i + 0.5);
      fprintf(f, "%.6e\t%.6e\n", mid, dN_pT[(size_t)is * prm.pT_bins + i] / (two_pi * 2.0 * Y_CUT * PT_WIDTH * mid * nev));
    }
    fclose(f);
    f = open_result("results/sampled/dN_dphipdy/dN_dphipdy_%d_test.dat", MCID[is]);
    for (int i = 0; i < prm.phip_bins; i++)
      fprintf(f, "%.6e\t%.6e\n", PHIP_WIDTH * ((double)i + 0.5), dN_dphip[(size_t)is * prm.phip_bins + i] / (2.0 * Y_CUT * PHIP_WIDTH * nev));
    fclose(f);
    f = open_result("results/sampled/vn/vn_%d_test.dat", MCID[is]);
    for (int i = 0; i < prm.pT_bins; i++) {
      fprintf(f, "%.6e", prm.pT_min + PT_WIDTH * ((double)i + 0.5));
      for (int k = 0; k < K; k++) {
        size_t j = ((size_t)k * ns + is) * prm.pT_bins + i;
        double v = std::abs(std::complex<double>(vn_re[j], vn_im[j])) / pT_count[(size_t)is * prm.pT_bins + i];
        if (std::isnan(v) || std::isinf(v)) v = 0.0;
        fprintf(f, "\t%.6e", v);
      }
      fprintf(f, "\n");
    }
    fclose(f);
    f = open_result("results/sampled/dN_2pirdrdy/dN_2pirdrdy_%d_test.dat", MCID[is]);
    for (int i = 0; i < prm.r_bins; i++) {
      double mid = prm.r_min + R_WIDTH * ((double)i + 0.5);
      fprintf(f, "%.6e\t%.6e\n", mid, dN_r[(size_t)is * prm.r_bins + i] / (two_pi * mid * R_WIDTH * nev * 2.0 * Y_CUT));
    }
    fclose(f);
    f = open_result("results/sampled/dN_taudtaudy/dN_taudtaudy_%d_test.dat", MCID[is]);
    for (int i = 0; i < prm.tau_bins; i++) {
      double mid = prm.tau_min + TAU_WIDTH * ((double)i + 0.5);
      fprintf(f, "%.6e\t%.6e\n", mid, dN_tau[(size_t)is * prm.tau_bins + i] / (mid * TAU_WIDTH * nev * 2.0 * Y_CUT));
    }
    fclose(f);
    f = open_result("results/sampled/dN_dphisdy/dN_dphisdy_%d_test.dat", MCID[is]);
    for (int i = 0; i < prm.phip_bins; i++)
      fprintf(f, "%.6e\t%.6e\n", PHIP_WIDTH * ((double)i + 0.5), dN_phis[(size_t)is * prm.phip_bins + i] / (PHIP_WIDTH * nev * 2.0 * Y_CUT));
    fclose(f);
  }
}

}  // namespace is3dhost
