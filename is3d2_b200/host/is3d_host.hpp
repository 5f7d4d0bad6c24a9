// Host side of the drop-in: the reference's L3/L4 classes re-implemented over a structure-of-arrays
// surface, with EmissionFunctionArray's compute members forwarding to the CUDA C ABI (include/is3d_b200.h).
// Class and method names mirror the reference so its callers (and JETSCAPE, via IS3D) read unchanged:
//   ParameterReader   reference src/cpp/ParameterReader.{h,cpp}
//   Table             reference src/cpp/Table.{h,cpp}, Arsenal.cpp:51-137
//   Gauss_Laguerre, Gauss_Legendre, Plasma, FO_data_reader, PDG_Data   reference src/cpp/readindata.{h,cpp}
//   Deltaf_Data       reference src/cpp/DeltafData.{h,cpp} (table loading, PTB tables, fast-mode densities)
//   EmissionFunctionArray   reference src/cpp/EmissionFunction.{h,cpp}
//   IS3D              reference src/cpp/iS3D.{h,cpp}
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/is3d_b200.h"

namespace is3dhost {

constexpr double hbarC = 0.197327053;            // GeV fm (reference iS3D.h:14)
constexpr double kPi = 3.14159265358979323846;
constexpr double two_pi = 2.0 * kPi;
// pow(M_PI,2) * pow(hbarC,3) exactly as the reference forms it (iS3D.h:16-17)
double two_pi2_hbarC3();
constexpr int Maxparticle = 600;                 // reference iS3D.h:21

// All relative paths ("iS3D_parameters.dat", "PDG/...", "tables/...", "input/surface.dat", "results/...") are
// resolved against this root ("" = current directory, the reference's behaviour).
void set_root(const std::string &root);
std::string path(const std::string &relative);

[[noreturn]] void fatal(const std::string &message);   // printf + exit(-1), the reference's error convention
// "all" | "0-7" | "0,2,3" -> CUDA ordinals; NULL / empty -> {fallback}
std::vector<int> parse_device_list(const char *spec, int fallback);

class ParameterReader {
 public:
  void readFromFile(const std::string &filename, const std::string &commentSymbol = "#");
  void readFromArguments(long argc, char *argv[], const std::string &commentSymbol = "#", long start_from = 1);
  bool exist(const std::string &name) const;
  void setVal(const std::string &name, double value);
  double getVal(const std::string &name) const;      // missing key is fatal (ParameterReader.cpp:142-155)
  void echo() const;
 private:
  void phraseOneLine(const std::string &line, const std::string &commentSymbol);
  long find(const std::string &name) const;
  std::vector<std::string> names_;
  std::vector<double> values_;
};

// whitespace-separated numeric block; 1-based (column, row) access
class Table {
 public:
  Table() = default;
  explicit Table(const std::string &filename) { loadTableFromFile(filename); }
  void loadTableFromFile(const std::string &filename);
  double get(long col, long row) const { return cols_[col - 1][row - 1]; }
  long getNumberOfRows() const { return rows_; }
  long getNumberOfCols() const { return (long)cols_.size(); }
  const std::vector<double> &column(long col) const { return cols_[col - 1]; }
 private:
  std::vector<std::vector<double>> cols_;
  long rows_ = 0;
};

struct Gauss_Laguerre {
  int alpha = 0, points = 0;
  std::vector<double> root, weight;                  // [alpha][points] row-major
  void load_roots_and_weights(const std::string &file_name);
  const double *roots(int a) const { return root.data() + (size_t)a * points; }
  const double *weights(int a) const { return weight.data() + (size_t)a * points; }
};

struct Gauss_Legendre {
  int points = 0;
  std::vector<double> root, weight;
  void load_roots_and_weights(const std::string &file_name);
};

struct Plasma {
  double temperature = 0, energy_density = 0, pressure = 0, baryon_chemical_potential = 0, net_baryon_density = 0;
  void load_thermodynamic_averages();
};

struct particle_info {
  long mc_id = 0;
  std::string name;
  double mass = 0, width = 0;
  int gspin = 0, baryon = 0, strange = 0, charm = 0, bottom = 0, gisospin = 0, charge = 0, decays = 0, stable = 0;
  int sign = 0;
  double equilibrium_density = 0, bulk_density = 0, diff_density = 0;
};

// Freezeout surface as structure-of-arrays in the column order of include/is3d_b200.h (IS3D_COL_*).
struct FO_surface {
  std::vector<double> col[IS3D_SURFACE_COLUMNS];
  std::vector<double> vorticity[6];                  // wtx wty wtn wxy wxn wyn (mode 5 files; polarization is out of scope)
  int64_t size() const { return (int64_t)col[0].size(); }
  void resize(int64_t n, bool with_vorticity = false);
};

// volume-weighted thermodynamic averages (readindata.cpp:330-366) + the side file the later stages re-read
void compute_thermodynamic_averages(const FO_surface &s, double avg[5]);
void compute_thermodynamic_sums(const FO_surface &s, double sums[6]);
void write_thermodynamic_averages(const double avg[5]);

class FO_data_reader {
 public:
  FO_data_reader(ParameterReader *paraRdr, const std::string &path_in);
  ~FO_data_reader();
  long get_number_cells();
  void read_freezeout_surface(FO_surface &surf);
  void read_surface_cpu_vh(FO_surface &surf);         // mode 1 (5 = with thermal vorticity)
  void read_surface_music(FO_surface &surf);          // mode 6
  void read_surface_hic_eventgen(FO_surface &surf);   // mode 7
 private:
  std::vector<double> slurp(long columns);
  int mode, dimension, include_baryon;
  long number_of_cells = 0;
  void unmap();
  bool load_cache(FO_surface &surf);                  // IS3D_SURFACE_CACHE=1: input/surface.dat.soa
  void store_cache(const FO_surface &surf);
  const char *map_ = nullptr;                         // surface.dat, memory-mapped once
  size_t map_size_ = 0;
  bool mapped_ = false;
};

class PDG_Data {
 public:
  explicit PDG_Data(ParameterReader *paraRdr);
  int read_resonances(std::vector<particle_info> &particle);
  int read_resonances_conventional(std::vector<particle_info> &particle, const std::string &pdg_filename);
  int read_resonances_smash_box(std::vector<particle_info> &particle, const std::string &pdg_filename);
 private:
  int hrg_eos;
};

class Deltaf_Data {
 public:
  explicit Deltaf_Data(ParameterReader *paraRdr);
  void load_df_coefficient_data();
  void compute_jonah_coefficients(const std::vector<particle_info> &particle_data);
  void compute_particle_densities(std::vector<particle_info> &particle_data);

  int hrg_eos, mode, df_mode, include_baryon;
  int points_T = 0, points_muB = 0;
  std::vector<double> T_array, muB_array;
  std::vector<double> tab[10];                       // c0 c1 c2 c3 c4 F G betabulk betaV betapi, [muB][T]
  static constexpr int jonah_points = 301;
  std::vector<double> lambda_squared_array, z_array, bulkPi_over_Peq_array;
  double bulkPi_over_Peq_max = -1.0;
  bool have_jonah = false;
 private:
  std::string hrg_eos_path;
};

struct Sampled_Particle {
  int chosen_index = 0, mcID = 0;
  double mass = 0, tau = 0, x = 0, y = 0, eta = 0, t = 0, z = 0, E = 0, px = 0, py = 0, pz = 0;
};

class EmissionFunctionArray {
 public:
  // `ready_group`: contexts created ahead of time with params_from(paraRdr) (the executable builds them while surface.dat is
  // parsed); NULL = create them here
  EmissionFunctionArray(ParameterReader *paraRdr_in, Table *chosen_particles, Table *pT_tab_in, Table *phi_tab_in,
                        Table *y_tab_in, Table *eta_tab_in, std::vector<particle_info> *particles_in,
                        FO_surface *surf_in, Deltaf_Data *df_data_in, is3d_group *ready_group = nullptr,
                        const is3d_params *ready_params = nullptr);
  // the run-time switches of iS3D_parameters.dat (+ the IS3D_* environment knobs) as the C ABI takes them
  static is3d_params params_from(ParameterReader *paraRdr, int *polzn_file_compat = nullptr);
  // IS3D_DEVICES / IS3D_DEVICE -> contexts (+ communicator); a failure is fatal unless `error` is given (then NULL + message)
  static is3d_group *create_group(const is3d_params &prm, std::string *error = nullptr);
  ~EmissionFunctionArray();

  void calculate_spectra(std::vector<std::vector<Sampled_Particle>> &particle_event_list_in);

  // compute members (EmissionFunction.h:147-179): thin calls into the CUDA C ABI
  void calculate_dN_pTdpTdphidy();          // df_mode 1,2 / 3,4 / 5 selected by the context
  void calculate_dN_dX();                   // df_mode 1-4
  double calculate_total_yield();
  void sample_dN_pTdpTdphidy();

  // writers (EmissionFunction.cpp:406-975)
  void write_dN_pTdpTdphidy_toFile();
  void write_dN_dphidy_toFile();
  void write_dN_twopipTdpTdy_toFile();
  void write_dN_dy_toFile();
  void write_continuous_vn_toFile();
  void write_dN_dX_toFile();
  void calculate_spin_polzn();
  void write_polzn_vector_toFile();
  void write_particle_list_OSC();
  void write_sampled_tests_to_file();

  is3d_ctx *context() { return ctx; }
  is3d_group *group() { return grp; }
  void set_surface_on_device();             // (re)upload the SoA surface

  std::vector<double> dN_pTdpTdphidy;       // Ns*NpT*Nphi*Ny, same indexing as the reference
  std::vector<double> St, Sx, Sy, Sn, Snorm; // spin polarization (mode-5 surfaces), spectra layout
  int polzn_file_compat = 1;                // reproduce the reference's storage / read index mismatch in results/S*.dat
  std::vector<double> dN_taudtaudy, dN_twopirdrdy, dN_dphisdy;   // dN/dX histograms, Ns x bins
  std::vector<std::vector<Sampled_Particle>> particle_event_list;
  long Nevents = 1;
  is3d_stats stats{};
  double seconds_compute = 0;
  int number_of_chosen_particles = 0;
  std::vector<int> MCID;
  long pT_tab_length, phi_tab_length, y_tab_length, eta_tab_length;

 private:
  void check(is3d_status st, const char *what);
  ParameterReader *paraRdr;
  is3d_params prm{};
  is3d_group *grp = nullptr;                // one context per GPU of this run (IS3D_DEVICES); owns the communicator
  is3d_ctx *ctx = nullptr;                  // the group's first context (single-GPU runs: the only one)
  int OPERATION, MODE, DF_MODE, DIMENSION, OVERSAMPLE, TEST_SAMPLER;
  double MIN_NUM_HADRONS, MAX_NUM_SAMPLES;
  Table *pT_tab, *phi_tab, *y_tab, *eta_tab;
  std::vector<particle_info> *particles;
  std::vector<int> chosen_particles_sampling_table;
  FO_surface *surf;
  Deltaf_Data *df_data;
  std::vector<double> sampled_hist[10];
};

class IS3D {
 public:
  // surface handed over in memory (JETSCAPE entry, reference iS3D.h:80-103)
  std::vector<double> tau, x, y, eta, dsigma_tau, dsigma_x, dsigma_y, dsigma_eta, E, T, P, ux, uy, un, pixx, pixy, pixn,
      piyy, piyn, pinn, Pi;
  std::vector<std::vector<Sampled_Particle>> final_particles_;

  void run_particlization(int fo_from_file);
  void read_fo_surf_from_memory(std::vector<double> tau_in, std::vector<double> x_in, std::vector<double> y_in,
                                std::vector<double> eta_in, std::vector<double> dsigma_tau_in,
                                std::vector<double> dsigma_x_in, std::vector<double> dsigma_y_in,
                                std::vector<double> dsigma_eta_in, std::vector<double> E_in, std::vector<double> T_in,
                                std::vector<double> P_in, std::vector<double> ux_in, std::vector<double> uy_in,
                                std::vector<double> un_in, std::vector<double> pixx_in, std::vector<double> pixy_in,
                                std::vector<double> pixn_in, std::vector<double> piyy_in, std::vector<double> piyn_in,
                                std::vector<double> pinn_in, std::vector<double> Pi_in);
};

}  // namespace is3dhost
