// Internal context of the C ABI (include/is3d_b200.h).  Not part of the public interface.
#pragma once

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "../../include/is3d_b200.h"
#include "cellmath.cuh"
#include "dftables.cuh"

#define IS3D_CUDA_TRY(ctx, expr)                                                                          \
  do {                                                                                                    \
    cudaError_t err__ = (expr);                                                                           \
    if (err__ != cudaSuccess) {                                                                           \
      (ctx)->set_error(std::string(#expr) + ": " + cudaGetErrorString(err__) + " (" + __FILE__ + ":" +    \
                       std::to_string(__LINE__) + ")");                                                   \
      return IS3D_ERR_CUDA;                                                                               \
    }                                                                                                     \
  } while (0)

#define IS3D_TRY(expr)                      \
  do {                                      \
    is3d_status st__ = (expr);              \
    if (st__ != IS3D_OK) return st__;       \
  } while (0)

struct is3d_ctx {
  is3d_params prm;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // timing events of the compute calls (created once, destroyed with the context)
  // second compute stream: the single-class launch of the spectra kernels runs beside the pair launch and fills its tail
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // sampler copy pipeline (sampler.cu): pass k's D2H on copy_stream overlaps pass k + 1's kernels on `stream`
  volatile unsigned long long *h_words = nullptr;   // mapped pinned words: control scalars the device publishes to the host
  unsigned long long *d_words = nullptr;            // their device address
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_sorted[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  std::string err;
  int sm_count = 148;
  // multi-GPU (comm.cu): NCCL communicator this context's results are summed over (nullptr = single GPU)
  void *comm = nullptr;
  int comm_size = 1, comm_rank = 0;
  int64_t comm_collectives = 0;            // all-reduces issued so far
  double *d_exptab = nullptr;              // 2^(m/1024) table of fast_exp (common.cuh)

  // chosen species (host copies + device arrays)
  int ns = 0;
  std::vector<double> h_mass, h_sign, h_deg, h_baryon, h_neq, h_dnbulk, h_dndiff;
  std::vector<int> h_mcid;
  double *d_mass = nullptr, *d_sign = nullptr, *d_deg = nullptr, *d_baryon = nullptr;
  double *d_neq = nullptr, *d_dnbulk = nullptr, *d_dndiff = nullptr;
  int *d_mcid = nullptr;

  // PDG table (PTMA reconstruction)
  int npdg = 0;
  std::vector<double> h_pdg_mass, h_pdg_sign, h_pdg_deg, h_pdg_baryon;
  double *d_pdg_mass = nullptr, *d_pdg_sign = nullptr, *d_pdg_deg = nullptr, *d_pdg_baryon = nullptr;

  // momentum tables as read from file (nodes, weights)
  std::vector<double> pT, pTw, phi, phiw, y, yw, eta, etaw;
  bool have_momentum = false;
  // effective grid after the dimension rule (EmissionFunction.cpp:146-153)
  int NpT = 0, Nphi = 0, Ny = 0, Neta = 0;
  std::vector<double> yv, etav, etawv;     // effective y nodes / eta nodes+weights
  double *d_pT = nullptr, *d_pTw = nullptr, *d_cosphi = nullptr, *d_sinphi = nullptr, *d_phiw = nullptr;
  double *d_y = nullptr, *d_yw = nullptr, *d_eta = nullptr, *d_etaw = nullptr;

  // Gauss-Laguerre / Legendre
  int gla_alpha = 0, gla_pts = 0, leg_pts = 0;
  std::vector<double> h_gla_root, h_gla_weight, h_leg_root, h_leg_weight;
  double *d_gla_root = nullptr, *d_gla_weight = nullptr, *d_leg_root = nullptr, *d_leg_weight = nullptr;

  // thermodynamic averages
  bool have_avg = false;
  double T_avg = 0, E_avg = 0, P_avg = 0, muB_avg = 0, nB_avg = 0;

  // df tables (device pointers inside) + host copies for fast-mode coefficients
  is3d::DfTables tb;
  bool have_df = false, have_ptb = false;
  std::vector<double> h_T, h_muB, h_tab[10], h_spc[5];
  std::vector<double> h_ptb_x, h_ptb_l2, h_ptb_z, h_ptb_l2c, h_ptb_zc;

  // surface
  is3d::SurfaceView surf{};
  bool have_surface = false;
  bool surface_owned = false;
  int64_t global_offset = 0;
  double *d_surface_block = nullptr;       // one (grow-only) allocation holding all owned columns
  size_t surface_block_bytes = 0;

  // thermal vorticity (mode-5 surfaces): six columns, one allocation
  bool have_vorticity = false;
  int64_t vorticity_n = 0, vorticity_pitch = 0;
  double *d_vorticity = nullptr;

  // sampler histograms (device)
  std::map<std::string, double *> hist;

  // pinned host buffers holding particle lists handed to the caller (sampler.cu)
  struct HostList { void *ptr = nullptr; size_t capacity = 0; bool in_use = false; is3d_ctx *owner = nullptr; bool pinned = true; };
  std::vector<HostList *> host_lists;

  // every device allocation made by this context
  std::vector<void *> owned;
  // grow-only named scratch buffers
  std::map<std::string, std::pair<void *, size_t>> scratch;

  void set_error(const std::string &m) { err = m; }

  is3d_status dev_alloc(void **p, size_t bytes)
  {
    *p = nullptr;
    if (bytes == 0) bytes = 8;
    IS3D_CUDA_TRY(this, cudaMalloc(p, bytes));
    owned.push_back(*p);
    return IS3D_OK;
  }
  void dev_free(void *p)
  {
    if (!p) return;
    for (size_t i = 0; i < owned.size(); i++)
      if (owned[i] == p) { owned.erase(owned.begin() + i); break; }
    cudaFree(p);
  }
  template <class T>
  is3d_status upload(T **dst, const T *src, size_t n)
  {
    if (*dst) { dev_free(*dst); *dst = nullptr; }
    IS3D_TRY(dev_alloc((void **)dst, n * sizeof(T)));
    if (n) IS3D_CUDA_TRY(this, cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, stream));
    IS3D_CUDA_TRY(this, cudaStreamSynchronize(stream));
    return IS3D_OK;
  }
  is3d_status get_scratch(const char *name, size_t bytes, void **p)
  {
    auto it = scratch.find(name);
    if (it != scratch.end() && it->second.second >= bytes) { *p = it->second.first; return IS3D_OK; }
    if (it != scratch.end()) { dev_free(it->second.first); scratch.erase(it); }
    IS3D_TRY(dev_alloc(p, bytes));
    scratch[name] = {*p, bytes};
    return IS3D_OK;
  }
};

namespace is3d {

// Species classes: hadrons with the same (mass, quantum-statistics sign, baryon number) have the same Cooper-Frye
// integrand and differ only by the degeneracy factor in front (the SMASH list's 444 species are 193 classes: isospin
// multiplets share one mass).  The spectra kernels integrate one representative per class; the final reduction writes
// every species' bins as degeneracy x class sum, so the output is exactly what a per-species loop delivers.
struct SpeciesBins {
  int nclass = 0;
  const double *mT = nullptr, *pT = nullptr, *m2 = nullptr, *baryon = nullptr, *sign = nullptr;   // [nclass * NpT], device
  const double *c_mass = nullptr, *c_deg = nullptr, *c_baryon = nullptr, *c_sign = nullptr;      // [nclass] representatives, device
  const int *class_of = nullptr;                                                                 // [ns], device
};

// cells per pass of the continuous paths (bounds the cell-pack scratch); IS3D_PASS_CELLS is a test hook that forces
// small passes so that the multi-pass accumulation is exercised on small surfaces (rounded up to the 256-cell tile)
inline int64_t pass_cells(int64_t default_cells)
{
  if (const char *v = getenv("IS3D_PASS_CELLS")) { long long c = atoll(v); if (c > 0) return (c + 255) / 256 * 256; }
  return default_cells;
}

void species_classes(const is3d_ctx *ctx, std::vector<int> *class_of, std::vector<int> *rep);

// Launch order of the spectra kernels (spectra_df.cu; K1 and K2), host side:
//   column_order   thread columns (group, pT node) = group * NpT + ip in order of their smallest mT, so that the columns of a
//                  block span a narrow mT range and agree on which (cell, y) items are negligible.  `slots` lists the class ids
//                  of the thread groups, ids_per_group each (-1 = padding); rep[class] = a species of the class.
//   rapidity_order rapidity rows in order of |y - mean y| ascending: the rows that drop the fewest items are launched first
//   fill_bin_rows  bin_row[class * NpT + ip] = row0 + (position of the bin's column in `order`) / threads_per_block: the block
//                  row group whose dropped-term bounds belong to that bin (PruneCheck)
std::vector<int> column_order(const is3d_ctx *ctx, const std::vector<int> &slots, int ids_per_group, const std::vector<int> &rep);
std::vector<int> rapidity_order(const is3d_ctx *ctx);
void fill_bin_rows(const is3d_ctx *ctx, const std::vector<int> &slots, int ids_per_group, const std::vector<int> &order, int threads_per_block,
                   int row0, std::vector<int> *bin_row);

// compute paths (one translation unit each)
is3d_status run_spectra_df(is3d_ctx *ctx, double *out_dev, is3d_stats *stats);        // df_mode 1,2
is3d_status run_spectra_feqmod(is3d_ctx *ctx, double *out_dev, is3d_stats *stats);    // df_mode 3,4 and 5 (PTMA)
is3d_status run_dndx(is3d_ctx *ctx, double *tau_dev, double *r_dev, double *phi_dev, is3d_stats *stats);
is3d_status run_total_yield(is3d_ctx *ctx, double *ntotal, is3d_stats *stats);
is3d_status run_cell_yields(is3d_ctx *ctx, double *dn_tot_host, double *dn_list_host, is3d_stats *stats);
is3d_status run_sampler(is3d_ctx *ctx, int64_t nevents, int record_kind, void **particles, int64_t *total, int64_t *counts,
                        is3d_stats *stats);
void expand_compact(const is3d_ctx *ctx, const is3d_particle_compact *in, int64_t n, is3d_particle *out);
void release_host_lists_of(is3d_ctx *ctx);
// a plain (pageable) library-owned list, released by is3d_free_particles like the pinned ones (merged multi-GPU lists)
void *alloc_plain_list(size_t bytes);
// comm.cu: in-place SUM all-reduce over the attached communicator (no-op for a single GPU)
is3d_status comm_allreduce(is3d_ctx *ctx, double *dev, int64_t n);
is3d_status comm_allreduce_host(is3d_ctx *ctx, double *host, int n);
void comm_release(is3d_ctx *ctx);
is3d_status run_polarization(is3d_ctx *ctx, double *out_dev, is3d_stats *stats);
is3d_status measure_fp64_peak(is3d_ctx *ctx, double *tflops);
is3d_status probe_math(is3d_ctx *ctx, int64_t n, const double *x, double *out_exp, double *out_rcp, double *out_sqrt);
is3d_status probe_aniso_math(is3d_ctx *ctx, int64_t n, const double *x, double *out_atan, double *out_atanh, double *out_log);

}  // namespace is3d
