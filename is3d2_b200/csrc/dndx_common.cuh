// K4 shared pieces: mean spacetime distributions dN/dX (operation 0).
// Reference: calculate_dN_dX / calculate_dN_dX_feqmod, src/cpp/SpacetimeDistribution.cpp:31-517, :520-1246.
//
// Per (cell, species) the reference integrates the same integrand as the spectra over (pT, phi, y[, eta]) with the
// pT and phi table weights (the y nodes are summed UNweighted in 3+1d, :330-404) into one scalar dN_dy_cell and adds
// it to three 1-D histograms selected by the cell's (tau, r, phi_s) (:413-440).
//
// GPU mapping (K1's schedule turned cell-stationary).  A thread owns R species classes of ONE baryon number at ONE pT node
// (the slot table of spectra_df.cu); a 256-thread block covers gpb = 256 / NpT thread groups (5 x 51 = 255 threads for
// the shipped pT table) and walks a contiguous chunk of cells.  The (y, eta, phi) quadrature points of up to 16 cells are
// built cooperatively -- one thread per (cell, point) -> 128-item tiles in shared memory -- then every thread marches
// over a cell's items with broadcast LDS.128 exactly like the spectra kernels (df_eval_u / feqmod_accum_u: the pT
// products are formed once per (item, thread) and shared by the R evaluations).  After the last item of a cell the
// R partial sums of each thread are multiplied by the pT weight, reduced over the NpT threads of the group through
// shared memory + warp shuffles, and the (cell, class) scalar is scattered with three FP64 atomicAdd into histograms
// [class][bin], expanded to species x degeneracy at the end.
#pragma once

#include "cellmath.cuh"
#include "spectra_df.cuh"

namespace is3d {

constexpr int kDndxThreads = 256;     // threads per block
constexpr int kDndxTile = 128;        // items (cell x quadrature point) per shared-memory tile
constexpr int kDndxR = 4;             // species classes per thread
constexpr int kDndxPairR = 2;         // charge-conjugate pair slots per thread of the pair launches (= 4 classes per thread)
constexpr int kDndxMaxCells = 16;     // cells per tile
constexpr int kDndxMaxGroups = 8;     // thread groups per block

struct DndxGrid {
  int ns;                         // species CLASSES
  int ngroups, gpb;               // thread groups (kDndxR class slots, one baryon number each); groups per block
  const int *slot_class;          // [ngroups * kDndxR], -1 = padding
  int NpT, Nphi, Ny, Neta, dimension;
  const double *mT, *baryon, *sign;                       // [class * NpT + ipT]  (ctx.h SpeciesBins)
  const double *c_mass, *c_deg, *c_baryon, *c_sign;       // [class]
  const double *pT, *pTw;         // [NpT]
  const double *cosphi, *sinphi, *phiw, *yv, *etav, *etaw;
  // histograms
  double tau_min, tau_width, r_min, r_width, phi_width;
  int tau_bins, r_bins, phi_bins;
  double *hist_tau, *hist_r, *hist_phi;               // [class][bins]
  const double *exptab;                               // 2^(m/1024), global memory
  // dropping of negligible quadrature points (df_mode 1, 2; dndx_df_kernel)
  double margin;                                      // is3d_params.negligible_margin; <= 0: only the x >= 680 points are dropped
  unsigned long long *prune_counters;                 // [0] += (cell, class) scalars failing the bound test, [1] += points dropped
};

// histogram bins of a cell (SpacetimeDistribution.cpp:413-440), -1 = outside the histogram; evaluated once per cell
struct DndxCellBins { int itau, ir, iphi; };

IS3D_D DndxCellBins dndx_cell_bins(const DndxGrid &g, double tau, double x, double y)
{
  double r = sqrt(x * x + y * y);
  double phi = atan2(y, x);
  if (phi < 0.0) phi += kTwoPi;
  long itau = (int)floor((tau - g.tau_min) / g.tau_width);
  long ir = (int)floor((r - g.r_min) / g.r_width);
  long iphi = (int)floor(phi / g.phi_width);
  DndxCellBins b;
  b.itau = (itau >= 0 && itau < g.tau_bins) ? (int)itau : -1;
  b.ir = (ir >= 0 && ir < g.r_bins) ? (int)ir : -1;
  b.iphi = (iphi >= 0 && iphi < g.phi_bins) ? (int)iphi : -1;
  return b;
}

IS3D_D void dndx_scatter(const DndxGrid &g, int s, const DndxCellBins &b, double value)
{
#if defined(__CUDA_ARCH__)
  if (b.itau >= 0) atomicAdd(&g.hist_tau[(size_t)s * g.tau_bins + b.itau], value);
  if (b.ir >= 0) atomicAdd(&g.hist_r[(size_t)s * g.r_bins + b.ir], value);
  if (b.iphi >= 0) atomicAdd(&g.hist_phi[(size_t)s * g.phi_bins + b.iphi], value);
#endif
}

}  // namespace is3d
