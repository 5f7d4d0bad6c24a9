// K4 shared pieces: mean spacetime distributions dN/dX (operation 0).
// Reference: calculate_dN_dX / calculate_dN_dX_feqmod, src/cpp/SpacetimeDistribution.cpp:31-517, :520-1246.
//
// Per (cell, species) the reference integrates the same integrand as the spectra over (pT, phi, y[, eta]) with the
// pT and phi table weights (the y nodes are summed UNweighted in 3+1d, :330-404) into one scalar dN_dy_cell and adds
// it to three 1-D histograms selected by the cell's (tau, r, phi_s) (:413-440).
//
// GPU mapping: cell-stationary.  One thread = one species class (ctx.h SpeciesBins); a one-warp block walks a contiguous chunk of cells,
// all lanes on the same cell, so every per-cell / per-(y, phi) quantity is warp-uniform and is simply recomputed in
// registers (a few % of the 51 x Nphi x Ny evaluations it feeds).  The per-(pT, species) momentum constants come
// from four transposed tables [ipT][species] (coalesced, L1/L2 resident, 725 KB for 444 species).  The (cell,
// species) scalar is scattered with three FP64 atomicAdd into histograms [species][bin].
#pragma once

#include "cellmath.cuh"
#include "spectra_df.cuh"

namespace is3d {

constexpr int kDndxThreads = 32;    // one warp per block: species classes are padded to a multiple of this

struct DndxGrid {
  int ns, ns_pad;                 // species, padded to a multiple of kDndxThreads
  int NpT, Nphi, Ny, Neta, dimension;
  // [ipT][ns_pad]: mT * pT_weight, mT, mT^2, mT * pT
  const double *mTw, *mT, *mT2, *mTpT;
  const double *pT, *pTw;         // [NpT]
  const double *cosphi, *sinphi, *phiw, *yv, *etav, *etaw;
  const double *mass2, *baryon, *sign, *deg, *mass;   // [ns_pad]
  // histograms
  double tau_min, tau_width, r_min, r_width, phi_width;
  int tau_bins, r_bins, phi_bins;
  double *hist_tau, *hist_r, *hist_phi;               // [ns][bins]
  const double *exptab;                               // 2^(m/1024), global memory
};

// SpacetimeDistribution.cpp:413-440
IS3D_D void dndx_scatter(const DndxGrid &g, int s, double tau, double x, double y, double value)
{
  double r = sqrt(x * x + y * y);
  double phi = atan2(y, x);
  if (phi < 0.0) phi += kTwoPi;
  long itau = (int)floor((tau - g.tau_min) / g.tau_width);
  long ir = (int)floor((r - g.r_min) / g.r_width);
  long iphi = (int)floor(phi / g.phi_width);
#if defined(__CUDA_ARCH__)
  if (itau >= 0 && itau < g.tau_bins) atomicAdd(&g.hist_tau[(size_t)s * g.tau_bins + itau], value);
  if (ir >= 0 && ir < g.r_bins) atomicAdd(&g.hist_r[(size_t)s * g.r_bins + ir], value);
  if (iphi >= 0 && iphi < g.phi_bins) atomicAdd(&g.hist_phi[(size_t)s * g.phi_bins + iphi], value);
#endif
}

}  // namespace is3d
