// K4: mean spacetime distributions dN/dX (operation 0) for df_mode 1-4 on sm_100a.
// Replaces EmissionFunctionArray::calculate_dN_dX and calculate_dN_dX_feqmod
// (reference src/cpp/SpacetimeDistribution.cpp:31-517, :520-1246).  See dndx_common.cuh for the mapping.
#include <cstring>

#include "ctx.h"
#include "dndx_common.cuh"
#include "spectra_feqmod.cuh"

namespace is3d {

namespace {

__global__ void dndx_df_setup_kernel(SurfaceView surf, int64_t begin, int64_t count, DfTables tb, DfFlags fl,
                                     double *__restrict__ pack, int64_t stride, unsigned long long *counters)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Cell c = load_cell(surf, begin + i, fl.include_baryon != 0);
  double p[DP_SIZE];
  int st = df_setup_cell(c, tb, fl, p);
#pragma unroll
  for (int k = 0; k < DP_SIZE; k++) pack[k * stride + i] = p[k];
  if (st == CELL_SKIPPED) atomicAdd(&counters[0], 1ull);
  if (st == CELL_OUT_OF_TABLE) atomicAdd(&counters[1], 1ull);
}

__global__ void dndx_feqmod_setup_kernel(SurfaceView surf, int64_t begin, int64_t count, DfTables tb, FeqmodFlags fl,
                                         const double *__restrict__ gla_root, const double *__restrict__ gla_weight, int gla_pts,
                                         double *__restrict__ pack, int64_t stride, unsigned long long *counters)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Cell c = load_cell(surf, begin + i, fl.include_baryon != 0);
  double p[FP_SIZE];
  int st = feqmod_setup_cell(c, tb, fl, gla_root, gla_weight, gla_pts, p);
#pragma unroll
  for (int k = 0; k < FP_SIZE; k++) pack[k * stride + i] = p[k];
  if (st == CELL_SKIPPED) { atomicAdd(&counters[0], 1ull); return; }
  if (st == CELL_OUT_OF_TABLE) { atomicAdd(&counters[1], 1ull); return; }
  if (st & CELL_BREAKDOWN) atomicAdd(&counters[2], 1ull);
  if (st & CELL_PL_NEGATIVE) atomicAdd(&counters[3], 1ull);
}

// per-(pT, species) momentum constants of one thread at one pT node
struct DndxBin {
  DfBin b;
  double pT, pT2;
  double mTw, pTw;     // mT * pT_weight, pT * pT_weight (p.dsigma carries the quadrature weights)
};

__device__ __forceinline__ DndxBin dndx_load_bin(const DndxGrid &g, int ipT, int s, double m2, double baryon, double sign)
{
  DndxBin d;
  const int idx = ipT * g.ns_pad + s;
  const double pT = g.pT[ipT], w = g.pTw[ipT];
  d.b.mT = g.mT[idx]; d.b.mT2 = g.mT2[idx]; d.pT = pT; d.pT2 = pT * pT;
  d.b.m2 = m2; d.b.baryon = baryon; d.b.sign = sign;
  d.mTw = g.mTw[idx]; d.pTw = pT * w;
  return d;
}

// Both kernels: one-warp blocks, lane = species class, all lanes on the same cell.  The (y, eta, phi) points of a cell are
// taken in chunks of 32: lane j builds the item of point j (one sinh / cosh per lane instead of every lane rebuilding
// every item), the warp then loops pT OUTER -- the lane's three momentum constants are loaded once per (chunk, pT) --
// and the chunk's items INNER, read from shared memory with broadcast LDS.128.
static_assert(kDndxThreads == 32, "the dN/dX kernels synchronise with __syncwarp");

// flattened (iy, ie, iphi) point j of the cell's momentum-space quadrature
struct DndxPoint { double yval, eta, w, cphi, sphi; };

template <class PackFn>
__device__ __forceinline__ DndxPoint dndx_point(const DndxGrid &g, PackFn pk, int j)
{
  const int iphi = j % g.Nphi, rest = j / g.Nphi, ie = rest % g.Neta, iy = rest / g.Neta;
  DndxPoint p;
  p.yval = g.yv[iy];
  if (g.dimension == 3) { p.eta = pk(DP_ETA); p.w = 1.0; }
  else { p.eta = g.etav[ie]; p.w = g.etaw[ie]; }
  p.w *= g.phiw[iphi];        // phi weight folded into the p.dsigma coefficients (positive: the outflow test is unchanged)
  p.cphi = g.cosphi[iphi]; p.sphi = g.sinphi[iphi];
  return p;
}

// df_mode 1, 2 (SpacetimeDistribution.cpp:170-441)
template <int MODE, bool BARYON, bool REGULATE, bool OUTFLOW>
__global__ void __launch_bounds__(kDndxThreads)
dndx_df_kernel(const double *__restrict__ pack, int64_t stride, int64_t ncells, int64_t cells_per_block, SurfaceView surf,
               int64_t surf_begin, DndxGrid g)
{
  __shared__ double exptab[kExpTableSize];
  __shared__ DfItem items[kDndxThreads];
  load_exp_table(exptab, g.exptab);
  __syncthreads();
  const int lane = threadIdx.x;
  const int s = blockIdx.x * kDndxThreads + lane;
  const double m2 = g.mass2[s], baryon = g.baryon[s], sign = g.sign[s];
  const int npoints = g.Ny * g.Neta * g.Nphi;
  const int64_t c0 = (int64_t)blockIdx.y * cells_per_block;
  int64_t c1 = c0 + cells_per_block;
  if (c1 > ncells) c1 = ncells;
  for (int64_t cell = c0; cell < c1; cell++) {
    if (pack[DP_VALID * stride + cell] == 0.0) continue;
    auto pk = [&](int k) { return pack[k * stride + cell]; };
    double acc = 0.0;
    for (int j0 = 0; j0 < npoints; j0 += kDndxThreads) {
      if (j0 + lane < npoints) {
        const DndxPoint pt = dndx_point(g, pk, j0 + lane);
        const double d = pt.yval - pt.eta;
        items[lane] = df_make_item(pk, MODE, sinh(d), cosh(d), pt.cphi, pt.sphi, pt.w);
      }
      __syncwarp();
      const int nj = min(kDndxThreads, npoints - j0);
      for (int ipT = 0; ipT < g.NpT; ipT++) {
        const DndxBin bn = dndx_load_bin(g, ipT, s, m2, baryon, sign);
#pragma unroll 2
        for (int k = 0; k < nj; k++) {
          const DfItem it = items[k];
          const double pds = fma(bn.mTw, it.c1, bn.pTw * it.d1);
          double v = pds * df_distribution<MODE, BARYON, REGULATE>(it, df_share<BARYON>(it, bn.pT, bn.pT2), bn.b, exptab);
          if (OUTFLOW) v = (pds <= 0.0) ? 0.0 : v;
          acc += v;
        }
      }
      __syncwarp();
    }
    if (s < g.ns) {
      const int64_t gc = surf_begin + cell;
      dndx_scatter(g, s, surf.col[0][gc], surf.col[1][gc], surf.col[2][gc], kCooperFryePrefactor * acc);
    }
  }
}

union DndxItemSlot {
  DfItem lin;
  FeqmodItem mod;
  __device__ DndxItemSlot() {}
};

// df_mode 3, 4 (SpacetimeDistribution.cpp:676-1160)
template <bool BARYON, bool REGULATE, bool OUTFLOW, bool SPECIES_RENORM>
__global__ void __launch_bounds__(kDndxThreads)
dndx_feqmod_kernel(const double *__restrict__ pack, int64_t stride, int64_t ncells, int64_t cells_per_block, SurfaceView surf,
                   int64_t surf_begin, DndxGrid g, const double *__restrict__ gla_root, const double *__restrict__ gla_weight,
                   int gla_pts)
{
  __shared__ double exptab[kExpTableSize];
  __shared__ DndxItemSlot items[kDndxThreads];
  __shared__ unsigned char item_linear[kDndxThreads];
  __shared__ RenormNodes nodes;
  load_exp_table(exptab, g.exptab);
  if (SPECIES_RENORM) nodes.load(gla_root, gla_weight, gla_pts);
  __syncthreads();
  const int lane = threadIdx.x;
  const int s = blockIdx.x * kDndxThreads + lane;
  const double m2 = g.mass2[s], baryon = g.baryon[s], sign = g.sign[s], deg = g.deg[s], mass = g.mass[s];   // deg cancels in the renorm ratio
  const int npoints = g.Ny * g.Neta * g.Nphi;
  const int64_t c0 = (int64_t)blockIdx.y * cells_per_block;
  int64_t c1 = c0 + cells_per_block;
  if (c1 > ncells) c1 = ncells;
  for (int64_t cell = c0; cell < c1; cell++) {
    if (pack[DP_VALID * stride + cell] == 0.0) continue;
    auto pk = [&](int k) { return pack[k * stride + cell]; };
    double rn = pk(FP_RENORM);
    if (SPECIES_RENORM) rn = feqmod_renorm_ptm_fused(pk, mass, deg, baryon, sign, nodes, gla_pts, exptab);
    const bool breakdown = pk(FP_BREAKDOWN) != 0.0;
    const double detA = pk(FP_DETA), eta_scale = pk(FP_ETA_SCALE);
    double acc = 0.0;
    for (int j0 = 0; j0 < npoints; j0 += kDndxThreads) {
      if (j0 + lane < npoints) {
        const DndxPoint pt = dndx_point(g, pk, j0 + lane);
        bool linear = breakdown;
        if (g.dimension == 3 && !linear && detA < 0.01 && fabs(pt.yval - pt.eta) < detA) linear = true;
        const double d = linear ? (pt.yval - pt.eta) : (pt.yval - eta_scale * pt.eta);
        const double sh = sinh(d), ch = cosh(d);
        if (linear) items[lane].lin = feqmod_make_linear_item(pk, sh, ch, pt.cphi, pt.sphi, pt.w, true);
        else items[lane].mod = feqmod_make_item(pk, sh, ch, pt.cphi, pt.sphi, pt.w, true);
        item_linear[lane] = linear ? 1 : 0;
      }
      __syncwarp();
      const int nj = min(kDndxThreads, npoints - j0);
      for (int ipT = 0; ipT < g.NpT; ipT++) {
        const DndxBin bn = dndx_load_bin(g, ipT, s, m2, baryon, sign);
        for (int k = 0; k < nj; k++) {
          double pds, v;
          if (item_linear[k]) {
            const DfItem it = items[k].lin;
            pds = fma(bn.mTw, it.c1, bn.pTw * it.d1);
            v = pds * df_distribution<2, BARYON, REGULATE, true>(it, df_share<BARYON>(it, bn.pT, bn.pT2), bn.b, exptab);
          } else {
            const FeqmodItem it = items[k].mod;
            pds = fma(bn.mTw, it.c1, bn.pTw * it.d1);
            v = pds * feqmod_distribution<BARYON>(it, feqmod_share(it, bn.pT, bn.pT2), bn.b, rn, exptab);
          }
          if (OUTFLOW) v = (pds <= 0.0) ? 0.0 : v;
          acc += v;
        }
      }
      __syncwarp();
    }
    // a NaN / inf renormalisation skips the (cell, species) in both branches (SpacetimeDistribution.cpp:955-959)
    if (s < g.ns && rn != 0.0) {
      const int64_t gc = surf_begin + cell;
      dndx_scatter(g, s, surf.col[0][gc], surf.col[1][gc], surf.col[2][gc], kCooperFryePrefactor * acc);
    }
  }
}

__global__ void dndx_expand_kernel(const double *__restrict__ class_hist, const int *__restrict__ class_of, const double *__restrict__ deg,
                                   int bins, int64_t total, double *__restrict__ out)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t s = i / bins;
  out[i] = deg[s] * class_hist[(int64_t)class_of[s] * bins + (i - s * bins)];
}

// transposed momentum tables [ipT][ns_pad] and padded species arrays
is3d_status build_dndx_grid(is3d_ctx *ctx, DndxGrid *g, const int **class_of_dev)
{
  const is3d_params &p = ctx->prm;
  // one thread per species CLASS (ctx.h SpeciesBins): the histograms are filled per class and expanded to species at the end
  std::vector<int> class_of, rep;
  species_classes(ctx, &class_of, &rep);
  const int ns = (int)rep.size(), nsp = (ns + kDndxThreads - 1) / kDndxThreads * kDndxThreads, npT = ctx->NpT;
  std::vector<double> h((size_t)4 * npT * nsp + 5 * nsp, 0.0);
  double *mTw = h.data(), *mT = mTw + (size_t)npT * nsp, *mT2 = mT + (size_t)npT * nsp, *mTpT = mT2 + (size_t)npT * nsp;
  double *mass2 = mTpT + (size_t)npT * nsp, *baryon = mass2 + nsp, *sign = baryon + nsp, *deg = sign + nsp, *mass = deg + nsp;
  for (int s = 0; s < nsp; s++) {
    int ss = rep[s < ns ? s : ns - 1];                    // class representative; padding lanes repeat the last class (never written)
    double m = ctx->h_mass[ss];
    mass[s] = m; mass2[s] = m * m; baryon[s] = ctx->h_baryon[ss]; sign[s] = ctx->h_sign[ss]; deg[s] = ctx->h_deg[ss];
    for (int ip = 0; ip < npT; ip++) {
      double pT = ctx->pT[ip], v = sqrt(m * m + pT * pT);
      size_t idx = (size_t)ip * nsp + s;
      mT[idx] = v; mTw[idx] = v * ctx->pTw[ip]; mT2[idx] = v * v; mTpT[idx] = v * pT;
    }
  }
  void *d = nullptr, *dm = nullptr;
  IS3D_TRY(ctx->get_scratch("dndx_tables", h.size() * sizeof(double), &d));
  IS3D_TRY(ctx->get_scratch("class_of", (size_t)ctx->ns * sizeof(int), &dm));
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(dm, class_of.data(), (size_t)ctx->ns * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  *class_of_dev = (const int *)dm;
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const double *b = (const double *)d;
  g->ns = ns; g->ns_pad = nsp; g->NpT = npT; g->Nphi = ctx->Nphi; g->Ny = ctx->Ny; g->Neta = ctx->Neta; g->dimension = p.dimension;
  g->mTw = b; g->mT = b + (size_t)npT * nsp; g->mT2 = b + (size_t)2 * npT * nsp; g->mTpT = b + (size_t)3 * npT * nsp;
  const double *tail = b + (size_t)4 * npT * nsp;
  g->mass2 = tail; g->baryon = tail + nsp; g->sign = tail + 2 * nsp; g->deg = tail + 3 * nsp; g->mass = tail + 4 * nsp;
  g->pT = ctx->d_pT; g->pTw = ctx->d_pTw;
  g->cosphi = ctx->d_cosphi; g->sinphi = ctx->d_sinphi; g->phiw = ctx->d_phiw;
  g->yv = ctx->d_y; g->etav = ctx->d_eta; g->etaw = ctx->d_etaw;
  g->tau_min = p.tau_min; g->tau_width = (p.tau_max - p.tau_min) / (double)p.tau_bins; g->tau_bins = p.tau_bins;
  g->r_min = p.r_min; g->r_width = (p.r_max - p.r_min) / (double)p.r_bins; g->r_bins = p.r_bins;
  g->phi_width = kTwoPi / (double)p.phip_bins; g->phi_bins = p.phip_bins;
  return IS3D_OK;
}

}  // namespace

// hist_*_dev: device buffers of ns x bins doubles (zeroed here)
is3d_status run_dndx(is3d_ctx *ctx, double *tau_dev, double *r_dev, double *phi_dev, is3d_stats *stats)
{
  const is3d_params &p = ctx->prm;
  if (p.df_mode == 5) { ctx->set_error("no spacetime distribution routine for famod yet (reference EmissionFunction.cpp:1184-1189)"); return IS3D_ERR_UNSUPPORTED; }
  const bool feqmod = (p.df_mode == 3 || p.df_mode == 4);
  if (feqmod && ctx->gla_pts <= 0) { ctx->set_error("Gauss-Laguerre tables not set"); return IS3D_ERR_INVALID; }
  if (feqmod && (ctx->gla_pts > kRenormMaxPts || ctx->gla_alpha < 3)) {
    ctx->set_error("Gauss-Laguerre tables: need alpha = 0..2 with at most 64 points");
    return IS3D_ERR_INVALID;
  }
  const int64_t n = ctx->surf.n;
  DndxGrid g;
  const int *class_of_dev = nullptr;
  IS3D_TRY(build_dndx_grid(ctx, &g, &class_of_dev));
  const size_t class_bins = (size_t)g.ns * (p.tau_bins + p.r_bins + p.phip_bins);
  void *class_hist = nullptr;
  IS3D_TRY(ctx->get_scratch("dndx_class_hist", class_bins * sizeof(double), &class_hist));
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(class_hist, 0, class_bins * sizeof(double), ctx->stream));
  g.hist_tau = (double *)class_hist; g.hist_r = g.hist_tau + (size_t)g.ns * p.tau_bins; g.hist_phi = g.hist_r + (size_t)g.ns * p.r_bins;
  g.exptab = ctx->d_exptab;

  DfFlags dfl;
  dfl.df_mode = p.df_mode; dfl.dimension = p.dimension; dfl.include_baryon = p.include_baryon;
  dfl.include_bulk = p.include_bulk_deltaf; dfl.include_shear = p.include_shear_deltaf; dfl.include_baryondiff = p.include_baryondiff_deltaf;
  FeqmodFlags ffl;
  ffl.df_mode = p.df_mode; ffl.dimension = p.dimension; ffl.include_baryon = p.include_baryon;
  ffl.include_bulk = p.include_bulk_deltaf; ffl.include_shear = p.include_shear_deltaf; ffl.include_baryondiff = p.include_baryondiff_deltaf;
  ffl.deta_min = p.deta_min; ffl.mass_pion0 = p.mass_pion0; ffl.bulkPi_over_P_max = ctx->tb.bulkPi_over_P_max;
  ffl.clamp_inclusive = 1;
  const bool species_renorm = (p.df_mode == 3 && p.include_bulk_deltaf);
  const int pack_size = feqmod ? (int)FP_SIZE : (int)DP_SIZE;

  const int64_t macro = pass_cells(2 << 20);
  const int64_t stride = n < macro ? n : macro;
  void *pack = nullptr, *counters = nullptr;
  IS3D_TRY(ctx->get_scratch("cell_pack", (size_t)pack_size * stride * sizeof(double), &pack));
  IS3D_TRY(ctx->get_scratch("counters", 16 * sizeof(unsigned long long), &counters));
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(counters, 0, 16 * sizeof(unsigned long long), ctx->stream));

  const int nslices = g.ns_pad / kDndxThreads;
  // ~8 waves of blocks: one-warp blocks, up to 16 resident per SM
  int64_t want_blocks = 8LL * 16 * ctx->sm_count;
  cudaEvent_t e0, e1;
  IS3D_CUDA_TRY(ctx, cudaEventCreate(&e0));
  IS3D_CUDA_TRY(ctx, cudaEventCreate(&e1));
  float ms_total = 0.f;
  int64_t launches = 0;
  const bool reg = p.regulate_deltaf != 0, outflow = p.outflow != 0, baryon = p.include_baryon != 0;
  for (int64_t begin = 0; begin < n; begin += macro) {
    int64_t count = n - begin < macro ? n - begin : macro;
    int64_t nchunks = (want_blocks + nslices - 1) / nslices;
    if (nchunks > count) nchunks = count;
    if (nchunks > 65535) nchunks = 65535;
    int64_t cpb = (count + nchunks - 1) / nchunks;
    nchunks = (count + cpb - 1) / cpb;
    dim3 grid(nslices, (unsigned)nchunks);
    IS3D_CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
    if (!feqmod) {
      dndx_df_setup_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(ctx->surf, begin, count, ctx->tb, dfl, (double *)pack, stride, (unsigned long long *)counters);
      IS3D_CUDA_TRY(ctx, cudaGetLastError());
#define IS3D_DNDX_DF(M, B, R, O) dndx_df_kernel<M, B, R, O><<<grid, kDndxThreads, 0, ctx->stream>>>((double *)pack, stride, count, cpb, ctx->surf, begin, g)
#define IS3D_DNDX_DF2(M, B) do { if (reg && outflow) IS3D_DNDX_DF(M, B, true, true); else if (reg) IS3D_DNDX_DF(M, B, true, false); else if (outflow) IS3D_DNDX_DF(M, B, false, true); else IS3D_DNDX_DF(M, B, false, false); } while (0)
      if (p.df_mode == 1) { if (baryon) IS3D_DNDX_DF2(1, true); else IS3D_DNDX_DF2(1, false); }
      else { if (baryon) IS3D_DNDX_DF2(2, true); else IS3D_DNDX_DF2(2, false); }
#undef IS3D_DNDX_DF2
#undef IS3D_DNDX_DF
    } else {
      dndx_feqmod_setup_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(ctx->surf, begin, count, ctx->tb, ffl, ctx->d_gla_root, ctx->d_gla_weight, ctx->gla_pts, (double *)pack, stride, (unsigned long long *)counters);
      IS3D_CUDA_TRY(ctx, cudaGetLastError());
#define IS3D_DNDX_FM(B, R, O, S) dndx_feqmod_kernel<B, R, O, S><<<grid, kDndxThreads, 0, ctx->stream>>>((double *)pack, stride, count, cpb, ctx->surf, begin, g, ctx->d_gla_root, ctx->d_gla_weight, ctx->gla_pts)
#define IS3D_DNDX_FM2(B, S) do { if (reg && outflow) IS3D_DNDX_FM(B, true, true, S); else if (reg) IS3D_DNDX_FM(B, true, false, S); else if (outflow) IS3D_DNDX_FM(B, false, true, S); else IS3D_DNDX_FM(B, false, false, S); } while (0)
      if (baryon) { if (species_renorm) IS3D_DNDX_FM2(true, true); else IS3D_DNDX_FM2(true, false); }
      else { if (species_renorm) IS3D_DNDX_FM2(false, true); else IS3D_DNDX_FM2(false, false); }
#undef IS3D_DNDX_FM2
#undef IS3D_DNDX_FM
    }
    IS3D_CUDA_TRY(ctx, cudaGetLastError());
    IS3D_CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaEventSynchronize(e1));
    float ms = 0.f;
    IS3D_CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
    ms_total += ms;
    launches += 2;
  }
  // species s = degeneracy_s x its class (SpacetimeDistribution.cpp:408: dN_dy_cell carries the degeneracy)
  const int bins3[3] = {p.tau_bins, p.r_bins, p.phip_bins};
  const double *src3[3] = {g.hist_tau, g.hist_r, g.hist_phi};
  double *dst3[3] = {tau_dev, r_dev, phi_dev};
  for (int k = 0; k < 3; k++) {
    const int64_t tot = (int64_t)ctx->ns * bins3[k];
    dndx_expand_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(src3[k], class_of_dev, ctx->d_deg, bins3[k], tot, dst3[k]);
    IS3D_CUDA_TRY(ctx, cudaGetLastError());
  }
  launches += 3;
  unsigned long long h_counters[16];
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (stats) {
    stats->cells_total = n;
    stats->cells_skipped = (int64_t)h_counters[0];
    stats->cells_out_of_table = (int64_t)h_counters[1];
    stats->cells_breakdown = (int64_t)h_counters[2];
    stats->cells_pl_negative = (int64_t)h_counters[3];
    stats->kernel_ms = ms_total;
    stats->kernel_launches = launches;
  }
  if (h_counters[1] != 0) {
    ctx->set_error(std::to_string(h_counters[1]) + " cell(s) outside the df coefficient tables (the reference aborts here)");
    return IS3D_ERR_TABLE_RANGE;
  }
  return IS3D_OK;
}

}  // namespace is3d

extern "C" {

is3d_status is3d_dndx_device(is3d_ctx *ctx, double *tau_dev, double *r_dev, double *phi_dev, is3d_stats *stats)
{
  if (!ctx) return IS3D_ERR_INVALID;
  IS3D_CUDA_TRY(ctx, cudaSetDevice(ctx->prm.device));
  if (!tau_dev || !r_dev || !phi_dev) { ctx->set_error("dndx: NULL output"); return IS3D_ERR_INVALID; }
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (ctx->ns <= 0 || !ctx->have_momentum || !ctx->have_surface || !ctx->have_df) { ctx->set_error("dndx: species / tables / surface not set"); return IS3D_ERR_INVALID; }
  if (ctx->prm.df_mode == 4 && !ctx->have_ptb) { ctx->set_error("PTB tables not set"); return IS3D_ERR_INVALID; }
  if (ctx->surf.n == 0) {
    IS3D_CUDA_TRY(ctx, cudaMemsetAsync(tau_dev, 0, (size_t)ctx->ns * ctx->prm.tau_bins * sizeof(double), ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaMemsetAsync(r_dev, 0, (size_t)ctx->ns * ctx->prm.r_bins * sizeof(double), ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaMemsetAsync(phi_dev, 0, (size_t)ctx->ns * ctx->prm.phip_bins * sizeof(double), ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return IS3D_OK;
  }
  return is3d::run_dndx(ctx, tau_dev, r_dev, phi_dev, stats);
}

is3d_status is3d_dndx(is3d_ctx *ctx, double *tau_hist, double *r_hist, double *phi_hist, is3d_stats *stats)
{
  if (!ctx) return IS3D_ERR_INVALID;
  IS3D_CUDA_TRY(ctx, cudaSetDevice(ctx->prm.device));
  if (!tau_hist || !r_hist || !phi_hist) { ctx->set_error("dndx: NULL output"); return IS3D_ERR_INVALID; }
  const size_t nt = (size_t)ctx->ns * ctx->prm.tau_bins, nr = (size_t)ctx->ns * ctx->prm.r_bins, np = (size_t)ctx->ns * ctx->prm.phip_bins;
  void *d = nullptr;
  IS3D_TRY(ctx->get_scratch("dndx_hist", (nt + nr + np + 3) * sizeof(double), &d));
  double *dt = (double *)d, *dr = dt + nt, *dp = dr + nr;
  IS3D_TRY(is3d_dndx_device(ctx, dt, dr, dp, stats));
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(tau_hist, dt, nt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(r_hist, dr, nr * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(phi_hist, dp, np * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return IS3D_OK;
}

}  // extern "C"
