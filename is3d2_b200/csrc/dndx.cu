// K4: mean spacetime distributions dN/dX (operation 0) for df_mode 1-4 on sm_100a.
// Replaces EmissionFunctionArray::calculate_dN_dX and calculate_dN_dX_feqmod
// (reference src/cpp/SpacetimeDistribution.cpp:31-517, :520-1246).  See dndx_common.cuh for the mapping.
#include <cstring>

#include "ctx.h"
#include "dndx_common.cuh"
#include "spectra_feqmod.cuh"

namespace is3d {

is3d_status build_bin_arrays(is3d_ctx *ctx, SpeciesBins *out);
bool build_slot_table(const is3d_ctx *ctx, int R, std::vector<int> *slots);
bool pair_tables_core(const std::vector<int> &rep, const double *mass, const double *sign, const double *baryon, int R, int R_pair,
                      std::vector<int> *singles, std::vector<int> *pairs);
void pick_chunks(int64_t ncells, int64_t blocks_per_chunk, int64_t resident, int64_t granule, int64_t max_chunks,
                 int *nchunks, int64_t *cells_per_chunk);

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

// thread constants: column (thread group, pT node) of the block, see dndx_common.cuh
// PAIR: a slot is a charge-conjugate pair of classes (two class ids per slot; spectra_df.cu pair_tables_core): the thread's
// constants are those of the b > 0 member, the partner differs by b -> -b only
template <int R, bool PAIR = false>
struct DndxThread {
  bool active;                 // the column exists (group < ngroups, pT node < NpT)
  int gl, ip;                  // group inside the block, pT node
  double pT, pT2, wpT, b;      // b = the group's baryon number
  double mT[R], mT2[R], sgn[R];
  int eslot;

  __device__ __forceinline__ void load(const DndxGrid &g, bool baryon_on)
  {
    constexpr int S = PAIR ? 2 : 1;
    const int t = threadIdx.x;
    int gl_ = t / g.NpT;
    ip = t - gl_ * g.NpT;
    int grp = blockIdx.x * g.gpb + gl_;
    active = gl_ < g.gpb && grp < g.ngroups;
    if (!active) { gl_ = 0; grp = blockIdx.x * g.gpb; }          // idle threads shadow a valid column and are never summed
    gl = gl_;
    const int cls0 = g.slot_class[S * grp * R];                    // slot 0 of a group is never padding
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int cls = g.slot_class[S * (grp * R + r)];
      const int jj = (cls >= 0 ? cls : cls0) * g.NpT + ip;
      const double m = g.mT[jj];
      mT[r] = m; mT2[r] = m * m; sgn[r] = g.sign[jj];
    }
    pT = g.pT[ip]; pT2 = pT * pT; wpT = g.pTw[ip];
    b = baryon_on ? g.baryon[cls0 * g.NpT + ip] : 0.0;
    eslot = kMaxBaryon + (int)b;
    asm volatile("" : "+r"(eslot));
  }
};

// End of a cell: thread partials (x pT weight x `gate`) -> sums over the NpT threads of every group -> histograms.
// red is double-buffered by the caller (one __syncthreads per cell); bins = the cell's histogram bins (shared memory).
// N = values per thread = class ids per thread group: R (single classes) or 2 R (pairs, acc[2 r] = the b > 0 member of slot r,
// acc[2 r + 1] = its partner -- the order of the flat slot list)
template <int N>
__device__ __forceinline__ void dndx_flush(double (&acc)[N], double factor, double (*red)[kDndxThreads], const DndxGrid &g,
                                           const DndxCellBins &bins, double dropped_bound = 0.0, const double *rn_row = nullptr)
{
  constexpr int R = N;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
#pragma unroll
  for (int r = 0; r < R; r++) { red[r][t] = acc[r] * factor; acc[r] = 0.0; }
  __syncthreads();
  const int nsum = g.gpb * R;
  for (int j = warp; j < nsum; j += kDndxThreads / 32) {
    const int gl = j / R, r = j - gl * R;
    const int grp = blockIdx.x * g.gpb + gl;
    if (grp >= g.ngroups) continue;
    const int cls = g.slot_class[grp * R + r];
    if (cls < 0) continue;
    double v = 0.0;
    for (int i = lane; i < g.NpT; i += 32) v += red[r][gl * g.NpT + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) {
      // a-posteriori test of the cell's dropped quadrature points: their summed bound (times the summed |pT weights|, without the
      // Cooper-Frye prefactor, like v) must vanish against the (cell, class) scalar itself; run_dndx repeats the call without the
      // margin otherwise
      // (rn_row: the PTM renormalisation of this class multiplies every term of the cell, the dropped ones included)
      const double bound = dropped_bound * (rn_row ? fabs(rn_row[gl * R + r]) : 1.0);
      if (bound != 0.0 && !(bound <= 1e-13 * fabs(v) + 1e-280)) atomicAdd(g.prune_counters, 1ull);
      dndx_scatter(g, cls, bins, kCooperFryePrefactor * v);
    }
  }
}

// tile geometry shared by both kernels: cells per tile and the (cell, point) a thread builds
struct DndxTiling {
  int npoints, cpt;
  __device__ __forceinline__ DndxTiling(const DndxGrid &g)
  {
    npoints = g.Ny * g.Neta * g.Nphi;
    cpt = npoints >= kDndxTile ? 1 : kDndxTile / npoints;
    if (cpt > kDndxMaxCells) cpt = kDndxMaxCells;
  }
};

// df_mode 1, 2 (SpacetimeDistribution.cpp:170-441)
template <int MODE, bool BARYON, bool REGULATE, bool OUTFLOW, bool PAIR>
__global__ void __launch_bounds__(kDndxThreads, 2)
dndx_df_kernel(const double *__restrict__ pack, int64_t stride, int64_t ncells, int64_t cells_per_block, SurfaceView surf,
               int64_t surf_begin, DndxGrid g)
{
  static_assert(!PAIR || BARYON, "pairs exist only with baryon terms");
  constexpr int R = PAIR ? kDndxPairR : kDndxR;       // slots per thread
  constexpr int N = PAIR ? 2 * R : R;                   // classes (accumulators) per thread
  __shared__ double exptab[kExpTableSize];
  __shared__ DfItemU items[kDndxTile];
  __shared__ double red[2][N][kDndxThreads];
  __shared__ int cell_ok[kDndxMaxCells];
  __shared__ DndxCellBins cell_bins[kDndxMaxCells];
  load_exp_table(exptab, g.exptab);
  const int t = threadIdx.x;
  DndxThread<R, PAIR> th;
  th.load(g, BARYON);
  DfThreadU tu;
  tu.pT = th.pT; tu.pT2 = th.pT2; tu.b = th.b; tu.bpT = th.b * th.pT; tu.eslot = th.eslot;
  DfThreadU tum = tu;                                   // the antibaryon partners of a pair slot
  if (PAIR) { tum.b = -tu.b; tum.bpT = -tu.bpT; tum.eslot = 2 * kMaxBaryon - tu.eslot; }
  double acc[N];
#pragma unroll
  for (int r = 0; r < N; r++) acc[r] = 0.0;
  const DndxTiling tl(g);
  int buf = 0;

  // Dropping of negligible quadrature points (the scheme of the spectra kernels, spectra_df.cu, per CELL): every exponent
  // x = xE - b alpha_B the block's columns (all pT nodes of its thread groups) can form for a point lies in [mT_lo (aT - max(bT, 0))
  // - 2|alpha_B|, mT_hi aT + pT_hi max(-bT, 0) + 2|alpha_B|].  A point whose lower bound is >= 680, or exceeds the smallest upper bound
  // among the cell's points of this tile by more than g.margin, is not marched over; its term bound (df_item_term_bound) is summed
  // per cell and tested against every (cell, class) scalar in dndx_flush.
  __shared__ double blk_range[3 * (kDndxThreads / 32)];
  __shared__ double blk_lohi[4];                         // mT_lo, mT_hi, pT_hi of the block's columns, sum of the pT weights
  __shared__ unsigned long long cell_xmin[kDndxMaxCells];
  __shared__ double cell_bound[kDndxMaxCells];
  __shared__ int cell_cnt[kDndxMaxCells];
  __shared__ unsigned long long blk_dropped;
  {
    const int lane = t & 31, warp = t >> 5;
    double lo = th.mT[0], hi = th.mT[0], ph = th.pT;
#pragma unroll
    for (int r = 1; r < R; r++) { lo = fmin(lo, th.mT[r]); hi = fmax(hi, th.mT[r]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      ph = fmax(ph, __shfl_xor_sync(0xffffffffu, ph, o));
    }
    if (lane == 0) { blk_range[3 * warp] = lo; blk_range[3 * warp + 1] = hi; blk_range[3 * warp + 2] = ph; }
    __syncthreads();
    if (t == 0) {
      for (int w = 1; w < kDndxThreads / 32; w++) { lo = fmin(lo, blk_range[3 * w]); hi = fmax(hi, blk_range[3 * w + 1]); ph = fmax(ph, blk_range[3 * w + 2]); }
      double wsum = 0.0;
      for (int i = 0; i < g.NpT; i++) wsum += fabs(g.pTw[i]);
      blk_lohi[0] = lo; blk_lohi[1] = hi; blk_lohi[2] = ph; blk_lohi[3] = wsum;
      blk_dropped = 0;
    }
    __syncthreads();
  }
  constexpr unsigned long long kHuge = 0x7f7f7f7f7f7f7f7full;

  const int64_t c0 = (int64_t)blockIdx.y * cells_per_block;
  int64_t c1 = c0 + cells_per_block;
  if (c1 > ncells) c1 = ncells;
  for (int64_t cell0 = c0; cell0 < c1; cell0 += tl.cpt) {
    for (int p0 = 0; p0 < tl.npoints; p0 += kDndxTile) {
      const int np_tile = min(kDndxTile, tl.npoints - p0);
      __syncthreads();                                   // previous tile consumed (first pass: exp table loaded)
      if (t < kDndxMaxCells) { cell_cnt[t] = 0; cell_xmin[t] = kHuge; if (p0 == 0) cell_bound[t] = 0.0; }
      __syncthreads();
      const int cl = t / np_tile, j = p0 + t - cl * np_tile;
      const int64_t cell = cell0 + cl;
      bool mine = false;
      double sh = 0.0, ch = 1.0, xe_lo = 0.0, xe_hi = 0.0, cphi = 1.0, sphi = 0.0, w = 1.0;
      auto pk = [&](int k) { return pack[k * stride + cell]; };
      if (cl < tl.cpt && cell < c1) {
        const bool ok = pack[DP_VALID * stride + cell] != 0.0;
        if (j == p0) {
          cell_ok[cl] = ok ? 1 : 0;
          const int64_t gc = surf_begin + cell;
          if (ok) cell_bins[cl] = dndx_cell_bins(g, surf.col[0][gc], surf.col[1][gc], surf.col[2][gc]);
        }
        if (ok) {
          mine = true;
          const DndxPoint pt = dndx_point(g, pk, j);
          const double d = pt.yval - pt.eta;
          sh = sinh(d); ch = cosh(d); cphi = pt.cphi; sphi = pt.sphi; w = pt.w;
          // aT, bT exactly as df_make_item_u forms them
          const double aT = ch * pk(DP_UTT) - sh * pk(DP_TUNT), bT = cphi * pk(DP_UXT) + sphi * pk(DP_UYT);
          const double shift = BARYON ? kMaxBaryon * fabs(pk(DP_ALPHAB)) : 0.0;
          const volatile double *range = blk_lohi;
          xe_lo = range[0] * (aT - fmax(bT, 0.0)) - shift;
          xe_hi = fma(range[1], aT, range[2] * fmax(-bT, 0.0)) + shift;
          if (xe_hi > 0.0) atomicMin(&cell_xmin[cl], (unsigned long long)__double_as_longlong(xe_hi));
        }
      }
      __syncthreads();
      if (mine) {
        double thr = 680.0;
        if (g.margin > 0.0) thr = fmin(thr, __longlong_as_double((long long)cell_xmin[cl]) + g.margin);
        const DfItemU item = df_make_item_u<MODE, BARYON, (MODE == 2 && !REGULATE)>(pk, sh, ch, cphi, sphi, w);
        if (xe_lo >= thr) {                               // NaN: kept
          const volatile double *range = blk_lohi;
          const double shift = BARYON ? kMaxBaryon * fabs(pk(DP_ALPHAB)) : 0.0;
          atomicAdd(&cell_bound[cl], range[3] * df_item_term_bound<MODE, BARYON, REGULATE>(item, xe_lo + shift, range[1], range[2], exptab));
        } else {
          items[cl * np_tile + atomicAdd(&cell_cnt[cl], 1)] = item;
        }
      }
      __syncthreads();
      const bool last_tile = p0 + kDndxTile >= tl.npoints;
      if (t == 0)
        for (int c = 0; c < tl.cpt && cell0 + c < c1; c++) if (cell_ok[c]) blk_dropped += (unsigned)(np_tile - cell_cnt[c]);
      for (int c = 0; c < tl.cpt && cell0 + c < c1; c++) {
        if (!cell_ok[c]) continue;
        if (th.active) {
          const int n_kept = cell_cnt[c];
#pragma unroll 1
          for (int k = 0; k < n_kept; k++) {
            const DfItemU &it = items[c * np_tile + k];
            const DfSharedU shd = df_share_u<MODE, BARYON, (MODE == 2 && !REGULATE)>(it, tu);
            if (!PAIR) {
#pragma unroll
              for (int r = 0; r < R; r++) acc[r] += df_eval_u<MODE, BARYON, REGULATE, OUTFLOW>(it, shd, th.mT[r], th.mT2[r], th.sgn[r], exptab);
            } else {
              const DfSharedU shm = df_share_u<MODE, BARYON, (MODE == 2 && !REGULATE)>(it, tum);
#pragma unroll
              for (int r = 0; r < R; r++) {           // one x_E and one exponential per pair (spectra_df.cuh)
                const double xE = df_eval_u_x(it, shd, th.mT[r]);
                const double e = fast_exp<false>(xE, exptab);
                acc[2 * r] += df_eval_u_tail<MODE, BARYON, REGULATE, OUTFLOW>(it, shd, th.mT[r], th.mT2[r], th.sgn[r], xE, e);
                acc[2 * r + 1] += df_eval_u_tail<MODE, BARYON, REGULATE, OUTFLOW>(it, shm, th.mT[r], th.mT2[r], th.sgn[r], xE, e);
              }
            }
          }
        }
        if (last_tile) { dndx_flush<N>(acc, th.wpT, red[buf], g, cell_bins[c], cell_bound[c]); buf ^= 1; }
      }
    }
  }
  __syncthreads();
  if (t == 0 && blk_dropped) atomicAdd(g.prune_counters + 1, blk_dropped);
}

union DndxItemSlot {
  DfItem lin;
  FeqmodItem mod;
  __device__ DndxItemSlot() {}
};

// df_mode 3, 4 (SpacetimeDistribution.cpp:676-1160)
template <bool BARYON, bool REGULATE, bool OUTFLOW, bool SPECIES_RENORM, bool PAIR>
__global__ void __launch_bounds__(kDndxThreads, 2)
dndx_feqmod_kernel(const double *__restrict__ pack, int64_t stride, int64_t ncells, int64_t cells_per_block, SurfaceView surf,
                   int64_t surf_begin, DndxGrid g, const double *__restrict__ gla_root, const double *__restrict__ gla_weight,
                   int gla_pts)
{
  static_assert(!PAIR || BARYON, "pairs exist only with baryon terms");
  constexpr int R = PAIR ? kDndxPairR : kDndxR;       // slots per thread
  constexpr int N = PAIR ? 2 * R : R;                   // classes (accumulators, renorm entries) per thread
  __shared__ double exptab[kExpTableSize];
  __shared__ DndxItemSlot items[kDndxTile];
  __shared__ unsigned char item_linear[kDndxTile];
  __shared__ double red[2][N][kDndxThreads];
  __shared__ double cell_rn[kDndxMaxCells];                                   // the cell's |renorm| (0: skip the cell)
  __shared__ int cell_lin[kDndxMaxCells];                                     // the cell may hold linear-df items (breakdown / window)
  __shared__ double class_rn[SPECIES_RENORM ? kDndxMaxCells : 1][kDndxMaxGroups * N];   // PTM: |renorm| per (cell, class of the flat slot list)
  __shared__ int cell_ok[kDndxMaxCells];
  __shared__ DndxCellBins cell_bins[kDndxMaxCells];
  __shared__ RenormNodes nodes;
  load_exp_table(exptab, g.exptab);
  if (SPECIES_RENORM) nodes.load(gla_root, gla_weight, gla_pts);
  __syncthreads();
  const int t = threadIdx.x;
  DndxThread<R, PAIR> th;
  th.load(g, BARYON);
  const int eslotm = 2 * kMaxBaryon - th.eslot;         // the antibaryon partners of a pair slot
  DfBin bin[R];                                         // linear-df fallback of breakdown cells
#pragma unroll
  for (int r = 0; r < R; r++) { bin[r].mT = th.mT[r]; bin[r].mT2 = th.mT2[r]; bin[r].baryon = th.b; bin[r].sign = th.sgn[r]; }
  double acc[N];
#pragma unroll
  for (int r = 0; r < N; r++) acc[r] = 0.0;
  const DndxTiling tl(g);
  const int nsum = g.gpb * N;
  int buf = 0;

  // dropping of negligible quadrature points as in dndx_df_kernel, for cells whose points all take the modified distribution
  // (feqmod_item_range / feqmod_item_term_bound, spectra_feqmod.cuh); cells that may hold linear-df items keep every point
  __shared__ double blk_range[4 * (kDndxThreads / 32)];
  __shared__ double blk_lohi[5];                         // mT_lo, mT_hi, pT_hi, m2_lo of the block's columns, sum of the pT weights
  __shared__ unsigned long long cell_xmin[kDndxMaxCells];
  __shared__ double cell_bound[kDndxMaxCells];
  __shared__ int cell_cnt[kDndxMaxCells];
  __shared__ unsigned long long blk_dropped;
  {
    const int lane = t & 31, warp = t >> 5;
    double lo = th.mT[0], hi = th.mT[0], ph = th.pT, m2 = th.mT2[0] - th.pT2;
#pragma unroll
    for (int r = 1; r < R; r++) { lo = fmin(lo, th.mT[r]); hi = fmax(hi, th.mT[r]); m2 = fmin(m2, th.mT2[r] - th.pT2); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      ph = fmax(ph, __shfl_xor_sync(0xffffffffu, ph, o));
      m2 = fmin(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    }
    if (lane == 0) { blk_range[4 * warp] = lo; blk_range[4 * warp + 1] = hi; blk_range[4 * warp + 2] = ph; blk_range[4 * warp + 3] = m2; }
    __syncthreads();
    if (t == 0) {
      for (int w = 1; w < kDndxThreads / 32; w++) {
        lo = fmin(lo, blk_range[4 * w]); hi = fmax(hi, blk_range[4 * w + 1]); ph = fmax(ph, blk_range[4 * w + 2]); m2 = fmin(m2, blk_range[4 * w + 3]);
      }
      double wsum = 0.0;
      for (int i = 0; i < g.NpT; i++) wsum += fabs(g.pTw[i]);
      blk_lohi[0] = lo; blk_lohi[1] = hi; blk_lohi[2] = ph; blk_lohi[3] = fmax(m2 * (1.0 - 1e-12) - 1e-12, 0.0); blk_lohi[4] = wsum;
      blk_dropped = 0;
    }
    __syncthreads();
  }
  constexpr unsigned long long kHuge = 0x7f7f7f7f7f7f7f7full;

  const int64_t c0 = (int64_t)blockIdx.y * cells_per_block;
  int64_t c1 = c0 + cells_per_block;
  if (c1 > ncells) c1 = ncells;
  for (int64_t cell0 = c0; cell0 < c1; cell0 += tl.cpt) {
    for (int p0 = 0; p0 < tl.npoints; p0 += kDndxTile) {
      const int np_tile = min(kDndxTile, tl.npoints - p0);
      __syncthreads();                                   // previous tile consumed
      if (t < kDndxMaxCells) { cell_cnt[t] = 0; cell_xmin[t] = kHuge; if (p0 == 0) cell_bound[t] = 0.0; }
      __syncthreads();
      bool mine_mod = false;                             // this thread holds a modified-distribution point of a valid cell
      double sh_ = 0.0, ch_ = 1.0, x_lo = 0.0;
      const int cl_ = t / np_tile, j_ = p0 + t - cl_ * np_tile;
      const int64_t cell_ = cell0 + cl_;
      {
        const int cl = cl_, j = j_;
        const int64_t cell = cell_;
        if (cl < tl.cpt && cell < c1) {
          const bool ok = pack[DP_VALID * stride + cell] != 0.0;
          auto pk = [&](int k) { return pack[k * stride + cell]; };
          if (j == p0) {
            cell_ok[cl] = ok ? 1 : 0; cell_rn[cl] = ok ? pk(FP_RENORM) : 0.0;
            cell_lin[cl] = (ok && (pk(FP_BREAKDOWN) != 0.0 || (g.dimension == 3 && pk(FP_DETA) < 0.01))) ? 1 : 0;
            const int64_t gc = surf_begin + cell;
            if (ok) cell_bins[cl] = dndx_cell_bins(g, surf.col[0][gc], surf.col[1][gc], surf.col[2][gc]);
          }
          if (ok) {
            const DndxPoint pt = dndx_point(g, pk, j);
            bool linear = pk(FP_BREAKDOWN) != 0.0;
            const double detA = pk(FP_DETA);
            if (g.dimension == 3 && !linear && detA < 0.01 && fabs(pt.yval - pt.eta) < detA) linear = true;
            const double d = linear ? (pt.yval - pt.eta) : (pt.yval - pk(FP_ETA_SCALE) * pt.eta);
            const double sh = sinh(d), ch = cosh(d);
            item_linear[t] = linear ? 1 : 0;
            if (linear) {
              items[t].lin = feqmod_make_linear_item(pk, sh, ch, pt.cphi, pt.sphi, pt.w, true);   // its cell keeps every point in place
            } else {
              mine_mod = true; sh_ = sh; ch_ = ch;
              const FeqmodItem item = feqmod_make_item(pk, sh, ch, pt.cphi, pt.sphi, pt.w, true, BARYON, !SPECIES_RENORM);
              const volatile double *range = blk_lohi;
              const FeqmodRange x = feqmod_item_range(item, pk(FP_IT2), range[0], range[1], range[2], range[3],
                                                      BARYON ? kMaxBaryon * fabs(item.alphaB_mod) : 0.0);
              x_lo = x.lo;
              if (x.hi > 0.0) atomicMin(&cell_xmin[cl], (unsigned long long)__double_as_longlong(x.hi));
            }
          }
        }
        // PTM with bulk: n_linear / n_mod of every (cell, class slot) of this tile, once per cell (first tile of the cell)
        if (SPECIES_RENORM && p0 == 0) {
          for (int task = t; task < tl.cpt * nsum; task += kDndxThreads) {
            const int tc = task / nsum, js = task - tc * nsum;
            const int64_t rcell = cell0 + tc;
            const int grp = blockIdx.x * g.gpb + js / N;
            double rn = 0.0;
            if (rcell < c1 && grp < g.ngroups && pack[DP_VALID * stride + rcell] != 0.0) {
              const int cls = g.slot_class[grp * N + (js % N)];
              if (cls >= 0) {
                auto pkr = [&](int k) { return pack[k * stride + rcell]; };
                rn = feqmod_renorm_ptm_fused(pkr, g.c_mass[cls], g.c_deg[cls], g.c_baryon[cls], g.c_sign[cls], nodes, gla_pts, exptab);
              }
            }
            class_rn[tc][js] = rn;
          }
        }
      }
      __syncthreads();
      if (mine_mod) {
        const int cl = cl_;
        const int64_t cell = cell_;
        auto pk = [&](int k) { return pack[k * stride + cell]; };
        const DndxPoint pt = dndx_point(g, pk, j_);
        const FeqmodItem item = feqmod_make_item(pk, sh_, ch_, pt.cphi, pt.sphi, pt.w, true, BARYON, !SPECIES_RENORM);
        if (cell_lin[cl]) {
          items[t].mod = item;                           // a cell with linear-df points: every point stays in place
        } else {
          double thr = 680.0;
          if (g.margin > 0.0) thr = fmin(thr, __longlong_as_double((long long)cell_xmin[cl]) + g.margin);
          if (x_lo >= thr) {                             // NaN: kept
            const volatile double *range = blk_lohi;
            atomicAdd(&cell_bound[cl], range[4] * feqmod_item_term_bound(item, x_lo, range[1], range[2], 1.0, exptab));
          } else {
            items[cl * np_tile + atomicAdd(&cell_cnt[cl], 1)].mod = item;
          }
        }
      }
      __syncthreads();
      const bool last_tile = p0 + kDndxTile >= tl.npoints;
      if (t == 0)
        for (int c = 0; c < tl.cpt && cell0 + c < c1; c++) if (cell_ok[c] && !cell_lin[c]) blk_dropped += (unsigned)(np_tile - cell_cnt[c]);
      for (int cl = 0; cl < tl.cpt && cell0 + cl < c1; cl++) {
        if (!cell_ok[cl]) continue;
        double rn[N];                                    // index = accumulator index (pairs: 2 r = the b > 0 member, 2 r + 1 = its partner)
#pragma unroll
        for (int r = 0; r < N; r++) rn[r] = SPECIES_RENORM ? class_rn[cl][th.gl * N + r] : 1.0;
        // one modified-distribution item for the thread's slots
        auto modified = [&](const FeqmodItem &it) {
          const FeqmodShared sh = feqmod_share(it, th.pT, th.pT2);
          const double eb = BARYON ? it.eb[th.eslot] : 1.0;
          if (PAIR) {
            const double ebm = it.eb[eslotm];
#pragma unroll
            for (int r = 0; r < R; r++)
              feqmod_accum_pair_u<OUTFLOW, !SPECIES_RENORM>(acc[PAIR ? 2 * r : r], acc[PAIR ? 2 * r + 1 : r], it, sh, eb, ebm, th.mT[r], th.mT2[r],
                                                            th.sgn[r], rn[PAIR ? 2 * r : r], rn[PAIR ? 2 * r + 1 : r], exptab);
          } else {
#pragma unroll
            for (int r = 0; r < R; r++)
              feqmod_accum_u<BARYON, OUTFLOW, !SPECIES_RENORM>(acc[r], it, sh, eb, th.mT[r], th.mT2[r], th.sgn[r], rn[r], exptab);
          }
        };
        if (th.active && !cell_lin[cl]) {
          // common case, no linear-df item in this cell: lean loop over the points that were kept, the R evaluations interleave
          const int n_kept = cell_cnt[cl];
#pragma unroll 1
          for (int k = 0; k < n_kept; k++) modified(items[cl * np_tile + k].mod);
        } else if (th.active) {
#pragma unroll 1
          for (int k = 0; k < np_tile; k++) {
            const int slot = cl * np_tile + k;
            if (!item_linear[slot]) {
              modified(items[slot].mod);
            } else {
              const DfItem it = items[slot].lin;
              const DfShared sh = df_share<BARYON>(it, th.pT, th.pT2);
#pragma unroll
              for (int r = 0; r < R; r++) {
                acc[PAIR ? 2 * r : r] += df_eval<2, BARYON, REGULATE, OUTFLOW, true>(it, sh, bin[r], exptab);
                if (PAIR) {                             // the rare fallback items: the partner is evaluated on its own
                  DfBin bm = bin[r];
                  bm.baryon = -bm.baryon;
                  acc[PAIR ? 2 * r + 1 : r] += df_eval<2, BARYON, REGULATE, OUTFLOW, true>(it, sh, bm, exptab);
                }
              }
            }
          }
        }
        if (last_tile) {
          // a NaN / inf renormalisation (stored as 0) skips the (cell, species) in both branches (SpacetimeDistribution.cpp:955-959)
          if (SPECIES_RENORM) {
#pragma unroll
            for (int r = 0; r < N; r++) acc[r] = (rn[r] != 0.0) ? acc[r] : 0.0;
          }
          const double factor = (SPECIES_RENORM || cell_rn[cl] != 0.0) ? th.wpT : 0.0;
          dndx_flush<N>(acc, factor, red[buf], g, cell_bins[cl], cell_lin[cl] ? 0.0 : cell_bound[cl], SPECIES_RENORM ? class_rn[cl] : nullptr);
          buf ^= 1;
        }
      }
    }
  }
  __syncthreads();
  if (t == 0 && blk_dropped) atomicAdd(g.prune_counters + 1, blk_dropped);
}

__global__ void dndx_expand_kernel(const double *__restrict__ class_hist, const int *__restrict__ class_of, const double *__restrict__ deg,
                                   int bins, int64_t total, double *__restrict__ out)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t s = i / bins;
  out[i] = deg[s] * class_hist[(int64_t)class_of[s] * bins + (i - s * bins)];
}

// species classes, their (class, pT) bin arrays and the uniform-baryon slot table (shared with the spectra kernels)
// *g: the single-class launch; *gp: the charge-conjugate pair launch (gp->ngroups = 0 without baryon terms / without pairs),
// whose slot list holds two class ids per slot and kDndxPairR slots per thread group
is3d_status build_dndx_grid(is3d_ctx *ctx, DndxGrid *g, DndxGrid *gp, const int **class_of_dev)
{
  const is3d_params &p = ctx->prm;
  SpeciesBins sb;
  IS3D_TRY(build_bin_arrays(ctx, &sb));
  if (ctx->NpT > kDndxThreads) { ctx->set_error("dN/dX: pT table longer than 256 points"); return IS3D_ERR_UNSUPPORTED; }
  std::vector<int> slots, pair_slots;
  bool ok;
  if (p.include_baryon) {
    std::vector<int> class_of, rep;
    species_classes(ctx, &class_of, &rep);
    ok = pair_tables_core(rep, ctx->h_mass.data(), ctx->h_sign.data(), ctx->h_baryon.data(), kDndxR, kDndxPairR, &slots, &pair_slots);
  } else {
    ok = build_slot_table(ctx, kDndxR, &slots);
  }
  if (!ok) {
    ctx->set_error("species list holds a baryon number outside -2..2 (the reference's PDG readers produce hadrons and the deuteron only)");
    return IS3D_ERR_INVALID;
  }
  void *d_slots = nullptr;
  std::vector<int> both(slots);
  both.insert(both.end(), pair_slots.begin(), pair_slots.end());
  IS3D_TRY(ctx->get_scratch("dndx_slots", (both.size() + 1) * sizeof(int), &d_slots));
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(d_slots, both.data(), both.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));       // `both` is pageable host memory
  *class_of_dev = sb.class_of;
  g->ns = sb.nclass;
  g->ngroups = (int)(slots.size() / kDndxR);
  g->gpb = kDndxThreads / ctx->NpT;
  if (g->gpb > kDndxMaxGroups) g->gpb = kDndxMaxGroups;
  g->slot_class = (const int *)d_slots;
  g->NpT = ctx->NpT; g->Nphi = ctx->Nphi; g->Ny = ctx->Ny; g->Neta = ctx->Neta; g->dimension = p.dimension;
  g->mT = sb.mT; g->baryon = sb.baryon; g->sign = sb.sign;
  g->c_mass = sb.c_mass; g->c_deg = sb.c_deg; g->c_baryon = sb.c_baryon; g->c_sign = sb.c_sign;
  g->pT = ctx->d_pT; g->pTw = ctx->d_pTw;
  g->cosphi = ctx->d_cosphi; g->sinphi = ctx->d_sinphi; g->phiw = ctx->d_phiw;
  g->yv = ctx->d_y; g->etav = ctx->d_eta; g->etaw = ctx->d_etaw;
  g->tau_min = p.tau_min; g->tau_width = (p.tau_max - p.tau_min) / (double)p.tau_bins; g->tau_bins = p.tau_bins;
  g->r_min = p.r_min; g->r_width = (p.r_max - p.r_min) / (double)p.r_bins; g->r_bins = p.r_bins;
  g->phi_width = kTwoPi / (double)p.phip_bins; g->phi_bins = p.phip_bins;
  *gp = *g;
  gp->slot_class = (const int *)d_slots + slots.size();
  gp->ngroups = (int)(pair_slots.size() / (2 * kDndxPairR));
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
  DndxGrid g, gp;
  const int *class_of_dev = nullptr;
  IS3D_TRY(build_dndx_grid(ctx, &g, &gp, &class_of_dev));
  const size_t class_bins = (size_t)g.ns * (p.tau_bins + p.r_bins + p.phip_bins);
  void *class_hist = nullptr;
  IS3D_TRY(ctx->get_scratch("dndx_class_hist", class_bins * sizeof(double), &class_hist));
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(class_hist, 0, class_bins * sizeof(double), ctx->stream));
  g.hist_tau = (double *)class_hist; g.hist_r = g.hist_tau + (size_t)g.ns * p.tau_bins; g.hist_phi = g.hist_r + (size_t)g.ns * p.r_bins;
  g.exptab = ctx->d_exptab;
  gp.hist_tau = g.hist_tau; gp.hist_r = g.hist_r; gp.hist_phi = g.hist_phi; gp.exptab = g.exptab;

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

  const int nslices = (g.ngroups + g.gpb - 1) / g.gpb, nslices_pair = (gp.ngroups + gp.gpb - 1) / gp.gpb;
  const int64_t resident = 2LL * ctx->sm_count;      // two 256-thread blocks per SM (launch bounds)
  cudaEvent_t e0 = ctx->ev0, e1 = ctx->ev1;             // owned by the context: nothing to release on an error path
  float ms_total = 0.f;
  int64_t launches = 0, prune_reruns = 0;
  const bool reg = p.regulate_deltaf != 0, outflow = p.outflow != 0, baryon = p.include_baryon != 0;
  unsigned long long h_counters[16];
  // attempt 0 drops quadrature points below the margin; if the bound test fails for any (cell, class) scalar,
  // attempt 1 repeats the call without it (counters[4] = failures, [5] = points dropped)
  for (int attempt = 0; attempt < 2; attempt++) {
  g.margin = gp.margin = attempt == 0 ? p.negligible_margin : 0.0;
  g.prune_counters = (unsigned long long *)counters + 4;
  gp.prune_counters = (unsigned long long *)counters + 4;
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(class_hist, 0, class_bins * sizeof(double), ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(counters, 0, 16 * sizeof(unsigned long long), ctx->stream));
  for (int64_t begin = 0; begin < n; begin += macro) {
    int64_t count = n - begin < macro ? n - begin : macro;
    // whole waves of equal blocks (spectra_df.cu pick_chunks); a chunk = whole cells, in multiples of the cells of one tile
    int nchunks = 1;
    int64_t cpb = count;
    {
      const int npoints = g.Ny * g.Neta * g.Nphi;
      int64_t cpt = npoints >= kDndxTile ? 1 : kDndxTile / npoints;
      if (cpt > kDndxMaxCells) cpt = kDndxMaxCells;
      pick_chunks(count, nslices + nslices_pair, resident, cpt, 65535, &nchunks, &cpb);
    }
    dim3 grid(nslices, (unsigned)nchunks), grid_pair(nslices_pair, (unsigned)nchunks);
    IS3D_CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
    if (!feqmod) {
      dndx_df_setup_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(ctx->surf, begin, count, ctx->tb, dfl, (double *)pack, stride, (unsigned long long *)counters);
      IS3D_CUDA_TRY(ctx, cudaGetLastError());
      // charge-conjugate pairs first (longer blocks), then the single classes: disjoint class histograms
#define IS3D_DNDX_DF(M, B, R, O, P, GRID, G) dndx_df_kernel<M, B, R, O, P><<<GRID, kDndxThreads, 0, ctx->stream>>>((double *)pack, stride, count, cpb, ctx->surf, begin, G)
#define IS3D_DNDX_DF2(M, B, P, GRID, G) do { if (reg && outflow) IS3D_DNDX_DF(M, B, true, true, P, GRID, G); else if (reg) IS3D_DNDX_DF(M, B, true, false, P, GRID, G); else if (outflow) IS3D_DNDX_DF(M, B, false, true, P, GRID, G); else IS3D_DNDX_DF(M, B, false, false, P, GRID, G); } while (0)
      if (p.df_mode == 1) {
        if (baryon) { if (nslices_pair) IS3D_DNDX_DF2(1, true, true, grid_pair, gp); if (nslices) IS3D_DNDX_DF2(1, true, false, grid, g); }
        else IS3D_DNDX_DF2(1, false, false, grid, g);
      } else {
        if (baryon) { if (nslices_pair) IS3D_DNDX_DF2(2, true, true, grid_pair, gp); if (nslices) IS3D_DNDX_DF2(2, true, false, grid, g); }
        else IS3D_DNDX_DF2(2, false, false, grid, g);
      }
#undef IS3D_DNDX_DF2
#undef IS3D_DNDX_DF
    } else {
      dndx_feqmod_setup_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(ctx->surf, begin, count, ctx->tb, ffl, ctx->d_gla_root, ctx->d_gla_weight, ctx->gla_pts, (double *)pack, stride, (unsigned long long *)counters);
      IS3D_CUDA_TRY(ctx, cudaGetLastError());
#define IS3D_DNDX_FM(B, R, O, S, P, GRID, G) dndx_feqmod_kernel<B, R, O, S, P><<<GRID, kDndxThreads, 0, ctx->stream>>>((double *)pack, stride, count, cpb, ctx->surf, begin, G, ctx->d_gla_root, ctx->d_gla_weight, ctx->gla_pts)
#define IS3D_DNDX_FM2(B, S, P, GRID, G) do { if (reg && outflow) IS3D_DNDX_FM(B, true, true, S, P, GRID, G); else if (reg) IS3D_DNDX_FM(B, true, false, S, P, GRID, G); else if (outflow) IS3D_DNDX_FM(B, false, true, S, P, GRID, G); else IS3D_DNDX_FM(B, false, false, S, P, GRID, G); } while (0)
      if (baryon) {
        if (nslices_pair) { if (species_renorm) IS3D_DNDX_FM2(true, true, true, grid_pair, gp); else IS3D_DNDX_FM2(true, false, true, grid_pair, gp); }
        if (nslices) { if (species_renorm) IS3D_DNDX_FM2(true, true, false, grid, g); else IS3D_DNDX_FM2(true, false, false, grid, g); }
      } else {
        if (species_renorm) IS3D_DNDX_FM2(false, true, false, grid, g); else IS3D_DNDX_FM2(false, false, false, grid, g);
      }
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
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (h_counters[4] == 0 || h_counters[1] != 0 || !(g.margin > 0.0)) break;
  prune_reruns++;
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
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (stats) {
    stats->cells_total = n;
    stats->cells_skipped = (int64_t)h_counters[0];
    stats->cells_out_of_table = (int64_t)h_counters[1];
    stats->cells_breakdown = (int64_t)h_counters[2];
    stats->cells_pl_negative = (int64_t)h_counters[3];
    stats->kernel_ms = ms_total;
    stats->kernel_launches = launches;
    // quadrature points not marched over x the block's thread slots (kDndxThreads columns x 4 classes per thread in both launches)
    stats->evals_dropped = (int64_t)h_counters[5] * kDndxThreads * kDndxR;
    stats->prune_reruns = prune_reruns;
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
  const int64_t nt = (int64_t)ctx->ns * ctx->prm.tau_bins, nr = (int64_t)ctx->ns * ctx->prm.r_bins, np = (int64_t)ctx->ns * ctx->prm.phip_bins;
  is3d_status st = IS3D_OK;
  if (ctx->surf.n == 0) {
    IS3D_CUDA_TRY(ctx, cudaMemsetAsync(tau_dev, 0, (size_t)nt * sizeof(double), ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaMemsetAsync(r_dev, 0, (size_t)nr * sizeof(double), ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaMemsetAsync(phi_dev, 0, (size_t)np * sizeof(double), ctx->stream));
  } else {
    st = is3d::run_dndx(ctx, tau_dev, r_dev, phi_dev, stats);
  }
  // sharded surface: the three histograms are summed over the GPUs (a failed rank still joins, see is3d_spectra_device)
  is3d_status sc = is3d::comm_allreduce(ctx, tau_dev, nt);
  if (sc == IS3D_OK) sc = is3d::comm_allreduce(ctx, r_dev, nr);
  if (sc == IS3D_OK) sc = is3d::comm_allreduce(ctx, phi_dev, np);
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return st != IS3D_OK ? st : sc;
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
  if (ctx->prm.dndx_bug_compat) {
    // The reference clears its per-species accumulators with memset(ptr, 0.0, CORES * bins): `bins` BYTES, i.e. only
    // the first bins/8 doubles (SpacetimeDistribution.cpp:166-168, :671-673).  Every later bin keeps the previous
    // species' sum, so its files are cumulative over species there (and a bin cut in half by the byte count keeps
    // its upper 32 bits).  Serial build (CORES = 1) reproduced literally on the clean histograms.
    auto emulate = [&](double *h, int bins) {
      std::vector<double> all((size_t)bins, 0.0);
      for (int s = 0; s < ctx->ns; s++) {
        std::memset(all.data(), 0, (size_t)bins);
        for (int b = 0; b < bins; b++) { all[b] += h[(size_t)s * bins + b]; h[(size_t)s * bins + b] = all[b]; }
      }
    };
    emulate(tau_hist, ctx->prm.tau_bins);
    emulate(r_hist, ctx->prm.r_bins);
    emulate(phi_hist, ctx->prm.phip_bins);
  }
  return IS3D_OK;
}

}  // extern "C"
