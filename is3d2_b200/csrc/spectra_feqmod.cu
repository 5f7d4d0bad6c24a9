// K2: continuous Cooper-Frye spectra for df_mode 3 (PTM) and 4 (PTB) modified equilibrium distributions on sm_100a.
// Replaces EmissionFunctionArray::calculate_dN_pTdpTdphidy_feqmod (reference src/cpp/MomentumSpectra.cpp:419-1044).
//
// Schedule (same output-stationary design as K1, see spectra_df.cu):
//   1. feqmod_setup_kernel   one thread per cell: LRF boost, A_ij, refined inverse, breakdown test (two 32-point
//                            Gauss-Laguerre sums for PTM), linear-df fallback coefficients -> 57-double cell pack
//   2. feqmod_renorm_kernel  PTM with bulk only: one thread per (cell, species): n_linear / n_mod from four
//                            32-point Gauss-Laguerre sums -> renorm[cell][species]
//   3. feqmod_spectra_kernel blocks = ((thread group, pT) slice, (y, phi), cell chunk), a thread = 4 classes of one baryon
//                            number at one pT node (build_slot_table); per 256-cell tile every thread builds one cell's
//                            item for the block's (y, phi); the inner loop takes a warp-uniform branch per item:
//                            modified distribution (3 FMA + sqrt + exp + FMA + rcp, feqmod_accum_u) or linear-df fallback
//   4. reduce_partials_kernel (shared with K1)
#include <algorithm>

#include "ctx.h"
#include "spectra_feqmod.cuh"

namespace is3d {

__global__ void reduce_partials_kernel(const double *__restrict__ partial, int nchunks, int64_t total_class, int64_t per_species,
                                       const int *__restrict__ class_of, const double *__restrict__ deg, int64_t total,
                                       double *__restrict__ out, PruneCheck chk);
is3d_status build_bin_arrays(is3d_ctx *ctx, SpeciesBins *out);
bool build_slot_table(const is3d_ctx *ctx, int R, std::vector<int> *slots);
bool pair_tables_core(const std::vector<int> &rep, const double *mass, const double *sign, const double *baryon, int R, int R_pair,
                      std::vector<int> *singles, std::vector<int> *pairs);
void choose_chunks(const is3d_ctx *ctx, int64_t ncells, int64_t blocks_per_chunk, int64_t total, int tile, int blocks_per_sm,
                   int *nchunks, int64_t *cells_per_chunk);
// df_mode 5 per-cell stage (spectra_famod.cu): fills the same pack layout, counters[8] = reconstruction failures,
// counters[9] = Newton iterations
is3d_status famod_setup_pass(is3d_ctx *ctx, int64_t begin, int64_t count, double *pack, int64_t stride, unsigned long long *counters,
                             int64_t *launches, bool sampler_rules);

namespace {

// 128 threads x 5 blocks per SM with R = 3 classes per thread (<= 102 registers, 20 warps per SM): the modified-distribution loop waits
// on its table / renormalisation loads, so occupancy pays more than amortising the item loads over a fourth class
// (profiles/r02_k1_variants_notes.txt)
#ifndef IS3D_K2_THREADS
#define IS3D_K2_THREADS 128
#endif
#ifndef IS3D_K2_MINBLOCKS
#define IS3D_K2_MINBLOCKS 5
#endif
constexpr int kThreads = IS3D_K2_THREADS;
constexpr int kTile = kThreads;
#ifndef IS3D_K2_R
#define IS3D_K2_R 3
#endif
constexpr int kBins = IS3D_K2_R;     // species classes per thread (R)

__global__ void feqmod_setup_kernel(SurfaceView surf, int64_t begin, int64_t count, DfTables tb, FeqmodFlags fl,
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
  if (st & CELL_BREAKDOWN) { atomicAdd(&counters[2], 1ull); atomicMax(&counters[4], (unsigned long long)(begin + i + 1)); }
  if (st & CELL_PL_NEGATIVE) { atomicAdd(&counters[3], 1ull); atomicMax(&counters[5], (unsigned long long)(begin + i + 1)); }
}

// PTM renormalisation n_linear / n_mod per (cell, class) (MomentumSpectra.cpp:795-826): feqmod_renorm_ptm_fused.
// Rows are written in SLOT order, renorm[cell][group * R + r] (ns = number of slots, padding slots = 0), so that a thread
// of the spectra kernel reads the R values of its group as one aligned 32-byte segment.
__global__ void __launch_bounds__(128)
feqmod_renorm_kernel(const double *__restrict__ pack, int64_t stride, int64_t count, int ns, const int *__restrict__ slot_class,
                     const double *__restrict__ mass, const double *__restrict__ deg,
                     const double *__restrict__ baryon, const double *__restrict__ sign,
                     const double *__restrict__ gla_root, const double *__restrict__ gla_weight, int gla_pts,
                     const double *__restrict__ exptab_g, double *__restrict__ renorm)
{
  __shared__ double exptab[kExpTableSize];
  __shared__ RenormNodes nodes;
  load_exp_table(exptab, exptab_g);
  nodes.load(gla_root, gla_weight, gla_pts);
  __syncthreads();
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count * ns) return;
  int64_t cell = idx / ns;
  const int s = slot_class[(int)(idx - cell * ns)];
  double r = 0.0;
  if (s >= 0 && pack[DP_VALID * stride + cell] != 0.0) {
    auto pk = [&](int k) { return pack[k * stride + cell]; };
    r = feqmod_renorm_ptm_fused(pk, mass[s], deg[s], baryon[s], sign[s], nodes, gla_pts, exptab);
  }
  renorm[idx] = r;
}

struct FeqGrid {
  const double *mT, *pT, *baryon, *sign;
  int ncols, NpT, ns, nslots;           // ns = number of species CLASSES; ncols = NpT * ngroups thread columns; nslots = ngroups * kBins
  const int *slot_class;                // [ngroups * kBins]: class of slot r of a thread group (-1 = padding), one baryon
                                        // number per group (build_slot_table, spectra_df.cu)
  int Ny, Nphi, Neta, dimension;
  const double *yv, *cosphi, *sinphi, *etav, *etaw;
  const double *exptab;
  bool w_on_dan;      // eta weight multiplies the whole p.dsigma (famod, MomentumSpectra.cpp:1617) instead of the feqmod placement
  // as in K1 (spectra_df.cu): thread columns sorted by their smallest mT, rapidity rows by expected work, dropping of negligible items
  const int *col_map, *y_order;
  unsigned long long *items_done;       // [0] += items marched, [4] += items dropped
  const unsigned long long *amin_bits;  // [Ny + 1] from feqmod_amin_kernel
  double margin;
  double *bsum;                         // [slices of this launch][Ny * Nphi]
  const double *renorm_max;             // [cell]: largest PTM renormalisation of the cell's classes (SPECIES_RENORM), else nullptr
};

constexpr unsigned long long kHugeBits = 0x7f7f7f7f7f7f7f7full;   // 1.4e306: what cudaMemset(0x7f) leaves
constexpr double kXeNegligible = 680.0;                           // fast_exp's range guard (see spectra_df.cu)

// Per y: the smallest S = sqrt((|g1| + |A3| + |A4|)^2 + 1/T'^2) over the pass's valid cells that take the modified distribution;
// mT_hi S (+ kMaxBaryon max|alphaB'|) bounds the smallest exponent E'/T' - b alphaB' of a block row from above (feqmod_item_range
// with pT <= mT): the scale of the bins' leading terms, as df_amin_kernel does for K1.
__global__ void feqmod_amin_kernel(const double *__restrict__ pack, int64_t stride, int64_t count, int Ny, const double *__restrict__ yv,
                                   int dimension, int Neta, const double *__restrict__ etav, unsigned long long *__restrict__ amin_bits)
{
  extern __shared__ unsigned long long s_min[];       // [Ny + 1]
  for (int k = threadIdx.x; k <= Ny; k += blockDim.x) s_min[k] = k < Ny ? kHugeBits : 0ull;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t step = (int64_t)gridDim.x * blockDim.x, rounded = (count + 31) / 32 * 32;
  for (int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; cell < rounded; cell += step) {
    const bool valid = cell < count && pack[DP_VALID * stride + cell] != 0.0 && pack[FP_BREAKDOWN * stride + cell] == 0.0;
    double a1[3] = {0, 0, 0}, a2[3] = {0, 0, 0}, G2 = 0, iT2 = 0, eta = 0, scale = 1, alpha = 0;
    if (valid) {
      double n3 = 0, n4 = 0;
      for (int i = 0; i < 3; i++) {
        a1[i] = pack[(FP_A1X + i) * stride + cell]; a2[i] = pack[(FP_A2X + i) * stride + cell];
        const double a3 = pack[(FP_A3X + i) * stride + cell], a4 = pack[(FP_A4X + i) * stride + cell];
        n3 += a3 * a3; n4 += a4 * a4;
      }
      G2 = sqrt(n3) + sqrt(n4);
      iT2 = pack[FP_IT2 * stride + cell];
      eta = pack[DP_ETA * stride + cell]; scale = pack[FP_ETA_SCALE * stride + cell];
      alpha = fabs(pack[FP_ALPHAB_MOD * stride + cell]);
    }
    unsigned long long am = (unsigned long long)__double_as_longlong(alpha);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xffffffffu, am, o); am = v > am ? v : am; }
    if (lane == 0 && am) atomicMax(&s_min[Ny], am);
    for (int iy = 0; iy < Ny; iy++) {
      double S = __longlong_as_double((long long)kHugeBits);
      if (valid) {
        for (int ie = 0; ie < (dimension == 3 ? 1 : Neta); ie++) {
          const double d = yv[iy] - scale * (dimension == 3 ? eta : etav[ie]), sh = sinh(d), ch = cosh(d);
          double n1 = 0;
          for (int i = 0; i < 3; i++) { const double v = ch * a1[i] + sh * a2[i]; n1 += v * v; }
          const double g = sqrt(n1) + G2, v = sqrt(g * g + iT2);
          if (v < S) S = v;
        }
        if (!(S > 0.0)) S = 0.0;
      }
      unsigned long long b = (unsigned long long)__double_as_longlong(S);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xffffffffu, b, o); b = v < b ? v : b; }
      if (lane == 0 && b != kHugeBits) atomicMin(&s_min[iy], b);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k <= Ny; k += blockDim.x) {
    if (k < Ny) { if (s_min[k] != kHugeBits) atomicMin(&amin_bits[k], s_min[k]); }
    else if (s_min[k]) atomicMax(&amin_bits[k], s_min[k]);
  }
}

// largest renormalisation of a cell over all class slots (single slots, then pair members)
__global__ void feqmod_renorm_max_kernel(const double *__restrict__ renorm, int nslots, const double *__restrict__ renorm_pair, int nslots_pair,
                                         int64_t count, double *__restrict__ out)
{
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= count) return;
  double m = 0.0;
  for (int k = 0; k < nslots; k++) m = fmax(m, fabs(renorm[cell * nslots + k]));
  for (int k = 0; k < nslots_pair; k++) m = fmax(m, fabs(renorm_pair[cell * nslots_pair + k]));
  out[cell] = m;
}


union ItemSlot {
  DfItem lin;
  FeqmodItem mod;
  __device__ ItemSlot() {}
};

// tile_linear[tile] != 0 <=> some cell of the 256-cell tile can take the linear-df branch for some y (breakdown cell, or
// detA < 0.01 on a 3+1d surface: the narrow y - eta window); lets the LINEAR launch skip every other tile after one load
__global__ void feqmod_tile_flags_kernel(const double *__restrict__ pack, int64_t stride, int64_t count, int dimension,
                                         int *__restrict__ tile_linear)
{
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool lin = false;
  if (cell < count && pack[DP_VALID * stride + cell] != 0.0)
    lin = pack[FP_BREAKDOWN * stride + cell] != 0.0 || (dimension == 3 && pack[FP_DETA * stride + cell] < 0.01);
  const unsigned any = __ballot_sync(0xffffffffu, lin);
  if ((threadIdx.x & 31) == 0 && any) atomicOr(&tile_linear[cell / kTile], 1);
}

// thread constants of feqmod_spectra_kernel's momentum loop
template <int R, bool PAIR>
struct FeqThread {
  DfBin bin[R];
  double pT, pT2;
  int eslot, eslotm, grp;
};

// The momentum loop over n items starting at items[0] / item_cell[0].  CLAMP = false (modified branch only): every exponent of
// these items is below kXePairShared on the block's columns -- no range guard, one reciprocal per charge-conjugate pair.
template <bool LINEAR, bool BARYON, bool REGULATE, bool OUTFLOW, bool SPECIES_RENORM, int R, bool PAIR, bool CLAMP>
__device__ __forceinline__ void feqmod_item_loop(const ItemSlot *__restrict__ items, const int *__restrict__ item_cell, int n_items,
                                                 const FeqThread<R, PAIR> &th, const double *__restrict__ renorm, int nslots,
                                                 double (&acc)[R], double (&accm)[PAIR ? R : 1], const double *__restrict__ exptab)
{
  constexpr int S = PAIR ? 2 : 1;
#pragma unroll 1
  for (int k = 0; k < n_items; k++) {
    double rn[R], rnm[PAIR ? R : 1];
    if (SPECIES_RENORM) {      // L2-resident row; a software prefetch of item k + 1's row measured 11 % slower (profiles/r01_summary.md)
      const double *row = renorm + (int64_t)item_cell[k] * nslots + S * th.grp * R;
      if (PAIR) {              // (member b > 0, member b < 0) of every slot side by side
#pragma unroll
        for (int r = 0; r < R; r++) { const double2 v = *reinterpret_cast<const double2 *>(row + 2 * r); rn[r] = v.x; rnm[r] = v.y; }
      } else if (R % 2 == 0) {
#pragma unroll
        for (int r = 0; r < R; r += 2) { const double2 v = *reinterpret_cast<const double2 *>(row + r); rn[r] = v.x; rn[r + 1] = v.y; }
      } else {
#pragma unroll
        for (int r = 0; r < R; r++) rn[r] = row[r];
      }
    }
    if (!LINEAR) {
      const FeqmodItem &it = items[k].mod;        // shared memory: broadcast LDS.128 + one LDS.64 for eb[eslot]
      const FeqmodShared sh = feqmod_share(it, th.pT, th.pT2);
      const double eb = BARYON ? it.eb[th.eslot] : 1.0;
      if (PAIR) {
        const double ebm = it.eb[th.eslotm];
#pragma unroll
        for (int r = 0; r < R; r++)
          feqmod_accum_pair_u<OUTFLOW, !SPECIES_RENORM, CLAMP>(acc[r], accm[r], it, sh, eb, ebm, th.bin[r].mT, th.bin[r].mT2, th.bin[r].sign,
                                                               SPECIES_RENORM ? rn[r] : 1.0, SPECIES_RENORM ? rnm[r] : 1.0, exptab);
      } else {
#pragma unroll
        for (int r = 0; r < R; r++)
          feqmod_accum_u<BARYON, OUTFLOW, !SPECIES_RENORM, CLAMP>(acc[r], it, sh, eb, th.bin[r].mT, th.bin[r].mT2, th.bin[r].sign,
                                                                  SPECIES_RENORM ? rn[r] : 1.0, exptab);
      }
    } else {
      const DfItem it = items[k].lin;
      const DfShared sh = df_share<BARYON>(it, th.pT, th.pT2);
#pragma unroll
      for (int r = 0; r < R; r++) {
        double v = df_eval<2, BARYON, REGULATE, OUTFLOW, true>(it, sh, th.bin[r], exptab);
        if (SPECIES_RENORM) v = (rn[r] != 0.0) ? v : 0.0;   // NaN renorm: the reference skips the species (:828-832)
        acc[r] += v;
        if (PAIR) {            // the rare fallback items: the partner is evaluated on its own
          DfBin bm = th.bin[r];
          bm.baryon = -bm.baryon;
          double vm = df_eval<2, BARYON, REGULATE, OUTFLOW, true>(it, sh, bm, exptab);
          if (SPECIES_RENORM) vm = (rnm[r] != 0.0) ? vm : 0.0;
          accm[r] += vm;
        }
      }
    }
  }
}

// One instantiation per BRANCH of the reference's per-momentum choice (MomentumSpectra.cpp:932-1040): LINEAR = false takes the
// items of the modified distribution, LINEAR = true the linear-df fallback items (breakdown cells and the narrow y - eta
// window); each launch compacts the other kind away together with the u.dsigma <= 0 cells, both add into `partial`.
// Two lean loops instead of one loop with a per-item branch: the modified loop alone needs far fewer registers, so ptxas
// interleaves the R evaluations (with both branches in one body it fell back to two at a time) and the per-item flag
// load / test / branch / reconvergence instructions disappear.
// PAIR = true: a slot holds a charge-conjugate pair of classes (two class ids per slot, two renorm entries per slot; see
// spectra_df.cu pair_tables_core and feqmod_accum_pair_u).
// Modified-branch items are classified like K1's (spectra_df.cu): negligible for the whole block (dropped, bound summed into
// g.bsum), cold (no range guard) at the front of the tile, hot at the back.
template <bool LINEAR, bool BARYON, bool REGULATE, bool OUTFLOW, bool SPECIES_RENORM, int R, bool PAIR>
__global__ void __launch_bounds__(kThreads, IS3D_K2_MINBLOCKS)
feqmod_spectra_kernel(const double *__restrict__ pack, int64_t stride, int64_t ncells, int64_t cells_per_chunk,
                      const double *__restrict__ renorm, const int *__restrict__ tile_linear, FeqGrid g,
                      double *__restrict__ partial, int64_t total)
{
  __shared__ ItemSlot items[kTile];
  __shared__ double exptab[kExpTableSize];
  load_exp_table(exptab, g.exptab);
  __shared__ int item_cell[kTile];
  __shared__ int warp_count[2][kThreads / 32];
  __shared__ double blk_range[4 * (kThreads / 32)];
  __shared__ double blk_lohi[5];                        // mT_lo, mT_hi, pT_hi, m2_lo of the block's columns; the drop threshold
  __shared__ unsigned long long blk_items, blk_dropped;
  __shared__ double blk_bound;

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int iyr = blockIdx.y / g.Nphi, iphi = blockIdx.y - iyr * g.Nphi, iy = g.y_order[iyr];
  const double yval = g.yv[iy], cphi = g.cosphi[iphi], sphi = g.sinphi[iphi];

  // column = (thread group, pT node): the R classes of a group share the thread's pT and one baryon number (spectra_df.cu)
  const int col = blockIdx.x * kThreads + t;
  const int colc = g.col_map[col < g.ncols ? col : g.ncols - 1];
  const int grp = colc / g.NpT, ip = colc - grp * g.NpT;
  static_assert(!PAIR || BARYON, "pairs exist only with baryon terms");
  constexpr int S = PAIR ? 2 : 1;                       // class ids (and renorm entries) per slot
  FeqThread<R, PAIR> th;
  th.grp = grp;
  double acc[R], accm[PAIR ? R : 1];
  int jbin[R], jbinm[PAIR ? R : 1];
  const int cls0 = g.slot_class[S * grp * R];           // slot 0 of a group is never padding
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int cls = g.slot_class[S * (grp * R + r)];
    const int jj = (cls >= 0 ? cls : cls0) * g.NpT + ip;
    jbin[r] = (col < g.ncols && cls >= 0) ? jj : -1;
    const double mT = g.mT[jj];
    th.bin[r].mT = mT; th.bin[r].mT2 = mT * mT; th.bin[r].baryon = g.baryon[jj]; th.bin[r].sign = g.sign[jj];
    asm volatile("" : "+d"(th.bin[r].mT2));   // opaque: ptxas otherwise re-multiplies mT^2 (and pT^2) per item to save registers
    acc[r] = 0.0;
    if (PAIR) {
      const int clsm = g.slot_class[S * (grp * R + r) + 1];
      jbinm[r] = (col < g.ncols && clsm >= 0) ? clsm * g.NpT + ip : -1;
      accm[r] = 0.0;
    }
  }
  th.pT = g.pT[ip];
  th.pT2 = th.pT * th.pT;
  asm volatile("" : "+d"(th.pT2));
  th.eslot = kMaxBaryon + (BARYON ? (int)th.bin[0].baryon : 0);
  asm volatile("" : "+r"(th.eslot));   // opaque: keeps the slot index in a register (ptxas otherwise re-derives it with F2I per item)
  th.eslotm = 2 * kMaxBaryon - th.eslot;  // the antibaryon partners of a pair slot
  asm volatile("" : "+r"(th.eslotm));

  {   // range of the block's columns (shared memory, re-read per tile) and the drop threshold of this block row
    double lo = th.bin[0].mT, hi = th.bin[0].mT, ph = th.pT, m2 = th.bin[0].mT2 - th.pT2;
#pragma unroll
    for (int r = 1; r < R; r++) { lo = fmin(lo, th.bin[r].mT); hi = fmax(hi, th.bin[r].mT); m2 = fmin(m2, th.bin[r].mT2 - th.pT2); }
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
      for (int w = 1; w < kThreads / 32; w++) {
        lo = fmin(lo, blk_range[4 * w]); hi = fmax(hi, blk_range[4 * w + 1]); ph = fmax(ph, blk_range[4 * w + 2]); m2 = fmin(m2, blk_range[4 * w + 3]);
      }
      blk_lohi[0] = lo; blk_lohi[1] = hi; blk_lohi[2] = ph;
      blk_lohi[3] = fmax(m2 * (1.0 - 1e-12) - 1e-12, 0.0);      // mT^2 - pT^2 carries the rounding of the two squares
      double thr = kXeNegligible;
      if (!LINEAR && g.margin > 0.0) {
        const double amin = __longlong_as_double((long long)g.amin_bits[iy]);
        const double amax = __longlong_as_double((long long)g.amin_bits[g.Ny]);
        thr = fmin(thr, fma(hi, amin, (BARYON ? kMaxBaryon * amax : 0.0) + g.margin));
      }
      blk_lohi[4] = thr;
      blk_items = 0; blk_dropped = 0; blk_bound = 0.0;
    }
    __syncthreads();
  }
  int flip = 0;

  const int64_t chunk_begin = (int64_t)blockIdx.z * cells_per_chunk;
  int64_t chunk_end = chunk_begin + cells_per_chunk;
  if (chunk_end > ncells) chunk_end = ncells;

  for (int64_t tile = chunk_begin; tile < chunk_end; tile += kTile) {
    if (LINEAR && tile_linear[tile / kTile] == 0) continue;      // chunks start on tile boundaries (choose_chunks)
    const int64_t cell = tile + t;
    bool valid = (cell < chunk_end) && (pack[DP_VALID * stride + cell] != 0.0);
    if (valid) {
      // which branch this (cell, y) takes; independent of the eta node (the window test applies to 3+1d surfaces only)
      bool linear = pack[FP_BREAKDOWN * stride + cell] != 0.0;
      if (g.dimension == 3 && !linear) {                         // narrow (y - eta) window, MomentumSpectra.cpp:865-871
        const double detA = pack[FP_DETA * stride + cell];
        if (detA < 0.01 && fabs(yval - pack[DP_ETA * stride + cell]) < detA) linear = true;
      }
      valid = (linear == LINEAR);
    }
    for (int ie = 0; ie < g.Neta; ie++) {
      bool cold = false, hot = valid, dropped = false;
      double sh = 0.0, ch = 1.0, w = 1.0, dropped_bound = 0.0;
      auto pk = [&](int k) { return pack[k * stride + cell]; };
      if (valid) {
        double eta;
        if (g.dimension == 3) { eta = pk(DP_ETA); w = 1.0; }
        else { eta = g.etav[ie]; w = g.etaw[ie]; }
        const double d = LINEAR ? yval - eta : yval - pk(FP_ETA_SCALE) * eta;
        sh = sinh(d); ch = cosh(d);
        if (!LINEAR) {
          const FeqmodItem item = feqmod_make_item(pk, sh, ch, cphi, sphi, w, g.w_on_dan, BARYON, !SPECIES_RENORM);
          const volatile double *range = blk_lohi;
          const double mT_hi = range[1], pT_hi = range[2];
          const double shift = BARYON ? kMaxBaryon * fabs(item.alphaB_mod) : 0.0;
          const FeqmodRange x = feqmod_item_range(item, pk(FP_IT2), range[0], mT_hi, pT_hi, range[3], shift);
          dropped = x.lo >= range[4];                              // NaN: false (kept, and hot)
          cold = !dropped && x.hi < kXePairShared;
          hot = !dropped && !cold;
          if (dropped) dropped_bound = feqmod_item_term_bound(item, x.lo, mT_hi, pT_hi, SPECIES_RENORM ? g.renorm_max[cell] : 1.0, exptab);
        }
      }
      const unsigned b_cold = __ballot_sync(0xffffffffu, cold), b_hot = __ballot_sync(0xffffffffu, hot);
      const unsigned b_dropped = __ballot_sync(0xffffffffu, dropped);
      if (b_dropped) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dropped_bound += __shfl_xor_sync(0xffffffffu, dropped_bound, o);
        if (lane == 0) { atomicAdd(&blk_bound, dropped_bound); atomicAdd(&blk_dropped, (unsigned long long)__popc(b_dropped)); }
      }
      if (lane == 0) warp_count[flip][warp] = __popc(b_cold) | (__popc(b_hot) << 16);
      __syncthreads();                       // previous tile fully consumed, counts visible
      int base_cold = 0, base_hot = 0, n_cold = 0, n_hot = 0;
#pragma unroll
      for (int w_ = 0; w_ < kThreads / 32; w_++) {
        const int c = warp_count[flip][w_], cc = c & 0xffff, ch_ = c >> 16;
        if (w_ < warp) { base_cold += cc; base_hot += ch_; }
        n_cold += cc; n_hot += ch_;
      }
      if (cold || hot) {
        const unsigned below = (1u << lane) - 1u;
        const int slot = cold ? base_cold + __popc(b_cold & below) : kTile - 1 - (base_hot + __popc(b_hot & below));
        if (LINEAR) items[slot].lin = feqmod_make_linear_item(pk, sh, ch, cphi, sphi, w, g.w_on_dan);
        else items[slot].mod = feqmod_make_item(pk, sh, ch, cphi, sphi, w, g.w_on_dan, BARYON, !SPECIES_RENORM);
        item_cell[slot] = (int)(cell - 0);   // index inside this pass's pack / renorm arrays
      }
      __syncthreads();
      if (!LINEAR) feqmod_item_loop<LINEAR, BARYON, REGULATE, OUTFLOW, SPECIES_RENORM, R, PAIR, false>(items, item_cell, n_cold, th, renorm, g.nslots, acc, accm, exptab);
      if (n_hot) feqmod_item_loop<LINEAR, BARYON, REGULATE, OUTFLOW, SPECIES_RENORM, R, PAIR, true>(items + (kTile - n_hot), item_cell + (kTile - n_hot), n_hot, th, renorm, g.nslots, acc, accm, exptab);
      if (t == 0) blk_items += (unsigned)(n_cold + n_hot);
      flip ^= 1;
    }
  }
  __syncthreads();
  if (t == 0) {
    atomicAdd(g.items_done, blk_items); atomicAdd(g.items_done + 4, blk_dropped);
    if (blk_bound != 0.0) atomicAdd(&g.bsum[(int64_t)blockIdx.x * gridDim.y + (iy * g.Nphi + iphi)], blk_bound);
  }

  const int64_t pbase = (int64_t)blockIdx.z * total;
#pragma unroll
  for (int r = 0; r < R; r++) {
    if (jbin[r] >= 0) {
      int64_t idx = iy + (int64_t)g.Ny * (iphi + (int64_t)g.Nphi * jbin[r]);
      partial[pbase + idx] += acc[r];
    }
    if (PAIR && jbinm[r] >= 0) {
      int64_t idx = iy + (int64_t)g.Ny * (iphi + (int64_t)g.Nphi * jbinm[r]);
      partial[pbase + idx] += accm[r];
    }
  }
}

template <bool BARYON, bool SPECIES_RENORM, bool PAIR>
void launch_feqmod(bool reg, bool outflow, dim3 grid, cudaStream_t st, const double *pack, int64_t stride, int64_t n, int64_t cpc,
                   const double *renorm, const int *tile_linear, const FeqGrid &g, double *partial, int64_t total)
{
  // modified-distribution items (regulate_deltaf does not reach them), then the linear-df fallback items
#define IS3D_LAUNCH(LIN, REG, OUT) feqmod_spectra_kernel<LIN, BARYON, REG, OUT, SPECIES_RENORM, kBins, PAIR><<<grid, kThreads, 0, st>>>(pack, stride, n, cpc, renorm, tile_linear, g, partial, total)
  if (outflow) IS3D_LAUNCH(false, false, true);
  else IS3D_LAUNCH(false, false, false);
  if (reg && outflow) IS3D_LAUNCH(true, true, true);
  else if (reg) IS3D_LAUNCH(true, true, false);
  else if (outflow) IS3D_LAUNCH(true, false, true);
  else IS3D_LAUNCH(true, false, false);
#undef IS3D_LAUNCH
}

}  // namespace

is3d_status run_spectra_feqmod(is3d_ctx *ctx, double *out_dev, is3d_stats *stats)
{
  const is3d_params &p = ctx->prm;
  if (p.df_mode != 5 && ctx->gla_pts <= 0) { ctx->set_error("Gauss-Laguerre tables not set"); return IS3D_ERR_INVALID; }
  if (p.df_mode != 5 && (ctx->gla_pts > kRenormMaxPts || ctx->gla_alpha < 3)) {
    ctx->set_error("Gauss-Laguerre tables: need alpha = 0..2 with at most 64 points");
    return IS3D_ERR_INVALID;
  }
  if (p.df_mode == 5 && ctx->npdg <= 0) { ctx->set_error("PDG table not set (is3d_set_pdg)"); return IS3D_ERR_INVALID; }
  const int64_t n = ctx->surf.n;
  const int64_t total = (int64_t)ctx->ns * ctx->NpT * ctx->Nphi * ctx->Ny;
  FeqmodFlags fl;
  fl.df_mode = p.df_mode; fl.dimension = p.dimension; fl.include_baryon = p.include_baryon;
  fl.include_bulk = p.include_bulk_deltaf; fl.include_shear = p.include_shear_deltaf;
  fl.include_baryondiff = p.include_baryondiff_deltaf;
  fl.deta_min = p.deta_min; fl.mass_pion0 = p.mass_pion0; fl.bulkPi_over_P_max = ctx->tb.bulkPi_over_P_max;
  const bool species_renorm = (p.df_mode == 3 && p.include_bulk_deltaf);

  FeqGrid g;
  SpeciesBins sb;
  IS3D_TRY(build_bin_arrays(ctx, &sb));
  g.mT = sb.mT; g.pT = sb.pT; g.baryon = sb.baryon; g.sign = sb.sign;
  // thread groups: single classes and, with baryon terms, charge-conjugate pairs (spectra_df.cu pair_tables_core)
  std::vector<int> slots, pair_slots;
  bool ok;
  if (p.include_baryon) {
    std::vector<int> class_of, rep;
    species_classes(ctx, &class_of, &rep);
    ok = pair_tables_core(rep, ctx->h_mass.data(), ctx->h_sign.data(), ctx->h_baryon.data(), kBins, kBins, &slots, &pair_slots);
  } else {
    ok = build_slot_table(ctx, kBins, &slots);
  }
  if (!ok) {
    ctx->set_error("species list holds a baryon number outside -2..2 (the reference's PDG readers produce hadrons and the deuteron only)");
    return IS3D_ERR_INVALID;
  }
  // thread columns in order of their smallest mT, rapidity rows by expected work (as K1, spectra_df.cu)
  std::vector<int> class_of_all, rep_all;
  species_classes(ctx, &class_of_all, &rep_all);
  const std::vector<int> order_single = column_order(ctx, slots, kBins, rep_all), order_pair = column_order(ctx, pair_slots, 2 * kBins, rep_all);
  const std::vector<int> yo = rapidity_order(ctx);
  void *d_slots = nullptr;
  std::vector<int> both(slots);
  both.insert(both.end(), pair_slots.begin(), pair_slots.end());
  both.insert(both.end(), order_single.begin(), order_single.end());
  both.insert(both.end(), order_pair.begin(), order_pair.end());
  both.insert(both.end(), yo.begin(), yo.end());
  IS3D_TRY(ctx->get_scratch("k2_slots", (both.size() + 1) * sizeof(int), &d_slots));
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(d_slots, both.data(), both.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));       // `both` is pageable host memory
  g.slot_class = (const int *)d_slots;
  g.col_map = (const int *)d_slots + slots.size() + pair_slots.size();
  g.y_order = g.col_map + order_single.size() + order_pair.size();
  g.NpT = ctx->NpT; g.ns = sb.nclass; g.nslots = (int)slots.size(); g.ncols = ctx->NpT * (int)(slots.size() / kBins);
  const int64_t per_species = (int64_t)ctx->NpT * ctx->Nphi * ctx->Ny;
  const int64_t total_class = (int64_t)sb.nclass * per_species;
  g.Ny = ctx->Ny; g.Nphi = ctx->Nphi; g.Neta = ctx->Neta; g.dimension = p.dimension;
  g.yv = ctx->d_y; g.cosphi = ctx->d_cosphi; g.sinphi = ctx->d_sinphi; g.etav = ctx->d_eta; g.etaw = ctx->d_etaw;
  g.w_on_dan = (p.df_mode == 5);
  g.exptab = ctx->d_exptab;

  FeqGrid gp = g;                                                // the pair launch: two class ids / renorm entries per slot
  gp.slot_class = (const int *)d_slots + slots.size();
  gp.nslots = (int)pair_slots.size();
  gp.col_map = g.col_map + order_single.size();
  gp.ncols = ctx->NpT * (int)(pair_slots.size() / (2 * kBins));
  const int nslices = (g.ncols + kThreads - 1) / kThreads, nslices_pair = (gp.ncols + kThreads - 1) / kThreads;
  const int64_t blocks_per_chunk = (int64_t)(nslices + nslices_pair) * ctx->Ny * ctx->Nphi;
  if ((int64_t)ctx->Ny * ctx->Nphi > 65535) { ctx->set_error("Ny*Nphi exceeds 65535"); return IS3D_ERR_INVALID; }

  // cells per pass: bounds the pack (440 B/cell) and the PTM renorm table (8 Ns B/cell) to ~2 GB
  int64_t macro = pass_cells(2 << 20);
  const int64_t all_slots = (int64_t)g.nslots + gp.nslots;
  if (species_renorm) { int64_t m2 = ((int64_t)1 << 31) / (8 * all_slots); if (m2 < macro) macro = m2; }
  macro = macro / kTile * kTile;
  if (macro < kTile) macro = kTile;
  const int64_t stride = n < macro ? n : macro;
  int nchunks; int64_t cpc;
  choose_chunks(ctx, stride, blocks_per_chunk, total_class, kTile, IS3D_K2_MINBLOCKS, &nchunks, &cpc);

  void *pack = nullptr, *partial = nullptr, *counters = nullptr, *renorm = nullptr, *tile_linear = nullptr, *prune = nullptr, *rmax = nullptr;
  const size_t ntile_flags = (size_t)((stride + kTile - 1) / kTile);
  IS3D_TRY(ctx->get_scratch("tile_linear", ntile_flags * sizeof(int), &tile_linear));
  IS3D_TRY(ctx->get_scratch("cell_pack", (size_t)FP_SIZE * stride * sizeof(double), &pack));
  IS3D_TRY(ctx->get_scratch("partial", (size_t)nchunks * total_class * sizeof(double), &partial));
  IS3D_TRY(ctx->get_scratch("counters", 32 * sizeof(unsigned long long), &counters));
  if (species_renorm) IS3D_TRY(ctx->get_scratch("renorm", ((size_t)stride * all_slots + 2) * sizeof(double), &renorm));
  if (species_renorm) IS3D_TRY(ctx->get_scratch("renorm_max", (size_t)stride * sizeof(double), &rmax));
  // the pair table starts on a 16-byte boundary (its rows are read as double2; an odd R makes stride * nslots odd)
  double *renorm_pair = species_renorm ? (double *)renorm + ((size_t)stride * g.nslots + 1) / 2 * 2 : nullptr;
  // dropping of negligible items (K1's scheme, spectra_df.cu): row scales | bounds of the dropped terms per block row | bin -> row
  const int NyNphi = ctx->Ny * ctx->Nphi, nrows = nslices + nslices_pair;
  const size_t amin_bytes = (size_t)(ctx->Ny + 1) * 8, bsum_bytes = (size_t)nrows * NyNphi * 8;
  std::vector<int> bin_row((size_t)sb.nclass * ctx->NpT, 0);
  fill_bin_rows(ctx, slots, kBins, order_single, kThreads, 0, &bin_row);
  fill_bin_rows(ctx, pair_slots, 2 * kBins, order_pair, kThreads, nslices, &bin_row);
  IS3D_TRY(ctx->get_scratch("k2_prune", amin_bytes + bsum_bytes + bin_row.size() * sizeof(int), &prune));
  unsigned long long *d_amin = (unsigned long long *)prune;
  double *d_bsum = (double *)((char *)prune + amin_bytes);
  int *d_bin_row = (int *)((char *)prune + amin_bytes + bsum_bytes);
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(d_bin_row, bin_row.data(), bin_row.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));       // pageable host memory
  // counters: 0..9 as before; 16 / 17 items marched (single, pair launches), 20 / 21 items dropped, 24 bins failing the bound test
  g.items_done = (unsigned long long *)counters + 16;
  gp.items_done = (unsigned long long *)counters + 17;
  g.amin_bits = gp.amin_bits = d_amin;
  g.bsum = d_bsum; gp.bsum = d_bsum + (size_t)nslices * NyNphi;
  g.renorm_max = gp.renorm_max = (const double *)rmax;

  cudaEvent_t e0 = ctx->ev0, e1 = ctx->ev1;             // owned by the context: nothing to release on an error path
  float ms_total = 0.f;
  int64_t launches = 0, prune_reruns = 0;
  unsigned long long h_counters[32];
  // attempt 0 drops items below the margin; if the a-posteriori test fails for any bin, attempt 1 repeats the call without it.
  // The serial-chain parity mode of df_mode 5 keeps its chain state across calls: no second attempt there, so no margin.
  const double margin0 = (p.df_mode == 5 && p.famod_chain) ? 0.0 : p.negligible_margin;
  for (int attempt = 0; attempt < 2; attempt++) {
  g.margin = gp.margin = attempt == 0 ? margin0 : 0.0;
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(partial, 0, (size_t)nchunks * total_class * sizeof(double), ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(counters, 0, 32 * sizeof(unsigned long long), ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(d_bsum, 0, bsum_bytes, ctx->stream));
  for (int64_t begin = 0; begin < n; begin += macro) {
    int64_t count = n - begin < macro ? n - begin : macro;
    IS3D_CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
    if (p.df_mode == 5) {
      IS3D_TRY(famod_setup_pass(ctx, begin, count, (double *)pack, stride, (unsigned long long *)counters, &launches, false));
    } else {
      feqmod_setup_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(
          ctx->surf, begin, count, ctx->tb, fl, ctx->d_gla_root, ctx->d_gla_weight, ctx->gla_pts, (double *)pack, stride,
          (unsigned long long *)counters);
      IS3D_CUDA_TRY(ctx, cudaGetLastError());
      launches++;
    }
    if (species_renorm) {
      for (int pass = 0; pass < 2; pass++) {           // single slots, then the flat list of pair members
        const FeqGrid &q = pass ? gp : g;
        const int64_t work = count * q.nslots;
        if (!work) continue;
        feqmod_renorm_kernel<<<(unsigned)((work + 127) / 128), 128, 0, ctx->stream>>>(
            (double *)pack, stride, count, q.nslots, q.slot_class, sb.c_mass, sb.c_deg, sb.c_baryon, sb.c_sign, ctx->d_gla_root,
            ctx->d_gla_weight, ctx->gla_pts, ctx->d_exptab, pass ? renorm_pair : (double *)renorm);
        IS3D_CUDA_TRY(ctx, cudaGetLastError());
        launches++;
      }
      feqmod_renorm_max_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>((double *)renorm, g.nslots, renorm_pair, gp.nslots, count, (double *)rmax);
      IS3D_CUDA_TRY(ctx, cudaGetLastError());
      launches++;
    }
    if (g.margin > 0.0) {
      IS3D_CUDA_TRY(ctx, cudaMemsetAsync(d_amin, 0x7f, amin_bytes - 8, ctx->stream));
      IS3D_CUDA_TRY(ctx, cudaMemsetAsync(d_amin + ctx->Ny, 0, 8, ctx->stream));
      int64_t ab = (count + 255) / 256, ab_max = 8 * (int64_t)ctx->sm_count;
      feqmod_amin_kernel<<<(unsigned)(ab < ab_max ? ab : ab_max), 256, amin_bytes, ctx->stream>>>((double *)pack, stride, count, ctx->Ny, ctx->d_y, p.dimension, ctx->Neta, ctx->d_eta, d_amin);
      IS3D_CUDA_TRY(ctx, cudaGetLastError());
      launches++;
    }
    IS3D_CUDA_TRY(ctx, cudaMemsetAsync(tile_linear, 0, ntile_flags * sizeof(int), ctx->stream));
    feqmod_tile_flags_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>((double *)pack, stride, count, p.dimension, (int *)tile_linear);
    IS3D_CUDA_TRY(ctx, cudaGetLastError());
    launches++;
    int nch = (int)((count + cpc - 1) / cpc);
    dim3 grid(nslices, ctx->Ny * ctx->Nphi, nch), grid_pair(nslices_pair, ctx->Ny * ctx->Nphi, nch);
    const bool reg = p.regulate_deltaf != 0, outflow = p.outflow != 0;
    if (p.include_baryon) {
      // the pair launch first (its blocks are the longer ones); the single-class launches run beside it on a second stream
      const bool two = nslices_pair && nslices;
      cudaStream_t s_single = two ? ctx->side_stream : ctx->stream;
      if (two) {
        IS3D_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
        IS3D_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
      }
      if (nslices_pair) {
        if (species_renorm) launch_feqmod<true, true, true>(reg, outflow, grid_pair, ctx->stream, (double *)pack, stride, count, cpc, renorm_pair, (const int *)tile_linear, gp, (double *)partial, total_class);
        else launch_feqmod<true, false, true>(reg, outflow, grid_pair, ctx->stream, (double *)pack, stride, count, cpc, renorm_pair, (const int *)tile_linear, gp, (double *)partial, total_class);
        launches += 2;
      }
      if (nslices) {
        if (species_renorm) launch_feqmod<true, true, false>(reg, outflow, grid, s_single, (double *)pack, stride, count, cpc, (double *)renorm, (const int *)tile_linear, g, (double *)partial, total_class);
        else launch_feqmod<true, false, false>(reg, outflow, grid, s_single, (double *)pack, stride, count, cpc, (double *)renorm, (const int *)tile_linear, g, (double *)partial, total_class);
        launches += 2;
      }
      if (two) {
        IS3D_CUDA_TRY(ctx, cudaGetLastError());
        IS3D_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join, ctx->side_stream));
        IS3D_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
      }
    } else {
      if (species_renorm) launch_feqmod<false, true, false>(reg, outflow, grid, ctx->stream, (double *)pack, stride, count, cpc, (double *)renorm, (const int *)tile_linear, g, (double *)partial, total_class);
      else launch_feqmod<false, false, false>(reg, outflow, grid, ctx->stream, (double *)pack, stride, count, cpc, (double *)renorm, (const int *)tile_linear, g, (double *)partial, total_class);
      launches += 2;
    }
    IS3D_CUDA_TRY(ctx, cudaGetLastError());
    IS3D_CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaEventSynchronize(e1));
    float ms = 0.f;
    IS3D_CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
    ms_total += ms;
  }
  PruneCheck chk;                                      // without a margin only the < 1e-295 items are dropped: nothing to test
  if (g.margin > 0.0) chk.bsum = d_bsum;
  chk.bin_row = d_bin_row; chk.Ny = ctx->Ny; chk.NyNphi = NyNphi; chk.eps = 1e-13;
  chk.violations = (unsigned long long *)counters + 24;
  reduce_partials_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((double *)partial, nchunks, total_class, per_species,
                                                                                sb.class_of, ctx->d_deg, total, out_dev, chk);
  IS3D_CUDA_TRY(ctx, cudaGetLastError());
  launches++;
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (h_counters[24] == 0 || h_counters[1] != 0 || !(g.margin > 0.0)) break;
  prune_reruns++;
  }
  if (stats) {
    stats->cells_total = n;
    stats->cells_skipped = (int64_t)h_counters[0];
    stats->cells_out_of_table = (int64_t)h_counters[1];
    stats->cells_breakdown = (int64_t)h_counters[2];
    stats->cells_pl_negative = (int64_t)h_counters[3];
    // "until t = ..." of the reference's printout: tau of the last (highest-index) such cell
    double tau_b = 0.0, tau_p = 0.0;
    if (h_counters[4]) IS3D_CUDA_TRY(ctx, cudaMemcpy(&tau_b, ctx->surf.col[0] + (h_counters[4] - 1), sizeof(double), cudaMemcpyDeviceToHost));
    if (h_counters[5]) IS3D_CUDA_TRY(ctx, cudaMemcpy(&tau_p, ctx->surf.col[0] + (h_counters[5] - 1), sizeof(double), cudaMemcpyDeviceToHost));
    stats->tau_breakdown = tau_b;
    stats->tau_pl_negative = tau_p;
    stats->reconstruction_failures = (int64_t)h_counters[8];
    stats->newton_iterations = (int64_t)h_counters[9];
    stats->kernel_ms = ms_total;
    stats->kernel_launches = launches;
    stats->pair_evals_executed = 2 * (int64_t)h_counters[17] * kThreads * kBins;
    stats->evals_executed = (int64_t)h_counters[16] * kThreads * kBins + stats->pair_evals_executed;
    stats->evals_dropped = ((int64_t)h_counters[20] * kBins + 2 * (int64_t)h_counters[21] * kBins) * kThreads;
    stats->prune_reruns = prune_reruns;
  }
  if (h_counters[1] != 0) {
    ctx->set_error(std::to_string(h_counters[1]) + " cell(s) outside the df coefficient tables (the reference aborts here)");
    return IS3D_ERR_TABLE_RANGE;
  }
  return IS3D_OK;
}

}  // namespace is3d
