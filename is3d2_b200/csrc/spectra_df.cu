// K1: continuous Cooper-Frye spectra for df_mode 1 (Grad 14-moment) and 2 (RTA Chapman-Enskog) on sm_100a.
// Replaces EmissionFunctionArray::calculate_dN_pTdpTdphidy (reference src/cpp/MomentumSpectra.cpp:32-415).
//
// Layout / schedule
//   1. df_setup_kernel: one thread per cell; reads the 20-25 SoA surface columns (coalesced), evaluates the df
//      coefficients and writes a 32-double "cell pack" as SoA into HBM (256 B / cell).
//   2. df_spectra_kernel: output-stationary.  blockIdx.x = slice of (thread group, pT) columns, blockIdx.y =
//      (iy, iphi), blockIdx.z = contiguous chunk of cells.  A thread owns R species classes of ONE baryon number at
//      ONE pT node (build_slot_table), so every product of an item constant with pT or b is formed once per item and
//      shared by its R evaluations (spectra_df.cuh: DfItemU, df_share_u, df_eval_u).
//      A block streams its chunk in tiles of 256 cells: each thread turns one cell pack into the 18-double item
//      constants for the block's (y, phi) -- one sinh per cell per tile -- with invalid (u.dsigma <= 0) cells
//      compacted away by ballot/prefix; then every thread marches over the tile, reading the warp-uniform item
//      with broadcast LDS.128 and updating its R register accumulators.
//      FP64-pipe bound: HBM traffic is 256 B per cell per block against >= 1024 * ~27 DFMA per cell per block.
//   3. reduce_partials_kernel: deterministic sum over the cell chunks.
#include <algorithm>

#include "ctx.h"
#include "spectra_df.cuh"

namespace is3d {

// bounds below this are the < 1e-295 items dropped by the range guard itself (e^-680 times any prefactor a surface can produce)
constexpr double kPruneFloor = 1e-280;

namespace {

// launch shape (tunable at build time for the sweeps recorded in profiles/): threads per block, resident blocks per SM
// the register allocation is bounded for, species per thread
// 128 x 3 (12 warps per SM, up to 168 registers) measured 8 % faster than 256 x 2 once negligible items are dropped: narrower blocks
// agree on more droppable items and the momentum loop keeps its constants in registers (profiles/r02_k1_variants_notes.txt)
#ifndef IS3D_K1_THREADS
#define IS3D_K1_THREADS 128
#endif
#ifndef IS3D_K1_MINBLOCKS
#define IS3D_K1_MINBLOCKS 3
#endif
#ifndef IS3D_K1_R
#define IS3D_K1_R 4
#endif
#ifndef IS3D_K1_ITEM_UNROLL
#define IS3D_K1_ITEM_UNROLL 1
#endif
constexpr int kItemUnroll = IS3D_K1_ITEM_UNROLL;   // items per trip of the momentum loop
constexpr int kThreads = IS3D_K1_THREADS;
constexpr int kTile = kThreads;  // cells per shared-memory tile = threads per block
constexpr int kDfBinsPerThread = IS3D_K1_R;   // species per thread (R)
#ifndef IS3D_K1_PAIR_R
#define IS3D_K1_PAIR_R IS3D_K1_R
#endif
constexpr int kDfPairsPerThread = IS3D_K1_PAIR_R;   // charge-conjugate pair slots per thread of the pair launch
// range classes of an item's xE = u.p/T over the columns of a block (df_spectra_kernel)
constexpr double kXeNegligible = 680.0;             // = fast_exp's range guard: feq < 1e-295 beyond, the reference's exp overflows at 709.8
constexpr double kXeCold = 600.0;                   // below: no range guard needed
constexpr double kPruneEps = 1e-13;                 // a-posteriori test of the dropped items: bound <= kPruneEps |bin| for every bin

__global__ void df_setup_kernel(SurfaceView surf, int64_t begin, int64_t count, DfTables tb, DfFlags fl,
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

struct DfGrid {
  const double *mT, *pT, *baryon, *sign;              // per (species class, pT) bin, [ns * NpT]
  int ns, NpT, ncols;                                 // ns = number of species CLASSES; ncols = NpT * ngroups thread columns
  const int *slot_class;                              // [ngroups * R]: class of slot r of a thread group, -1 = padding; the
                                                      // valid slots of a group carry ONE baryon number (build_slot_table)
  int Ny, Nphi, Neta, dimension;
  const double *yv, *cosphi, *sinphi, *etav, *etaw;
  const double *exptab;                               // 2^(m/1024), global memory (ctx->d_exptab)
  const int *col_map;                                 // [ncols]: thread column -> group * NpT + ip, sorted by the column's smallest
                                                      // mT so that the columns of a block see the same cells as negligible
  const int *y_order;                                 // [Ny]: rapidity index of block row rank k
  unsigned long long *items_done;                     // += items a block has marched over (executed-work statistic); [4]: += items dropped
  // dropping of negligible items (see df_spectra_kernel)
  const unsigned long long *amin_bits;                // [Ny + 1] from df_amin_kernel: per y the smallest A = aT + |u_perp|/T of the pass
                                                      // (bits of a positive double), then the largest |alpha_B|
  double margin;                                      // Delta: items whose xE exceeds the row's smallest possible xE by more are dropped; <= 0: off
  double *bsum;                                       // [slices of this launch][Ny * Nphi] += bounds of the dropped terms
};

constexpr unsigned long long kHugeBits = 0x7f7f7f7f7f7f7f7full;   // 1.4e306: what cudaMemset(0x7f) leaves

// Per y: the smallest A_i = aT_i + |u_perp,i| / T_i over the valid cells (and eta nodes) of the pass, where aT_i = (u^tau cosh(y - eta)
// - tau u^eta sinh(y - eta)) / T.  For a block whose columns have mT <= mT_hi, mT_hi A_min (+ kMaxBaryon max|alpha_B|) is an upper
// bound of the smallest exponent x = xE - b alpha_B any of its bins sees: the scale of the bins' dominant terms.
__global__ void df_amin_kernel(const double *__restrict__ pack, int64_t stride, int64_t count, int Ny, const double *__restrict__ yv,
                               int dimension, int Neta, const double *__restrict__ etav, unsigned long long *__restrict__ amin_bits)
{
  extern __shared__ unsigned long long s_min[];       // [Ny + 1]
  for (int k = threadIdx.x; k <= Ny; k += blockDim.x) s_min[k] = k < Ny ? kHugeBits : 0ull;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t step = (int64_t)gridDim.x * blockDim.x, rounded = (count + 31) / 32 * 32;
  for (int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; cell < rounded; cell += step) {
    const bool valid = cell < count && pack[DP_VALID * stride + cell] != 0.0;
    double utt = 0, tunt = 0, uperp = 0, eta = 0, alpha = 0;
    if (valid) {
      utt = pack[DP_UTT * stride + cell]; tunt = pack[DP_TUNT * stride + cell];
      const double uxt = pack[DP_UXT * stride + cell], uyt = pack[DP_UYT * stride + cell];
      uperp = sqrt(uxt * uxt + uyt * uyt);
      eta = pack[DP_ETA * stride + cell];
      alpha = fabs(pack[DP_ALPHAB * stride + cell]);
    }
    unsigned long long am = (unsigned long long)__double_as_longlong(alpha);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xffffffffu, am, o); am = v > am ? v : am; }
    if (lane == 0 && am) atomicMax(&s_min[Ny], am);
    for (int iy = 0; iy < Ny; iy++) {
      double A = __longlong_as_double((long long)kHugeBits);
      if (valid) {
        for (int ie = 0; ie < (dimension == 3 ? 1 : Neta); ie++) {
          const double sh = sinh(yv[iy] - (dimension == 3 ? eta : etav[ie])), ch = sqrt(1.0 + sh * sh);
          const double a = ch * utt - sh * tunt + uperp;
          if (a < A) A = a;                            // NaN: ignored
        }
        if (!(A > 0.0)) A = 0.0;
      }
      unsigned long long b = (unsigned long long)__double_as_longlong(A);
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

// The momentum loop over the items of one tile.  CLAMP = false: no item of the tile can reach the exp range guard.
template <int MODE, bool BARYON, bool REGULATE, bool OUTFLOW, int R, bool PAIR, bool CLAMP>
__device__ __forceinline__ void df_item_loop(const DfItemU *__restrict__ items, int n_items, const DfThreadU &th, const DfThreadU &thm,
                                             const double (&mT)[R], const double (&mT2)[R], const double (&sgn)[R],
                                             double (&acc)[R], double (&accm)[PAIR ? R : 1], const double *__restrict__ exptab)
{
#pragma unroll kItemUnroll
  for (int k = 0; k < n_items; k++) {
    const DfItemU &it = items[k];          // shared memory: fields arrive as broadcast LDS.128, eb[eslot] as one LDS.64
    const DfSharedU sh = df_share_u<MODE, BARYON, (MODE == 2 && !REGULATE)>(it, th);
    if (!PAIR) {
#pragma unroll
      for (int r = 0; r < R; r++) acc[r] += df_eval_u<MODE, BARYON, REGULATE, OUTFLOW, CLAMP>(it, sh, mT[r], mT2[r], sgn[r], exptab);
    } else {
      const DfSharedU shm = df_share_u<MODE, BARYON, (MODE == 2 && !REGULATE)>(it, thm);   // common sub-expressions are shared by the compiler
#pragma unroll
      for (int r = 0; r < R; r++) {
        const double xE = df_eval_u_x<CLAMP>(it, sh, mT[r]);
        int spare;
        const double e = fast_exp_k<false>(xE, exptab, spare);
        if (!CLAMP) {              // cold items of the pair launch lie below kXePairShared: one reciprocal for both members
          df_eval_u_pair_shared<MODE, REGULATE, OUTFLOW>(it, sh, shm, mT[r], mT2[r], sgn[r], xE, e, spare, acc[r], accm[r]);
        } else {
          acc[r] += df_eval_u_tail<MODE, BARYON, REGULATE, OUTFLOW>(it, sh, mT[r], mT2[r], sgn[r], xE, e, spare);
          accm[r] += df_eval_u_tail<MODE, BARYON, REGULATE, OUTFLOW>(it, shm, mT[r], mT2[r], sgn[r], xE, e, spare);
        }
      }
    }
  }
}

// PAIR = true: a slot is a charge-conjugate PAIR of classes (baryon class, its antibaryon class: same mass, same statistics,
// b and -b).  x_E = u.p/T and exp(x_E) do not depend on b, so one exponential (7 of the ~21 FP64 instructions and 6 of the 8
// integer / shared-memory instructions of an evaluation) serves both; g.slot_class then holds two class ids per slot.
template <int MODE, bool BARYON, bool REGULATE, bool OUTFLOW, int R, bool PAIR>
__global__ void __launch_bounds__(kThreads, IS3D_K1_MINBLOCKS)
df_spectra_kernel(const double *__restrict__ pack, int64_t stride, int64_t ncells, int64_t cells_per_chunk, DfGrid g,
                  double *__restrict__ partial, int64_t total)
{
  static_assert(!PAIR || BARYON, "pairs exist only with baryon terms");
  __shared__ DfItemU items[kTile];
  __shared__ double exptab[kExpTableSize];
  __shared__ int warp_count[2][kThreads / 32];      // double-buffered: two barriers per tile
  load_exp_table(exptab, g.exptab);                 // visible after the first __syncthreads of the tile loop

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  // rapidity rows in order of decreasing expected work (y_order: |y - y_mid| ascending; far rapidities drop most items): the last
  // blocks of the launch are then the short ones
  const int iyr = blockIdx.y / g.Nphi, iphi = blockIdx.y - iyr * g.Nphi, iy = g.y_order[iyr];
  const double yval = g.yv[iy], cphi = g.cosphi[iphi], sphi = g.sinphi[iphi];
  const double ey = exp(yval), emy = exp(-yval);

  // column = (thread group, pT node): the R classes of a group share the thread's pT and one baryon number
  const int col = blockIdx.x * kThreads + t;
  const int colc = g.col_map[col < g.ncols ? col : g.ncols - 1];
  const int grp = colc / g.NpT, ip = colc - grp * g.NpT;
  constexpr int S = PAIR ? 2 : 1;                       // class ids per slot
  double mT[R], mT2[R], sgn[R];
  double acc[R], accm[PAIR ? R : 1];                    // accm: the antibaryon partners of a pair slot
  int jbin[R], jbinm[PAIR ? R : 1];                     // (class, pT) bin index, -1 = padding
  const int cls0 = g.slot_class[S * grp * R];           // slot 0 of a group is never padding
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int cls = g.slot_class[S * (grp * R + r)];
    const int jj = (cls >= 0 ? cls : cls0) * g.NpT + ip;
    jbin[r] = (col < g.ncols && cls >= 0) ? jj : -1;
    const double m = g.mT[jj];
    mT[r] = m; mT2[r] = m * m; sgn[r] = g.sign[jj];
    if (MODE == 2) asm volatile("" : "+d"(mT2[r]));   // opaque: ptxas otherwise re-multiplies mT^2 per item when registers are tight
    acc[r] = 0.0;
    if (PAIR) {
      const int clsm = g.slot_class[S * (grp * R + r) + 1];
      jbinm[r] = (col < g.ncols && clsm >= 0) ? clsm * g.NpT + ip : -1;
      accm[r] = 0.0;
    }
  }
  DfThreadU th;
  th.pT = g.pT[ip]; th.pT2 = th.pT * th.pT;            // bin arrays are [class][pT]: entry ip = class 0
  th.b = BARYON ? g.baryon[cls0 * g.NpT + ip] : 0.0;
  th.bpT = th.b * th.pT;
  if (MODE == 2) asm volatile("" : "+d"(th.pT2), "+d"(th.bpT));
  th.eslot = kMaxBaryon + (int)th.b;
  DfThreadU thm = th;                                   // the antibaryon partners: b -> -b
  if (PAIR) { thm.b = -th.b; thm.bpT = -th.bpT; thm.eslot = kMaxBaryon - (int)th.b; }

  // range of the block's columns: every xE = mT aT - pT bT the block can form for an item lies in [mT_lo (aT - max(bT, 0)),
  // mT_hi aT + pT_hi max(-bT, 0)]  (aT >= |bT| for a time-like flow velocity, pT <= mT)
  // (kept in shared memory and re-read per tile: the momentum loop needs every register)
  __shared__ double blk_range[3 * (kThreads / 32)];
  __shared__ double blk_lohi[4];                      // mT_lo, mT_hi, pT_hi of the block's columns; the drop threshold of this row
  __shared__ unsigned long long blk_items, blk_dropped;
  __shared__ double blk_bound;                        // sum of the term bounds of the items this block dropped
  {
    double lo = mT[0], hi = mT[0], ph = th.pT;
#pragma unroll
    for (int r = 1; r < R; r++) { lo = fmin(lo, mT[r]); hi = fmax(hi, mT[r]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      ph = fmax(ph, __shfl_xor_sync(0xffffffffu, ph, o));
    }
    if (lane == 0) { blk_range[3 * warp] = lo; blk_range[3 * warp + 1] = hi; blk_range[3 * warp + 2] = ph; }
    __syncthreads();
    if (t == 0) {
      for (int w = 1; w < kThreads / 32; w++) {
        lo = fmin(lo, blk_range[3 * w]); hi = fmax(hi, blk_range[3 * w + 1]); ph = fmax(ph, blk_range[3 * w + 2]);
      }
      blk_lohi[0] = lo; blk_lohi[1] = hi; blk_lohi[2] = ph;
      // Items are dropped when every exponent x = xE - b alpha_B they can produce in this block is >= the threshold:
      //   kXeNegligible always (feq < 1e-295), and, with g.margin > 0, the row's dominant exponent + margin: such terms are
      //   < e^-margin of the bins' leading terms.  The bounds of everything dropped are summed per row (g.bsum) and compared
      //   with the finished spectra by reduce_partials_kernel; run_spectra_df repeats the call without the margin if any bin
      //   fails that test, so the margin is a speed heuristic, not a precision knob.
      double thr = kXeNegligible;
      if (g.margin > 0.0) {
        const double amin = __longlong_as_double((long long)g.amin_bits[iy]);
        const double amax = __longlong_as_double((long long)g.amin_bits[g.Ny]);
        thr = fmin(thr, fma(hi, amin, (BARYON ? kMaxBaryon * amax : 0.0) + g.margin));
      }
      blk_lohi[3] = thr;
      blk_items = 0; blk_dropped = 0; blk_bound = 0.0;
    }
    __syncthreads();
  }
  int flip = 0;

  const int64_t chunk_begin = (int64_t)blockIdx.z * cells_per_chunk;
  int64_t chunk_end = chunk_begin + cells_per_chunk;
  if (chunk_end > ncells) chunk_end = ncells;

  for (int64_t tile = chunk_begin; tile < chunk_end; tile += kTile) {
    const int64_t cell = tile + t;
    const bool valid = (cell < chunk_end) && (pack[DP_VALID * stride + cell] != 0.0);
    for (int ie = 0; ie < g.Neta; ie++) {
      // The cell's item for this block's (y, phi), classified by the range of xE over the block's columns:
      //   negligible: every xE >= kXeNegligible -- the Bose/Fermi factor of every evaluation is below 1e-295 (where fast_exp's range
      //               guard saturates and the reference's exp overflows to feq = 0): the item is dropped like a u.dsigma <= 0 cell;
      //   cold:       every xE < kXeCold (pair launch: every exponent < kXePairShared): the loop without the range guard;  hot: the rest.
      // Cold items fill the tile from the front, hot items from the back.
      bool cold = false, hot = false, dropped = false;
      double sh = 0.0, ch = 1.0, w = 1.0, dropped_bound = 0.0;
      if (valid) {
        double eta;
        if (g.dimension == 3) { eta = pack[DP_ETA * stride + cell]; w = 1.0; }
        else { eta = g.etav[ie]; w = g.etaw[ie]; }
        // 3+1d: sinh(y - eta) from the cell's e^{+-eta} (pack) and the block's e^{+-y}: 4 flops instead of a libm call per (cell, y);
        // the absolute error near y = eta (1e-16) is what every use of sh tolerates (it multiplies tau u^eta, dsigma_eta, pi^{mu eta})
        sh = g.dimension == 3 ? 0.5 * (ey * pack[DP_EMETA * stride + cell] - emy * pack[DP_EPETA * stride + cell]) : sinh(yval - eta);
        ch = sqrt(1.0 + sh * sh);            // the reference's cosh (MomentumSpectra.cpp:307-308)
        // aT, bT exactly as df_make_item_u forms them
        const double aT = ch * pack[DP_UTT * stride + cell] - sh * pack[DP_TUNT * stride + cell];
        const double bT = cphi * pack[DP_UXT * stride + cell] + sphi * pack[DP_UYT * stride + cell];
        const volatile double *range = blk_lohi;
        const double mT_hi = range[1], pT_hi = range[2];
        const double xe_lo = range[0] * (aT - fmax(bT, 0.0)), xe_hi = fma(mT_hi, aT, pT_hi * fmax(-bT, 0.0));
        const double shift = BARYON ? kMaxBaryon * fabs(pack[DP_ALPHAB * stride + cell]) : 0.0;
        const bool negligible = xe_lo - shift >= range[3];       // NaN: false (kept, and hot)
        cold = !negligible && xe_hi + (PAIR ? shift : 0.0) < (PAIR ? kXePairShared : kXeCold);
        hot = !negligible && !cold;
        dropped = negligible;
        if (negligible && xe_lo - shift < kXeNegligible) {       // beyond the range guard a term is < 1e-295: nothing to bound
          auto pk = [&](int k) { return pack[k * stride + cell]; };
          const DfItemU item = df_make_item_u<MODE, BARYON, (MODE == 2 && !REGULATE)>(pk, sh, ch, cphi, sphi, w);
          dropped_bound = df_item_term_bound<MODE, BARYON, REGULATE>(item, xe_lo, mT_hi, pT_hi, exptab);
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
        auto pk = [&](int k) { return pack[k * stride + cell]; };
        const int slot = cold ? base_cold + __popc(b_cold & below) : kTile - 1 - (base_hot + __popc(b_hot & below));
        items[slot] = df_make_item_u<MODE, BARYON, (MODE == 2 && !REGULATE)>(pk, sh, ch, cphi, sphi, w);
      }
      __syncthreads();
      df_item_loop<MODE, BARYON, REGULATE, OUTFLOW, R, PAIR, false>(items, n_cold, th, thm, mT, mT2, sgn, acc, accm, exptab);
      if (n_hot) df_item_loop<MODE, BARYON, REGULATE, OUTFLOW, R, PAIR, true>(items + (kTile - n_hot), n_hot, th, thm, mT, mT2, sgn, acc, accm, exptab);
      if (t == 0) blk_items += (unsigned)(n_cold + n_hot);
      flip ^= 1;
    }
  }
  if (t == 0) { atomicAdd(g.items_done, blk_items); atomicAdd(g.items_done + 4, blk_dropped); }
  __syncthreads();
  if (t == 0 && blk_bound != 0.0) atomicAdd(&g.bsum[(int64_t)blockIdx.x * gridDim.y + (iy * g.Nphi + iphi)], blk_bound);

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

}  // namespace

// Deterministic sum over the cell chunks and expansion of the species classes: species s takes the bins of its class,
// times (2 pi hbarc)^-3 and its own degeneracy (MomentumSpectra.cpp:38, :399-401).
__global__ void reduce_partials_kernel(const double *__restrict__ partial, int nchunks, int64_t total_class, int64_t per_species,
                                       const int *__restrict__ class_of, const double *__restrict__ deg, int64_t total,
                                       double *__restrict__ out, PruneCheck chk)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t sp = i / per_species, src = (int64_t)class_of[sp] * per_species + (i - sp * per_species);
  double s = 0.0;
  for (int c = 0; c < nchunks; c++) s += partial[(int64_t)c * total_class + src];
  out[i] = kCooperFryePrefactor * deg[sp] * s;
  if (chk.bsum) {
    // a-posteriori test of the dropped items: the bounds of everything dropped for this bin's block row must vanish against the
    // bin itself (src = iy + Ny (iphi + Nphi (class NpT + ipT)); rows are numbered iy Nphi + iphi like blockIdx.y)
    const int64_t jbin = src / chk.NyNphi;
    const int rem = (int)(src - jbin * chk.NyNphi), iy = rem % chk.Ny, iphi = rem / chk.Ny;
    const double b = chk.bsum[(int64_t)chk.bin_row[jbin] * chk.NyNphi + (iy * (chk.NyNphi / chk.Ny) + iphi)];
    if (!(b <= chk.eps * fabs(s) + kPruneFloor)) atomicAdd(chk.violations, 1ull);
  }
}

namespace {

template <int MODE, bool BARYON, bool REGULATE, bool OUTFLOW, bool PAIR>
void launch_df(dim3 grid, cudaStream_t st, const double *pack, int64_t stride, int64_t n, int64_t cpc, const DfGrid &g,
               double *partial, int64_t total)
{
  df_spectra_kernel<MODE, BARYON, REGULATE, OUTFLOW, (PAIR ? kDfPairsPerThread : kDfBinsPerThread), PAIR><<<grid, kThreads, 0, st>>>(pack, stride, n, cpc, g, partial, total);
}

template <int MODE, bool BARYON, bool PAIR>
void dispatch_df2(bool reg, bool outflow, dim3 grid, cudaStream_t st, const double *pack, int64_t stride, int64_t n,
                  int64_t cpc, const DfGrid &g, double *partial, int64_t total)
{
  if (reg && outflow) launch_df<MODE, BARYON, true, true, PAIR>(grid, st, pack, stride, n, cpc, g, partial, total);
  else if (reg) launch_df<MODE, BARYON, true, false, PAIR>(grid, st, pack, stride, n, cpc, g, partial, total);
  else if (outflow) launch_df<MODE, BARYON, false, true, PAIR>(grid, st, pack, stride, n, cpc, g, partial, total);
  else launch_df<MODE, BARYON, false, false, PAIR>(grid, st, pack, stride, n, cpc, g, partial, total);
}

}  // namespace


// class_of[s] = class of species s, rep[c] = first species of class c (classes numbered in order of first appearance).
// The baryon number separates classes only when baryon terms are switched on.
void species_classes_core(int ns, const double *mass, const double *sign, const double *baryon, bool baryon_on,
                          std::vector<int> *class_of, std::vector<int> *rep)
{
  class_of->assign(ns, 0);
  rep->clear();
  for (int s = 0; s < ns; s++) {
    int c = -1;
    const double bs = baryon_on ? baryon[s] : 0.0;
    for (size_t k = 0; k < rep->size() && c < 0; k++) {
      const int r = (*rep)[k];
      const double br = baryon_on ? baryon[r] : 0.0;
      if (mass[r] == mass[s] && sign[r] == sign[s] && br == bs) c = (int)k;
    }
    if (c < 0) { c = (int)rep->size(); rep->push_back(s); }
    (*class_of)[s] = c;
  }
}

void species_classes(const is3d_ctx *ctx, std::vector<int> *class_of, std::vector<int> *rep)
{
  species_classes_core(ctx->ns, ctx->h_mass.data(), ctx->h_sign.data(), ctx->h_baryon.data(), ctx->prm.include_baryon != 0, class_of, rep);
}

// Thread groups of the spectra / dN/dX kernels: R class slots per group, the valid slots of a group carrying ONE baryon
// number (the kernels fold b into per-(item, thread) coefficients, spectra_df.cuh).  Classes are taken per baryon number in
// order of first appearance and each run is padded to a multiple of R with -1; without baryon terms all classes form one
// run.  false: a baryon number outside -kMaxBaryon .. kMaxBaryon (or not an integer).
bool slot_table_core(const std::vector<int> &rep, const double *baryon, bool baryon_on, int R, std::vector<int> *slots)
{
  std::vector<double> bvals;
  for (int r : rep) {
    const double b = baryon_on ? baryon[r] : 0.0;
    bool seen = false;
    for (double v : bvals) seen = seen || (v == b);
    if (!seen) bvals.push_back(b);
  }
  slots->clear();
  for (double b : bvals) {
    if (fabs(b) > (double)kMaxBaryon || b != (double)(int)b) return false;
    for (size_t c = 0; c < rep.size(); c++)
      if ((baryon_on ? baryon[rep[c]] : 0.0) == b) slots->push_back((int)c);
    while (slots->size() % (size_t)R) slots->push_back(-1);
  }
  return true;
}

// Charge-conjugate pairs (df_spectra_kernel<..., PAIR = true>): with baryon terms a baryon class and its antibaryon class
// -- same mass, same statistics, opposite baryon number -- differ only in the b-dependent pieces of df and in exp(-b alpha_B),
// so they are evaluated together.  pairs = 2 class ids per slot (b > 0 first), groups of R slots with ONE |b|, padded with
// (-1, -1); singles = the slot table of every class without a partner (mesons, the deuteron, unmatched baryons).
void species_classes_core(int ns, const double *mass, const double *sign, const double *baryon, bool baryon_on,
                          std::vector<int> *class_of, std::vector<int> *rep);

bool pair_tables_core(const std::vector<int> &rep, const double *mass, const double *sign, const double *baryon, int R, int R_pair,
                      std::vector<int> *singles, std::vector<int> *pairs)
{
  const size_t nc = rep.size();
  std::vector<int> partner(nc, -1);
  for (size_t c = 0; c < nc; c++) {
    const int s = rep[c];
    if (!(baryon[s] > 0.0) || partner[c] >= 0) continue;
    for (size_t d = 0; d < nc; d++) {
      const int t = rep[d];
      if (partner[d] < 0 && d != c && baryon[t] == -baryon[s] && mass[t] == mass[s] && sign[t] == sign[s]) { partner[c] = (int)d; partner[d] = (int)c; break; }
    }
  }
  // pairs, one run per |b|
  pairs->clear();
  std::vector<double> bvals;
  for (size_t c = 0; c < nc; c++)
    if (partner[c] >= 0 && baryon[rep[c]] > 0.0) {
      const double b = baryon[rep[c]];
      bool seen = false;
      for (double v : bvals) seen = seen || (v == b);
      if (!seen) bvals.push_back(b);
    }
  for (double b : bvals) {
    if (b > (double)kMaxBaryon || b != (double)(int)b) return false;
    for (size_t c = 0; c < nc; c++)
      if (partner[c] >= 0 && baryon[rep[c]] == b) { pairs->push_back((int)c); pairs->push_back(partner[c]); }
    while ((pairs->size() / 2) % (size_t)R_pair) { pairs->push_back(-1); pairs->push_back(-1); }
  }
  // singles: the ordinary slot table over the classes without a partner
  std::vector<double> sb;
  for (size_t c = 0; c < nc; c++)
    if (partner[c] < 0) {
      const double b = baryon[rep[c]];
      bool seen = false;
      for (double v : sb) seen = seen || (v == b);
      if (!seen) sb.push_back(b);
    }
  singles->clear();
  for (double b : sb) {
    if (fabs(b) > (double)kMaxBaryon || b != (double)(int)b) return false;
    for (size_t c = 0; c < nc; c++)
      if (partner[c] < 0 && baryon[rep[c]] == b) singles->push_back((int)c);
    while (singles->size() % (size_t)R) singles->push_back(-1);
  }
  return true;
}

bool build_slot_table(const is3d_ctx *ctx, int R, std::vector<int> *slots)
{
  std::vector<int> class_of, rep;
  species_classes(ctx, &class_of, &rep);
  return slot_table_core(rep, ctx->h_baryon.data(), ctx->prm.include_baryon != 0, R, slots);
}

// Builds the species classes and their per-(class, pT) bin arrays shared by all spectra kernels (device pointers in `out`).
is3d_status build_bin_arrays(is3d_ctx *ctx, SpeciesBins *out)
{
  const int ns = ctx->ns, npT = ctx->NpT;
  std::vector<int> class_of, rep;
  species_classes(ctx, &class_of, &rep);
  const int nc = (int)rep.size(), nb = nc * npT;
  std::vector<double> h(5 * (size_t)nb + 4 * (size_t)nc);
  for (int c = 0; c < nc; c++) {
    const int s = rep[c];
    const double m = ctx->h_mass[s], mass2 = m * m;
    for (int ip = 0; ip < npT; ip++) {
      const int j = c * npT + ip;
      const double p = ctx->pT[ip];
      h[0 * nb + j] = sqrt(mass2 + p * p);      // mT, MomentumSpectra.cpp:266
      h[1 * nb + j] = p;
      h[2 * nb + j] = mass2;
      h[3 * nb + j] = ctx->h_baryon[s];
      h[4 * nb + j] = ctx->h_sign[s];
    }
    double *cr = h.data() + 5 * (size_t)nb;
    cr[0 * nc + c] = m; cr[1 * nc + c] = ctx->h_deg[s]; cr[2 * nc + c] = ctx->h_baryon[s]; cr[3 * nc + c] = ctx->h_sign[s];
  }
  void *d = nullptr, *dm = nullptr;
  IS3D_TRY(ctx->get_scratch("bin_arrays", h.size() * sizeof(double), &d));
  IS3D_TRY(ctx->get_scratch("class_of", (size_t)ns * sizeof(int), &dm));
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(dm, class_of.data(), (size_t)ns * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const double *b = (const double *)d;
  out->nclass = nc;
  out->mT = b; out->pT = b + nb; out->m2 = b + 2 * nb; out->baryon = b + 3 * nb; out->sign = b + 4 * nb;
  const double *cr = b + 5 * (size_t)nb;
  out->c_mass = cr; out->c_deg = cr + nc; out->c_baryon = cr + 2 * nc; out->c_sign = cr + 3 * nc;
  out->class_of = (const int *)dm;
  return IS3D_OK;
}

// Chunking policy shared by the spectra kernels.  All blocks of a launch do the same amount of work, so the grid runs in
// waves of `resident` blocks and a partly filled last wave costs a whole block time (the first policy, "about 16 waves",
// gave 4830 blocks = 16.3 waves for the headline launch: 17 block times for 16.3 of work).  Among the chunk counts that give
// 32 .. 64 waves the one with the best wave efficiency blocks / (resident ceil(blocks / resident)) is taken (ties: fewer
// chunks); partial sums stay <= 1 GiB and chunks start on tile boundaries.
void pick_chunks(int64_t ncells, int64_t blocks_per_chunk, int64_t resident, int64_t granule, int64_t max_chunks,
                 int *nchunks, int64_t *cells_per_chunk)
{
  if (ncells <= 0) { *nchunks = 1; *cells_per_chunk = granule; return; }
  if (max_chunks < 1) max_chunks = 1;
  const int64_t by_cells = (ncells + granule - 1) / granule;
  if (max_chunks > by_cells) max_chunks = by_cells < 1 ? 1 : by_cells;
  int64_t lo = (32 * resident + blocks_per_chunk - 1) / blocks_per_chunk, hi = (64 * resident + blocks_per_chunk - 1) / blocks_per_chunk;
  if (lo < 1) lo = 1;
  if (lo > max_chunks) lo = max_chunks;
  if (hi > max_chunks) hi = max_chunks;
  double best_eff = -1.0;
  int64_t best_nc = lo, best_cpc = ncells;
  for (int64_t nc = lo; nc <= hi; nc++) {
    int64_t cpc = (ncells + nc - 1) / nc;
    cpc = (cpc + granule - 1) / granule * granule;
    const int64_t nca = (ncells + cpc - 1) / cpc;
    const int64_t blocks = nca * blocks_per_chunk, waves = (blocks + resident - 1) / resident;
    const double eff = (double)blocks / (double)(waves * resident);
    if (eff > best_eff + 1e-12) { best_eff = eff; best_nc = nca; best_cpc = cpc; }
  }
  *nchunks = (int)best_nc;
  *cells_per_chunk = best_cpc;
}

void choose_chunks(const is3d_ctx *ctx, int64_t ncells, int64_t blocks_per_chunk, int64_t total, int tile, int blocks_per_sm,
                   int *nchunks, int64_t *cells_per_chunk)
{
  const int64_t resident = blocks_per_sm * (int64_t)ctx->sm_count;
  int64_t max_by_mem = ((int64_t)1 << 30) / (total * 8);
  if (max_by_mem > 65535) max_by_mem = 65535;
  pick_chunks(ncells, blocks_per_chunk, resident, tile, max_by_mem, nchunks, cells_per_chunk);
}

// ctx-free cores of the launch order (also behind the host-only helper is3d_launch_order)
std::vector<int> column_order_core(const std::vector<int> &slots, int ids_per_group, const double *class_mass, int NpT, const double *pT)
{
  const int ngroups = (int)(slots.size() / ids_per_group);
  std::vector<double> key((size_t)ngroups * NpT);
  for (int gi = 0; gi < ngroups; gi++) {
    double m_min = 1e300;
    for (int k = 0; k < ids_per_group; k++) {
      const int cls = slots[(size_t)gi * ids_per_group + k];
      if (cls >= 0) m_min = fmin(m_min, fabs(class_mass[cls]));
    }
    for (int ip = 0; ip < NpT; ip++) key[(size_t)gi * NpT + ip] = sqrt(m_min * m_min + pT[ip] * pT[ip]);
  }
  std::vector<int> order(key.size());
  for (size_t k = 0; k < order.size(); k++) order[k] = (int)k;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
  return order;
}

void fill_bin_rows_core(const std::vector<int> &slots, int ids_per_group, const std::vector<int> &order, int NpT, int threads_per_block,
                        int row0, std::vector<int> *bin_row)
{
  for (size_t c = 0; c < order.size(); c++) {
    const int grp = order[c] / NpT, ip = order[c] - grp * NpT;
    for (int k = 0; k < ids_per_group; k++) {
      const int cls = slots[(size_t)grp * ids_per_group + k];
      if (cls >= 0) (*bin_row)[(size_t)cls * NpT + ip] = row0 + (int)(c / threads_per_block);
    }
  }
}

std::vector<int> column_order(const is3d_ctx *ctx, const std::vector<int> &slots, int ids_per_group, const std::vector<int> &rep)
{
  std::vector<double> class_mass(rep.size());
  for (size_t c = 0; c < rep.size(); c++) class_mass[c] = ctx->h_mass[rep[c]];
  return column_order_core(slots, ids_per_group, class_mass.data(), ctx->NpT, ctx->pT.data());
}

std::vector<int> rapidity_order(const is3d_ctx *ctx)
{
  std::vector<int> yo(ctx->Ny);
  double ymid = 0.0;
  for (int k = 0; k < ctx->Ny; k++) { yo[k] = k; ymid += ctx->yv[k] / ctx->Ny; }
  std::stable_sort(yo.begin(), yo.end(), [&](int a, int b) { return fabs(ctx->yv[a] - ymid) < fabs(ctx->yv[b] - ymid); });
  return yo;
}

void fill_bin_rows(const is3d_ctx *ctx, const std::vector<int> &slots, int ids_per_group, const std::vector<int> &order, int threads_per_block,
                   int row0, std::vector<int> *bin_row)
{
  fill_bin_rows_core(slots, ids_per_group, order, ctx->NpT, threads_per_block, row0, bin_row);
}

is3d_status run_spectra_df(is3d_ctx *ctx, double *out_dev, is3d_stats *stats)
{
  const is3d_params &p = ctx->prm;
  int64_t prune_reruns = 0;
  const int64_t n = ctx->surf.n;
  const int64_t total = (int64_t)ctx->ns * ctx->NpT * ctx->Nphi * ctx->Ny;
  DfFlags fl;
  fl.df_mode = p.df_mode; fl.dimension = p.dimension; fl.include_baryon = p.include_baryon;
  fl.include_bulk = p.include_bulk_deltaf; fl.include_shear = p.include_shear_deltaf;
  fl.include_baryondiff = p.include_baryondiff_deltaf;

  DfGrid g;
  SpeciesBins sb;
  IS3D_TRY(build_bin_arrays(ctx, &sb));
  g.mT = sb.mT; g.pT = sb.pT; g.baryon = sb.baryon; g.sign = sb.sign;
  // thread groups: single classes (one baryon number per group) and, with baryon terms, charge-conjugate pairs
  std::vector<int> slots, pair_slots;
  bool ok;
  if (p.include_baryon) {
    std::vector<int> class_of, rep;
    species_classes(ctx, &class_of, &rep);
    ok = pair_tables_core(rep, ctx->h_mass.data(), ctx->h_sign.data(), ctx->h_baryon.data(), kDfBinsPerThread, kDfPairsPerThread, &slots, &pair_slots);
  } else {
    ok = build_slot_table(ctx, kDfBinsPerThread, &slots);
  }
  if (!ok) {
    ctx->set_error("species list holds a baryon number outside -2..2 (the reference's PDG readers produce hadrons and the deuteron only)");
    return IS3D_ERR_INVALID;
  }
  // thread columns (group, pT) in order of their smallest mT: the columns of a block then agree on which cells are negligible
  std::vector<int> class_of_all, rep_all;
  species_classes(ctx, &class_of_all, &rep_all);
  const std::vector<int> order_single = column_order(ctx, slots, kDfBinsPerThread, rep_all), order_pair = column_order(ctx, pair_slots, 2 * kDfPairsPerThread, rep_all);
  void *d_slots = nullptr;
  std::vector<int> both(slots);
  both.insert(both.end(), pair_slots.begin(), pair_slots.end());
  both.insert(both.end(), order_single.begin(), order_single.end());
  both.insert(both.end(), order_pair.begin(), order_pair.end());
  {
    const std::vector<int> yo = rapidity_order(ctx);
    both.insert(both.end(), yo.begin(), yo.end());
  }
  IS3D_TRY(ctx->get_scratch("k1_slots", (both.size() + 1) * sizeof(int), &d_slots));
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(d_slots, both.data(), both.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));       // `both` is pageable host memory
  g.slot_class = (const int *)d_slots;
  g.col_map = (const int *)d_slots + slots.size() + pair_slots.size();
  g.y_order = g.col_map + order_single.size() + order_pair.size();
  g.ns = sb.nclass; g.NpT = ctx->NpT; g.ncols = ctx->NpT * (int)(slots.size() / kDfBinsPerThread);
  const int64_t per_species = (int64_t)ctx->NpT * ctx->Nphi * ctx->Ny;
  const int64_t total_class = (int64_t)sb.nclass * per_species;
  g.Ny = ctx->Ny; g.Nphi = ctx->Nphi; g.Neta = ctx->Neta; g.dimension = p.dimension;
  g.yv = ctx->d_y; g.cosphi = ctx->d_cosphi; g.sinphi = ctx->d_sinphi; g.etav = ctx->d_eta; g.etaw = ctx->d_etaw; g.exptab = ctx->d_exptab;
  DfGrid gp = g;                                                 // the pair launch: two class ids per slot
  gp.slot_class = (const int *)d_slots + slots.size();
  gp.col_map = g.col_map + order_single.size();
  gp.ncols = ctx->NpT * (int)(pair_slots.size() / (2 * kDfPairsPerThread));
  const int nslices = (g.ncols + kThreads - 1) / kThreads, nslices_pair = (gp.ncols + kThreads - 1) / kThreads;
  // a pair block does about 1.6x the work of a single block; the wave policy counts blocks
  const int64_t blocks_per_chunk = (int64_t)(nslices + nslices_pair) * ctx->Ny * ctx->Nphi;
  if ((int64_t)ctx->Ny * ctx->Nphi > 65535) { ctx->set_error("Ny*Nphi exceeds 65535"); return IS3D_ERR_INVALID; }

  const int64_t macro = pass_cells(4 << 20);          // cells per pass: bounds the pack scratch to ~1 GB
  const int64_t stride = n < macro ? n : macro;
  int nchunks; int64_t cpc;
  choose_chunks(ctx, stride, blocks_per_chunk, total_class, kTile, IS3D_K1_MINBLOCKS, &nchunks, &cpc);

  void *pack = nullptr, *partial = nullptr, *counters = nullptr, *prune = nullptr;
  IS3D_TRY(ctx->get_scratch("cell_pack", (size_t)DP_SIZE * stride * sizeof(double), &pack));
  IS3D_TRY(ctx->get_scratch("partial", (size_t)nchunks * total_class * sizeof(double), &partial));
  IS3D_TRY(ctx->get_scratch("counters", 16 * sizeof(unsigned long long), &counters));
  // dropping of negligible items: [Ny + 1] row scales from df_amin_kernel | bounds of the dropped terms per block row | bin -> row
  const int NyNphi = ctx->Ny * ctx->Nphi, nrows = nslices + nslices_pair;
  const size_t amin_bytes = (size_t)(ctx->Ny + 1) * 8, bsum_bytes = (size_t)nrows * NyNphi * 8;
  std::vector<int> bin_row((size_t)sb.nclass * ctx->NpT, 0);
  fill_bin_rows(ctx, slots, kDfBinsPerThread, order_single, kThreads, 0, &bin_row);
  fill_bin_rows(ctx, pair_slots, 2 * kDfPairsPerThread, order_pair, kThreads, nslices, &bin_row);
  IS3D_TRY(ctx->get_scratch("k1_prune", amin_bytes + bsum_bytes + bin_row.size() * sizeof(int), &prune));
  unsigned long long *d_amin = (unsigned long long *)prune;
  double *d_bsum = (double *)((char *)prune + amin_bytes);
  int *d_bin_row = (int *)((char *)prune + amin_bytes + bsum_bytes);
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(d_bin_row, bin_row.data(), bin_row.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));       // pageable host memory
  g.items_done = (unsigned long long *)counters + 2;
  gp.items_done = (unsigned long long *)counters + 3;
  g.amin_bits = gp.amin_bits = d_amin;
  g.bsum = d_bsum; gp.bsum = d_bsum + (size_t)nslices * NyNphi;
  cudaEvent_t e0 = ctx->ev0, e1 = ctx->ev1;             // owned by the context: nothing to release on an error path
  float ms_total = 0.f;
  int64_t launches = 0;
  unsigned long long h_counters[16];

  // attempt 0 drops items below the margin; if the a-posteriori test fails for any bin, attempt 1 repeats the call without it
  for (int attempt = 0; attempt < 2; attempt++) {
  g.margin = gp.margin = attempt == 0 ? p.negligible_margin : 0.0;
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(partial, 0, (size_t)nchunks * total_class * sizeof(double), ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(counters, 0, 16 * sizeof(unsigned long long), ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(d_bsum, 0, bsum_bytes, ctx->stream));

  for (int64_t begin = 0; begin < n; begin += macro) {
    int64_t count = n - begin < macro ? n - begin : macro;
    df_setup_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(
        ctx->surf, begin, count, ctx->tb, fl, (double *)pack, stride, (unsigned long long *)counters);
    IS3D_CUDA_TRY(ctx, cudaGetLastError());
    int nch = (int)((count + cpc - 1) / cpc);
    dim3 grid(nslices, ctx->Ny * ctx->Nphi, nch), grid_pair(nslices_pair, ctx->Ny * ctx->Nphi, nch);
    IS3D_CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
    if (g.margin > 0.0) {
      IS3D_CUDA_TRY(ctx, cudaMemsetAsync(d_amin, 0x7f, amin_bytes - 8, ctx->stream));
      IS3D_CUDA_TRY(ctx, cudaMemsetAsync(d_amin + ctx->Ny, 0, 8, ctx->stream));
      int64_t ab = (count + 255) / 256, ab_max = 8 * (int64_t)ctx->sm_count;
      df_amin_kernel<<<(unsigned)(ab < ab_max ? ab : ab_max), 256, amin_bytes, ctx->stream>>>((double *)pack, stride, count, ctx->Ny, ctx->d_y, p.dimension, ctx->Neta, ctx->d_eta, d_amin);
      IS3D_CUDA_TRY(ctx, cudaGetLastError());
      launches += 1;
    }
    const bool reg = p.regulate_deltaf != 0, outflow = p.outflow != 0;
    // the pair launch first (its blocks are the longer ones); the single-class launch runs beside it on a second stream and
    // fills the tail of its last wave (disjoint bins of `partial`)
    const bool two = p.include_baryon && nslices_pair && nslices;
    cudaStream_t s_single = two ? ctx->side_stream : ctx->stream;
    if (two) {
      IS3D_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
      IS3D_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
    }
    if (p.df_mode == 1) {
      if (p.include_baryon && nslices_pair) dispatch_df2<1, true, true>(reg, outflow, grid_pair, ctx->stream, (double *)pack, stride, count, cpc, gp, (double *)partial, total_class);
      if (p.include_baryon && nslices) dispatch_df2<1, true, false>(reg, outflow, grid, s_single, (double *)pack, stride, count, cpc, g, (double *)partial, total_class);
      if (!p.include_baryon) dispatch_df2<1, false, false>(reg, outflow, grid, ctx->stream, (double *)pack, stride, count, cpc, g, (double *)partial, total_class);
    } else {
      if (p.include_baryon && nslices_pair) dispatch_df2<2, true, true>(reg, outflow, grid_pair, ctx->stream, (double *)pack, stride, count, cpc, gp, (double *)partial, total_class);
      if (p.include_baryon && nslices) dispatch_df2<2, true, false>(reg, outflow, grid, s_single, (double *)pack, stride, count, cpc, g, (double *)partial, total_class);
      if (!p.include_baryon) dispatch_df2<2, false, false>(reg, outflow, grid, ctx->stream, (double *)pack, stride, count, cpc, g, (double *)partial, total_class);
    }
    IS3D_CUDA_TRY(ctx, cudaGetLastError());
    if (two) {
      IS3D_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join, ctx->side_stream));
      IS3D_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    }
    IS3D_CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaEventSynchronize(e1));
    float ms = 0.f;
    IS3D_CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
    ms_total += ms;
    launches += 2 + (nslices_pair && nslices ? 1 : 0);
  }
  PruneCheck chk;                                      // without a margin only the < 1e-295 items are dropped: nothing to test
  if (g.margin > 0.0) chk.bsum = d_bsum;
  chk.bin_row = d_bin_row; chk.Ny = ctx->Ny; chk.NyNphi = NyNphi; chk.eps = kPruneEps;
  chk.violations = (unsigned long long *)counters + 4;     // counters: 2, 3 items marched (single, pair launch), 6, 7 items dropped
  reduce_partials_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((double *)partial, nchunks, total_class, per_species,
                                                                                sb.class_of, ctx->d_deg, total, out_dev, chk);
  IS3D_CUDA_TRY(ctx, cudaGetLastError());
  launches += 1;
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (h_counters[4] == 0 || h_counters[1] != 0 || !(g.margin > 0.0)) break;
  prune_reruns++;
  }
  if (stats) {
    stats->cells_total = n;
    stats->cells_skipped = (int64_t)h_counters[0];
    stats->cells_out_of_table = (int64_t)h_counters[1];
    stats->kernel_ms = ms_total;
    stats->kernel_launches = launches;
    // items the blocks marched over (negligible items are dropped at tile-build time) x kThreads thread columns x R slots
    stats->pair_evals_executed = 2 * (int64_t)h_counters[3] * kThreads * kDfPairsPerThread;
    stats->evals_executed = (int64_t)h_counters[2] * kThreads * kDfBinsPerThread + stats->pair_evals_executed;
    stats->evals_dropped = ((int64_t)h_counters[6] * kDfBinsPerThread + 2 * (int64_t)h_counters[7] * kDfPairsPerThread) * kThreads;
    stats->prune_reruns = prune_reruns;
  }
  if (h_counters[1] != 0) {
    ctx->set_error(std::to_string(h_counters[1]) + " cell(s) outside the df coefficient tables (the reference aborts here)");
    return IS3D_ERR_TABLE_RANGE;
  }
  return IS3D_OK;
}

}  // namespace is3d

extern "C" int is3d_species_pairs(int ns, const double *mass, const double *sign, const double *baryon, int slots_per_group,
                                  int *single_slots, int single_capacity, int *pair_slots, int pair_capacity, int *n_single, int *n_pair)
{
  if (ns <= 0 || !mass || !sign || !baryon || slots_per_group <= 0 || !single_slots || !pair_slots || !n_single || !n_pair) return -1;
  std::vector<int> cls, rep, singles, pairs;
  is3d::species_classes_core(ns, mass, sign, baryon, true, &cls, &rep);
  if (!is3d::pair_tables_core(rep, mass, sign, baryon, slots_per_group, slots_per_group, &singles, &pairs)) return -3;
  if ((int)singles.size() > single_capacity || (int)pairs.size() > pair_capacity) return -2;
  for (size_t k = 0; k < singles.size(); k++) single_slots[k] = singles[k];
  for (size_t k = 0; k < pairs.size(); k++) pair_slots[k] = pairs[k];
  *n_single = (int)singles.size();
  *n_pair = (int)pairs.size();
  return (int)rep.size();
}

extern "C" int is3d_launch_order(int nslots, const int *slots, int ids_per_group, int nclass, const double *class_mass, int NpT, const double *pT,
                                 int threads_per_block, int *order, int *bin_row)
{
  if (nslots <= 0 || !slots || ids_per_group <= 0 || nslots % ids_per_group || nclass <= 0 || !class_mass || NpT <= 0 || !pT ||
      threads_per_block <= 0 || !order || !bin_row) return -1;
  std::vector<int> sl(slots, slots + nslots);
  for (int c : sl) if (c >= nclass) return -1;
  const std::vector<int> ord = is3d::column_order_core(sl, ids_per_group, class_mass, NpT, pT);
  std::vector<int> rows((size_t)nclass * NpT, -1);
  is3d::fill_bin_rows_core(sl, ids_per_group, ord, NpT, threads_per_block, 0, &rows);
  for (size_t k = 0; k < ord.size(); k++) order[k] = ord[k];
  for (size_t k = 0; k < rows.size(); k++) bin_row[k] = rows[k];
  return (int)ord.size();
}

extern "C" int is3d_species_groups(int ns, const double *mass, const double *sign, const double *baryon, int include_baryon,
                                   int slots_per_group, int *class_of, int *slot_class, int capacity, int *nclass)
{
  if (ns <= 0 || !mass || !sign || !baryon || slots_per_group <= 0 || !class_of || !slot_class || !nclass) return -1;
  std::vector<int> cls, rep, slots;
  is3d::species_classes_core(ns, mass, sign, baryon, include_baryon != 0, &cls, &rep);
  if (!is3d::slot_table_core(rep, baryon, include_baryon != 0, slots_per_group, &slots)) return -3;
  if ((int)slots.size() > capacity) return -2;
  for (int s = 0; s < ns; s++) class_of[s] = cls[s];
  for (size_t k = 0; k < slots.size(); k++) slot_class[k] = slots[k];
  *nclass = (int)rep.size();
  return (int)slots.size();
}
