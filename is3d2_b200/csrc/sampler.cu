// K5 + K6: mean yields and the Monte-Carlo particle sampler (operation 2, df_mode 1-4) on sm_100a.
// Replaces calculate_total_yield, sample_dN_pTdpTdphidy and the BinSampledParticle counters
// (reference src/cpp/ParticleSampler.cpp:447-1134, BinSampledParticle.cpp).
//
// The reference loops cell -> event -> Poisson(dn_tot) hadrons.  Independent Poisson draws per event are
// equivalent to ONE Poisson draw of mean E * dn_tot per cell for a BLOCK of E events followed by a uniform event label
// inside the block per hadron, so the GPU pipeline is flat in (event block, cell, hadron):
//   1. sampler_setup_kernel   thread per cell: LRF quantities, df coefficients, breakdown test, mean yield dn_tot
//                             -> 59-double pack (once per call)
//   then, per PASS = a run of consecutive 64-event blocks (kEventBlock is part of the random-stream keying, the pass
//   size is not: the sampled set does not depend on how the call is cut into passes, kernels or GPUs):
//   2. sampler_count_kernel   thread per (block, cell): N ~ Poisson(E dn_tot) from the (cell, block) Philox stream;
//      exclusive scan of the counts (cub::DeviceScan, plumbing) -> proposal offsets
//   3. sampler_hadron_kernel  thread per proposed hadron: (block, cell) by binary search in the offsets, event label, species
//                             by inverse CDF over the (cell-independent) cumulative density tables, thermal momentum
//                             by rejection, viscous/flux weights, accept -> record or self-test histograms
//   4. stable radix sort of (event, proposal index) (cub, plumbing) + gather -> the pass's particles grouped by event, in a
//      deterministic order; passes cover ascending event ranges, so the gathered records of pass k are the next
//      contiguous piece of the final list and go to the host on a second stream while pass k + 1 samples.
#include <cub/cub.cuh>

#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#include "ctx.h"
#include "aniso.cuh"
#include "sampler.cuh"

namespace is3d {

namespace {

// events per Poisson block.  Part of the random-stream keying: changing it changes the sampled set (not its distribution).
constexpr int kEventBlock = 64;

struct SamplerTables {
  int ns;
  const double *mass, *sign, *baryon;
  const int *mcid;
  const double *cumA, *cumB;      // inclusive cumulative sums over species of neq and dn_bulk
  const double *cell_cdf;         // [cell][ns] inclusive cumulative species densities of each cell, or NULL (fast mode)
  double totA, totB, totD;        // sums of neq, dn_bulk, dn_diff
};

__global__ void sampler_setup_kernel(SurfaceView surf, int64_t begin, int64_t count, int64_t global_offset, DfTables tb,
                                     SamplerFlags fl, const double *__restrict__ gla_root, const double *__restrict__ gla_weight,
                                     int gla_pts, SamplerTables st, double nevents, uint64_t seed, double *__restrict__ pack,
                                     int64_t stride, unsigned long long *__restrict__ ncount, double *__restrict__ yield,
                                     unsigned long long *counters)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Cell c = load_cell(surf, begin + i, fl.include_baryon != 0);
  double p[SP_SIZE];
  int status = sampler_setup_cell(c, tb, fl, gla_root, gla_weight, gla_pts, st.totA, st.totB, p);
#pragma unroll
  for (int k = 0; k < SP_SIZE; k++) pack[k * stride + i] = p[k];
  unsigned long long n = 0;
  double y = 0.0;
  if (status == CELL_SKIPPED) atomicAdd(&counters[0], 1ull);
  else if (status == CELL_OUT_OF_TABLE) atomicAdd(&counters[1], 1ull);
  else {
    if (status & CELL_BREAKDOWN) atomicAdd(&counters[2], 1ull);
    y = cell_mean_yield(p, fl.df_mode, st.totA, st.totB, st.totD);
  }
  (void)nevents; (void)seed; (void)global_offset;
  if (ncount) ncount[i] = n;
  if (yield) yield[i] = y;
}

struct HistGrid {
  int test_sampler;
  double y_cut, y_width, eta_cut, eta_width, pT_min, pT_width, phi_width, tau_min, tau_width, r_min, r_width;
  int y_bins, eta_bins, pT_bins, phi_bins, tau_bins, r_bins, ns;
  double *dN_dy, *dN_deta, *dN_dphip, *dN_pT, *pT_count, *vn_re, *vn_im, *dN_tau, *dN_r, *dN_phis;
};

// BinSampledParticle.cpp:9-133
__device__ void bin_particle(const HistGrid &h, int s, const LabParticle &q, double tau, double x, double y)
{
  int iy = (int)floor((q.rapidity + h.y_cut) / h.y_width);
  if (iy >= 0 && iy < h.y_bins) atomicAdd(&h.dN_dy[(size_t)s * h.y_bins + iy], 1.0);
  int ieta = (int)floor((q.eta + h.eta_cut) / h.eta_width);
  if (ieta >= 0 && ieta < h.eta_bins) atomicAdd(&h.dN_deta[(size_t)s * h.eta_bins + ieta], 1.0);
  double phip = atan2(q.py, q.px);
  if (phip < 0.0) phip += kTwoPi;
  int iphip = (int)floor(phip / h.phi_width);
  if (iphip >= 0 && iphip < h.phi_bins) atomicAdd(&h.dN_dphip[(size_t)s * h.phi_bins + iphip], 1.0);
  double pT = sqrt(q.px * q.px + q.py * q.py);
  int ipT = (int)floor((pT - h.pT_min) / h.pT_width);
  if (ipT >= 0 && ipT < h.pT_bins) {
    atomicAdd(&h.dN_pT[(size_t)s * h.pT_bins + ipT], 1.0);
    atomicAdd(&h.pT_count[(size_t)s * h.pT_bins + ipT], 1.0);
    for (int k = 0; k < 7; k++) {
      size_t j = ((size_t)k * h.ns + s) * h.pT_bins + ipT;
      atomicAdd(&h.vn_re[j], cos(((double)k + 1.0) * phip));
      atomicAdd(&h.vn_im[j], sin(((double)k + 1.0) * phip));
    }
  }
  double r = sqrt(x * x + y * y), phis = atan2(y, x);
  if (phis < 0.0) phis += kTwoPi;
  int itau = (int)floor((tau - h.tau_min) / h.tau_width), ir = (int)floor((r - h.r_min) / h.r_width);
  int iphis = (int)floor(phis / h.phi_width);
  if (itau >= 0 && itau < h.tau_bins) atomicAdd(&h.dN_tau[(size_t)s * h.tau_bins + itau], 1.0);
  if (ir >= 0 && ir < h.r_bins) atomicAdd(&h.dN_r[(size_t)s * h.r_bins + ir], 1.0);
  if (iphis >= 0 && iphis < h.phi_bins) atomicAdd(&h.dN_phis[(size_t)s * h.phi_bins + iphis], 1.0);
}

// accepted hadrons of one pass, compacted on the device
struct SamplerOut {
  is3d_particle *rec;                 // [capacity] records in acceptance (= scheduling) order
  unsigned long long *key;            // [capacity] (event << 40) | proposal index: sorting by it gives a geometry-independent order
  unsigned long long *nacc;           // running number of accepted records
  unsigned long long *event_counts;   // [nevents] accepted hadrons per event (all passes)
};

// one instantiation per df_mode: the branches of the other modes would only add instruction-cache pressure
template <int DF_MODE>
__global__ void __launch_bounds__(128)
sampler_hadron_kernel(const double *__restrict__ pack, int64_t stride, int64_t ncells, int64_t cell_global0, int block0, int nblocks,
                      const unsigned long long *__restrict__ offsets, unsigned long long nprop,
                      SamplerTables st, int dimension, double y_cut, long nevents, uint64_t seed, HistGrid hg,
                      SamplerOut out, unsigned long long *counters)
{
  const unsigned long long jg = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;   // proposal index in this pass
  bool accept = false;
  long samples = 0;
  int event = 0;
  is3d_particle r;
  if (jg < nprop) {
    // (block, cell) = last entry v with offsets[v] <= jg
    int64_t lo = 0, hi = (int64_t)nblocks * ncells;
    while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (offsets[mid] <= jg) lo = mid; else hi = mid; }
    const int blk = block0 + (int)(lo / ncells);
    const int64_t cell = lo - (int64_t)(blk - block0) * ncells;
    const uint32_t n = (uint32_t)(jg - offsets[lo]);
    auto pk = [&](int k) { return pack[k * stride + cell]; };
    Philox rng;
    rng.init(seed, (uint64_t)(cell_global0 + cell), n, (uint32_t)blk);
    // uniform event label inside the block (the reference draws every event's Poisson number separately, :919-922)
    const long first = (long)blk * kEventBlock;
    const long ev = nevents - first < kEventBlock ? nevents - first : kEventBlock;
    int within = (int)(rng.canonical() * (double)ev);
    if (within >= ev) within = (int)ev - 1;
    event = (int)first + within;
    int s;
    if (st.cell_cdf) {
      // per-cell discrete distribution over species (fast = 0 and df_mode 5): inclusive cumulative row of this cell
      const double *cdf = st.cell_cdf + (size_t)cell * st.ns;
      const double target = rng.canonical() * cdf[st.ns - 1];
      int a = 0, b = st.ns - 1;
      while (a < b) { int m = (a + b) >> 1; if (cdf[m] > target) b = m; else a = m + 1; }
      s = a;
    } else {
      // species by inverse CDF of w_s = WA neq_s + WB dn_bulk_s (discrete_distribution, :919-931)
      const double WA = pk(SP_WA), WB = pk(SP_WB);
      const double target = rng.canonical() * (WA * st.totA + WB * st.totB);
      int a = 0, b = st.ns - 1;
      while (a < b) { int m = (a + b) >> 1; if (WA * st.cumA[m] + WB * st.cumB[m] > target) b = m; else a = m + 1; }
      s = a;
    }
    const double mass = st.mass[s], sign = st.sign[s], baryon = st.baryon[s];
    LrfMomentum p;
    accept = sample_hadron<DF_MODE>(rng, pk, mass, sign, baryon, &samples, &p);
    if (accept) {
      const double y_max = (dimension == 2) ? y_cut : 0.5;
      LabParticle q = boost_to_lab(rng, pk, p, mass, dimension, y_max);
      if (hg.test_sampler) {
        bin_particle(hg, s, q, pk(SP_TAU), pk(SP_X), pk(SP_Y));
      } else {
        r.chosen_index = s; r.mcid = st.mcid[s]; r.event = event; r.pad_ = 0;
        r.mass = mass; r.tau = pk(SP_TAU); r.x = pk(SP_X); r.y = pk(SP_Y); r.eta = q.eta;
        r.t = q.t; r.z = q.z; r.E = q.E; r.px = q.px; r.py = q.py; r.pz = q.pz;
      }
    }
  }
  // warp-aggregated counters: proposals of the rejection loops, accepted hadrons, output slots
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  unsigned long long wsamples = (unsigned long long)samples;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wsamples += __shfl_xor_sync(full, wsamples, o);
  const unsigned amask = __ballot_sync(full, accept);
  if (lane == 0) {
    atomicAdd(&counters[6], wsamples);
    if (amask) atomicAdd(&counters[7], (unsigned long long)__popc(amask));
  }
  if (!hg.test_sampler && amask) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(out.nacc, (unsigned long long)__popc(amask));
    base = __shfl_sync(full, base, 0);
    if (accept) {
      const unsigned long long slot = base + __popc(amask & ((1u << lane) - 1u));
      out.rec[slot] = r;
      out.key[slot] = ((unsigned long long)event << 40) | jg;
      atomicAdd(&out.event_counts[event], 1ull);
    }
  }
}

__global__ void iota_kernel(unsigned int *v, unsigned long long n)
{
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (unsigned int)i;
}

__global__ void gather_particles_kernel(const is3d_particle *__restrict__ in, const unsigned int *__restrict__ idx,
                                        unsigned long long n, is3d_particle *__restrict__ out)
{
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[idx[i]];
}

__global__ void yield_reduce_kernel(const double *__restrict__ v, int64_t n, double *__restrict__ block_sums)
{
  __shared__ double sh[256];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += v[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) { if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w]; __syncthreads(); }
  if (threadIdx.x == 0) block_sums[blockIdx.x] = sh[0];
}

// everything the sampler kernels need besides the surface
struct SamplerSetup {
  SamplerFlags fl;
  SamplerTables st;
};

is3d_status prepare_sampler(is3d_ctx *ctx, SamplerSetup *ss)
{
  const is3d_params &p = ctx->prm;
  if (p.df_mode == 5 && ctx->npdg <= 0) { ctx->set_error("PDG table not set (is3d_set_pdg)"); return IS3D_ERR_INVALID; }
  if (ctx->gla_pts <= 0) { ctx->set_error("Gauss-Laguerre tables not set"); return IS3D_ERR_INVALID; }
  if (!ctx->have_avg) { ctx->set_error("thermodynamic averages not set"); return IS3D_ERR_INVALID; }
  SamplerFlags &fl = ss->fl;
  fl.df_mode = p.df_mode; fl.dimension = p.dimension; fl.include_baryon = p.include_baryon;
  fl.include_bulk = p.include_bulk_deltaf; fl.include_shear = p.include_shear_deltaf; fl.include_baryondiff = p.include_baryondiff_deltaf;
  fl.fast = p.fast; fl.deta_min = p.deta_min; fl.mass_pion0 = p.mass_pion0; fl.bulkPi_over_P_max = ctx->tb.bulkPi_over_P_max;
  fl.y_cut = p.y_cut; fl.T_avg = ctx->T_avg; fl.F_avg = 0.0; fl.betabulk_avg = 1.0;
  if (p.df_mode == 3 && p.fast) {
    // df coefficients at the surface averages (ParticleSampler.cpp:660-669), evaluated on the host copies of the tables
    DfTables ht = ctx->tb;
    std::vector<double> cF(ctx->h_T.size()), cB(ctx->h_T.size()), cP(ctx->h_T.size());
    ht.T = ctx->h_T.data(); ht.muB = ctx->h_muB.data();
    for (int k = 0; k < 10; k++) ht.tab[k] = ctx->h_tab[k].data();
    const int nT = (int)ctx->h_T.size();
    natural_cspline_coefficients(ctx->h_T.data(), ctx->h_tab[TAB_F].data(), nT, cF.data());
    natural_cspline_coefficients(ctx->h_T.data(), ctx->h_tab[TAB_BETABULK].data(), nT, cB.data());
    natural_cspline_coefficients(ctx->h_T.data(), ctx->h_tab[TAB_BETAPI].data(), nT, cP.data());
    ht.sp_F = {ctx->h_T.data(), ctx->h_tab[TAB_F].data(), cF.data(), nT};
    ht.sp_betabulk = {ctx->h_T.data(), ctx->h_tab[TAB_BETABULK].data(), cB.data(), nT};
    ht.sp_betapi = {ctx->h_T.data(), ctx->h_tab[TAB_BETAPI].data(), cP.data(), nT};
    DfCoeff d;
    if (!evaluate_df_coefficients(ht, 3, p.include_baryon, ctx->T_avg, ctx->muB_avg, 0.0, 0.0, 0.0, &d)) {
      ctx->set_error("surface-averaged (T, muB) outside the df coefficient tables");
      return IS3D_ERR_TABLE_RANGE;
    }
    fl.F_avg = d.F; fl.betabulk_avg = d.betabulk;
  }
  // cumulative species tables
  const int ns = ctx->ns;
  std::vector<double> cum(2 * (size_t)ns);
  double a = 0.0, b = 0.0, d = 0.0;
  for (int s = 0; s < ns; s++) { a += ctx->h_neq[s]; b += ctx->h_dnbulk[s]; d += ctx->h_dndiff[s]; cum[s] = a; cum[ns + s] = b; }
  void *dc = nullptr;
  IS3D_TRY(ctx->get_scratch("sampler_cum", cum.size() * sizeof(double), &dc));
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(dc, cum.data(), cum.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  SamplerTables &st = ss->st;
  st.ns = ns; st.mass = ctx->d_mass; st.sign = ctx->d_sign; st.baryon = ctx->d_baryon; st.mcid = ctx->d_mcid;
  st.cumA = (const double *)dc; st.cumB = (const double *)dc + ns;
  st.cell_cdf = nullptr;
  st.totA = a; st.totB = b; st.totD = d;
  return IS3D_OK;
}

is3d_status fill_stats(is3d_ctx *ctx, void *counters, is3d_stats *stats, float ms, int64_t launches)
{
  unsigned long long h[16];
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (stats) {
    stats->cells_total = ctx->surf.n;
    stats->cells_skipped = (int64_t)h[0]; stats->cells_out_of_table = (int64_t)h[1]; stats->cells_breakdown = (int64_t)h[2];
    stats->cells_pl_negative = (int64_t)h[3]; stats->reconstruction_failures = (int64_t)h[8]; stats->newton_iterations = (int64_t)h[9];
    stats->sampler_proposals = (int64_t)h[6]; stats->sampler_accepted = (int64_t)h[7];
    stats->kernel_ms = ms; stats->kernel_launches = launches;
  }
  if (h[1] != 0) {
    ctx->set_error(std::to_string(h[1]) + " cell(s) outside the df coefficient tables (the reference aborts here)");
    return IS3D_ERR_TABLE_RANGE;
  }
  return IS3D_OK;
}

// the famod stage counts into its own array (same slot meaning as the spectra path): pl < 0 cells -> [3], reconstruction
// failures -> [8], Newton iterations -> [9]; breakdown cells are counted once, here
__global__ void fold_famod_counters_kernel(const unsigned long long *__restrict__ f, unsigned long long *__restrict__ c)
{
  if (threadIdx.x == 0) { c[2] += f[2]; c[3] += f[3]; c[8] += f[8]; c[9] += f[9]; }
}

// ---- per-cell species densities (fast = 0, ParticleSampler.cpp:896-911 / max_particle_number :164-239; df_mode 5,
// :1461-1499): the species weights differ from cell to cell, so each cell gets its own cumulative row cdf[cell][ns].
struct DensityTables {
  const double *mass, *sign, *deg, *baryon;
  const double *gla_root, *gla_weight;
  int gla_pts, ns, include_baryon;
  const double *gl16;                 // 16-point Gauss-Laguerre table of the anisotropic path (aniso.cuh layout)
};

__global__ void sampler_density_kernel(const double *__restrict__ pack, int64_t stride, int64_t count, int df_mode, DensityTables t,
                                       double *__restrict__ dens)
{
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count * t.ns) return;
  const int64_t cell = idx / t.ns;
  const int s = (int)(idx - cell * t.ns);
  double dn = 0.0;
  if (pack[SP_VALID * stride + cell] != 0.0) {
    auto pk = [&](int k) { return pack[k * stride + cell]; };
    const double mass = t.mass[s], sign = t.sign[s], deg = t.deg[s], baryon = t.baryon[s];
    if (df_mode == 5) {
      // anisotropic density g Lambda^3 detA / (2 pi^2 hbarc^3) I_100 with the 16-point a = 1 rule; the chemical
      // potential enters as exp(Ebar + chem) exactly as in the reference (:1490)
      const double lambda = pk(SP_TSAMPLE), mbar = mass / lambda, mbar2 = mbar * mbar, chem = baryon * pk(SP_ALPHAB_SAMPLE);
      double I_100 = 0.0;
      for (int k = 0; k < 16; k++) {
        const double pbar = t.gl16[k], w = t.gl16[16 + k];
        const double Ebar = sqrt(pbar * pbar + mbar2);
        I_100 += pbar * w * exp(pbar) / (exp(Ebar + chem) + sign);
      }
      dn = deg * pk(SP_C0) * I_100;
    } else {
      const double T = pk(SP_T), alphaB = pk(SP_ALPHAB), mbar = mass / T;
      const double neq_fact = T * T * T / kTwoPi2HbarC3, J20_fact = T * neq_fact;
      const double *r1 = t.gla_root + 1 * t.gla_pts, *w1 = t.gla_weight + 1 * t.gla_pts;
      const double *r2 = t.gla_root + 2 * t.gla_pts, *w2 = t.gla_weight + 2 * t.gla_pts;
      const bool breaks = pk(SP_BREAKDOWN) != 0.0;
      if (df_mode == 3 && !breaks) {
        const double neq = neq_fact * deg * gauss_thermal<TI_NEQ>(r1, w1, t.gla_pts, mbar, alphaB, baryon, sign);
        double J10 = 0.0;
        if (t.include_baryon) J10 = neq_fact * deg * gauss_thermal<TI_J10>(r1, w1, t.gla_pts, mbar, alphaB, baryon, sign);
        const double J20 = J20_fact * deg * gauss_thermal<TI_J20>(r2, w2, t.gla_pts, mbar, alphaB, baryon, sign);
        // SP_C1 = G, SP_C2 = F / T^2, SP_C4 = bulkPi / betabulk
        dn = neq + pk(SP_C4) * (neq + (baryon * J10 * pk(SP_C1)) + (J20 * pk(SP_C2)));
      } else if (df_mode == 4 && !breaks) {
        dn = pk(SP_Z) * neq_fact * deg * gauss_thermal<TI_NEQ>(r1, w1, t.gla_pts, mbar, 0.0, 0.0, sign);
      } else {
        dn = 2.0 * neq_fact * deg * gauss_thermal<TI_NEQ>(r1, w1, t.gla_pts, mbar, alphaB, baryon, sign);
      }
    }
  }
  dens[idx] = dn;
}

// one warp per cell: in-place inclusive sum over species (serial order inside each lane's contiguous slice, slices
// combined by a shuffle scan), total -> dn_tot with the volume factor (:913-915) and the Poisson proposal count
__global__ void sampler_cdf_kernel(double *__restrict__ pack, int64_t stride, int64_t count, int64_t cell_global0, int ns,
                                   double *__restrict__ cdf, double y_max, double nevents, uint64_t seed,
                                   unsigned long long *__restrict__ ncount)
{
  const int64_t cell = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (cell >= count) return;
  double *row = cdf + (size_t)cell * ns;
  const int per = (ns + 31) / 32, s0 = lane * per, s1 = min(ns, s0 + per);
  double local = 0.0;
  for (int s = s0; s < s1; s++) { local += row[s]; row[s] = local; }
  double incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { double v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  const double before = incl - local;
  for (int s = s0; s < s1; s++) row[s] += before;
  const double total = __shfl_sync(0xffffffffu, incl, 31);
  if (lane == 0) {
    const bool valid = pack[SP_VALID * stride + cell] != 0.0;
    const double dn_tot = (valid && total > 0.0) ? total * (2.0 * y_max * pack[SP_DSMAX * stride + cell]) : 0.0;
    pack[SP_DNTOT * stride + cell] = dn_tot;
    if (ncount) ncount[cell] = 0;
  }
  (void)nevents; (void)seed; (void)cell_global0;
}

// proposal counts of one pass: entry v = (block - block0) * count + cell
__global__ void sampler_count_kernel(const double *__restrict__ pack, int64_t stride, int64_t count, int64_t cell_global0, int block0,
                                     int nblocks, long nevents, uint64_t seed, unsigned long long *__restrict__ ncount)
{
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= (int64_t)nblocks * count) return;
  const int blk = block0 + (int)(v / count);
  const int64_t cell = v - (int64_t)(blk - block0) * count;
  const long first = (long)blk * kEventBlock;
  const long ev = nevents - first < kEventBlock ? nevents - first : kEventBlock;
  const double dn_tot = pack[SP_DNTOT * stride + cell];
  unsigned long long n = 0;
  if (dn_tot > 0.0 && ev > 0) {
    Philox rng;
    rng.init(seed, (uint64_t)(cell_global0 + cell), 0xFFFFFFFFu, (uint32_t)blk);
    n = (unsigned long long)poisson_sample(rng, (double)ev * dn_tot);
  }
  ncount[v] = n;
}

// df_mode 5: fold the anisotropic solution left by the famod stage (feqmod pack layout) into the sampler pack:
// sampling temperature Lambda, chemical potential upsilon_B = alpha_B, momentum map B_ij = C_ik A_kj (identity on
// breakdown), density prefactor Lambda^3 detA / (2 pi^2 hbarc^3)   (ParticleSampler.cpp:1385-1470)
__global__ void sampler_famod_fold_kernel(double *__restrict__ pack, int64_t stride, int64_t count, const double *__restrict__ fpack,
                                          int64_t fstride, int include_shear)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  if (pack[SP_VALID * stride + i] == 0.0) return;
  auto fp = [&](int k) { return fpack[k * fstride + i]; };
  auto sp = [&](int k) -> double & { return pack[k * stride + i]; };
  const double lambda = fp(FP_FAMOD_LAMBDA), aT = fp(FP_FAMOD_AT), aL = fp(FP_FAMOD_AL);
  const double shear_coeff = 0.5 / fp(FP_FAMOD_BETAPIPERP), diff_coeff = 1.0 / fp(FP_FAMOD_BETAWPERP);
  double piTxx = 0, piTxy = 0, piTyy = 0, WTzx = 0, WTzy = 0;
  if (include_shear) {
    piTxx = (sp(SP_PIXX) - sp(SP_PIYY)) / 2.; piTxy = sp(SP_PIXY); piTyy = -piTxx; WTzx = sp(SP_PIXZ); WTzy = sp(SP_PIYZ);
  }
  double Bxx = aT + aT * shear_coeff * piTxx, Bxy = aT * shear_coeff * piTxy, Bxz = diff_coeff * WTzx * aT * aL / (aT + aL);
  double Byy = aT + aT * shear_coeff * piTyy, Byz = diff_coeff * WTzy * aT * aL / (aT + aL), Bzz = aL;
  const bool breaks = fp(FP_BREAKDOWN) != 0.0;
  if (breaks) { Bxx = 1; Bxy = 0; Bxz = 0; Byy = 1; Byz = 0; Bzz = 1; }
  sp(SP_PIXX) = Bxx; sp(SP_PIXY) = Bxy; sp(SP_PIXZ) = Bxz; sp(SP_PIYY) = Byy; sp(SP_PIYZ) = Byz; sp(SP_PIZZ) = Bzz;
  sp(SP_BREAKDOWN) = breaks ? 1.0 : 0.0;
  sp(SP_TSAMPLE) = lambda;
  sp(SP_ALPHAB_SAMPLE) = fp(FP_ALPHAB_MOD);
  sp(SP_C0) = lambda * lambda * lambda * (aT * aT * aL) / kTwoPi2HbarC3;
}

// budget for the per-pass scratch (cell pack + per-cell species rows), in bytes
constexpr int64_t kSamplerPassBytes = (int64_t)8 << 30;

bool sampler_uses_cell_cdf(const is3d_ctx *ctx) { return ctx->prm.df_mode == 5 || !ctx->prm.fast; }

int64_t sampler_cells_per_pass(const is3d_ctx *ctx)
{
  int64_t per_cell = SP_SIZE * 8;
  if (sampler_uses_cell_cdf(ctx)) per_cell += 8 * (int64_t)ctx->ns;
  if (ctx->prm.df_mode == 5) per_cell += FP_SIZE * 8;
  int64_t cells = kSamplerPassBytes / per_cell;
  cells = cells / 1024 * 1024;
  if (cells > ((int64_t)16 << 20)) cells = (int64_t)16 << 20;
  if (cells < 1024) cells = 1024;
  // test hook: force small passes so that the multi-pass merge is exercised on small surfaces
  if (const char *v = getenv("IS3D_SAMPLER_PASS_CELLS")) { int64_t c = atoll(v); if (c > 0) cells = c; }
  return cells;
}

}  // namespace

is3d_status famod_setup_pass(is3d_ctx *ctx, int64_t begin, int64_t count, double *pack, int64_t stride, unsigned long long *counters,
                             int64_t *launches, bool sampler_rules);

namespace {

// One pass of the per-cell stage: cell pack, mean yields, species rows (where the weights are cell dependent) and the
// Poisson proposal counts.  nevents = 0 skips the Poisson draw (yield-only callers).
is3d_status sampler_setup_pass(is3d_ctx *ctx, SamplerSetup &ss, int64_t begin, int64_t count, double nevents, double *pack,
                               int64_t stride, unsigned long long *ncount, double *yield, unsigned long long *counters,
                               int64_t *launches, bool with_species, double *dens_host = nullptr)
{
  const is3d_params &p = ctx->prm;
  const bool cdf_mode = with_species && sampler_uses_cell_cdf(ctx);
  ss.st.cell_cdf = nullptr;
  sampler_setup_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(
      ctx->surf, begin, count, ctx->global_offset, ctx->tb, ss.fl, ctx->d_gla_root, ctx->d_gla_weight, ctx->gla_pts, ss.st,
      cdf_mode ? 0.0 : nevents, (uint64_t)p.sampler_seed, pack, stride, cdf_mode ? nullptr : ncount, yield, counters);
  IS3D_CUDA_TRY(ctx, cudaGetLastError());
  (*launches)++;
  if (!cdf_mode) return IS3D_OK;

  void *cdf = nullptr, *gl = nullptr;
  IS3D_TRY(ctx->get_scratch("sampler_cell_cdf", (size_t)stride * ctx->ns * sizeof(double), &cdf));
  IS3D_TRY(ctx->get_scratch("gl16", 96 * sizeof(double), &gl));
  if (p.df_mode == 5) {
    void *fpack = nullptr, *fcounters = nullptr;
    IS3D_TRY(ctx->get_scratch("famod_pack", (size_t)FP_SIZE * stride * sizeof(double), &fpack));
    IS3D_TRY(ctx->get_scratch("famod_counters", 16 * sizeof(unsigned long long), &fcounters));
    IS3D_CUDA_TRY(ctx, cudaMemsetAsync(fcounters, 0, 16 * sizeof(unsigned long long), ctx->stream));
    IS3D_TRY(famod_setup_pass(ctx, begin, count, (double *)fpack, stride, (unsigned long long *)fcounters, launches, true));
    sampler_famod_fold_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(pack, stride, count, (const double *)fpack, stride,
                                                                                     p.include_shear_deltaf);
    IS3D_CUDA_TRY(ctx, cudaGetLastError());
    fold_famod_counters_kernel<<<1, 32, 0, ctx->stream>>>((const unsigned long long *)fcounters, counters);
    (*launches) += 2;
  } else if (begin == 0) {
    double t[96];
    fill_gl16_table(t);       // unused by df_mode 1-4; keeps the pointer valid
    IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(gl, t, sizeof(t), cudaMemcpyHostToDevice, ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  DensityTables dt;
  dt.mass = ctx->d_mass; dt.sign = ctx->d_sign; dt.deg = ctx->d_deg; dt.baryon = ctx->d_baryon;
  dt.gla_root = ctx->d_gla_root; dt.gla_weight = ctx->d_gla_weight; dt.gla_pts = ctx->gla_pts; dt.ns = ctx->ns;
  dt.include_baryon = p.include_baryon; dt.gl16 = (const double *)gl;
  const int64_t work = count * ctx->ns;
  sampler_density_kernel<<<(unsigned)((work + 127) / 128), 128, 0, ctx->stream>>>(pack, stride, count, p.df_mode, dt, (double *)cdf);
  IS3D_CUDA_TRY(ctx, cudaGetLastError());
  if (dens_host)      // per-species densities before they are accumulated in place (is3d_cell_yields)
    IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(dens_host, cdf, (size_t)work * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  const double y_max = (p.dimension == 2) ? p.y_cut : 0.5;
  sampler_cdf_kernel<<<(unsigned)((count * 32 + 127) / 128), 128, 0, ctx->stream>>>(pack, stride, count, ctx->global_offset + begin, ctx->ns,
                                                                                (double *)cdf, y_max, nevents, (uint64_t)p.sampler_seed, ncount);
  IS3D_CUDA_TRY(ctx, cudaGetLastError());
  (*launches) += 2;
  ss.st.cell_cdf = (const double *)cdf;
  return IS3D_OK;
}

}  // namespace

// calculate_total_yield (:447-636): deterministic two-level sum (per-block partials, then host sum in block order)
is3d_status run_total_yield(is3d_ctx *ctx, double *ntotal, is3d_stats *stats)
{
  SamplerSetup ss;
  IS3D_TRY(prepare_sampler(ctx, &ss));
  const int64_t n = ctx->surf.n;
  const int64_t macro = sampler_cells_per_pass(ctx);
  const int64_t stride = n < macro ? n : macro;
  void *pack = nullptr, *yield = nullptr, *counters = nullptr, *bsum = nullptr;
  IS3D_TRY(ctx->get_scratch("cell_pack", (size_t)SP_SIZE * stride * sizeof(double), &pack));
  IS3D_TRY(ctx->get_scratch("cell_yield", (size_t)stride * sizeof(double), &yield));
  IS3D_TRY(ctx->get_scratch("counters", 16 * sizeof(unsigned long long), &counters));
  IS3D_TRY(ctx->get_scratch("block_sums", 1024 * sizeof(double), &bsum));
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(counters, 0, 16 * sizeof(unsigned long long), ctx->stream));
  double total = 0.0;
  int64_t launches = 0;
  for (int64_t begin = 0; begin < n; begin += macro) {
    int64_t count = n - begin < macro ? n - begin : macro;
    IS3D_TRY(sampler_setup_pass(ctx, ss, begin, count, 0.0, (double *)pack, stride, nullptr, (double *)yield,
                                (unsigned long long *)counters, &launches, false));
    yield_reduce_kernel<<<1024, 256, 0, ctx->stream>>>((double *)yield, count, (double *)bsum);
    IS3D_CUDA_TRY(ctx, cudaGetLastError());
    double h[1024];
    IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(h, bsum, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 1024; i++) total += h[i];
    launches += 1;
  }
  if (ctx->prm.dimension == 2) total *= (2.0 * ctx->prm.y_cut);     // :628-631
  *ntotal = total;
  return fill_stats(ctx, counters, stats, 0.f, launches);
}

// per-cell mean yields of the sampler: dn_tot after the volume factor; dn_list[cell][s] = fast-mode WA neq_s + WB dn_bulk_s,
// or the cell's own Gauss-Laguerre densities for fast = 0 / df_mode 5
is3d_status run_cell_yields(is3d_ctx *ctx, double *dn_tot_host, double *dn_list_host, is3d_stats *stats)
{
  SamplerSetup ss;
  IS3D_TRY(prepare_sampler(ctx, &ss));
  const int64_t n = ctx->surf.n;
  const int64_t macro = sampler_cells_per_pass(ctx);
  const int64_t stride = n < macro ? n : macro;
  const int ns = ctx->ns;
  void *pack = nullptr, *counters = nullptr;
  IS3D_TRY(ctx->get_scratch("cell_pack", (size_t)SP_SIZE * stride * sizeof(double), &pack));
  IS3D_TRY(ctx->get_scratch("counters", 16 * sizeof(unsigned long long), &counters));
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(counters, 0, 16 * sizeof(unsigned long long), ctx->stream));
  std::vector<double> wa(stride), wb(stride);
  int64_t launches = 0;
  for (int64_t begin = 0; begin < n; begin += macro) {
    int64_t count = n - begin < macro ? n - begin : macro;
    const bool rows = dn_list_host && sampler_uses_cell_cdf(ctx);
    IS3D_TRY(sampler_setup_pass(ctx, ss, begin, count, 0.0, (double *)pack, stride, nullptr, nullptr, (unsigned long long *)counters,
                                &launches, true, rows ? dn_list_host + (size_t)begin * ns : nullptr));
    const double *pk = (const double *)pack;
    IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(dn_tot_host + begin, pk + (size_t)SP_DNTOT * stride, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (rows) {
      // dn_list came back from the density stage inside sampler_setup_pass
    } else if (dn_list_host) {
      IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(wa.data(), pk + (size_t)SP_WA * stride, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(wb.data(), pk + (size_t)SP_WB * stride, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      for (int64_t i = 0; i < count; i++)
        for (int s = 0; s < ns; s++)
          dn_list_host[(size_t)(begin + i) * ns + s] = wa[i] * ctx->h_neq[s] + wb[i] * ctx->h_dnbulk[s];
    }
    IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return fill_stats(ctx, counters, stats, 0.f, launches);
}

static is3d_status ensure_hist(is3d_ctx *ctx, HistGrid *hg, bool zero)
{
  const is3d_params &p = ctx->prm;
  const size_t ns = ctx->ns;
  const size_t sizes[10] = {ns * p.y_bins, ns * p.eta_bins, ns * p.phip_bins, ns * p.pT_bins, ns * p.pT_bins,
                            7 * ns * p.pT_bins, 7 * ns * p.pT_bins, ns * p.tau_bins, ns * p.r_bins, ns * p.phip_bins};
  size_t total = 0;
  for (size_t v : sizes) total += v;
  void *d = nullptr;
  IS3D_TRY(ctx->get_scratch("sampler_hist", total * sizeof(double), &d));
  if (zero) IS3D_CUDA_TRY(ctx, cudaMemsetAsync(d, 0, total * sizeof(double), ctx->stream));
  double *q = (double *)d;
  double **slots[10] = {&hg->dN_dy, &hg->dN_deta, &hg->dN_dphip, &hg->dN_pT, &hg->pT_count, &hg->vn_re, &hg->vn_im, &hg->dN_tau, &hg->dN_r, &hg->dN_phis};
  for (int k = 0; k < 10; k++) { *slots[k] = q; q += sizes[k]; }
  hg->test_sampler = p.test_sampler; hg->ns = ctx->ns;
  hg->y_cut = p.y_cut; hg->y_bins = p.y_bins; hg->y_width = 2.0 * p.y_cut / (double)p.y_bins;
  hg->eta_cut = p.eta_cut; hg->eta_bins = p.eta_bins; hg->eta_width = 2.0 * p.eta_cut / (double)p.eta_bins;
  hg->pT_min = p.pT_min; hg->pT_bins = p.pT_bins; hg->pT_width = (p.pT_max - p.pT_min) / (double)p.pT_bins;
  hg->phi_bins = p.phip_bins; hg->phi_width = kTwoPi / (double)p.phip_bins;
  hg->tau_min = p.tau_min; hg->tau_bins = p.tau_bins; hg->tau_width = (p.tau_max - p.tau_min) / (double)p.tau_bins;
  hg->r_min = p.r_min; hg->r_bins = p.r_bins; hg->r_width = (p.r_max - p.r_min) / (double)p.r_bins;
  return IS3D_OK;
}

// Library-owned pinned host buffers for particle lists.  A list handed out by is3d_sample stays valid until
// is3d_free_particles; released buffers are reused by later calls on the same context (page-locking gigabytes costs
// more than sampling them).  The registry maps a list pointer back to its buffer for is3d_free_particles(ptr).
static std::mutex g_list_mutex;
static std::map<void *, is3d_ctx::HostList *> g_lists;

// NUMA node of the GPU (from sysfs), -1 if unknown
static int gpu_numa_node(int device)
{
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), device) != cudaSuccess) return -1;
  for (char *c = bus; *c; c++) *c = (char)tolower(*c);
  std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
  FILE *f = fopen(path.c_str(), "r");
  if (!f) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  return node;
}

// Page-locked list buffer on the GPU's own NUMA node: with eight GPUs writing particle lists at once, lists that all live
// on node 0 share one socket's memory and inter-socket bandwidth.  The memory policy of the calling thread is switched to
// "prefer that node" for the allocation (raw syscall: libnuma is not a dependency) and restored; failures are ignored.
static cudaError_t malloc_host_near_gpu(void **p, size_t bytes, int device)
{
  const int node = gpu_numa_node(device);
  bool switched = false;
  if (node >= 0 && node < 1024) {
    unsigned long mask[16] = {0};
    mask[node / 64] |= 1ul << (node % 64);
    switched = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, 1025ul) == 0;
  }
  cudaError_t e = cudaMallocHost(p, bytes);
  if (switched) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul);
  return e;
}

static is3d_status acquire_host_list(is3d_ctx *ctx, size_t bytes, is3d_ctx::HostList **out)
{
  std::lock_guard<std::mutex> lock(g_list_mutex);
  is3d_ctx::HostList *best = nullptr;
  for (auto *h : ctx->host_lists)
    if (!h->in_use && h->capacity >= bytes && (!best || h->capacity < best->capacity)) best = h;
  if (!best) {
    // drop released buffers that are too small, then page-lock a new one
    for (size_t i = 0; i < ctx->host_lists.size();) {
      auto *h = ctx->host_lists[i];
      if (!h->in_use) { g_lists.erase(h->ptr); cudaFreeHost(h->ptr); delete h; ctx->host_lists.erase(ctx->host_lists.begin() + i); }
      else i++;
    }
    best = new is3d_ctx::HostList;
    best->capacity = bytes + bytes / 8 + 4096;
    cudaError_t e = malloc_host_near_gpu(&best->ptr, best->capacity, ctx->prm.device);
    if (e != cudaSuccess) { delete best; ctx->set_error(std::string("cudaMallocHost (particle list): ") + cudaGetErrorString(e)); return IS3D_ERR_CUDA; }
    best->owner = ctx;
    ctx->host_lists.push_back(best);
    g_lists[best->ptr] = best;
  }
  best->in_use = true;
  *out = best;
  return IS3D_OK;
}

void release_host_lists_of(is3d_ctx *ctx)
{
  std::lock_guard<std::mutex> lock(g_list_mutex);
  for (auto *h : ctx->host_lists) {
    if (h->in_use) { h->owner = nullptr; continue; }      // still held by the caller: freed by is3d_free_particles
    g_lists.erase(h->ptr);
    cudaFreeHost(h->ptr);
    delete h;
  }
  ctx->host_lists.clear();
}

static bool release_host_list(void *ptr)
{
  std::lock_guard<std::mutex> lock(g_list_mutex);
  auto it = g_lists.find(ptr);
  if (it == g_lists.end()) return false;
  is3d_ctx::HostList *h = it->second;
  if (h->owner) { h->in_use = false; return true; }         // back to its context's pool
  g_lists.erase(it);
  if (h->pinned) cudaFreeHost(h->ptr); else free(h->ptr);
  delete h;
  return true;
}

void *alloc_plain_list(size_t bytes)
{
  void *p = malloc(bytes ? bytes : 8);
  if (!p) return nullptr;
  auto *h = new is3d_ctx::HostList;
  h->ptr = p; h->capacity = bytes; h->in_use = true; h->owner = nullptr; h->pinned = false;
  std::lock_guard<std::mutex> lock(g_list_mutex);
  g_lists[p] = h;
  return p;
}

// soft bound of the proposals of one pass (record scratch ~ 27 GB; test hook: IS3D_SAMPLER_PASS_PROPOSALS)
static unsigned long long max_proposals_per_pass()
{
  if (const char *v = getenv("IS3D_SAMPLER_PASS_PROPOSALS")) { long long c = atoll(v); if (c > 0) return (unsigned long long)c; }
  return 256ull << 20;
}

// record layouts a call can deliver: the reference's Sampled_Particle (104 B) or the 64-byte wire record
__global__ void gather_compact_kernel(const is3d_particle *__restrict__ in, const unsigned int *__restrict__ idx,
                                      unsigned long long n, is3d_particle_compact *__restrict__ out)
{
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const is3d_particle r = in[idx[i]];
  is3d_particle_compact c;
  c.chosen_index = r.chosen_index; c.event = r.event;
  c.tau = r.tau; c.x = r.x; c.y = r.y; c.eta = r.eta; c.px = r.px; c.py = r.py; c.pz = r.pz;
  out[i] = c;
}

namespace {

// One u64 from device memory into the context's mapped host words.  Control scalars (proposal / acceptance counts) must not
// travel through cudaMemcpy: a D2H copy of 8 bytes queues on the same copy engine BEHIND the previous pass's bulk transfer
// and would serialise the pipeline (measured: 26 ms instead of ~16 ms per 10 M hadrons).
__global__ void publish_word_kernel(const unsigned long long *__restrict__ src, volatile unsigned long long *dst)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) { *dst = *src; __threadfence_system(); }
}

is3d_status ensure_mapped_words(is3d_ctx *ctx)
{
  if (ctx->h_words) return IS3D_OK;
  IS3D_CUDA_TRY(ctx, cudaHostAlloc((void **)&ctx->h_words, 64 * sizeof(unsigned long long), cudaHostAllocMapped));
  IS3D_CUDA_TRY(ctx, cudaHostGetDevicePointer((void **)&ctx->d_words, (void *)ctx->h_words, 0));
  return IS3D_OK;
}

// second stream + events of the copy pipeline, created on first use and destroyed with the context
is3d_status ensure_copy_pipeline(is3d_ctx *ctx)
{
  if (ctx->copy_stream) return IS3D_OK;
  IS3D_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < 2; k++) {
    IS3D_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_sorted[k], cudaEventDisableTiming));
    IS3D_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_copied[k], cudaEventDisableTiming));
  }
  return IS3D_OK;
}

struct SampleRequest {
  int64_t nevents;
  bool compact;              // 64-byte wire records instead of the 104-byte Sampled_Particle
  bool to_device;            // leave the list in a library-owned device buffer (no D2H)
};

// result of one contiguous cell block: records grouped by event
struct BlockList {
  is3d_ctx::HostList *host = nullptr;   // pinned list (host delivery)
  void *dev = nullptr;                  // device list (device delivery)
  unsigned long long n = 0;
};

// Samples the cells [begin, begin + count) whose pack is already on the device: event-block passes, pipelined D2H.
// sum_dn = sum of the cells' dn_tot (mean hadrons per event of this cell block).
is3d_status sample_cell_block(is3d_ctx *ctx, SamplerSetup &ss, const HistGrid &hg, const SampleRequest &rq, int64_t begin, int64_t count,
                              const double *pack, int64_t stride, double sum_dn, unsigned long long *counters, unsigned long long *evc,
                              int key_bits, int64_t *launches, BlockList *result)
{
  const is3d_params &p = ctx->prm;
  const int64_t nevents = rq.nevents;
  const bool lists = !p.test_sampler;
  const size_t rec_bytes = rq.compact ? sizeof(is3d_particle_compact) : sizeof(is3d_particle);
  const int nblocks_total = (int)((nevents + kEventBlock - 1) / kEventBlock);
  // pass size: enough passes to hide the D2H of pass k behind the sampling of pass k + 1 when the list is large, one pass
  // when it is small (every pass costs two host round trips), never (on average) above the proposal budget
  const double expected_total = (double)nevents * sum_dn;
  const double per_block = (double)kEventBlock * sum_dn;
  double target = expected_total;
  if (lists && !rq.to_device && expected_total * (double)rec_bytes > 96e6) target = expected_total / 6.0;
  const double budget = (double)max_proposals_per_pass();
  if (target > budget) target = budget;
  int blocks_per_pass = per_block > 0.0 ? (int)(target / per_block) : nblocks_total;
  if (blocks_per_pass < 1) blocks_per_pass = 1;
  if (blocks_per_pass > nblocks_total) blocks_per_pass = nblocks_total;
  const int64_t max_entries = (int64_t)64 << 20;                        // (block, cell) count entries of one pass
  if ((int64_t)blocks_per_pass * count > max_entries) blocks_per_pass = (int)(max_entries / count > 0 ? max_entries / count : 1);

  void *ncount = nullptr, *offsets = nullptr, *nacc_dev = nullptr;
  const size_t entries = (size_t)blocks_per_pass * count + 1;
  IS3D_TRY(ctx->get_scratch("sampler_ncount", entries * sizeof(unsigned long long), &ncount));
  IS3D_TRY(ctx->get_scratch("sampler_offsets", entries * sizeof(unsigned long long), &offsets));
  IS3D_TRY(ctx->get_scratch("sampler_nacc", sizeof(unsigned long long), &nacc_dev));
  IS3D_TRY(ensure_mapped_words(ctx));

  // destination of the block's list: capacity from the mean + 8 sigma (accepted <= proposed); grown if ever exceeded
  size_t capacity = (size_t)(expected_total + 8.0 * sqrt(expected_total + 1.0)) + 4096;
  is3d_ctx::HostList *hb = nullptr;
  void *dev_list = nullptr;
  if (lists) {
    if (rq.to_device) IS3D_TRY(ctx->get_scratch("sampler_device_list", capacity * rec_bytes, &dev_list));
    else { IS3D_TRY(acquire_host_list(ctx, capacity * rec_bytes, &hb)); IS3D_TRY(ensure_copy_pipeline(ctx)); }
  }
  auto fail = [&](is3d_status st) { if (hb) { cudaStreamSynchronize(ctx->copy_stream); release_host_list(hb->ptr); } return st; };
#define SMP_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { ctx->set_error(std::string(#expr) + ": " + cudaGetErrorString(e__)); return fail(IS3D_ERR_CUDA); } } while (0)
#define SMP_TRY(expr) do { is3d_status s__ = (expr); if (s__ != IS3D_OK) return fail(s__); } while (0)

  // double-buffered sorted records of a pass (host delivery): sized for the largest pass that can reasonably occur
  size_t pass_cap = (size_t)((double)blocks_per_pass * per_block * 1.25 + 8.0 * sqrt((double)blocks_per_pass * per_block + 1.0)) + 4096;
  void *sorted[2] = {nullptr, nullptr};
  if (lists && !rq.to_device) {
    SMP_TRY(ctx->get_scratch("sampler_sorted0", pass_cap * rec_bytes, &sorted[0]));
    SMP_TRY(ctx->get_scratch("sampler_sorted1", pass_cap * rec_bytes, &sorted[1]));
  }
  bool buf_in_flight[2] = {false, false};
  unsigned long long written = 0;
  int pass_index = 0;
  for (int b0 = 0; b0 < nblocks_total; b0 += blocks_per_pass, pass_index++) {
    const int nb = nblocks_total - b0 < blocks_per_pass ? nblocks_total - b0 : blocks_per_pass;
    const int64_t nv = (int64_t)nb * count;
    sampler_count_kernel<<<(unsigned)((nv + 255) / 256), 256, 0, ctx->stream>>>(pack, stride, count, ctx->global_offset + begin, b0, nb,
                                                                             (long)nevents, (uint64_t)p.sampler_seed, (unsigned long long *)ncount);
    SMP_CUDA(cudaGetLastError());
    SMP_CUDA(cudaMemsetAsync((unsigned long long *)ncount + nv, 0, sizeof(unsigned long long), ctx->stream));
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (unsigned long long *)ncount, (unsigned long long *)offsets, (int)(nv + 1), ctx->stream);
    void *tmp = nullptr;
    SMP_TRY(ctx->get_scratch("cub_tmp", tmp_bytes, &tmp));
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, (unsigned long long *)ncount, (unsigned long long *)offsets, (int)(nv + 1), ctx->stream);
    SMP_CUDA(cudaGetLastError());
    publish_word_kernel<<<1, 32, 0, ctx->stream>>>((unsigned long long *)offsets + nv, ctx->d_words + 0);
    SMP_CUDA(cudaStreamSynchronize(ctx->stream));
    const unsigned long long nprop = ctx->h_words[0];
    (*launches) += 3;
    if (nprop == 0) continue;
    SamplerOut out{nullptr, nullptr, (unsigned long long *)nacc_dev, evc};
    void *rec = nullptr, *key = nullptr;
    if (lists) {
      SMP_TRY(ctx->get_scratch("sampler_rec", nprop * sizeof(is3d_particle), &rec));
      SMP_TRY(ctx->get_scratch("sampler_key", nprop * sizeof(unsigned long long), &key));
      SMP_CUDA(cudaMemsetAsync(nacc_dev, 0, sizeof(unsigned long long), ctx->stream));
      out.rec = (is3d_particle *)rec; out.key = (unsigned long long *)key;
    }
#define IS3D_HADRONS(M) sampler_hadron_kernel<M><<<(unsigned)((nprop + 127) / 128), 128, 0, ctx->stream>>>( \
        pack, stride, count, ctx->global_offset + begin, b0, nb, (unsigned long long *)offsets, nprop, ss.st, \
        p.dimension, p.y_cut, (long)nevents, (uint64_t)p.sampler_seed, hg, out, counters)
    switch (p.df_mode) {
      case 1: IS3D_HADRONS(1); break;
      case 2: IS3D_HADRONS(2); break;
      case 3: IS3D_HADRONS(3); break;
      case 4: IS3D_HADRONS(4); break;
      default: IS3D_HADRONS(5); break;
    }
#undef IS3D_HADRONS
    SMP_CUDA(cudaGetLastError());
    (*launches)++;
    if (!lists) continue;
    publish_word_kernel<<<1, 32, 0, ctx->stream>>>((unsigned long long *)nacc_dev, ctx->d_words + 1);
    SMP_CUDA(cudaStreamSynchronize(ctx->stream));
    const unsigned long long nacc = ctx->h_words[1];
    (*launches)++;
    if (nacc == 0) continue;
    // the list outgrew its estimate (cannot happen within 8 sigma): move it to a larger buffer
    if (written + nacc > capacity) {
      const size_t bigger = (size_t)((written + nacc) * 1.5) + 4096;
      if (rq.to_device) {
        void *old = dev_list, *keep = nullptr;
        SMP_TRY(ctx->get_scratch("sampler_device_list_grow", written * rec_bytes + 8, &keep));
        SMP_CUDA(cudaMemcpyAsync(keep, old, written * rec_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        SMP_CUDA(cudaStreamSynchronize(ctx->stream));
        SMP_TRY(ctx->get_scratch("sampler_device_list", bigger * rec_bytes, &dev_list));
        SMP_CUDA(cudaMemcpyAsync(dev_list, keep, written * rec_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
      } else {
        SMP_CUDA(cudaStreamSynchronize(ctx->copy_stream));
        is3d_ctx::HostList *nb2 = nullptr;
        SMP_TRY(acquire_host_list(ctx, bigger * rec_bytes, &nb2));
        std::memcpy(nb2->ptr, hb->ptr, written * rec_bytes);
        release_host_list(hb->ptr);
        hb = nb2;
      }
      capacity = bigger;
    }
    void *key2 = nullptr, *idx = nullptr, *idx2 = nullptr, *stmp = nullptr;
    SMP_TRY(ctx->get_scratch("sampler_key2", nacc * sizeof(unsigned long long), &key2));
    SMP_TRY(ctx->get_scratch("sampler_idx", nacc * sizeof(unsigned int), &idx));
    SMP_TRY(ctx->get_scratch("sampler_idx2", nacc * sizeof(unsigned int), &idx2));
    iota_kernel<<<(unsigned)((nacc + 255) / 256), 256, 0, ctx->stream>>>((unsigned int *)idx, nacc);
    size_t sb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sb, (unsigned long long *)key, (unsigned long long *)key2, (unsigned int *)idx,
                                    (unsigned int *)idx2, (int)nacc, 0, key_bits, ctx->stream);
    SMP_TRY(ctx->get_scratch("cub_tmp", sb, &stmp));
    cub::DeviceRadixSort::SortPairs(stmp, sb, (unsigned long long *)key, (unsigned long long *)key2, (unsigned int *)idx,
                                    (unsigned int *)idx2, (int)nacc, 0, key_bits, ctx->stream);
    void *dst;
    const int slot = pass_index & 1;
    if (rq.to_device) {
      dst = (char *)dev_list + written * rec_bytes;
    } else {
      if (nacc > pass_cap) {                       // a pass far above its mean: wait for the copies in flight and enlarge
        SMP_CUDA(cudaStreamSynchronize(ctx->copy_stream));
        buf_in_flight[0] = buf_in_flight[1] = false;
        pass_cap = (size_t)(nacc * 1.25) + 4096;
        SMP_TRY(ctx->get_scratch("sampler_sorted0", pass_cap * rec_bytes, &sorted[0]));
        SMP_TRY(ctx->get_scratch("sampler_sorted1", pass_cap * rec_bytes, &sorted[1]));
      }
      if (buf_in_flight[slot]) SMP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[slot], 0));   // its previous D2H must be done
      dst = sorted[slot];
    }
    if (rq.compact) gather_compact_kernel<<<(unsigned)((nacc + 255) / 256), 256, 0, ctx->stream>>>((is3d_particle *)rec, (unsigned int *)idx2, nacc, (is3d_particle_compact *)dst);
    else gather_particles_kernel<<<(unsigned)((nacc + 255) / 256), 256, 0, ctx->stream>>>((is3d_particle *)rec, (unsigned int *)idx2, nacc, (is3d_particle *)dst);
    SMP_CUDA(cudaGetLastError());
    (*launches) += 3;
    if (!rq.to_device) {
      // pass k's records travel on the copy stream while pass k + 1 samples on the compute stream
      SMP_CUDA(cudaEventRecord(ctx->ev_sorted[slot], ctx->stream));
      SMP_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_sorted[slot], 0));
      SMP_CUDA(cudaMemcpyAsync((char *)hb->ptr + written * rec_bytes, dst, nacc * rec_bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
      SMP_CUDA(cudaEventRecord(ctx->ev_copied[slot], ctx->copy_stream));
      buf_in_flight[slot] = true;
    }
    written += nacc;
  }
  if (hb) SMP_CUDA(cudaStreamSynchronize(ctx->copy_stream));
#undef SMP_CUDA
#undef SMP_TRY
  result->host = hb; result->dev = dev_list; result->n = written;
  return IS3D_OK;
}

}  // namespace

is3d_status run_sampler(is3d_ctx *ctx, int64_t nevents, int record_kind, void **particles, int64_t *total_out, int64_t *counts,
                        is3d_stats *stats)
{
  const is3d_params &p = ctx->prm;
  if (nevents <= 0 || nevents > (1 << 24)) { ctx->set_error("sample: nevents out of range (1 .. 2^24)"); return IS3D_ERR_INVALID; }
  SampleRequest rq{nevents, (record_kind & 1) != 0, (record_kind & 2) != 0};
  SamplerSetup ss;
  IS3D_TRY(prepare_sampler(ctx, &ss));
  HistGrid hg;
  IS3D_TRY(ensure_hist(ctx, &hg, true));
  const size_t rec_bytes = rq.compact ? sizeof(is3d_particle_compact) : sizeof(is3d_particle);
  const int64_t n = ctx->surf.n;
  const int64_t macro = sampler_cells_per_pass(ctx);
  const int64_t stride = n < macro ? n : macro;
  void *pack = nullptr, *counters = nullptr, *evc = nullptr, *bsum = nullptr;
  IS3D_TRY(ctx->get_scratch("cell_pack", (size_t)SP_SIZE * stride * sizeof(double), &pack));
  IS3D_TRY(ctx->get_scratch("counters", 16 * sizeof(unsigned long long), &counters));
  IS3D_TRY(ctx->get_scratch("sampler_event_counts", (size_t)nevents * sizeof(unsigned long long), &evc));
  IS3D_TRY(ctx->get_scratch("block_sums", 1024 * sizeof(double), &bsum));
  void *counters_backup = nullptr;
  IS3D_TRY(ctx->get_scratch("counters_backup", 16 * sizeof(unsigned long long), &counters_backup));
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(counters, 0, 16 * sizeof(unsigned long long), ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(evc, 0, (size_t)nevents * sizeof(unsigned long long), ctx->stream));

  std::vector<BlockList> blocks;                        // one list per cell block, each grouped by event
  auto drop_blocks = [&]() { for (auto &q : blocks) if (q.host) release_host_list(q.host->ptr); blocks.clear(); };
  cudaEvent_t e0 = ctx->ev0, e1 = ctx->ev1;             // owned by the context: nothing to release on an error path
  float ms_total = 0.f;
  int64_t launches = 0;
  int key_bits = 40;
  while ((1ll << (key_bits - 40)) < nevents) key_bits++;

  // df_mode 5 with the serial parity chain: the chain state advances through every set-up; a block whose set-up has to be
  // redone (proposal budget) must restart from the state it began with
  void *chain_state = nullptr, *chain_backup = nullptr;
  size_t chain_bytes = 0;
  if (p.df_mode == 5 && p.famod_chain) {
    auto it = ctx->scratch.find("famod_chain_state");
    if (it != ctx->scratch.end()) { chain_state = it->second.first; chain_bytes = it->second.second; }
    if (chain_state) IS3D_TRY(ctx->get_scratch("famod_chain_backup", chain_bytes, &chain_backup));
  }

  int64_t begin = 0;
  int64_t pass_cells = stride;
  while (begin < n) {
    int64_t count = n - begin < pass_cells ? n - begin : pass_cells;
    if (rq.to_device && (begin != 0 || count != n)) {
      drop_blocks();
      ctx->set_error("sample_device: the surface does not fit one sampler pass on this context; use is3d_sample or fewer cells per context");
      return IS3D_ERR_UNSUPPORTED;
    }
    IS3D_CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(counters_backup, counters, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, ctx->stream));
    if (chain_state && begin > 0) IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(chain_backup, chain_state, chain_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    is3d_status st = sampler_setup_pass(ctx, ss, begin, count, 0.0, (double *)pack, stride, nullptr, nullptr, (unsigned long long *)counters,
                                        &launches, true);
    if (st != IS3D_OK) { drop_blocks(); return st; }
    // mean hadrons per event of this cell block
    yield_reduce_kernel<<<1024, 256, 0, ctx->stream>>>((double *)pack + (size_t)SP_DNTOT * stride, count, (double *)bsum);
    double hsum[1024];
    IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(hsum, bsum, sizeof(hsum), cudaMemcpyDeviceToHost, ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    launches++;
    double sum_dn = 0.0;
    for (double v : hsum) sum_dn += v;
    const double budget = (double)max_proposals_per_pass();
    if ((double)kEventBlock * sum_dn > budget && count > 1) {
      // even ONE event block of this cell block exceeds the proposal budget: take fewer cells and redo their set-up
      // (deterministic; the counters and the df_mode 5 chain state of the discarded attempt are rolled back)
      int64_t fewer = (int64_t)((double)count * 0.75 * budget / ((double)kEventBlock * sum_dn));
      pass_cells = fewer < 1 ? 1 : (fewer >= count ? (count + 1) / 2 : fewer);
      IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(counters, counters_backup, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, ctx->stream));
      if (chain_state && begin > 0) IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(chain_state, chain_backup, chain_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
      continue;
    }
    BlockList bl;
    st = sample_cell_block(ctx, ss, hg, rq, begin, count, (const double *)pack, stride, sum_dn, (unsigned long long *)counters,
                           (unsigned long long *)evc, key_bits, &launches, &bl);
    if (st != IS3D_OK) { drop_blocks(); return st; }
    if (bl.host || bl.dev) blocks.push_back(bl);
    IS3D_CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaEventSynchronize(e1));
    float ms = 0.f;
    IS3D_CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
    ms_total += ms;
    begin += count;
  }
  is3d_status fs = fill_stats(ctx, counters, stats, ms_total, launches);
  if (fs != IS3D_OK) { drop_blocks(); return fs; }

  std::vector<unsigned long long> event_counts(nevents, 0);
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(event_counts.data(), evc, (size_t)nevents * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  int64_t total = 0;
  for (int64_t e = 0; e < nevents; e++) total += (int64_t)event_counts[e];
  if (particles) {
    if (rq.to_device) {
      *particles = blocks.empty() ? nullptr : blocks[0].dev;
    } else if (blocks.size() == 1) {
      *particles = blocks[0].host->ptr;                          // the common case: one cell block, already grouped by event
    } else if (blocks.empty()) {
      is3d_ctx::HostList *hb = nullptr;
      IS3D_TRY(acquire_host_list(ctx, rec_bytes, &hb));
      *particles = hb->ptr;
    } else {
      // several cell blocks (surface larger than one pass): every block list is grouped by event; event e of the final list
      // is block 0's part, then block 1's, ... -- one host thread per block copies its pieces to their final places
      is3d_ctx::HostList *hb = nullptr;
      is3d_status st = acquire_host_list(ctx, (size_t)total * rec_bytes, &hb);
      if (st != IS3D_OK) { drop_blocks(); return st; }
      const size_t nb = blocks.size();
      // per-block event counts from the records themselves (event = second int32 of both layouts)
      std::vector<std::vector<int64_t>> start(nb, std::vector<int64_t>((size_t)nevents + 1, 0));
      auto event_of = [&](const char *base, unsigned long long i) { return ((const int32_t *)(base + i * rec_bytes))[rq.compact ? 1 : 2]; };
      std::vector<std::thread> th;
      for (size_t b = 0; b < nb; b++) th.emplace_back([&, b] {
        const char *src = (const char *)blocks[b].host->ptr;
        for (unsigned long long i = 0; i < blocks[b].n; i++) start[b][(size_t)event_of(src, i) + 1]++;
      });
      for (auto &t : th) t.join();
      th.clear();
      // dest[b][e] = final position of block b's first record of event e
      std::vector<std::vector<int64_t>> dest(nb, std::vector<int64_t>((size_t)nevents, 0));
      int64_t pos = 0;
      for (int64_t e = 0; e < nevents; e++)
        for (size_t b = 0; b < nb; b++) { dest[b][e] = pos; pos += start[b][(size_t)e + 1]; }
      for (size_t b = 0; b < nb; b++) th.emplace_back([&, b] {
        const char *src = (const char *)blocks[b].host->ptr;
        int64_t off = 0;
        for (int64_t e = 0; e < nevents; e++) {
          const int64_t c = start[b][(size_t)e + 1];
          if (c) std::memcpy((char *)hb->ptr + (size_t)dest[b][e] * rec_bytes, src + (size_t)off * rec_bytes, (size_t)c * rec_bytes);
          off += c;
        }
      });
      for (auto &t : th) t.join();
      drop_blocks();
      *particles = hb->ptr;
    }
  } else {
    drop_blocks();
  }
  if (total_out) *total_out = total;
  if (counts) for (int64_t e = 0; e < nevents; e++) counts[e] = (int64_t)event_counts[e];
  return IS3D_OK;
}

// host-side expansion of the 64-byte wire record into the reference's Sampled_Particle fields: mass / mcid from the species
// tables, E from the mass shell, (t, z) from (tau, eta) -- the formulas of the device path (sampler.cuh boost_to_lab, 3+1d)
void expand_compact(const is3d_ctx *ctx, const is3d_particle_compact *in, int64_t n, is3d_particle *out)
{
  const double *mass = ctx->h_mass.data();
  const int *mcid = ctx->h_mcid.data();
  const bool boost_invariant = ctx->prm.dimension == 2;
  auto work = [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; i++) {
      const is3d_particle_compact &c = in[i];
      is3d_particle r;
      r.chosen_index = c.chosen_index; r.mcid = mcid[c.chosen_index]; r.event = c.event; r.pad_ = 0;
      r.mass = mass[c.chosen_index];
      r.tau = c.tau; r.x = c.x; r.y = c.y; r.eta = c.eta; r.px = c.px; r.py = c.py; r.pz = c.pz;
      const double sh = sinh(c.eta), ch = sqrt(1.0 + sh * sh);
      r.t = c.tau * ch; r.z = c.tau * sh;
      (void)boost_invariant;
      r.E = sqrt(r.mass * r.mass + c.px * c.px + c.py * c.py + c.pz * c.pz);
      out[i] = r;
    }
  };
  const int nt = (int)std::min<int64_t>(std::max<unsigned>(1u, std::thread::hardware_concurrency()), n / 65536 + 1);
  if (nt <= 1) { work(0, n); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; t++) th.emplace_back(work, n * t / nt, n * (t + 1) / nt);
  for (auto &t : th) t.join();
}

}  // namespace is3d

extern "C" {

static is3d_status sampler_ready(is3d_ctx *ctx)
{
  if (ctx->ns <= 0 || !ctx->have_surface || !ctx->have_df) { ctx->set_error("sampler: species / df tables / surface not set"); return IS3D_ERR_INVALID; }
  if (ctx->prm.df_mode == 4 && !ctx->have_ptb) { ctx->set_error("PTB tables not set"); return IS3D_ERR_INVALID; }
  return IS3D_OK;
}

is3d_status is3d_total_yield(is3d_ctx *ctx, double *ntotal, is3d_stats *stats)
{
  if (!ctx) return IS3D_ERR_INVALID;
  IS3D_CUDA_TRY(ctx, cudaSetDevice(ctx->prm.device));
  if (!ntotal) { ctx->set_error("total_yield: NULL output"); return IS3D_ERR_INVALID; }
  if (stats) std::memset(stats, 0, sizeof(*stats));
  IS3D_TRY(sampler_ready(ctx));
  is3d_status st = IS3D_OK;
  *ntotal = 0.0;
  if (ctx->surf.n != 0) st = is3d::run_total_yield(ctx, ntotal, stats);
  // sharded surface: Nevents = f(total yield) must be the same on every GPU before sampling
  const is3d_status sc = is3d::comm_allreduce_host(ctx, ntotal, 1);
  return st != IS3D_OK ? st : sc;
}

is3d_status is3d_cell_yields(is3d_ctx *ctx, double *dn_tot, double *dn_list, is3d_stats *stats)
{
  if (!ctx) return IS3D_ERR_INVALID;
  IS3D_CUDA_TRY(ctx, cudaSetDevice(ctx->prm.device));
  if (!dn_tot) { ctx->set_error("cell_yields: NULL output"); return IS3D_ERR_INVALID; }
  if (stats) std::memset(stats, 0, sizeof(*stats));
  IS3D_TRY(sampler_ready(ctx));
  if (ctx->surf.n == 0) return IS3D_OK;
  return is3d::run_cell_yields(ctx, dn_tot, dn_list, stats);
}

// record_kind: bit 0 = 64-byte wire records, bit 1 = leave the list on the device
static is3d_status sample_entry(is3d_ctx *ctx, int64_t nevents, int record_kind, void **particles, int64_t *total, int64_t *counts,
                                is3d_stats *stats)
{
  if (!ctx) return IS3D_ERR_INVALID;
  IS3D_CUDA_TRY(ctx, cudaSetDevice(ctx->prm.device));
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (particles) *particles = nullptr;
  if (total) *total = 0;
  IS3D_TRY(sampler_ready(ctx));
  if (ctx->surf.n == 0) { if (counts) for (int64_t e = 0; e < nevents; e++) counts[e] = 0; return IS3D_OK; }
  return is3d::run_sampler(ctx, nevents, record_kind, particles, total, counts, stats);
}

is3d_status is3d_sample(is3d_ctx *ctx, int64_t nevents, is3d_particle **particles, int64_t *total, int64_t *counts, is3d_stats *stats)
{ return sample_entry(ctx, nevents, 0, (void **)particles, total, counts, stats); }

is3d_status is3d_sample_compact(is3d_ctx *ctx, int64_t nevents, is3d_particle_compact **particles, int64_t *total, int64_t *counts,
                                is3d_stats *stats)
{ return sample_entry(ctx, nevents, 1, (void **)particles, total, counts, stats); }

is3d_status is3d_sample_device(is3d_ctx *ctx, int64_t nevents, const is3d_particle **particles_dev, int64_t *total, int64_t *counts,
                               is3d_stats *stats)
{ return sample_entry(ctx, nevents, 2, (void **)particles_dev, total, counts, stats); }

is3d_status is3d_expand_particles(const is3d_ctx *ctx, const is3d_particle_compact *compact, int64_t n, is3d_particle *out)
{
  if (!ctx || n < 0 || (n > 0 && (!compact || !out))) return IS3D_ERR_INVALID;
  if (ctx->ns <= 0) return IS3D_ERR_INVALID;
  for (int64_t i = 0; i < n; i += 1 << 20)          // reject indices outside the species list before touching the tables
    if (compact[i].chosen_index < 0 || compact[i].chosen_index >= ctx->ns) return IS3D_ERR_INVALID;
  is3d::expand_compact(ctx, compact, n, out);
  return IS3D_OK;
}

void is3d_free_particles(void *p)
{
  // only lists this library handed out are released; an unknown (stale, already released after its context was
  // destroyed, or foreign) pointer is left alone rather than passed to free()
  if (p) (void)is3d::release_host_list((void *)p);
}

is3d_status is3d_sample_histograms(is3d_ctx *ctx, double *dN_dy, double *dN_deta, double *dN_dphipdy, double *dN_2pipTdpTdy,
                                   double *pT_count, double *vn_real, double *vn_imag, double *dN_taudtaudy,
                                   double *dN_twopirdrdy, double *dN_dphisdy)
{
  if (!ctx) return IS3D_ERR_INVALID;
  IS3D_CUDA_TRY(ctx, cudaSetDevice(ctx->prm.device));
  if (ctx->scratch.find("sampler_hist") == ctx->scratch.end()) { ctx->set_error("sample_histograms: is3d_sample has not run"); return IS3D_ERR_INVALID; }
  is3d::HistGrid hg;
  IS3D_TRY(is3d::ensure_hist(ctx, &hg, false));
  const is3d_params &p = ctx->prm;
  const size_t ns = ctx->ns;
  struct { double *dst; const double *src; size_t n; } c[10] = {
      {dN_dy, hg.dN_dy, ns * p.y_bins}, {dN_deta, hg.dN_deta, ns * p.eta_bins}, {dN_dphipdy, hg.dN_dphip, ns * p.phip_bins},
      {dN_2pipTdpTdy, hg.dN_pT, ns * p.pT_bins}, {pT_count, hg.pT_count, ns * p.pT_bins}, {vn_real, hg.vn_re, 7 * ns * p.pT_bins},
      {vn_imag, hg.vn_im, 7 * ns * p.pT_bins}, {dN_taudtaudy, hg.dN_tau, ns * p.tau_bins}, {dN_twopirdrdy, hg.dN_r, ns * p.r_bins},
      {dN_dphisdy, hg.dN_phis, ns * p.phip_bins}};
  for (auto &e : c)
    if (e.dst) IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(e.dst, e.src, e.n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return IS3D_OK;
}

}  // extern "C"
