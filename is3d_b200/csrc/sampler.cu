// K5 + K6: mean yields and the Monte-Carlo particle sampler (operation 2, df_mode 1-4) on sm_100a.
// Replaces calculate_total_yield, sample_dN_pTdpTdphidy and the BinSampledParticle counters
// (reference src/cpp/ParticleSampler.cpp:447-1134, BinSampledParticle.cpp).
//
// The reference loops cell -> event -> Poisson(dn_tot) hadrons.  Independent Poisson draws per event are
// equivalent to ONE Poisson draw of mean Nevents * dn_tot per cell followed by a uniform event label per hadron, so
// the GPU pipeline is flat in (cell, hadron):
//   1. sampler_setup_kernel   thread per cell: LRF quantities, df coefficients, breakdown test, mean yield dn_tot,
//                             N ~ Poisson(Nevents dn_tot) with the cell's own Philox stream -> 57-double pack, count
//   2. exclusive scan of the counts (cub::DeviceScan, plumbing) -> proposal offsets
//   3. sampler_hadron_kernel  thread per proposed hadron: cell by binary search in the offsets, event label, species
//                             by inverse CDF over the (cell-independent) cumulative density tables, thermal momentum
//                             by rejection, viscous/flux weights, accept -> record or self-test histograms
//   4. stable radix sort of (event, proposal index) (cub, plumbing) + gather -> particles grouped by event, in a
//      deterministic order that does not depend on the launch geometry.
#include <cub/cub.cuh>

#include <cstring>

#include "ctx.h"
#include "sampler.cuh"

namespace is3d {

namespace {

struct SamplerTables {
  int ns;
  const double *mass, *sign, *baryon;
  const int *mcid;
  const double *cumA, *cumB;      // inclusive cumulative sums over species of neq and dn_bulk
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
    if (nevents > 0.0 && p[SP_DNTOT] > 0.0) {
      Philox rng;
      rng.init(seed, (uint64_t)(global_offset + begin + i), 0xFFFFFFFFu);
      n = (unsigned long long)poisson_sample(rng, nevents * p[SP_DNTOT]);
    }
  }
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

__global__ void __launch_bounds__(128)
sampler_hadron_kernel(const double *__restrict__ pack, int64_t stride, int64_t ncells, int64_t cell_global0,
                      const unsigned long long *__restrict__ offsets, unsigned long long first, unsigned long long nprop,
                      SamplerTables st, int df_mode, int dimension, double y_cut, long nevents, uint64_t seed, HistGrid hg,
                      is3d_particle *__restrict__ out, unsigned int *__restrict__ keys, unsigned long long *counters)
{
  unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nprop) return;
  const unsigned long long jg = first + j;            // index in this pass's proposal numbering
  // cell = last index with offsets[cell] <= jg
  int64_t lo = 0, hi = ncells;
  while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (offsets[mid] <= jg) lo = mid; else hi = mid; }
  const int64_t cell = lo;
  const uint32_t n = (uint32_t)(jg - offsets[cell]);
  auto pk = [&](int k) { return pack[k * stride + cell]; };
  Philox rng;
  rng.init(seed, (uint64_t)(cell_global0 + cell), n);
  int event = (int)(rng.canonical() * (double)nevents);
  if (event >= nevents) event = (int)nevents - 1;
  // species by inverse CDF of w_s = WA neq_s + WB dn_bulk_s (discrete_distribution, :919-931)
  const double WA = pk(SP_WA), WB = pk(SP_WB);
  const double target = rng.canonical() * (WA * st.totA + WB * st.totB);
  int a = 0, b = st.ns - 1;
  while (a < b) { int m = (a + b) >> 1; if (WA * st.cumA[m] + WB * st.cumB[m] > target) b = m; else a = m + 1; }
  const int s = a;
  const double mass = st.mass[s], sign = st.sign[s], baryon = st.baryon[s];
  long samples = 0;
  LrfMomentum p;
  const bool accept = sample_hadron(rng, pk, df_mode, mass, sign, baryon, &samples, &p);
  atomicAdd(&counters[6], (unsigned long long)samples);
  unsigned int key = 0xFFFFFFFFu;
  if (accept) {
    atomicAdd(&counters[7], 1ull);
    const double y_max = (dimension == 2) ? y_cut : 0.5;
    LabParticle q = boost_to_lab(rng, pk, p, mass, dimension, y_max);
    if (hg.test_sampler) {
      bin_particle(hg, s, q, pk(SP_TAU), pk(SP_X), pk(SP_Y));
    } else {
      is3d_particle r;
      r.chosen_index = s; r.mcid = st.mcid[s]; r.event = event; r.pad_ = 0;
      r.mass = mass; r.tau = pk(SP_TAU); r.x = pk(SP_X); r.y = pk(SP_Y); r.eta = q.eta;
      r.t = q.t; r.z = q.z; r.E = q.E; r.px = q.px; r.py = q.py; r.pz = q.pz;
      out[j] = r;
      key = (unsigned int)event;
    }
  }
  if (keys) keys[j] = key;
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
  if (p.df_mode == 5) { ctx->set_error("sampler df_mode 5 (PTMA): CUDA kernel not implemented in this build"); return IS3D_ERR_UNSUPPORTED; }
  if (!p.fast) { ctx->set_error("sampler with fast = 0 (per-cell Gauss-Laguerre densities): CUDA kernel not implemented in this build"); return IS3D_ERR_UNSUPPORTED; }
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
    stats->sampler_proposals = (int64_t)h[6]; stats->sampler_accepted = (int64_t)h[7];
    stats->kernel_ms = ms; stats->kernel_launches = launches;
  }
  if (h[1] != 0) {
    ctx->set_error(std::to_string(h[1]) + " cell(s) outside the df coefficient tables (the reference aborts here)");
    return IS3D_ERR_TABLE_RANGE;
  }
  return IS3D_OK;
}

constexpr int64_t kSamplerMacro = 4 << 20;     // cells per pass (pack: 456 B / cell)

}  // namespace

// calculate_total_yield (:447-636): deterministic two-level sum (per-block partials, then host sum in block order)
is3d_status run_total_yield(is3d_ctx *ctx, double *ntotal, is3d_stats *stats)
{
  SamplerSetup ss;
  IS3D_TRY(prepare_sampler(ctx, &ss));
  const int64_t n = ctx->surf.n;
  const int64_t stride = n < kSamplerMacro ? n : kSamplerMacro;
  void *pack = nullptr, *yield = nullptr, *counters = nullptr, *bsum = nullptr;
  IS3D_TRY(ctx->get_scratch("cell_pack", (size_t)SP_SIZE * stride * sizeof(double), &pack));
  IS3D_TRY(ctx->get_scratch("cell_yield", (size_t)stride * sizeof(double), &yield));
  IS3D_TRY(ctx->get_scratch("counters", 16 * sizeof(unsigned long long), &counters));
  IS3D_TRY(ctx->get_scratch("block_sums", 1024 * sizeof(double), &bsum));
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(counters, 0, 16 * sizeof(unsigned long long), ctx->stream));
  double total = 0.0;
  int64_t launches = 0;
  for (int64_t begin = 0; begin < n; begin += kSamplerMacro) {
    int64_t count = n - begin < kSamplerMacro ? n - begin : kSamplerMacro;
    sampler_setup_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(
        ctx->surf, begin, count, ctx->global_offset, ctx->tb, ss.fl, ctx->d_gla_root, ctx->d_gla_weight, ctx->gla_pts, ss.st, 0.0,
        (uint64_t)ctx->prm.sampler_seed, (double *)pack, stride, nullptr, (double *)yield, (unsigned long long *)counters);
    IS3D_CUDA_TRY(ctx, cudaGetLastError());
    yield_reduce_kernel<<<1024, 256, 0, ctx->stream>>>((double *)yield, count, (double *)bsum);
    IS3D_CUDA_TRY(ctx, cudaGetLastError());
    double h[1024];
    IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(h, bsum, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 1024; i++) total += h[i];
    launches += 2;
  }
  if (ctx->prm.dimension == 2) total *= (2.0 * ctx->prm.y_cut);     // :628-631
  *ntotal = total;
  return fill_stats(ctx, counters, stats, 0.f, launches);
}

// per-cell mean yields of the sampler (dn_tot after the volume factor; dn_list[cell][s] = WA neq_s + WB dn_bulk_s)
is3d_status run_cell_yields(is3d_ctx *ctx, double *dn_tot_host, double *dn_list_host, is3d_stats *stats)
{
  SamplerSetup ss;
  IS3D_TRY(prepare_sampler(ctx, &ss));
  const int64_t n = ctx->surf.n;
  const int64_t stride = n < kSamplerMacro ? n : kSamplerMacro;
  void *pack = nullptr, *counters = nullptr;
  IS3D_TRY(ctx->get_scratch("cell_pack", (size_t)SP_SIZE * stride * sizeof(double), &pack));
  IS3D_TRY(ctx->get_scratch("counters", 16 * sizeof(unsigned long long), &counters));
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(counters, 0, 16 * sizeof(unsigned long long), ctx->stream));
  std::vector<double> wa(stride), wb(stride);
  int64_t launches = 0;
  for (int64_t begin = 0; begin < n; begin += kSamplerMacro) {
    int64_t count = n - begin < kSamplerMacro ? n - begin : kSamplerMacro;
    sampler_setup_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(
        ctx->surf, begin, count, ctx->global_offset, ctx->tb, ss.fl, ctx->d_gla_root, ctx->d_gla_weight, ctx->gla_pts, ss.st, 0.0,
        (uint64_t)ctx->prm.sampler_seed, (double *)pack, stride, nullptr, nullptr, (unsigned long long *)counters);
    IS3D_CUDA_TRY(ctx, cudaGetLastError());
    launches++;
    const double *pk = (const double *)pack;
    IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(dn_tot_host + begin, pk + (size_t)SP_DNTOT * stride, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (dn_list_host) {
      IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(wa.data(), pk + (size_t)SP_WA * stride, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(wb.data(), pk + (size_t)SP_WB * stride, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (dn_list_host)
      for (int64_t i = 0; i < count; i++)
        for (int s = 0; s < ctx->ns; s++)
          dn_list_host[(size_t)(begin + i) * ctx->ns + s] = wa[i] * ctx->h_neq[s] + wb[i] * ctx->h_dnbulk[s];
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

is3d_status run_sampler(is3d_ctx *ctx, int64_t nevents, is3d_particle **particles, int64_t *total_out, int64_t *counts,
                        is3d_stats *stats)
{
  const is3d_params &p = ctx->prm;
  if (nevents <= 0 || nevents > 0x7FFFFFFF) { ctx->set_error("sample: nevents out of range"); return IS3D_ERR_INVALID; }
  SamplerSetup ss;
  IS3D_TRY(prepare_sampler(ctx, &ss));
  HistGrid hg;
  IS3D_TRY(ensure_hist(ctx, &hg, true));
  const int64_t n = ctx->surf.n;
  const int64_t stride = n < kSamplerMacro ? n : kSamplerMacro;
  void *pack = nullptr, *counters = nullptr, *ncount = nullptr, *offsets = nullptr;
  IS3D_TRY(ctx->get_scratch("cell_pack", (size_t)SP_SIZE * stride * sizeof(double), &pack));
  IS3D_TRY(ctx->get_scratch("counters", 16 * sizeof(unsigned long long), &counters));
  IS3D_TRY(ctx->get_scratch("sampler_ncount", (size_t)(stride + 1) * sizeof(unsigned long long), &ncount));
  IS3D_TRY(ctx->get_scratch("sampler_offsets", (size_t)(stride + 1) * sizeof(unsigned long long), &offsets));
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(counters, 0, 16 * sizeof(unsigned long long), ctx->stream));

  std::vector<std::vector<is3d_particle>> passes;      // each pass: particles sorted by event
  std::vector<int64_t> event_counts(nevents, 0);
  const unsigned long long kMaxProposalsPerLaunch = 16ull << 20;
  cudaEvent_t e0, e1;
  IS3D_CUDA_TRY(ctx, cudaEventCreate(&e0));
  IS3D_CUDA_TRY(ctx, cudaEventCreate(&e1));
  float ms_total = 0.f;
  int64_t launches = 0;

  for (int64_t begin = 0; begin < n; begin += kSamplerMacro) {
    int64_t count = n - begin < kSamplerMacro ? n - begin : kSamplerMacro;
    IS3D_CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
    sampler_setup_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(
        ctx->surf, begin, count, ctx->global_offset, ctx->tb, ss.fl, ctx->d_gla_root, ctx->d_gla_weight, ctx->gla_pts, ss.st,
        (double)nevents, (uint64_t)p.sampler_seed, (double *)pack, stride, (unsigned long long *)ncount, nullptr,
        (unsigned long long *)counters);
    IS3D_CUDA_TRY(ctx, cudaGetLastError());
    IS3D_CUDA_TRY(ctx, cudaMemsetAsync((unsigned long long *)ncount + count, 0, sizeof(unsigned long long), ctx->stream));
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (unsigned long long *)ncount, (unsigned long long *)offsets, (int)(count + 1), ctx->stream);
    void *tmp = nullptr;
    IS3D_TRY(ctx->get_scratch("cub_tmp", tmp_bytes, &tmp));
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, (unsigned long long *)ncount, (unsigned long long *)offsets, (int)(count + 1), ctx->stream);
    IS3D_CUDA_TRY(ctx, cudaGetLastError());
    unsigned long long nprop_total = 0;
    IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(&nprop_total, (unsigned long long *)offsets + count, sizeof(nprop_total), cudaMemcpyDeviceToHost, ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    launches += 3;

    for (unsigned long long first = 0; first < nprop_total; first += kMaxProposalsPerLaunch) {
      unsigned long long np = nprop_total - first < kMaxProposalsPerLaunch ? nprop_total - first : kMaxProposalsPerLaunch;
      void *rec = nullptr, *keys = nullptr, *keys2 = nullptr, *idx = nullptr, *idx2 = nullptr, *sorted = nullptr;
      const bool lists = !p.test_sampler;
      if (lists) {
        IS3D_TRY(ctx->get_scratch("sampler_rec", np * sizeof(is3d_particle), &rec));
        IS3D_TRY(ctx->get_scratch("sampler_keys", np * sizeof(unsigned int), &keys));
        IS3D_TRY(ctx->get_scratch("sampler_keys2", np * sizeof(unsigned int), &keys2));
        IS3D_TRY(ctx->get_scratch("sampler_idx", np * sizeof(unsigned int), &idx));
        IS3D_TRY(ctx->get_scratch("sampler_idx2", np * sizeof(unsigned int), &idx2));
      }
      sampler_hadron_kernel<<<(unsigned)((np + 127) / 128), 128, 0, ctx->stream>>>(
          (double *)pack, stride, count, ctx->global_offset + begin, (unsigned long long *)offsets, first, np, ss.st, p.df_mode,
          p.dimension, p.y_cut, (long)nevents, (uint64_t)p.sampler_seed, hg, (is3d_particle *)rec, (unsigned int *)keys,
          (unsigned long long *)counters);
      IS3D_CUDA_TRY(ctx, cudaGetLastError());
      launches++;
      if (lists) {
        iota_kernel<<<(unsigned)((np + 255) / 256), 256, 0, ctx->stream>>>((unsigned int *)idx, np);
        size_t sb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, sb, (unsigned int *)keys, (unsigned int *)keys2, (unsigned int *)idx, (unsigned int *)idx2, (int)np, 0, 32, ctx->stream);
        void *stmp = nullptr;
        IS3D_TRY(ctx->get_scratch("cub_tmp", sb, &stmp));
        cub::DeviceRadixSort::SortPairs(stmp, sb, (unsigned int *)keys, (unsigned int *)keys2, (unsigned int *)idx, (unsigned int *)idx2, (int)np, 0, 32, ctx->stream);
        IS3D_CUDA_TRY(ctx, cudaGetLastError());
        // accepted hadrons sort to the front (rejected carry key 0xFFFFFFFF): count them on the host from the keys
        std::vector<unsigned int> hk(np);
        IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(hk.data(), keys2, np * sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        unsigned long long nacc = 0;
        while (nacc < np && hk[nacc] != 0xFFFFFFFFu) { event_counts[hk[nacc]]++; nacc++; }
        launches += 3;
        if (nacc) {
          IS3D_TRY(ctx->get_scratch("sampler_sorted", nacc * sizeof(is3d_particle), &sorted));
          gather_particles_kernel<<<(unsigned)((nacc + 255) / 256), 256, 0, ctx->stream>>>((is3d_particle *)rec, (unsigned int *)idx2, nacc, (is3d_particle *)sorted);
          IS3D_CUDA_TRY(ctx, cudaGetLastError());
          passes.emplace_back(nacc);
          IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(passes.back().data(), sorted, nacc * sizeof(is3d_particle), cudaMemcpyDeviceToHost, ctx->stream));
          IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
          launches++;
        }
      }
    }
    IS3D_CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaEventSynchronize(e1));
    float ms = 0.f;
    IS3D_CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
    ms_total += ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  IS3D_TRY(fill_stats(ctx, counters, stats, ms_total, launches));

  // merge the per-pass event-sorted lists into one array grouped by event
  int64_t total = 0;
  for (int64_t e = 0; e < nevents; e++) total += event_counts[e];
  is3d_particle *outp = nullptr;
  if (particles) {
    outp = (is3d_particle *)malloc((size_t)(total > 0 ? total : 1) * sizeof(is3d_particle));
    if (!outp) { ctx->set_error("sample: out of host memory"); return IS3D_ERR_INVALID; }
    std::vector<int64_t> cursor(nevents, 0);
    int64_t acc = 0;
    for (int64_t e = 0; e < nevents; e++) { cursor[e] = acc; acc += event_counts[e]; }
    for (auto &v : passes)
      for (const is3d_particle &q : v) outp[cursor[q.event]++] = q;
    *particles = outp;
  }
  if (total_out) *total_out = total;
  if (counts) for (int64_t e = 0; e < nevents; e++) counts[e] = event_counts[e];
  return IS3D_OK;
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
  if (ctx->surf.n == 0) { *ntotal = 0.0; return IS3D_OK; }
  return is3d::run_total_yield(ctx, ntotal, stats);
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

is3d_status is3d_sample(is3d_ctx *ctx, int64_t nevents, is3d_particle **particles, int64_t *total, int64_t *counts, is3d_stats *stats)
{
  if (!ctx) return IS3D_ERR_INVALID;
  IS3D_CUDA_TRY(ctx, cudaSetDevice(ctx->prm.device));
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (particles) *particles = nullptr;
  if (total) *total = 0;
  IS3D_TRY(sampler_ready(ctx));
  if (ctx->surf.n == 0) { if (counts) for (int64_t e = 0; e < nevents; e++) counts[e] = 0; return IS3D_OK; }
  return is3d::run_sampler(ctx, nevents, particles, total, counts, stats);
}

void is3d_free_particles(is3d_particle *p) { free(p); }

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
