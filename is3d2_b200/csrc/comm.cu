// Multi-GPU side of the C ABI: surface cells are sharded, every GPU integrates its own block, and the
// partial spectra / dN/dX histograms / total yield are combined by ONE ncclAllReduce(sum, double) over NVLink.
// The reference has no counterpart (it is a single-process OpenMP code, iS3D.cpp:81-286).
//
// Two ways in:
//   * one process per GPU (torchrun, MPI): is3d_comm_unique_id on rank 0, ship the 128 bytes to every rank, then
//     is3d_comm_attach(ctx, id, nranks, rank).  From then on the single-context compute entries all-reduce their result.
//   * one process, many GPUs (the drop-in executable, JETSCAPE): is3d_group_create(params, ndev, devices) builds one context
//     per device and a communicator over them (ncclCommInitAll); the is3d_group_* calls split the cells into contiguous
//     blocks, run one host thread per device and return the combined result.
// NCCL is resolved at run time (dlopen "libnccl.so.2"): single-GPU users need no NCCL, and a host program that already
// loaded its own NCCL (PyTorch) keeps exactly one copy in the process.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <thread>

#include "ctx.h"

namespace is3d {

namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi *nccl_api()
{
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) { api.error = std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror(); return; }
    auto sym = [&](const char *name) {
      void *p = dlsym(api.handle, name);
      if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + name;
      return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  });
  return &api;
}

thread_local std::string g_comm_error;

}  // namespace

// In-place SUM all-reduce of n doubles at `dev` over the context's communicator, on the context's stream; no-op without one.
is3d_status comm_allreduce(is3d_ctx *ctx, double *dev, int64_t n)
{
  if (!ctx->comm || ctx->comm_size <= 1 || n <= 0) return IS3D_OK;
  NcclApi *api = nccl_api();
  ncclResult_t r = api->AllReduce(dev, dev, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)ctx->comm, ctx->stream);
  if (r != ncclSuccess) { ctx->set_error(std::string("ncclAllReduce: ") + api->GetErrorString(r)); return IS3D_ERR_CUDA; }
  ctx->comm_collectives++;
  return IS3D_OK;
}

// host scalars (total yield): staged through a small device buffer
is3d_status comm_allreduce_host(is3d_ctx *ctx, double *host, int n)
{
  if (!ctx->comm || ctx->comm_size <= 1 || n <= 0) return IS3D_OK;
  void *d = nullptr;
  IS3D_TRY(ctx->get_scratch("comm_scalars", 64 * sizeof(double), &d));
  if (n > 64) { ctx->set_error("comm_allreduce_host: too many scalars"); return IS3D_ERR_INVALID; }
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(d, host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  IS3D_TRY(comm_allreduce(ctx, (double *)d, n));
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(host, d, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return IS3D_OK;
}

void comm_release(is3d_ctx *ctx)
{
  if (ctx->comm) {
    NcclApi *api = nccl_api();
    if (api->CommDestroy) api->CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
  }
  ctx->comm_size = 1; ctx->comm_rank = 0;
}

}  // namespace is3d

using namespace is3d;

// ---------------------------------------------------------------------------------------------------------------------
struct is3d_group {
  std::vector<is3d_ctx *> ctx;
  std::vector<int64_t> begin, count;        // cell block of every device (set by is3d_group_set_surface)
  std::string err;
};

namespace {

// runs fn(i) on one host thread per device and returns the first non-OK status (message copied into the group)
template <class F>
is3d_status for_each_device(is3d_group *g, F fn)
{
  const int n = (int)g->ctx.size();
  std::vector<is3d_status> st(n, IS3D_OK);
  if (n == 1) st[0] = fn(0);
  else {
    std::vector<std::thread> th;
    for (int i = 0; i < n; i++) th.emplace_back([&, i] { st[i] = fn(i); });
    for (auto &t : th) t.join();
  }
  for (int i = 0; i < n; i++)
    if (st[i] != IS3D_OK) { g->err = "device " + std::to_string(g->ctx[i]->prm.device) + ": " + g->ctx[i]->err; return st[i]; }
  return IS3D_OK;
}

void add_stats(is3d_stats *total, const is3d_stats &s)
{
  total->cells_total += s.cells_total; total->cells_skipped += s.cells_skipped; total->cells_breakdown += s.cells_breakdown;
  total->cells_pl_negative += s.cells_pl_negative; total->reconstruction_failures += s.reconstruction_failures;
  total->newton_iterations += s.newton_iterations; total->cells_out_of_table += s.cells_out_of_table;
  total->sampler_proposals += s.sampler_proposals; total->sampler_accepted += s.sampler_accepted;
  if (s.tau_breakdown > total->tau_breakdown) total->tau_breakdown = s.tau_breakdown;
  if (s.tau_pl_negative > total->tau_pl_negative) total->tau_pl_negative = s.tau_pl_negative;
  if (s.kernel_ms > total->kernel_ms) total->kernel_ms = s.kernel_ms;          // devices run side by side: the slowest one
  total->kernel_launches += s.kernel_launches;
  total->evals_executed += s.evals_executed;
  total->pair_evals_executed += s.pair_evals_executed;
  total->evals_dropped += s.evals_dropped;
  total->prune_reruns += s.prune_reruns;
}

}  // namespace

extern "C" {

is3d_status is3d_comm_unique_id(char id[IS3D_COMM_ID_BYTES])
{
  static_assert(IS3D_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
  NcclApi *api = nccl_api();
  if (!api->error.empty()) { g_comm_error = api->error; return IS3D_ERR_UNSUPPORTED; }
  ncclUniqueId u;
  ncclResult_t r = api->GetUniqueId(&u);
  if (r != ncclSuccess) { g_comm_error = std::string("ncclGetUniqueId: ") + api->GetErrorString(r); return IS3D_ERR_CUDA; }
  std::memcpy(id, u.internal, IS3D_COMM_ID_BYTES);
  return IS3D_OK;
}

const char *is3d_comm_last_error(void) { return g_comm_error.c_str(); }

is3d_status is3d_comm_attach(is3d_ctx *ctx, const char id[IS3D_COMM_ID_BYTES], int nranks, int rank)
{
  if (!ctx) return IS3D_ERR_INVALID;
  if (nranks < 1 || rank < 0 || rank >= nranks || !id) { ctx->set_error("comm_attach: bad arguments"); return IS3D_ERR_INVALID; }
  IS3D_CUDA_TRY(ctx, cudaSetDevice(ctx->prm.device));
  comm_release(ctx);
  if (nranks == 1) return IS3D_OK;
  NcclApi *api = nccl_api();
  if (!api->error.empty()) { ctx->set_error(api->error); return IS3D_ERR_UNSUPPORTED; }
  ncclUniqueId u;
  std::memcpy(u.internal, id, IS3D_COMM_ID_BYTES);
  ncclComm_t c = nullptr;
  ncclResult_t r = api->CommInitRank(&c, nranks, u, rank);
  if (r != ncclSuccess) { ctx->set_error(std::string("ncclCommInitRank: ") + api->GetErrorString(r)); return IS3D_ERR_CUDA; }
  ctx->comm = c; ctx->comm_size = nranks; ctx->comm_rank = rank;
  return IS3D_OK;
}

void is3d_comm_detach(is3d_ctx *ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->prm.device);
  cudaStreamSynchronize(ctx->stream);
  comm_release(ctx);
}

int is3d_comm_size(const is3d_ctx *ctx) { return ctx ? ctx->comm_size : 0; }
int64_t is3d_comm_collectives(const is3d_ctx *ctx) { return ctx ? ctx->comm_collectives : 0; }

// ---- one process, many GPUs ---------------------------------------------------------------------------------------
const char *is3d_group_last_error(const is3d_group *g) { return g ? g->err.c_str() : g_comm_error.c_str(); }

is3d_status is3d_group_create(const is3d_params *p, int ndev, const int *devices, is3d_group **out)
{
  *out = nullptr;
  if (!p || ndev < 1 || ndev > 64) { g_comm_error = "group_create: bad arguments"; return IS3D_ERR_INVALID; }
  std::vector<int> dev(ndev);
  for (int i = 0; i < ndev; i++) dev[i] = devices ? devices[i] : i;
  for (int i = 0; i < ndev; i++)
    for (int j = 0; j < i; j++)
      if (dev[i] == dev[j]) { g_comm_error = "group_create: device listed twice"; return IS3D_ERR_INVALID; }
  is3d_group *g = new is3d_group;
  {
    // one host thread per device: creating a CUDA primary context costs a few hundred ms, eight in a row several seconds
    std::vector<is3d_ctx *> made(ndev, nullptr);
    std::vector<is3d_status> st(ndev, IS3D_OK);
    std::vector<std::string> msg(ndev);
    auto make = [&](int i) {
      is3d_params q = *p;
      q.device = dev[i];
      st[i] = is3d_create(&q, &made[i]);
      if (st[i] != IS3D_OK) msg[i] = is3d_last_error(nullptr);          // thread-local message of the failed create
    };
    if (ndev == 1) make(0);
    else {
      std::vector<std::thread> th;
      for (int i = 0; i < ndev; i++) th.emplace_back(make, i);
      for (auto &t : th) t.join();
    }
    for (int i = 0; i < ndev; i++) if (made[i]) g->ctx.push_back(made[i]);
    for (int i = 0; i < ndev; i++)
      if (st[i] != IS3D_OK) { g_comm_error = "device " + std::to_string(dev[i]) + ": " + msg[i]; is3d_group_destroy(g); return st[i]; }
  }
  g->begin.assign(ndev, 0); g->count.assign(ndev, 0);
  if (ndev > 1) {
    NcclApi *api = nccl_api();
    if (!api->error.empty()) { g_comm_error = api->error; is3d_group_destroy(g); return IS3D_ERR_UNSUPPORTED; }
    std::vector<ncclComm_t> comms(ndev, nullptr);
    ncclResult_t r = api->CommInitAll(comms.data(), ndev, dev.data());
    if (r != ncclSuccess) { g_comm_error = std::string("ncclCommInitAll: ") + api->GetErrorString(r); is3d_group_destroy(g); return IS3D_ERR_CUDA; }
    for (int i = 0; i < ndev; i++) { g->ctx[i]->comm = comms[i]; g->ctx[i]->comm_size = ndev; g->ctx[i]->comm_rank = i; }
  }
  *out = g;
  return IS3D_OK;
}

void is3d_group_destroy(is3d_group *g)
{
  if (!g) return;
  for (is3d_ctx *c : g->ctx) { cudaSetDevice(c->prm.device); cudaStreamSynchronize(c->stream); }
  for (is3d_ctx *c : g->ctx) is3d_destroy(c);          // releases the communicators
  delete g;
}

int is3d_group_size(const is3d_group *g) { return g ? (int)g->ctx.size() : 0; }
is3d_ctx *is3d_group_ctx(is3d_group *g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr; }

void is3d_group_cell_block(const is3d_group *g, int i, int64_t *begin, int64_t *count)
{
  if (begin) *begin = g->begin[i];
  if (count) *count = g->count[i];
}

// contiguous blocks whose sizes differ by at most one cell (the same rule as is3d2_b200/shard.py cell_range)
is3d_status is3d_group_set_surface(is3d_group *g, int64_t n, const double *const cols[IS3D_SURFACE_COLUMNS], int64_t global_offset)
{
  if (!g || n < 0 || !cols) { if (g) g->err = "group_set_surface: bad arguments"; return IS3D_ERR_INVALID; }
  const int nd = (int)g->ctx.size();
  const int64_t base = n / nd, extra = n % nd;
  for (int i = 0; i < nd; i++) {
    g->begin[i] = i * base + (i < extra ? i : extra);
    g->count[i] = base + (i < extra ? 1 : 0);
  }
  return for_each_device(g, [&](int i) {
    const double *sub[IS3D_SURFACE_COLUMNS];
    for (int k = 0; k < IS3D_SURFACE_COLUMNS; k++) sub[k] = cols[k] ? cols[k] + g->begin[i] : nullptr;
    return is3d_set_surface(g->ctx[i], g->count[i], sub, global_offset + g->begin[i]);
  });
}

is3d_status is3d_group_set_vorticity(is3d_group *g, int64_t n, const double *const w[6])
{
  if (!g || !w) return IS3D_ERR_INVALID;
  int64_t have = 0;
  for (int64_t c : g->count) have += c;
  if (have != n) { g->err = "group_set_vorticity: set the surface first (same number of cells)"; return IS3D_ERR_INVALID; }
  return for_each_device(g, [&](int i) {
    const double *sub[6];
    for (int k = 0; k < 6; k++) sub[k] = w[k] ? w[k] + g->begin[i] : nullptr;
    return is3d_set_vorticity(g->ctx[i], g->count[i], sub);
  });
}

// calculate_dN_pTdpTdphidy over all devices: every device integrates its block and joins the all-reduce inside
// is3d_spectra_device; device 0 copies the combined array to the host.  stats = counters summed over the devices.
is3d_status is3d_group_spectra(is3d_group *g, double *out, is3d_stats *stats)
{
  if (!g || !out) { if (g) g->err = "group_spectra: NULL output"; return IS3D_ERR_INVALID; }
  const int nd = (int)g->ctx.size();
  std::vector<is3d_stats> st(nd);
  is3d_status rc = for_each_device(g, [&](int i) {
    is3d_ctx *c = g->ctx[i];
    if (i == 0) return is3d_spectra(c, out, &st[i]);
    IS3D_CUDA_TRY(c, cudaSetDevice(c->prm.device));
    void *d = nullptr;
    const int64_t total = is3d_spectra_size(c);
    IS3D_TRY(c->get_scratch("spectra_out", (size_t)(total > 0 ? total : 1) * sizeof(double), &d));
    return is3d_spectra_device(c, (double *)d, &st[i]);
  });
  if (stats) { std::memset(stats, 0, sizeof(*stats)); for (auto &s : st) add_stats(stats, s); }
  return rc;
}

is3d_status is3d_group_dndx(is3d_group *g, double *tau_hist, double *r_hist, double *phi_hist, is3d_stats *stats)
{
  if (!g || !tau_hist || !r_hist || !phi_hist) { if (g) g->err = "group_dndx: NULL output"; return IS3D_ERR_INVALID; }
  const int nd = (int)g->ctx.size();
  std::vector<is3d_stats> st(nd);
  is3d_status rc = for_each_device(g, [&](int i) {
    is3d_ctx *c = g->ctx[i];
    if (i == 0) return is3d_dndx(c, tau_hist, r_hist, phi_hist, &st[i]);
    IS3D_CUDA_TRY(c, cudaSetDevice(c->prm.device));
    const size_t nt = (size_t)c->ns * c->prm.tau_bins, nr = (size_t)c->ns * c->prm.r_bins, np = (size_t)c->ns * c->prm.phip_bins;
    void *d = nullptr;
    IS3D_TRY(c->get_scratch("dndx_hist", (nt + nr + np + 3) * sizeof(double), &d));
    double *dt = (double *)d;
    return is3d_dndx_device(c, dt, dt + nt, dt + nt + nr, &st[i]);
  });
  if (stats) { std::memset(stats, 0, sizeof(*stats)); for (auto &s : st) add_stats(stats, s); }
  return rc;
}

is3d_status is3d_group_total_yield(is3d_group *g, double *ntotal, is3d_stats *stats)
{
  if (!g || !ntotal) { if (g) g->err = "group_total_yield: NULL output"; return IS3D_ERR_INVALID; }
  const int nd = (int)g->ctx.size();
  std::vector<is3d_stats> st(nd);
  std::vector<double> v(nd, 0.0);
  is3d_status rc = for_each_device(g, [&](int i) { return is3d_total_yield(g->ctx[i], &v[i], &st[i]); });
  *ntotal = v[0];                                       // all-reduced inside is3d_total_yield: every device holds the sum
  if (stats) { std::memset(stats, 0, sizeof(*stats)); for (auto &s : st) add_stats(stats, s); }
  return rc;
}

is3d_status is3d_group_polarization(is3d_group *g, double *St, double *Sx, double *Sy, double *Sn, double *Snorm, is3d_stats *stats)
{
  if (!g || !St || !Sx || !Sy || !Sn || !Snorm) { if (g) g->err = "group_polarization: NULL output"; return IS3D_ERR_INVALID; }
  const int nd = (int)g->ctx.size();
  std::vector<is3d_stats> st(nd);
  const size_t total = (size_t)is3d_spectra_size(g->ctx[0]);
  std::vector<std::vector<double>> tmp(nd);
  is3d_status rc = for_each_device(g, [&](int i) {
    if (i == 0) return is3d_polarization(g->ctx[0], St, Sx, Sy, Sn, Snorm, &st[0]);
    tmp[i].assign(5 * total, 0.0);                      // combined arrays arrive on every device; only device 0's are kept
    double *q = tmp[i].data();
    return is3d_polarization(g->ctx[i], q, q + total, q + 2 * total, q + 3 * total, q + 4 * total, &st[i]);
  });
  if (stats) { std::memset(stats, 0, sizeof(*stats)); for (auto &s : st) add_stats(stats, s); }
  return rc;
}

// sample_dN_pTdpTdphidy over all devices.  No collective: the Philox streams are keyed by the GLOBAL cell index, so the
// devices sample disjoint cell blocks independently; event e of the merged list holds device 0's hadrons of e, then device
// 1's, ... -- the (cell, draw) order a single GPU produces for the whole surface (ParticleSampler.cpp:1093-1120 appends per cell).
static is3d_status group_sample_any(is3d_group *g, int64_t nevents, bool compact, void **particles, int64_t *total, int64_t *counts,
                                    is3d_stats *stats)
{
  if (!g) return IS3D_ERR_INVALID;
  const int nd = (int)g->ctx.size();
  const size_t rec = compact ? sizeof(is3d_particle_compact) : sizeof(is3d_particle);
  auto sample_one = [&](int i, void **list, int64_t *tot, int64_t *cnt, is3d_stats *st) {
    return compact ? is3d_sample_compact(g->ctx[i], nevents, (is3d_particle_compact **)list, tot, cnt, st)
                   : is3d_sample(g->ctx[i], nevents, (is3d_particle **)list, tot, cnt, st);
  };
  if (nd == 1) {
    is3d_status rc = sample_one(0, particles, total, counts, stats);
    if (rc != IS3D_OK) g->err = g->ctx[0]->err;
    return rc;
  }
  if (nevents <= 0) { g->err = "group_sample: nevents out of range"; return IS3D_ERR_INVALID; }
  std::vector<is3d_stats> st(nd);
  std::vector<void *> lists(nd, nullptr);
  std::vector<int64_t> tot(nd, 0);
  std::vector<std::vector<int64_t>> cnt(nd, std::vector<int64_t>((size_t)nevents, 0));
  is3d_status rc = for_each_device(g, [&](int i) { return sample_one(i, &lists[i], &tot[i], cnt[i].data(), &st[i]); });
  if (stats) { std::memset(stats, 0, sizeof(*stats)); for (auto &s : st) add_stats(stats, s); }
  auto drop = [&] { for (auto *l : lists) if (l) is3d_free_particles(l); };
  if (rc != IS3D_OK) { drop(); return rc; }
  int64_t all = 0;
  for (int i = 0; i < nd; i++) all += tot[i];
  // start[i][e] = position of device i's hadrons of event e in the merged list
  std::vector<std::vector<int64_t>> start(nd, std::vector<int64_t>((size_t)nevents, 0));
  int64_t pos = 0;
  for (int64_t e = 0; e < nevents; e++) {
    int64_t ce = 0;
    for (int i = 0; i < nd; i++) { start[i][e] = pos; pos += cnt[i][e]; ce += cnt[i][e]; }
    if (counts) counts[e] = ce;
  }
  if (total) *total = all;
  if (particles) {
    char *out = (char *)alloc_plain_list((size_t)all * rec);
    if (!out) { drop(); g->err = "group_sample: out of host memory"; return IS3D_ERR_INVALID; }
    for_each_device(g, [&](int i) {                       // one host thread per device list: disjoint destination ranges
      const char *src = (const char *)lists[i];
      int64_t off = 0;
      for (int64_t e = 0; e < nevents; e++) {
        const int64_t c = cnt[i][e];
        if (c) std::memcpy(out + (size_t)start[i][e] * rec, src + (size_t)off * rec, (size_t)c * rec);
        off += c;
      }
      return IS3D_OK;
    });
    *particles = out;
  }
  drop();
  return IS3D_OK;
}

is3d_status is3d_group_sample(is3d_group *g, int64_t nevents, is3d_particle **particles, int64_t *total, int64_t *counts, is3d_stats *stats)
{ return group_sample_any(g, nevents, false, (void **)particles, total, counts, stats); }

is3d_status is3d_group_sample_compact(is3d_group *g, int64_t nevents, is3d_particle_compact **particles, int64_t *total, int64_t *counts,
                                      is3d_stats *stats)
{ return group_sample_any(g, nevents, true, (void **)particles, total, counts, stats); }

// self-test histograms of a sharded sampler run: sums of the per-device counters (BinSampledParticle.cpp counts are additive)
is3d_status is3d_group_sample_histograms(is3d_group *g, double *dN_dy, double *dN_deta, double *dN_dphipdy, double *dN_2pipTdpTdy,
                                         double *pT_count, double *vn_real, double *vn_imag, double *dN_taudtaudy,
                                         double *dN_twopirdrdy, double *dN_dphisdy)
{
  if (!g) return IS3D_ERR_INVALID;
  is3d_ctx *c0 = g->ctx[0];
  const is3d_params &p = c0->prm;
  const size_t ns = c0->ns;
  double *dst[10] = {dN_dy, dN_deta, dN_dphipdy, dN_2pipTdpTdy, pT_count, vn_real, vn_imag, dN_taudtaudy, dN_twopirdrdy, dN_dphisdy};
  const size_t sizes[10] = {ns * p.y_bins, ns * p.eta_bins, ns * p.phip_bins, ns * p.pT_bins, ns * p.pT_bins, 7 * ns * p.pT_bins,
                            7 * ns * p.pT_bins, ns * p.tau_bins, ns * p.r_bins, ns * p.phip_bins};
  for (size_t d = 0; d < g->ctx.size(); d++) {
    std::vector<std::vector<double>> tmp(10);
    double *q[10];
    for (int k = 0; k < 10; k++) {
      if (!dst[k]) { q[k] = nullptr; continue; }
      if (d == 0) q[k] = dst[k];
      else { tmp[k].assign(sizes[k], 0.0); q[k] = tmp[k].data(); }
    }
    is3d_status rc = is3d_sample_histograms(g->ctx[d], q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8], q[9]);
    if (rc != IS3D_OK) { g->err = g->ctx[d]->err; return rc; }
    if (d > 0)
      for (int k = 0; k < 10; k++)
        if (dst[k]) for (size_t j = 0; j < sizes[k]; j++) dst[k][j] += tmp[k][j];
  }
  return IS3D_OK;
}

}  // extern "C"
