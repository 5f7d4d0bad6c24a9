// C ABI entry points (include/is3d_b200.h): context lifetime, static tables, surface upload, dispatch.
#include <cmath>
#include <cstring>

#include "ctx.h"

using namespace is3d;

static thread_local std::string g_create_error;


extern "C" {

const char *is3d_version(void) { return "is3d_b200 0.2 (sm_100a)"; }

int is3d_device_count(void)
{
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

void is3d_default_params(is3d_params *p)
{
  std::memset(p, 0, sizeof(*p));
  p->operation = 1; p->dimension = 3; p->df_mode = 1;
  p->include_bulk_deltaf = 1; p->include_shear_deltaf = 1;
  p->deta_min = 1.e-5; p->mass_pion0 = 0.138;
  p->fast = 1; p->y_cut = 5.0; p->sampler_seed = 1; p->test_sampler = 0;
  p->pT_min = 0.0; p->pT_max = 3.0; p->pT_bins = 100; p->y_bins = 100; p->phip_bins = 100;
  p->eta_cut = 7.0; p->eta_bins = 140; p->tau_min = 0.0; p->tau_max = 12.0; p->tau_bins = 120;
  p->r_min = 0.0; p->r_max = 12.0; p->r_bins = 60;
  p->device = 0; p->famod_chain = 0; p->dndx_bug_compat = 0; p->polzn_chunk_compat = 0;
  p->negligible_margin = 60.0;
}

const char *is3d_last_error(const is3d_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

is3d_status is3d_create(const is3d_params *p, is3d_ctx **out)
{
  *out = nullptr;
  if (!p) { g_create_error = "params is NULL"; return IS3D_ERR_INVALID; }
  // the reference's own argument checks (EmissionFunction.cpp:146-187)
  if (p->dimension != 2 && p->dimension != 3) { g_create_error = "need to set dimension = (2,3)"; return IS3D_ERR_INVALID; }
  if (p->df_mode < 1 || p->df_mode > 5) { g_create_error = "need to set df_mode = (1,2,3,4,5)"; return IS3D_ERR_INVALID; }
  if (p->operation < 0 || p->operation > 2) { g_create_error = "need to set operation = (0, 1, 2)"; return IS3D_ERR_INVALID; }
  if (p->df_mode == 4 && p->include_baryon) {
    g_create_error = "PTB (df_mode 4) has no muB != 0 coefficient tables (reference DeltafData.cpp:480-484)";
    return IS3D_ERR_UNSUPPORTED;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)";
    return IS3D_ERR_NO_DEVICE;
  }
  if (p->device < 0 || p->device >= ndev) { g_create_error = "device ordinal out of range"; return IS3D_ERR_INVALID; }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, p->device);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return IS3D_ERR_CUDA; }
  if (prop.major != 10) {
    g_create_error = "device is sm_" + std::to_string(prop.major * 10 + prop.minor) + "; kernels are built for sm_100a only";
    return IS3D_ERR_NO_DEVICE;
  }
  e = cudaSetDevice(p->device);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return IS3D_ERR_CUDA; }
  is3d_ctx *ctx = new is3d_ctx;
  ctx->prm = *p;
  ctx->sm_count = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return IS3D_ERR_CUDA; }
  if (cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    g_create_error = "cudaEventCreate failed";
    is3d_destroy(ctx);
    return IS3D_ERR_CUDA;
  }
  {
    std::vector<double> tab(kExpTableSize);
    fill_exp_table(tab.data());
    if (ctx->upload(&ctx->d_exptab, tab.data(), tab.size()) != IS3D_OK) {
      g_create_error = ctx->err;
      is3d_destroy(ctx);
      return IS3D_ERR_CUDA;
    }
  }
  *out = ctx;
  return IS3D_OK;
}

void is3d_destroy(is3d_ctx *ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->prm.device);
  cudaStreamSynchronize(ctx->stream);
  comm_release(ctx);
  for (void *p : ctx->owned) cudaFree(p);
  release_host_lists_of(ctx);
  if (ctx->h_words) cudaFreeHost((void *)ctx->h_words);
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  for (int k = 0; k < 2; k++) {
    if (ctx->ev_sorted[k]) cudaEventDestroy(ctx->ev_sorted[k]);
    if (ctx->ev_copied[k]) cudaEventDestroy(ctx->ev_copied[k]);
  }
  if (ctx->side_stream) { cudaStreamSynchronize(ctx->side_stream); cudaStreamDestroy(ctx->side_stream); }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

void *is3d_stream(is3d_ctx *ctx) { return (void *)ctx->stream; }

#define CTX_ENTER(ctx)                                                  \
  if (!(ctx)) return IS3D_ERR_INVALID;                                  \
  IS3D_CUDA_TRY(ctx, cudaSetDevice((ctx)->prm.device));

is3d_status is3d_set_species(is3d_ctx *ctx, int n, const double *mass, const double *sign, const double *degeneracy,
                             const double *baryon, const int *mcid, const double *neq, const double *dn_bulk,
                             const double *dn_diff)
{
  CTX_ENTER(ctx);
  if (n <= 0 || !mass || !sign || !degeneracy || !baryon) { ctx->set_error("set_species: bad arguments"); return IS3D_ERR_INVALID; }
  ctx->ns = n;
  ctx->h_mass.assign(mass, mass + n); ctx->h_sign.assign(sign, sign + n);
  ctx->h_deg.assign(degeneracy, degeneracy + n); ctx->h_baryon.assign(baryon, baryon + n);
  ctx->h_mcid.assign(n, 0); if (mcid) ctx->h_mcid.assign(mcid, mcid + n);
  ctx->h_neq.assign(n, 0.0); if (neq) ctx->h_neq.assign(neq, neq + n);
  ctx->h_dnbulk.assign(n, 0.0); if (dn_bulk) ctx->h_dnbulk.assign(dn_bulk, dn_bulk + n);
  ctx->h_dndiff.assign(n, 0.0); if (dn_diff) ctx->h_dndiff.assign(dn_diff, dn_diff + n);
  IS3D_TRY(ctx->upload(&ctx->d_mass, ctx->h_mass.data(), n));
  IS3D_TRY(ctx->upload(&ctx->d_sign, ctx->h_sign.data(), n));
  IS3D_TRY(ctx->upload(&ctx->d_deg, ctx->h_deg.data(), n));
  IS3D_TRY(ctx->upload(&ctx->d_baryon, ctx->h_baryon.data(), n));
  IS3D_TRY(ctx->upload(&ctx->d_mcid, ctx->h_mcid.data(), n));
  IS3D_TRY(ctx->upload(&ctx->d_neq, ctx->h_neq.data(), n));
  IS3D_TRY(ctx->upload(&ctx->d_dnbulk, ctx->h_dnbulk.data(), n));
  IS3D_TRY(ctx->upload(&ctx->d_dndiff, ctx->h_dndiff.data(), n));
  return IS3D_OK;
}

is3d_status is3d_set_pdg(is3d_ctx *ctx, int n, const double *mass, const double *sign, const double *degeneracy,
                         const double *baryon)
{
  CTX_ENTER(ctx);
  if (n <= 0 || !mass || !sign || !degeneracy || !baryon) { ctx->set_error("set_pdg: bad arguments"); return IS3D_ERR_INVALID; }
  ctx->npdg = n;
  ctx->h_pdg_mass.assign(mass, mass + n); ctx->h_pdg_sign.assign(sign, sign + n);
  ctx->h_pdg_deg.assign(degeneracy, degeneracy + n); ctx->h_pdg_baryon.assign(baryon, baryon + n);
  IS3D_TRY(ctx->upload(&ctx->d_pdg_mass, mass, n));
  IS3D_TRY(ctx->upload(&ctx->d_pdg_sign, sign, n));
  IS3D_TRY(ctx->upload(&ctx->d_pdg_deg, degeneracy, n));
  IS3D_TRY(ctx->upload(&ctx->d_pdg_baryon, baryon, n));
  return IS3D_OK;
}

is3d_status is3d_set_momentum_tables(is3d_ctx *ctx, int npT, const double *pT, const double *pTw, int nphi,
                                     const double *phi, const double *phiw, int ny, const double *y, const double *yw,
                                     int neta, const double *eta, const double *etaw)
{
  CTX_ENTER(ctx);
  if (npT <= 0 || nphi <= 0 || !pT || !phi) { ctx->set_error("set_momentum_tables: bad arguments"); return IS3D_ERR_INVALID; }
  const int dim = ctx->prm.dimension;
  if (dim == 3 && (ny <= 0 || !y)) { ctx->set_error("set_momentum_tables: 3+1d needs the y table"); return IS3D_ERR_INVALID; }
  if (dim == 2 && (neta <= 0 || !eta || !etaw)) { ctx->set_error("set_momentum_tables: 2+1d needs the eta table"); return IS3D_ERR_INVALID; }
  ctx->pT.assign(pT, pT + npT);
  ctx->pTw.assign(npT, 1.0); if (pTw) ctx->pTw.assign(pTw, pTw + npT);
  ctx->phi.assign(phi, phi + nphi);
  ctx->phiw.assign(nphi, 1.0); if (phiw) ctx->phiw.assign(phiw, phiw + nphi);
  ctx->y.clear(); ctx->yw.clear(); ctx->eta.clear(); ctx->etaw.clear();
  if (y) { ctx->y.assign(y, y + ny); ctx->yw.assign(ny, 1.0); if (yw) ctx->yw.assign(yw, yw + ny); }
  if (eta) { ctx->eta.assign(eta, eta + neta); ctx->etaw.assign(neta, 1.0); if (etaw) ctx->etaw.assign(etaw, etaw + neta); }
  ctx->NpT = npT; ctx->Nphi = nphi;
  // dimension rule, EmissionFunction.cpp:146-153 and MomentumSpectra.cpp:73-91
  std::vector<double> yweff;
  if (dim == 2) {
    ctx->Ny = 1; ctx->yv.assign(1, 0.0); yweff.assign(1, 1.0);
    ctx->Neta = neta; ctx->etav = ctx->eta; ctx->etawv = ctx->etaw;
  } else {
    ctx->Ny = ny; ctx->yv = ctx->y; yweff = ctx->yw;
    ctx->Neta = 1; ctx->etav.assign(1, 0.0); ctx->etawv.assign(1, 1.0);
  }
  std::vector<double> c(nphi), s(nphi);
  for (int i = 0; i < nphi; i++) { c[i] = cos(phi[i]); s[i] = sin(phi[i]); }   // MomentumSpectra.cpp:53-58
  IS3D_TRY(ctx->upload(&ctx->d_pT, ctx->pT.data(), npT));
  IS3D_TRY(ctx->upload(&ctx->d_pTw, ctx->pTw.data(), npT));
  IS3D_TRY(ctx->upload(&ctx->d_cosphi, c.data(), nphi));
  IS3D_TRY(ctx->upload(&ctx->d_sinphi, s.data(), nphi));
  IS3D_TRY(ctx->upload(&ctx->d_phiw, ctx->phiw.data(), nphi));
  IS3D_TRY(ctx->upload(&ctx->d_y, ctx->yv.data(), ctx->yv.size()));
  IS3D_TRY(ctx->upload(&ctx->d_yw, yweff.data(), yweff.size()));
  IS3D_TRY(ctx->upload(&ctx->d_eta, ctx->etav.data(), ctx->etav.size()));
  IS3D_TRY(ctx->upload(&ctx->d_etaw, ctx->etawv.data(), ctx->etawv.size()));
  ctx->have_momentum = true;
  return IS3D_OK;
}

is3d_status is3d_set_gauss_tables(is3d_ctx *ctx, int n_alpha, int n_points, const double *root, const double *weight,
                                  int n_leg, const double *leg_root, const double *leg_weight)
{
  CTX_ENTER(ctx);
  if (n_alpha < 4 || n_points <= 0 || !root || !weight) { ctx->set_error("set_gauss_tables: need alpha >= 4 rows of Gauss-Laguerre data"); return IS3D_ERR_INVALID; }
  ctx->gla_alpha = n_alpha; ctx->gla_pts = n_points;
  ctx->h_gla_root.assign(root, root + (size_t)n_alpha * n_points);
  ctx->h_gla_weight.assign(weight, weight + (size_t)n_alpha * n_points);
  IS3D_TRY(ctx->upload(&ctx->d_gla_root, root, (size_t)n_alpha * n_points));
  IS3D_TRY(ctx->upload(&ctx->d_gla_weight, weight, (size_t)n_alpha * n_points));
  ctx->leg_pts = 0;
  if (n_leg > 0 && leg_root && leg_weight) {
    ctx->leg_pts = n_leg;
    ctx->h_leg_root.assign(leg_root, leg_root + n_leg);
    ctx->h_leg_weight.assign(leg_weight, leg_weight + n_leg);
    IS3D_TRY(ctx->upload(&ctx->d_leg_root, leg_root, n_leg));
    IS3D_TRY(ctx->upload(&ctx->d_leg_weight, leg_weight, n_leg));
  }
  return IS3D_OK;
}

is3d_status is3d_set_thermo_averages(is3d_ctx *ctx, double T, double E, double P, double muB, double nB)
{
  if (!ctx) return IS3D_ERR_INVALID;
  ctx->T_avg = T; ctx->E_avg = E; ctx->P_avg = P; ctx->muB_avg = muB; ctx->nB_avg = nB;
  ctx->have_avg = true;
  return IS3D_OK;
}

static is3d_status upload_spline(is3d_ctx *ctx, const std::vector<double> &x, const double *y, int n, Spline *sp,
                                 const double *d_x)
{
  std::vector<double> c(n);
  natural_cspline_coefficients(x.data(), y, n, c.data());
  double *dy = const_cast<double *>(sp->y), *dc = const_cast<double *>(sp->c);   // a repeated call releases the previous copies
  IS3D_TRY(ctx->upload(&dy, y, n));
  IS3D_TRY(ctx->upload(&dc, c.data(), n));
  sp->x = d_x; sp->y = dy; sp->c = dc; sp->n = n;
  return IS3D_OK;
}

is3d_status is3d_set_df_tables(is3d_ctx *ctx, int n_T, int n_muB, const double *T, const double *muB, const double *c0,
                               const double *c1, const double *c2, const double *c3, const double *c4, const double *F,
                               const double *G, const double *betabulk, const double *betaV, const double *betapi)
{
  CTX_ENTER(ctx);
  const double *tabs[10] = {c0, c1, c2, c3, c4, F, G, betabulk, betaV, betapi};
  if (n_T < 3 || n_muB < 1 || !T) { ctx->set_error("set_df_tables: bad arguments"); return IS3D_ERR_INVALID; }
  for (int k = 0; k < 10; k++) if (!tabs[k]) { ctx->set_error("set_df_tables: NULL table"); return IS3D_ERR_INVALID; }
  if (ctx->prm.include_baryon && (n_muB < 2 || !muB)) { ctx->set_error("set_df_tables: include_baryon needs the muB grid"); return IS3D_ERR_INVALID; }
  DfTables &tb = ctx->tb;
  tb.n_T = n_T; tb.n_muB = n_muB;
  ctx->h_T.assign(T, T + n_T);
  ctx->h_muB.assign(n_muB, 0.0); if (muB) ctx->h_muB.assign(muB, muB + n_muB);
  tb.T_min = T[0]; tb.muB_min = ctx->h_muB[0];
  tb.dT = fabs(T[1] - T[0]);                                   // DeltafData.cpp:199-204
  tb.dmuB = n_muB > 1 ? fabs(ctx->h_muB[1] - ctx->h_muB[0]) : 0.0;
  double *dT = const_cast<double *>(tb.T), *dB = const_cast<double *>(tb.muB);    // upload() frees what a previous call left
  IS3D_TRY(ctx->upload(&dT, ctx->h_T.data(), n_T));
  IS3D_TRY(ctx->upload(&dB, ctx->h_muB.data(), n_muB));
  tb.T = dT; tb.muB = dB;
  for (int k = 0; k < 10; k++) {
    ctx->h_tab[k].assign(tabs[k], tabs[k] + (size_t)n_T * n_muB);
    double *d = const_cast<double *>(tb.tab[k]);
    IS3D_TRY(ctx->upload(&d, tabs[k], (size_t)n_T * n_muB));
    tb.tab[k] = d;
  }
  // cubic splines of the muB = 0 row (construct_cubic_splines, DeltafData.cpp:298-321)
  IS3D_TRY(upload_spline(ctx, ctx->h_T, c0, n_T, &tb.sp_c0, dT));
  IS3D_TRY(upload_spline(ctx, ctx->h_T, c2, n_T, &tb.sp_c2, dT));
  IS3D_TRY(upload_spline(ctx, ctx->h_T, F, n_T, &tb.sp_F, dT));
  IS3D_TRY(upload_spline(ctx, ctx->h_T, betabulk, n_T, &tb.sp_betabulk, dT));
  IS3D_TRY(upload_spline(ctx, ctx->h_T, betapi, n_T, &tb.sp_betapi, dT));
  ctx->have_df = true;
  return IS3D_OK;
}

is3d_status is3d_set_ptb_tables(is3d_ctx *ctx, int n, const double *x, const double *l2, const double *z, double xmax)
{
  CTX_ENTER(ctx);
  if (n < 3 || !x || !l2 || !z) { ctx->set_error("set_ptb_tables: bad arguments"); return IS3D_ERR_INVALID; }
  ctx->h_ptb_x.assign(x, x + n); ctx->h_ptb_l2.assign(l2, l2 + n); ctx->h_ptb_z.assign(z, z + n);
  double *dx = const_cast<double *>(ctx->tb.sp_lambda2.x);
  IS3D_TRY(ctx->upload(&dx, x, n));
  IS3D_TRY(upload_spline(ctx, ctx->h_ptb_x, l2, n, &ctx->tb.sp_lambda2, dx));
  IS3D_TRY(upload_spline(ctx, ctx->h_ptb_x, z, n, &ctx->tb.sp_z, dx));
  ctx->tb.bulkPi_over_P_max = xmax;
  ctx->have_ptb = true;
  return IS3D_OK;
}

static is3d_status set_surface_impl(is3d_ctx *ctx, int64_t n, const double *const cols[IS3D_SURFACE_COLUMNS],
                                    int64_t global_offset, bool device_ptrs)
{
  CTX_ENTER(ctx);
  if (n < 0 || !cols) { ctx->set_error("set_surface: bad arguments"); return IS3D_ERR_INVALID; }
  const int ncol = ctx->prm.include_baryon ? 25 : 20;
  for (int k = 0; k < ncol; k++)
    if (!cols[k] && n > 0) { ctx->set_error("set_surface: NULL column " + std::to_string(k)); return IS3D_ERR_INVALID; }
  ctx->surface_owned = false;
  for (int k = 0; k < 25; k++) ctx->surf.col[k] = nullptr;
  ctx->surf.n = n;
  ctx->global_offset = global_offset;
  if (device_ptrs) {
    for (int k = 0; k < ncol; k++) ctx->surf.col[k] = cols[k];
  } else {
    // one block, columns padded to 256 B so every column start is aligned for vector loads
    // the block is grow-only: cudaFree of a large allocation synchronises the device and can take hundreds of ms
    int64_t pitch = (n + 31) / 32 * 32;
    const size_t need = (size_t)ncol * pitch * sizeof(double);
    if (!ctx->d_surface_block || ctx->surface_block_bytes < need) {
      if (ctx->d_surface_block) { ctx->dev_free(ctx->d_surface_block); ctx->d_surface_block = nullptr; ctx->surface_block_bytes = 0; }
      void *blk = nullptr;
      IS3D_TRY(ctx->dev_alloc(&blk, need));
      ctx->d_surface_block = (double *)blk;
      ctx->surface_block_bytes = need;
    }
    ctx->surface_owned = true;
    for (int k = 0; k < ncol; k++) {
      double *dst = ctx->d_surface_block + (size_t)k * pitch;
      if (n) IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(dst, cols[k], (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      ctx->surf.col[k] = dst;
    }
    IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  ctx->have_surface = true;
  return IS3D_OK;
}

is3d_status is3d_set_surface(is3d_ctx *ctx, int64_t n, const double *const cols[IS3D_SURFACE_COLUMNS], int64_t off)
{ return set_surface_impl(ctx, n, cols, off, false); }

is3d_status is3d_set_surface_device(is3d_ctx *ctx, int64_t n, const double *const cols[IS3D_SURFACE_COLUMNS], int64_t off)
{ return set_surface_impl(ctx, n, cols, off, true); }

int64_t is3d_spectra_size(const is3d_ctx *ctx)
{
  if (!ctx || !ctx->have_momentum || ctx->ns <= 0) return 0;
  return (int64_t)ctx->ns * ctx->NpT * ctx->Nphi * ctx->Ny;
}

static is3d_status check_ready(is3d_ctx *ctx, bool need_df)
{
  if (ctx->ns <= 0) { ctx->set_error("species not set"); return IS3D_ERR_INVALID; }
  if (!ctx->have_momentum) { ctx->set_error("momentum tables not set"); return IS3D_ERR_INVALID; }
  if (!ctx->have_surface) { ctx->set_error("surface not set"); return IS3D_ERR_INVALID; }
  if (need_df && !ctx->have_df) { ctx->set_error("df coefficient tables not set"); return IS3D_ERR_INVALID; }
  if (ctx->prm.df_mode == 4 && !ctx->have_ptb) { ctx->set_error("PTB tables not set"); return IS3D_ERR_INVALID; }
  return IS3D_OK;
}

is3d_status is3d_spectra_device(is3d_ctx *ctx, double *out_dev, is3d_stats *stats)
{
  CTX_ENTER(ctx);
  if (!out_dev) { ctx->set_error("spectra: NULL output"); return IS3D_ERR_INVALID; }
  if (stats) std::memset(stats, 0, sizeof(*stats));
  IS3D_TRY(check_ready(ctx, ctx->prm.df_mode != 5 || true));
  const int64_t total = is3d_spectra_size(ctx);
  is3d_status st = IS3D_OK;
  if (ctx->surf.n == 0) {
    IS3D_CUDA_TRY(ctx, cudaMemsetAsync(out_dev, 0, total * sizeof(double), ctx->stream));
  } else if (ctx->prm.df_mode <= 2) {
    st = run_spectra_df(ctx, out_dev, stats);
  } else {
    st = run_spectra_feqmod(ctx, out_dev, stats);
  }
  // cells sharded over GPUs: ONE all-reduce of the spectra.  A rank that failed locally still joins the
  // collective (its peers would otherwise wait forever) and then reports its own error.
  const is3d_status sc = comm_allreduce(ctx, out_dev, total);
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return st != IS3D_OK ? st : sc;
}

is3d_status is3d_spectra(is3d_ctx *ctx, double *out, is3d_stats *stats)
{
  CTX_ENTER(ctx);
  if (!out) { ctx->set_error("spectra: NULL output"); return IS3D_ERR_INVALID; }
  const int64_t total = is3d_spectra_size(ctx);
  void *d = nullptr;
  IS3D_TRY(ctx->get_scratch("spectra_out", (size_t)(total > 0 ? total : 1) * sizeof(double), &d));
  IS3D_TRY(is3d_spectra_device(ctx, (double *)d, stats));
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(out, d, total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return IS3D_OK;
}

is3d_status is3d_set_vorticity(is3d_ctx *ctx, int64_t n, const double *const w[6])
{
  CTX_ENTER(ctx);
  if (n < 0 || !w) { ctx->set_error("set_vorticity: bad arguments"); return IS3D_ERR_INVALID; }
  for (int k = 0; k < 6; k++) if (!w[k] && n > 0) { ctx->set_error("set_vorticity: NULL column " + std::to_string(k)); return IS3D_ERR_INVALID; }
  if (!ctx->have_surface || ctx->surf.n != n) { ctx->set_error("set_vorticity: set the surface first (same number of cells)"); return IS3D_ERR_INVALID; }
  const int64_t pitch = (n + 31) / 32 * 32;
  if (ctx->d_vorticity) { ctx->dev_free(ctx->d_vorticity); ctx->d_vorticity = nullptr; }
  void *blk = nullptr;
  IS3D_TRY(ctx->dev_alloc(&blk, (size_t)6 * (pitch ? pitch : 32) * sizeof(double)));
  ctx->d_vorticity = (double *)blk;
  for (int k = 0; k < 6 && n; k++)
    IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_vorticity + (size_t)k * pitch, w[k], (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->vorticity_n = n; ctx->vorticity_pitch = pitch; ctx->have_vorticity = true;
  return IS3D_OK;
}

is3d_status is3d_polarization(is3d_ctx *ctx, double *St, double *Sx, double *Sy, double *Sn, double *Snorm, is3d_stats *stats)
{
  CTX_ENTER(ctx);
  if (!St || !Sx || !Sy || !Sn || !Snorm) { ctx->set_error("polarization: NULL output"); return IS3D_ERR_INVALID; }
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (ctx->ns <= 0 || !ctx->have_momentum || !ctx->have_surface) { ctx->set_error("polarization: species / momentum tables / surface not set"); return IS3D_ERR_INVALID; }
  const int64_t total = is3d_spectra_size(ctx);
  double *outs[5] = {St, Sx, Sy, Sn, Snorm};
  void *d = nullptr;
  IS3D_TRY(ctx->get_scratch("pol_out", (size_t)5 * total * sizeof(double), &d));
  is3d_status st = IS3D_OK;
  if (ctx->surf.n == 0) IS3D_CUDA_TRY(ctx, cudaMemsetAsync(d, 0, (size_t)5 * total * sizeof(double), ctx->stream));
  else st = run_polarization(ctx, (double *)d, stats);
  const is3d_status sc = comm_allreduce(ctx, (double *)d, 5 * total);       // sharded surface: sums of the five arrays
  if (st != IS3D_OK) return st;
  IS3D_TRY(sc);
  for (int k = 0; k < 5; k++)
    IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(outs[k], (double *)d + (size_t)k * total, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return IS3D_OK;
}

is3d_status is3d_copy_from_device(is3d_ctx *ctx, void *host, const void *device, size_t bytes)
{
  CTX_ENTER(ctx);
  if (bytes && (!host || !device)) { ctx->set_error("copy_from_device: NULL pointer"); return IS3D_ERR_INVALID; }
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(host, device, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return IS3D_OK;
}

is3d_status is3d_measure_fp64_peak(is3d_ctx *ctx, double *tflops)
{
  CTX_ENTER(ctx);
  return measure_fp64_peak(ctx, tflops);
}

is3d_status is3d_probe_math(is3d_ctx *ctx, int64_t n, const double *x, double *out_exp, double *out_rcp, double *out_sqrt)
{
  CTX_ENTER(ctx);
  if (n <= 0 || !x || !out_exp || !out_rcp || !out_sqrt) { ctx->set_error("probe_math: bad arguments"); return IS3D_ERR_INVALID; }
  return probe_math(ctx, n, x, out_exp, out_rcp, out_sqrt);
}

is3d_status is3d_probe_aniso_math(is3d_ctx *ctx, int64_t n, const double *x, double *out_atan_over_s, double *out_atanh_over_s, double *out_log)
{
  CTX_ENTER(ctx);
  if (n <= 0 || !x || !out_atan_over_s || !out_atanh_over_s || !out_log) { ctx->set_error("probe_aniso_math: bad arguments"); return IS3D_ERR_INVALID; }
  return probe_aniso_math(ctx, n, x, out_atan_over_s, out_atanh_over_s, out_log);
}

}  // extern "C"
