// K7: spin polarization vector of the chosen hadrons from the thermal vorticity of a mode-5 surface.
// Replaces EmissionFunctionArray::calculate_spin_polzn (reference src/cpp/Polarization.cpp:25-263).
//
// Per (cell, species, pT, phi, y[, eta]) the reference accumulates  w p.dsigma f0 S_mu  (mu = t, x, y, n)  and the norm
// w p.dsigma f0, with f0 = 1 / (exp(u.p / T_avg) + sign) taken at the SURFACE-AVERAGED temperature (:76, :186) and
//     S_mu = -(1 / 8m) (1 - sign f0) 2 eps-contraction of the thermal vorticity with p  (:189-193).
// Every S_mu is linear in (mT, pT) once (cell, y, phi) are fixed, so the K1 schedule applies unchanged: a thread owns R
// species classes at one pT node, a block streams its chunk of cells in 256-cell tiles of 12-double items, five register
// accumulators per bin.  Species of the same (mass, sign) form one class (the degeneracy never enters).  Cells with
// u.dsigma <= 0 are NOT skipped here (the reference does not).
//
// Reference quirk kept under polzn_chunk_compat: the thermal vorticity is read with the index INSIDE the 10 000-cell chunk
// (wtx_fo[icell], :125-130) while every other column uses the global index; for surfaces above 10 000 cells the reference
// therefore pairs cell 10 000 + k with the vorticity of cell k.
#include "ctx.h"
#include "spectra_df.cuh"

namespace is3d {

namespace {

constexpr int kThreads = 256;
constexpr int kTile = kThreads;
constexpr int kR = 3;
constexpr int64_t kRefChunk = 10000;      // FO_chunk, Polarization.cpp:31

struct alignas(16) PolItem {
  double aT, bT, c1, d1;                  // u.p / T_avg = mT aT - pT bT;  w p.dsigma = mT c1 + pT d1
  double At, Bt, Ax, Bx;                  // S_mu / prefactor = mT A_mu + pT B_mu
  double Ay, By, An, Bn;
};

struct PolGrid {
  const double *mT, *pT, *sign;           // per (class, pT) bin
  int nclass, NpT, ncols;
  int Ny, Nphi, Neta, dimension;
  const double *yv, *cosphi, *sinphi, *etav, *etaw;
  double deta;                            // eta_table step: the 2+1d weights are w_eta * delta_eta (:57-70)
  double invT;                            // 1 / T_avg
  const double *exptab;
  const double *w[6];                     // wtx wty wtn wxy wxn wyn
  int chunk_compat;
};

template <int R>
__global__ void __launch_bounds__(kThreads, 2)
polarization_kernel(SurfaceView surf, int64_t ncells, int64_t cells_per_chunk, PolGrid g, double *__restrict__ partial, int64_t total)
{
  __shared__ PolItem items[kTile];
  __shared__ double exptab[kExpTableSize];
  load_exp_table(exptab, g.exptab);
  const int t = threadIdx.x;
  const int iy = blockIdx.y / g.Nphi, iphi = blockIdx.y - iy * g.Nphi;
  const double yval = g.yv[iy], cphi = g.cosphi[iphi], sphi = g.sinphi[iphi];
  const int col = blockIdx.x * kThreads + t;
  const int colc = col < g.ncols ? col : g.ncols - 1;
  const int grp = colc / g.NpT, ip = colc - grp * g.NpT;
  double mT[R], sign[R], acc[R][5];
  int jbin[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int c = grp * R + r;
    const int jj = (c < g.nclass ? c : g.nclass - 1) * g.NpT + ip;
    jbin[r] = (col < g.ncols && c < g.nclass) ? jj : -1;
    mT[r] = g.mT[jj]; sign[r] = g.sign[jj];
#pragma unroll
    for (int k = 0; k < 5; k++) acc[r][k] = 0.0;
  }
  const double pT = g.pT[ip];
  const int64_t chunk_begin = (int64_t)blockIdx.z * cells_per_chunk;
  int64_t chunk_end = chunk_begin + cells_per_chunk;
  if (chunk_end > ncells) chunk_end = ncells;

  for (int64_t tile = chunk_begin; tile < chunk_end; tile += kTile) {
    const int64_t cell = tile + t;
    const int n_items = (int)((chunk_end - tile) < kTile ? (chunk_end - tile) : kTile);
    for (int ie = 0; ie < g.Neta; ie++) {
      __syncthreads();
      if (cell < chunk_end) {
        const double tau = surf.col[IS3D_COL_TAU][cell], tau2 = tau * tau;
        const double ux = surf.col[IS3D_COL_UX][cell], uy = surf.col[IS3D_COL_UY][cell], un = surf.col[IS3D_COL_UN][cell];
        const double ut = sqrt(fabs(1.0 + ux * ux + uy * uy + tau2 * un * un));
        double eta, w;
        if (g.dimension == 3) { eta = surf.col[IS3D_COL_ETA][cell]; w = 1.0; }
        else { eta = g.etav[ie]; w = g.etaw[ie] * g.deta; }
        const double d = yval - eta, ch = cosh(d), sh = sinh(d);
        const int64_t wc = g.chunk_compat ? (cell % kRefChunk) : cell;
        const double wtx = g.w[0][wc], wty = g.w[1][wc], wtn = g.w[2][wc], wxy = g.w[3][wc], wxn = g.w[4][wc], wyn = g.w[5][wc];
        PolItem it;
        // p^tau = mT ch, p^eta = mT sh / tau, p^x = pT cphi, p^y = pT sphi
        it.aT = (ch * ut - sh * tau * un) * g.invT;
        it.bT = (cphi * ux + sphi * uy) * g.invT;
        it.c1 = w * (ch * surf.col[IS3D_COL_DAT][cell] + sh / tau * surf.col[IS3D_COL_DAN][cell]);
        it.d1 = w * (cphi * surf.col[IS3D_COL_DAX][cell] + sphi * surf.col[IS3D_COL_DAY][cell]);
        const double pn1 = sh / tau;                               // p^eta / mT
        it.At = wxy * pn1;            it.Bt = -wxn * sphi + wyn * cphi;
        it.Ax = wyn * ch + wty * pn1; it.Bx = -wtn * sphi;
        it.Ay = -wxn * ch - wtx * pn1; it.By = wtn * cphi;
        it.An = wxy * ch;             it.Bn = wtx * sphi - wty * cphi;
        items[t] = it;
      }
      __syncthreads();
#pragma unroll 1
      for (int k = 0; k < n_items; k++) {
        const PolItem it = items[k];
        const double pb = pT * it.bT, pd = pT * it.d1;
        const double pBt = pT * it.Bt, pBx = pT * it.Bx, pBy = pT * it.By, pBn = pT * it.Bn;
#pragma unroll
        for (int r = 0; r < R; r++) {
          const double x = fma(mT[r], it.aT, -pb);
          const double f0 = fast_rcp(fast_exp(x, exptab) + sign[r]);
          const double wgt = fma(mT[r], it.c1, pd) * f0;           // w p.dsigma f0
          const double gq = wgt * fma(-sign[r], f0, 1.0);          // ... (1 - sign f0)
          acc[r][0] = fma(gq, fma(mT[r], it.At, pBt), acc[r][0]);
          acc[r][1] = fma(gq, fma(mT[r], it.Ax, pBx), acc[r][1]);
          acc[r][2] = fma(gq, fma(mT[r], it.Ay, pBy), acc[r][2]);
          acc[r][3] = fma(gq, fma(mT[r], it.An, pBn), acc[r][3]);
          acc[r][4] += wgt;
        }
      }
    }
  }
  // partial[chunk][component][class bins]
  const int64_t pbase = (int64_t)blockIdx.z * 5 * total;
#pragma unroll
  for (int r = 0; r < R; r++) {
    if (jbin[r] >= 0) {
      const int64_t idx = iy + (int64_t)g.Ny * (iphi + (int64_t)g.Nphi * jbin[r]);
#pragma unroll
      for (int k = 0; k < 5; k++) partial[pbase + k * total + idx] += acc[r][k];
    }
  }
}

// sum over chunks, expansion of the classes to species, and the species prefactor -(1 / 8m) 2 of the four components
__global__ void polarization_reduce_kernel(const double *__restrict__ partial, int nchunks, int64_t total_class, int64_t per_species,
                                           const int *__restrict__ class_of, const double *__restrict__ mass, int64_t total,
                                           double *__restrict__ St, double *__restrict__ Sx, double *__restrict__ Sy,
                                           double *__restrict__ Sn, double *__restrict__ Snorm)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t sp = i / per_species, src = (int64_t)class_of[sp] * per_species + (i - sp * per_species);
  double s[5] = {0, 0, 0, 0, 0};
  for (int c = 0; c < nchunks; c++)
    for (int k = 0; k < 5; k++) s[k] += partial[((int64_t)c * 5 + k) * total_class + src];
  const double pref = -(1.0 / 8.0 / mass[sp]) * 2.0;
  St[i] = pref * s[0]; Sx[i] = pref * s[1]; Sy[i] = pref * s[2]; Sn[i] = pref * s[3]; Snorm[i] = s[4];
}

}  // namespace

void choose_chunks(const is3d_ctx *ctx, int64_t ncells, int64_t blocks_per_chunk, int64_t total, int tile, int blocks_per_sm,
                   int *nchunks, int64_t *cells_per_chunk);

// out_dev: five arrays of Ns NpT Nphi Ny doubles each (St, Sx, Sy, Sn, Snorm), index iy + Ny (iphi + Nphi (ipT + NpT is))
is3d_status run_polarization(is3d_ctx *ctx, double *out_dev, is3d_stats *stats)
{
  const is3d_params &p = ctx->prm;
  if (!ctx->have_vorticity || ctx->vorticity_n != ctx->surf.n) { ctx->set_error("polarization: thermal vorticity not set for this surface (is3d_set_vorticity)"); return IS3D_ERR_INVALID; }
  if (!ctx->have_avg) { ctx->set_error("polarization: thermodynamic averages not set"); return IS3D_ERR_INVALID; }
  const int64_t n = ctx->surf.n;
  const int ns = ctx->ns, npT = ctx->NpT;
  const int64_t per_species = (int64_t)npT * ctx->Nphi * ctx->Ny, total = (int64_t)ns * per_species;

  // classes by (mass, sign): neither the degeneracy nor the baryon number enters the polarization
  std::vector<int> class_of(ns), rep;
  for (int s = 0; s < ns; s++) {
    int c = -1;
    for (size_t k = 0; k < rep.size() && c < 0; k++)
      if (ctx->h_mass[rep[k]] == ctx->h_mass[s] && ctx->h_sign[rep[k]] == ctx->h_sign[s]) c = (int)k;
    if (c < 0) { c = (int)rep.size(); rep.push_back(s); }
    class_of[s] = c;
  }
  const int nc = (int)rep.size(), nb = nc * npT;
  std::vector<double> h(3 * (size_t)nb);
  for (int c = 0; c < nc; c++)
    for (int ip = 0; ip < npT; ip++) {
      const double m = ctx->h_mass[rep[c]], pT = ctx->pT[ip];
      h[c * npT + ip] = sqrt(m * m + pT * pT); h[nb + c * npT + ip] = pT; h[2 * nb + c * npT + ip] = ctx->h_sign[rep[c]];
    }
  void *d = nullptr, *dm = nullptr;
  IS3D_TRY(ctx->get_scratch("pol_bins", h.size() * sizeof(double), &d));
  IS3D_TRY(ctx->get_scratch("class_of", (size_t)ns * sizeof(int), &dm));
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(dm, class_of.data(), (size_t)ns * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));

  PolGrid g;
  g.mT = (const double *)d; g.pT = g.mT + nb; g.sign = g.mT + 2 * nb;
  g.nclass = nc; g.NpT = npT; g.ncols = npT * ((nc + kR - 1) / kR);
  g.Ny = ctx->Ny; g.Nphi = ctx->Nphi; g.Neta = ctx->Neta; g.dimension = p.dimension;
  g.yv = ctx->d_y; g.cosphi = ctx->d_cosphi; g.sinphi = ctx->d_sinphi; g.etav = ctx->d_eta; g.etaw = ctx->d_etaw;
  g.deta = ctx->eta.size() > 1 ? ctx->eta[1] - ctx->eta[0] : 0.0;
  g.invT = 1.0 / ctx->T_avg;
  g.exptab = ctx->d_exptab;
  for (int k = 0; k < 6; k++) g.w[k] = ctx->d_vorticity + (size_t)k * ctx->vorticity_pitch;
  g.chunk_compat = p.polzn_chunk_compat;
  if (g.chunk_compat && ctx->global_offset != 0) {
    // the reference reads rows (global cell) % 10 000 of the WHOLE vorticity array; a shard does not hold them
    ctx->set_error("polzn_chunk_compat = 1 needs the unsharded surface (global_offset = 0); sharded polarization runs with polzn_chunk_compat = 0");
    return IS3D_ERR_UNSUPPORTED;
  }

  const int64_t total_class = (int64_t)nc * per_species;
  const int nslices = (g.ncols + kThreads - 1) / kThreads;
  if ((int64_t)ctx->Ny * ctx->Nphi > 65535) { ctx->set_error("Ny*Nphi exceeds 65535"); return IS3D_ERR_INVALID; }
  int nchunks; int64_t cpc;
  choose_chunks(ctx, n, (int64_t)nslices * ctx->Ny * ctx->Nphi, 5 * total_class, kTile, 2, &nchunks, &cpc);
  void *partial = nullptr;
  IS3D_TRY(ctx->get_scratch("pol_partial", (size_t)nchunks * 5 * total_class * sizeof(double), &partial));
  IS3D_CUDA_TRY(ctx, cudaMemsetAsync(partial, 0, (size_t)nchunks * 5 * total_class * sizeof(double), ctx->stream));
  cudaEvent_t e0 = ctx->ev0, e1 = ctx->ev1;             // owned by the context: nothing to release on an error path
  IS3D_CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
  dim3 grid(nslices, ctx->Ny * ctx->Nphi, nchunks);
  polarization_kernel<kR><<<grid, kThreads, 0, ctx->stream>>>(ctx->surf, n, cpc, g, (double *)partial, total_class);
  IS3D_CUDA_TRY(ctx, cudaGetLastError());
  IS3D_CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
  polarization_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(
      (const double *)partial, nchunks, total_class, per_species, (const int *)dm, ctx->d_mass, total, out_dev, out_dev + total,
      out_dev + 2 * total, out_dev + 3 * total, out_dev + 4 * total);
  IS3D_CUDA_TRY(ctx, cudaGetLastError());
  IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  IS3D_CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
  if (stats) { stats->cells_total = n; stats->kernel_ms = ms; stats->kernel_launches = 2; }
  return IS3D_OK;
}

}  // namespace is3d
