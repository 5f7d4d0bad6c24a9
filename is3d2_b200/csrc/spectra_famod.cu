// K3: per-cell stage of the PTMA modified anisotropic distribution (df_mode 5) on sm_100a: reconstruction of the
// anisotropic variables (Lambda, alpha_T, alpha_L) by Newton iteration, famod coefficients, B_ij and its inverse ->
// the pack consumed by the K2 spectra kernel (spectra_feqmod.cu).
// Replaces the cell prologue of calculate_dN_pTdpTdphidy_famod (reference src/cpp/MomentumSpectra.cpp:1159-1481) and
// src/cpp/AnisoVariables.cpp.
//
// One warp per cell: the 320 x 16 quadrature terms of every F / Jacobian evaluation are strided over the lanes and
// reduced by butterfly shuffles; the Newton / line-search control flow is warp-uniform.
//   famod_chain = 0 (default): every cell starts from (T, 1, 1); cells are independent, warps grid-stride over them.
//   famod_chain = 1: the reference's serial chain (the previous cell's solution seeds the next one,
//                    MomentumSpectra.cpp:1308-1364) is reproduced by ONE warp walking the cells in order -- a parity
//                    mode for checking against the serial reference, ~1e4 cells/s.
#include "aniso.cuh"
#include "ctx.h"

namespace is3d {

namespace {

// p = the warp's staging row in SHARED memory: every lane of the warp computes the same scalars (the Newton control flow is
// warp-uniform) and stores the same values to the same addresses; a per-lane local array cost 14.6 KB of local-memory
// traffic per cell (ncu: 5 GB of DRAM writes per 200 k cells)
__device__ void famod_store(const double *p, int status, int iterations, int64_t i, int64_t begin, double *pack,
                            int64_t stride, unsigned long long *counters)
{
  const int lane = threadIdx.x & 31;
  __syncwarp();
  for (int k = lane; k < FP_SIZE; k += 32) pack[k * stride + i] = p[k];
  __syncwarp();
  if (lane == 0) {
    if (status == CELL_SKIPPED) { atomicAdd(&counters[0], 1ull); return; }
    if (status & CELL_BREAKDOWN) { atomicAdd(&counters[2], 1ull); atomicMax(&counters[4], (unsigned long long)(begin + i + 1)); }
    if (status & CELL_PL_NEGATIVE) { atomicAdd(&counters[3], 1ull); atomicMax(&counters[5], (unsigned long long)(begin + i + 1)); }
    if (status & CELL_RECONSTRUCTION_FAIL) atomicAdd(&counters[8], 1ull);
    atomicAdd(&counters[9], (unsigned long long)iterations);
  }
}

#ifndef IS3D_K3_MINBLOCKS
#define IS3D_K3_MINBLOCKS 6   // occupancy beats registers here: 128 regs -> 206 ms, 80 regs -> 199 ms per step (profiles/r01_k3_sweep.txt)
#endif
__global__ void __launch_bounds__(128, IS3D_K3_MINBLOCKS)
famod_setup_free_kernel(SurfaceView surf, int64_t begin, int64_t count, FamodFlags fl, AnisoHadrons h, double *__restrict__ pack,
                        int64_t stride, unsigned long long *counters)
{
  __shared__ double exptab[kExpTableSize];
  __shared__ double rows[4][FP_SIZE];                   // one staging row per warp of the 128-thread block
  load_exp_table(exptab, h.exptab);
  __syncthreads();
  h.exptab = exptab;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  WarpReducer red;
  double *p = rows[threadIdx.x >> 5];
  for (int64_t i = warp0; i < count; i += nwarps) {
    Cell c = load_cell(surf, begin + i, fl.include_baryon != 0);
    int iterations;
    int status = famod_setup_cell(red, c, fl, h, (FamodChain *)nullptr, p, &iterations);
    famod_store(p, status, iterations, i, begin, pack, stride, counters);
  }
}

__global__ void __launch_bounds__(32)
famod_setup_chain_kernel(SurfaceView surf, int64_t begin, int64_t count, FamodFlags fl, AnisoHadrons h, double *__restrict__ pack,
                         int64_t stride, unsigned long long *counters, FamodChain *chain_state)
{
  __shared__ double exptab[kExpTableSize];
  __shared__ double p[FP_SIZE];
  load_exp_table(exptab, h.exptab);
  __syncthreads();
  h.exptab = exptab;
  WarpReducer red;
  FamodChain chain = *chain_state;          // carried across passes
  for (int64_t i = 0; i < count; i++) {
    Cell c = load_cell(surf, begin + i, fl.include_baryon != 0);
    int iterations;
    int status = famod_setup_cell(red, c, fl, h, &chain, p, &iterations);
    famod_store(p, status, iterations, i, begin, pack, stride, counters);
  }
  if ((threadIdx.x & 31) == 0) *chain_state = chain;
}

}  // namespace

is3d_status famod_setup_pass(is3d_ctx *ctx, int64_t begin, int64_t count, double *pack, int64_t stride, unsigned long long *counters,
                             int64_t *launches, bool sampler_rules)
{
  const is3d_params &p = ctx->prm;
  FamodFlags fl;
  fl.dimension = p.dimension; fl.include_baryon = p.include_baryon; fl.include_shear = p.include_shear_deltaf;
  fl.include_baryondiff = p.include_baryondiff_deltaf; fl.deta_min = p.deta_min;
  fl.sampler = sampler_rules ? 1 : 0;
  void *gl = nullptr, *chain = nullptr;
  IS3D_TRY(ctx->get_scratch("gl16", 96 * sizeof(double), &gl));
  IS3D_TRY(ctx->get_scratch("famod_chain_state", sizeof(FamodChain), &chain));
  if (begin == 0) {
    double t[96];
    fill_gl16_table(t);
    IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(gl, t, sizeof(t), cudaMemcpyHostToDevice, ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaMemsetAsync(chain, 0, sizeof(FamodChain), ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  AnisoHadrons h;
  h.mass = ctx->d_pdg_mass; h.sign = ctx->d_pdg_sign; h.deg = ctx->d_pdg_deg;
  h.n = ctx->npdg < kAnisoMaxHadrons ? ctx->npdg : kAnisoMaxHadrons;
  h.gl16 = (const double *)gl;
  h.exptab = ctx->d_exptab;
  h.exact = p.famod_chain ? 1 : 0;          // the chain-faithful parity mode also keeps the reference's libm formulas
  if (!h.exact) {
    // Production mode: hadrons with the same (mass, sign) contribute identical terms to every F / J sum (no chemical
    // potential enters them), so the first min(320, N_pdg) table entries are merged into classes with summed degeneracy --
    // for the SMASH table 320 -> ~130 terms per quadrature node.  Same sums up to the order of the additions.
    std::vector<double> cm, cs, cd;
    for (int n = 0; n < h.n; n++) {
      const double m = ctx->h_pdg_mass[n], sg = ctx->h_pdg_sign[n], g = ctx->h_pdg_deg[n];
      if (m == 0.0) continue;                                       // photons are skipped by the sums anyway
      size_t k = 0;
      while (k < cm.size() && !(cm[k] == m && cs[k] == sg)) k++;
      if (k == cm.size()) { cm.push_back(m); cs.push_back(sg); cd.push_back(g); }
      else cd[k] += g;
    }
    const size_t nc = cm.size();
    void *d = nullptr;
    IS3D_TRY(ctx->get_scratch("aniso_merged_hadrons", 3 * (nc ? nc : 1) * sizeof(double), &d));
    if (begin == 0 && nc) {
      std::vector<double> hbuf(3 * nc);
      for (size_t k = 0; k < nc; k++) { hbuf[k] = cm[k]; hbuf[nc + k] = cs[k]; hbuf[2 * nc + k] = cd[k]; }
      IS3D_CUDA_TRY(ctx, cudaMemcpyAsync(d, hbuf.data(), hbuf.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      IS3D_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    h.mass = (const double *)d; h.sign = h.mass + nc; h.deg = h.mass + 2 * nc;
    h.n = (int)nc;
  }
  if (p.famod_chain) {
    famod_setup_chain_kernel<<<1, 32, 0, ctx->stream>>>(ctx->surf, begin, count, fl, h, pack, stride, counters, (FamodChain *)chain);
  } else {
    int64_t warps = count;
    int64_t max_warps = (int64_t)ctx->sm_count * 16 * 4;     // a few waves of 4-warp blocks (cells are strided over the warps)
    if (warps > max_warps) warps = max_warps;
    unsigned blocks = (unsigned)((warps + 3) / 4);
    famod_setup_free_kernel<<<blocks, 128, 0, ctx->stream>>>(ctx->surf, begin, count, fl, h, pack, stride, counters);
  }
  IS3D_CUDA_TRY(ctx, cudaGetLastError());
  (*launches)++;
  return IS3D_OK;
}

}  // namespace is3d
