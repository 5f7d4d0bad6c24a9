// DFMA micro-benchmark: the measured FP64 roofline denominator (MEASURED_PEAKS.json has no FP64 entry).
#include "ctx.h"

namespace is3d {

namespace {
constexpr int kChains = 8;
constexpr int kIters = 4096;

__global__ void __launch_bounds__(256) dfma_kernel(double *out, double a, double b)
{
  double x[kChains];
#pragma unroll
  for (int c = 0; c < kChains; c++) x[c] = (double)(threadIdx.x + c) * 1e-3;
#pragma unroll 1
  for (int i = 0; i < kIters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int c = 0; c < kChains; c++) x[c] = fma(x[c], a, b);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < kChains; c++) s += x[c];
  if (s == 123.456) out[0] = s;   // never true; keeps the chains alive
}
}  // namespace

is3d_status measure_fp64_peak(is3d_ctx *ctx, double *tflops)
{
  void *d = nullptr;
  IS3D_TRY(ctx->get_scratch("fp64_peak", 64, &d));
  const int blocks = ctx->sm_count * 8, threads = 256;
  cudaEvent_t e0, e1;
  IS3D_CUDA_TRY(ctx, cudaEventCreate(&e0));
  IS3D_CUDA_TRY(ctx, cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; rep++) {
    IS3D_CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
    dfma_kernel<<<blocks, threads, 0, ctx->stream>>>((double *)d, 0.999999, 1e-7);
    IS3D_CUDA_TRY(ctx, cudaGetLastError());
    IS3D_CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
    IS3D_CUDA_TRY(ctx, cudaEventSynchronize(e1));
    float ms = 0.f;
    IS3D_CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * (double)blocks * threads * kIters * 4 * kChains;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep >= 1 && tf > best) best = tf;   // first repetition is warm-up
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  return IS3D_OK;
}

}  // namespace is3d
