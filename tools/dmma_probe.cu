// Dev probe: does the FP64 tensor-core MMA (mma.sync m8n8k4 / m16n8k4 f64) run BESIDE the FP64 vector pipe on B200, or does it
// share it?  K1's low-rank pieces (u.p/T, p.dsigma and the delta-f quadratic forms are rank 2-4 products of cell items and
// (class, pT) columns) could move to DMMA only if it is extra throughput.  Prints TFLOP/s of DFMA alone, DMMA alone, and both
// in one instruction stream.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double &d0, double &d1, double &d2, double &d3, double a0, double a1, double b)
{
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3) : "d"(a0), "d"(a1), "d"(b));
}

// NF DFMA (independent chains) and NM m8n8k4 DMMAs (independent accumulators) per inner step
template <int NF, int NM, int BIG>
__global__ void probe_kernel(double *out, double a, double b, int iters)
{
  double x[NF > 0 ? NF : 1];
  double c[NM > 0 ? 4 * NM : 1];
#pragma unroll
  for (int k = 0; k < NF; k++) x[k] = (threadIdx.x + k) * 1e-3;
#pragma unroll
  for (int k = 0; k < 4 * NM; k++) c[k] = (threadIdx.x + k) * 1e-4;
  const double fa = 1e-3 * (threadIdx.x & 7), fb = 1e-3 * (threadIdx.x & 3);
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int k = 0; k < NM; k++) {
        if (BIG) dmma1684(c[4 * k], c[4 * k + 1], c[4 * k + 2], c[4 * k + 3], fa, fb, fa);
        else dmma884(c[4 * k], c[4 * k + 1], fa, fb);
      }
#pragma unroll
      for (int k = 0; k < NF; k++) x[k] = fma(x[k], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < NF; k++) s += x[k];
#pragma unroll
  for (int k = 0; k < 4 * NM; k++) s += c[k];
  if (s == 123.456) out[0] = s;
}

template <int NF, int NM, int BIG>
void run(int warps_per_sched, int sms)
{
  double *d;
  cudaMalloc(&d, 64);
  const int iters = 512, threads = 128, blocks = sms * warps_per_sched;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    probe_kernel<NF, NM, BIG><<<blocks, threads>>>(d, 0.999999, 1e-7, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  cudaError_t err = cudaGetLastError();
  cudaFree(d);
  const double steps = (double)blocks * (threads / 32) * iters * 8;                 // warp-level inner steps
  const double fma_flops = steps * NF * 32 * 2, mma_flops = steps * NM * (BIG ? 16 * 8 * 4 : 8 * 8 * 4) * 2;
  const double ns_per_step = best * 1e6 / (iters * 8.0);
  printf("  DFMA/step %2d  DMMA(%s)/step %d  w/sched %d : %8.3f ms  %6.1f ns/step  DFMA %6.2f + DMMA %6.2f = %6.2f TFLOP/s  %s\n", NF, BIG ? "m16n8k4" : "m8n8k4",
         NM, warps_per_sched, best, ns_per_step, fma_flops / (best * 1e-3) / 1e12, mma_flops / (best * 1e-3) / 1e12,
         (fma_flops + mma_flops) / (best * 1e-3) / 1e12, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, sms);
  for (int w = 2; w <= 8; w *= 2) {
    run<8, 0, 0>(w, sms);
    run<0, 1, 0>(w, sms); run<0, 2, 0>(w, sms); run<0, 4, 0>(w, sms);
    run<0, 1, 1>(w, sms); run<0, 2, 1>(w, sms); run<0, 4, 1>(w, sms);
    run<8, 1, 0>(w, sms); run<8, 2, 0>(w, sms); run<8, 4, 0>(w, sms);
    run<8, 1, 1>(w, sms); run<8, 2, 1>(w, sms);
  }
  return 0;
}
