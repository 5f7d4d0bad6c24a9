// Dev probe: does non-FP64 work issued between DFMAs cost FP64 throughput on B200?  Each thread runs CHAINS independent
// DFMA chains; per DFMA it also issues MIX/8 integer (IMAD) instructions on independent integer chains.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_mix_probe tools/fp64_mix_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, int MIX>   // MIX = integer instructions per 8 DFMAs
__global__ void mix_kernel(double *out, double a, double b, int m, int iters)
{
  double x[CHAINS];
  int y[8];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) x[c] = (threadIdx.x + c) * 1e-3;
#pragma unroll
  for (int c = 0; c < 8; c++) y[c] = threadIdx.x + c;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int c = 0; c < CHAINS; c++) x[c] = fma(x[c], a, b);
#pragma unroll
      for (int k = 0; k < (MIX * CHAINS) / 8; k++) y[k & 7] = y[k & 7] * m + u;     // IMAD on independent chains
    }
  }
  double s = 0;
  int t = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += x[c];
#pragma unroll
  for (int c = 0; c < 8; c++) t += y[c];
  if (s == 123.456 || t == 0x7fffffff) out[0] = s + t;
}

template <int CHAINS, int MIX>
double run(int warps_per_sched, int sms)
{
  double *d;
  cudaMalloc(&d, 64);
  const int iters = 1024, threads = 128, blocks = sms * warps_per_sched;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    mix_kernel<CHAINS, MIX><<<blocks, threads>>>(d, 0.999999, 1e-7, 3, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  cudaFree(d);
  return 2.0 * blocks * threads * (double)iters * 8 * CHAINS / (best * 1e-3) / 1e12;
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("%s: DFMA TFLOP/s with 3 chains x 4 warps/scheduler (K1's shape) and 8 x 8, as integer instructions are mixed in\n", p.name);
  printf("int per DFMA   3x4      8x8\n");
#define ROW(M) printf("   %5.3f    %7.2f  %7.2f\n", M / 8.0, run<3, M>(4, sms), run<8, M>(8, sms));
  ROW(0) ROW(2) ROW(4) ROW(6) ROW(8) ROW(12) ROW(16)
  return 0;
}
