// Dev probe: DFMA issue rate of one B200 as a function of independent chains per warp (ILP) and resident warps per
// scheduler (TLP).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_ilp_probe tools/fp64_ilp_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_chains(double *out, double a, double b, int iters)
{
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) x[c] = (threadIdx.x + c) * 1e-3;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int c = 0; c < CHAINS; c++) x[c] = fma(x[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += x[c];
  if (s == 123.456) out[0] = s;
}

template <int CHAINS>
double run(int warps_per_sm, int sms)
{
  double *d;
  cudaMalloc(&d, 64);
  const int iters = 2048;
  const int threads = 128;                 // 4 warps per block, one per scheduler
  const int blocks = sms * warps_per_sm / 4;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    dfma_chains<CHAINS><<<blocks, threads>>>(d, 0.999999, 1e-7, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  cudaFree(d);
  double flops = 2.0 * blocks * threads * (double)iters * 8 * CHAINS;
  return flops / (best * 1e-3) / 1e12;
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("%s, %d SMs; TFLOP/s (DFMA = 2 flops) for chains-per-warp x warps-per-scheduler\n", p.name, sms);
  printf("chains  w/s=1    w/s=2    w/s=4    w/s=8    w/s=16\n");
#define ROW(C) printf("%5d  %7.2f  %7.2f  %7.2f  %7.2f  %7.2f\n", C, run<C>(4, sms), run<C>(8, sms), run<C>(16, sms), run<C>(32, sms), run<C>(64, sms));
  ROW(1) ROW(2) ROW(3) ROW(4) ROW(6) ROW(8)
  return 0;
}
