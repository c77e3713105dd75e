// dfma_probe.cu — how many warps and how much ILP does the B200 FP64 pipe need?  (design probe, not product)
// Prints DFMA per clock per SM for warps/SMSP in {1,2,3,4,6,8,16} x independent chains per thread in {1,2,4,8}.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* sink, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x * 1e-9 + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 32 / ILP; ++r)
#pragma unroll
      for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += x[j];
  if (s == 123.456) sink[0] = s;
}
template <int ILP>
double run(int warps_per_sm, int sms, double* sink, double ghz) {
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP><<<sms, warps_per_sm * 32>>>(sink, 100, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  k<ILP><<<sms, warps_per_sm * 32>>>(sink, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fma_per_sm = 32.0 * iters * warps_per_sm * 32;   // thread-level FMAs per SM
  return fma_per_sm / (ms * 1e-3 * ghz * 1e9);
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  double* sink; cudaMalloc(&sink, 8);
  printf("SMs %d, clock %.3f GHz; DFMA thread-ops per clock per SM (peak would be 64)\n", sms, ghz);
  printf("warps/SMSP   ILP1    ILP2    ILP4    ILP8\n");
  for (int w : {1, 2, 3, 4, 6, 8, 16}) {
    printf("%9d  %6.1f  %6.1f  %6.1f  %6.1f\n", w, run<1>(4 * w, sms, sink, ghz), run<2>(4 * w, sms, sink, ghz),
           run<4>(4 * w, sms, sink, ghz), run<8>(4 * w, sms, sink, ghz));
  }
  return 0;
}
