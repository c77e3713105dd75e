// fp64_cost_probe.cu — scheduler cycles per FP64 warp instruction by operand pattern (diagnostic, not product code).
// One warp per scheduler (128 threads per CTA, one CTA per SM) and two warps per scheduler; 16 independent chains.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(128) probe(double* out, const double* in, int iters, double cpar) {
  double x[16], y[16], z[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { x[i] = in[i] + threadIdx.x * 1e-9; y[i] = in[16 + i]; z[i] = in[32 + i]; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) x[i] = fma(y[i], z[i], x[i]);            // three distinct registers
      if (MODE == 1) x[i] = fma(y[0], z[i], x[i]);            // one operand shared by consecutive instructions
      if (MODE == 2) x[i] = fma(cpar, z[i], x[i]);            // one operand from the constant bank
      if (MODE == 3) x[i] = fma(x[i], y[0], z[0]);            // the peak kernel's pattern: two shared operands
      if (MODE == 4) x[i] = x[i] + y[i];                      // DADD, two registers
      if (MODE == 5) x[i] = x[i] * y[i];                      // DMUL, two registers
      if (MODE == 6) x[i] = fma(x[i], -2.0, z[i]);            // immediate operand
      if (MODE == 7) x[i] = x[i] + y[0];                      // DADD with a shared operand
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(double* out, const double* in, int sms, const char* what) {
  const int iters = 20000;
  for (int ctas = 1; ctas <= 2; ++ctas) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<sms * ctas, 128>>>(out, in, 100, 1.0000001);
    cudaEventRecord(e0);
    probe<MODE><<<sms * ctas, 128>>>(out, in, iters, 1.0000001);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9 / ((double)iters * 16 * ctas);
    printf("%-52s %d warp(s)/scheduler: %.2f cycles per instruction per scheduler\n", what, ctas, cyc);
  }
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double h[48];
  for (int i = 0; i < 16; ++i) { h[i] = 1.0 + i * 1e-3; h[16 + i] = 1.0 + 1e-9 * (i + 1); h[32 + i] = 1e-9 * (i + 1); }
  double *in, *out; cudaMalloc(&in, sizeof h); cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
  cudaMalloc(&out, sizeof(double) * sms * 2 * 128);
  run<0>(out, in, sms, "DFMA three distinct registers");
  run<1>(out, in, sms, "DFMA one operand shared by neighbours");
  run<2>(out, in, sms, "DFMA one operand from the constant bank");
  run<3>(out, in, sms, "DFMA two shared operands (peak-kernel pattern)");
  run<6>(out, in, sms, "DFMA immediate operand");
  run<4>(out, in, sms, "DADD two registers");
  run<7>(out, in, sms, "DADD one shared operand");
  run<5>(out, in, sms, "DMUL two registers");
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
