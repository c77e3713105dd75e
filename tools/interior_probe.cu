// interior_probe.cu — how fast can the FP64 pipe run the interior update alone?  (diagnostic, not product code)
// Variants: K nodes per lane, with/without the halo shuffles, 1..4 CTAs of 128 threads per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/interior_probe tools/interior_probe.cu && tools/interior_probe
#include <cstdio>
#include <cuda_runtime.h>
constexpr unsigned FULL = 0xffffffffu;
__device__ __forceinline__ double shfl_up1(double x) {
  const int lo = __shfl_up_sync(FULL, __double2loint(x), 1), hi = __shfl_up_sync(FULL, __double2hiint(x), 1);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_down1(double x) {
  const int lo = __shfl_down_sync(FULL, __double2loint(x), 1), hi = __shfl_down_sync(FULL, __double2hiint(x), 1);
  return __hiloint2double(hi, lo);
}
enum { iSFK, aSFK, GAB1, pGAB1, GRB2, G2G1, G2PG1, SHP2, PG1S, G2PG1S, NCY };

template <int K, bool SHUF, int MINB>
__global__ void __launch_bounds__(128, MINB) probe(const double* __restrict__ par, double* out, int steps) {
  const int lane = threadIdx.x & 31;
  double u[NCY][K];
#pragma unroll
  for (int q = 0; q < NCY; ++q)
#pragma unroll
    for (int i = 0; i < K; ++i) u[q][i] = par[q] * (1.0 + 1e-3 * (lane * K + i));
  const double kS2f_t = par[10], kS2r_t = par[11], kG1f_t = par[12], kG1r_t = par[13], kG1p_t = par[14], kG1dp_t = par[15], kSi_t = par[16];
  const double l_Si = par[17], l_Sa = par[18], l_G1 = par[19], l_G2 = par[20], l_G2G1 = par[21], l_S2 = par[22], l_G1S2 = par[23], l_G2G1S2 = par[24];
  const double c_Si = 1 - 2 * l_Si, c_Sa = 1 - 2 * l_Sa - kSi_t, c_G1 = 1 - 2 * l_G1, c_G2 = 1 - 2 * l_G2, c_G2G1 = 1 - 2 * l_G2G1,
               c_S2 = 1 - 2 * l_S2, c_G1S2 = 1 - 2 * l_G1S2, c_G2G1S2 = 1 - 2 * l_G2G1S2;
  double cp[K], cm[K];
#pragma unroll
  for (int i = 0; i < K; ++i) { cp[i] = 1.0 + 1.0 / (lane * K + i + 1); cm[i] = 1.0 - 1.0 / (lane * K + i + 1); }
  for (int s = 0; s < steps; ++s) {
    double hr[NCY], carry[NCY];
#pragma unroll
    for (int q = 0; q < NCY; ++q) {
      carry[q] = cm[0] * (SHUF ? shfl_up1(u[q][K - 1]) : u[q][K - 1]);
      hr[q] = SHUF ? shfl_down1(u[q][0]) : u[q][0];
    }
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const double Si = u[iSFK][i], Sa = u[aSFK][i], G1 = u[GAB1][i], pG1 = u[pGAB1][i], G2 = u[GRB2][i],
                   g2g1 = u[G2G1][i], g2pg1 = u[G2PG1][i], S2 = u[SHP2][i], pg1s = u[PG1S][i], g2pg1s = u[G2PG1S][i];
      const double gb = kG1f_t * G2, ph = kG1p_t * Sa, sbd = kS2f_t * S2;
      const double v1 = fma(gb, G1, -(kG1r_t * g2g1));
      const double v3 = fma(gb, pG1, -(kG1r_t * g2pg1));
      const double v5 = fma(gb, pg1s, -(kG1r_t * g2pg1s));
      const double v2 = fma(ph, G1, -(kG1dp_t * pG1));
      const double v6 = fma(ph, g2g1, -(kG1dp_t * g2pg1));
      const double v4 = fma(sbd, pG1, -(kS2r_t * pg1s));
      const double v7 = fma(sbd, g2pg1, -(kS2r_t * g2pg1s));
      double ks[NCY];
      ks[iSFK] = fma(c_Si, Si, kSi_t * Sa);
      ks[aSFK] = c_Sa * Sa;
      ks[GAB1] = fma(c_G1, G1, -(v1 + v2));
      ks[pGAB1] = fma(c_G1, pG1, (v2 - v3) - v4);
      ks[GRB2] = fma(c_G2, G2, -((v1 + v3) + v5));
      ks[G2G1] = fma(c_G2G1, g2g1, v1 - v6);
      ks[G2PG1] = fma(c_G2G1, g2pg1, (v3 + v6) - v7);
      ks[SHP2] = fma(c_S2, S2, -(v4 + v7));
      ks[PG1S] = fma(c_G1S2, pg1s, v4 - v5);
      ks[G2PG1S] = fma(c_G2G1S2, g2pg1s, v5 + v7);
      const double lam[NCY] = {l_Si, l_Sa, l_G1, l_G1, l_G2, l_G2G1, l_G2G1, l_S2, l_G1S2, l_G2G1S2};
#pragma unroll
      for (int q = 0; q < NCY; ++q) {
        const double up = i + 1 < K ? u[q][i + 1] : hr[q];
        const double nb = fma(cp[i], up, carry[q]);
        if (i + 1 < K) carry[q] = cm[i + 1] * u[q][i];
        u[q][i] = fma(lam[q], nb, ks[q]);
      }
    }
  }
  double acc = 0;
#pragma unroll
  for (int q = 0; q < NCY; ++q)
#pragma unroll
    for (int i = 0; i < K; ++i) acc += u[q][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int K, bool SHUF, int MINB>
void run(const double* dpar, double* dout, int sms) {
  const int steps = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<K, SHUF, MINB><<<sms * MINB, 128>>>(dpar, dout, 100);
  cudaEventRecord(e0);
  probe<K, SHUF, MINB><<<sms * MINB, 128>>>(dpar, dout, steps);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double fp64_per_step = 69.0 * K;                       // warp instructions
  const double warps = (double)sms * MINB * 4;
  const double instr_per_s = fp64_per_step * steps * warps / (ms * 1e-3);
  const double peak = (double)sms * 4 * 0.5 * 1.965e9;         // one FP64 warp instruction per 2 cycles per scheduler
  const double cyc_per_warp_step = ms * 1e-3 * 1.965e9 / steps;
  printf("K=%d shuf=%d ctas/SM=%d: %.2f ms, %.1f cycles per warp-step, FP64 pipe %.1f%%\n", K, (int)SHUF, MINB, ms,
         cyc_per_warp_step, 100 * instr_per_s / peak);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double h[32];
  for (int i = 0; i < 10; ++i) h[i] = 100.0 + i;
  for (int i = 10; i < 17; ++i) h[i] = 1e-6 * (i - 8);
  for (int i = 17; i < 25; ++i) h[i] = 0.1 + 0.01 * i;
  double *dpar, *dout;
  cudaMalloc(&dpar, sizeof h); cudaMemcpy(dpar, h, sizeof h, cudaMemcpyHostToDevice);
  cudaMalloc(&dout, sizeof(double) * sms * 4 * 128);
  run<2, false, 1>(dpar, dout, sms); run<2, false, 2>(dpar, dout, sms); run<2, false, 3>(dpar, dout, sms);
  run<2, true, 1>(dpar, dout, sms);  run<2, true, 2>(dpar, dout, sms);  run<2, true, 3>(dpar, dout, sms);
  run<4, false, 1>(dpar, dout, sms); run<4, false, 2>(dpar, dout, sms);
  run<4, true, 1>(dpar, dout, sms);  run<4, true, 2>(dpar, dout, sms);
  run<1, true, 2>(dpar, dout, sms);  run<1, true, 4>(dpar, dout, sms);
  return 0;
}
