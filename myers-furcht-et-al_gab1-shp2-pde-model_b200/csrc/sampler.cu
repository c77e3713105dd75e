// sampler.cu — synthetic prior ensembles generated on the device.
//
// Stands in for the prior half of generate_ensemble (get_param_posteriors.jl:38-86): per parameter set, independent
// log-normal draws for the 7 diffusivities and kG1p, kG1dp, kSa, kSi, kp, kdp (:60-72) and (Kd, k_r) pairs for the
// binding reactions with k_f = k_r / Kd (:75-76); (mu, sigma) from get_param_priors.jl:19-198.  The reference's own
// random stream (Julia's default RNG) cannot be reproduced outside Julia, so the library defines one that any host can
// restate: Philox4x32-10, counter = (set index lo, hi, draw pair, 0), key = seed; two 53-bit uniforms per call;
// Box-Muller.  One thread per parameter set, 11 Philox calls, D and k written straight into the buffers the solver
// kernels read — a 10^5-set sweep then starts without any host-to-device copy of parameters.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gab1pde.h"
#include "launch.h"

namespace gab1 {
namespace {

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

struct PriorTable { double mu[GAB1_N_PRIOR_NORMALS], sigma[GAB1_N_PRIOR_NORMALS]; };

__global__ void sample_prior_kernel(long long S, unsigned long long seed, PriorTable t, double EGF, double Kdd, double* D, double* k) {
  const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (s >= S) return;
  double v[GAB1_N_PRIOR_NORMALS];
#pragma unroll
  for (int j = 0; j < GAB1_N_PRIOR_NORMALS / 2; ++j) {
    uint32_t c[4] = {(uint32_t)s, (uint32_t)((unsigned long long)s >> 32), (uint32_t)j, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u1 = ((double)(((unsigned long long)c[0] << 21) | (c[1] >> 11)) + 0.5) * 0x1p-53;
    const double u2 = ((double)(((unsigned long long)c[2] << 21) | (c[3] >> 11)) + 0.5) * 0x1p-53;
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincos(2.0 * 3.141592653589793 * u2, &sn, &cs);
    v[2 * j] = exp(fma(t.sigma[2 * j], rad * cs, t.mu[2 * j]));
    v[2 * j + 1] = exp(fma(t.sigma[2 * j + 1], rad * sn, t.mu[2 * j + 1]));
  }
  double* Ds = D + s * GAB1_N_D;
#pragma unroll
  for (int i = 0; i < GAB1_N_D; ++i) Ds[i] = v[i];
  // v[7..21] = Kd_S2, kS2r, Kd_G2, kG2r, kG1f, kG1r, kEGFf, kEGFr, kdf, kG1p, kG1dp, kSa, kSi, kp, kdp
  double* ks = k + s * GAB1_N_K;
  ks[0] = v[8] / v[7];  ks[1] = v[8];                  // kS2f = kS2r / Kd, kS2r       (get_param_posteriors.jl:75)
  ks[2] = v[11];        ks[3] = v[12];                 // kG1f, kG1r
  ks[4] = v[10] / v[9]; ks[5] = v[10];                 // kG2f = kG2r / Kd, kG2r
  ks[6] = v[16]; ks[7] = v[17]; ks[8] = v[18]; ks[9] = v[19];     // kG1p, kG1dp, kSa, kSi
  ks[10] = v[20]; ks[11] = v[21];                      // kp, kdp
  ks[12] = v[13]; ks[13] = v[14];                      // kEGFf, kEGFr
  ks[14] = EGF;                                        // get_param_priors.jl:14
  ks[15] = v[15]; ks[16] = v[15] * Kdd;                // kdf, kdr = kdf * Kdd
}

}  // namespace

int sample_prior_device(int device, cudaStream_t stream, long long S, unsigned long long seed, const double* mu,
                        const double* sigma, double EGF, double Kdd, double* D, double* k) {
  if (S < 0 || !mu || !sigma || !D || !k) return fail(-2, "bad arguments to gab1_sample_prior");
  if (S == 0) return 0;
  CUDA_TRY(cudaSetDevice(device));
  PriorTable t;
  for (int i = 0; i < GAB1_N_PRIOR_NORMALS; ++i) { t.mu[i] = mu[i]; t.sigma[i] = sigma[i]; }
  const int tb = 128;
  sample_prior_kernel<<<(unsigned)((S + tb - 1) / tb), tb, 0, stream>>>(S, seed, t, EGF, Kdd, D, k);
  count_launch();
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace gab1
