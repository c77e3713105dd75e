// kernels_duo.cu — instantiations of the one-set-per-warp kernel with the two-warps-per-set latency lane in front
// (duo_kernel.cuh) and its launcher.
#include <stdlib.h>

#include <mutex>

#include "duo_kernel.cuh"
#include "launch.h"

namespace gab1 {
namespace {
template <int K, int MODE>
int launch(const KernelArgs& args, int device, cudaStream_t stream) {
  static std::mutex mu;
  static int blocks_per_sm[64] = {0};
  static int sms[64] = {0};
  const size_t smem = (4 * (WS_HDR + 2 * (size_t)args.P_pad + WS_EX) + 2 * DUO_DX) * sizeof(double);
  auto kern = duo_solve_kernel<K, MODE>;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (device < 64 && blocks_per_sm[device] == 0) {
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      int n = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, 128, smem));
      if (n < 1) return fail(-5, "duo kernel does not fit on an SM (K=%d, smem=%zu)", K, smem);
      blocks_per_sm[device] = n;
      CUDA_TRY(cudaDeviceGetAttribute(&sms[device], cudaDevAttrMultiProcessorCount, device));
    }
  }
  int nb = 0, nsm = 0;
  if (device < 64) { nb = blocks_per_sm[device]; nsm = sms[device]; }
  if (nb == 0) {
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 128, smem));
    CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
  }
  // persistent grid: a multiple of the SM count; never more pairs than sets
  long long grid = (long long)nsm * nb;
  const long long need = (args.S + 1) / 2;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, 128, smem, stream>>>(args);
  count_launch();
  CUDA_TRY(cudaGetLastError());
  return 0;
}
}  // namespace

int launch_duo_kernel(int K, int mode, const KernelArgs& a, int device, cudaStream_t stream) {
  if (mode == MODE_STRICT) return fail(-6, "the duo kernel runs the fast arithmetic");
#define GAB1_LAUNCH(KK)                                                                              \
  case KK:                                                                                           \
    return mode == MODE_FAST_WHILE ? launch<KK, MODE_FAST_WHILE>(a, device, stream)        \
                                   : launch<KK, MODE_FAST_FOR>(a, device, stream);
  switch (K) {
    GAB1_LAUNCH(2)
    GAB1_LAUNCH(4)
  }
#undef GAB1_LAUNCH
  return fail(-6, "no duo kernel for K=%d", K);
}

}  // namespace gab1
