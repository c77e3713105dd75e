// kernels_pair.cu — instantiations of the two-sets-per-warp kernels (pair_kernel.cuh) and their launcher.
#include <mutex>

#include "launch.h"
#include "pair_kernel.cuh"

namespace gab1 {
namespace {
template <int K, int MODE, bool MIRROR>
int launch_pair(const KernelArgs& args, int device, cudaStream_t stream) {
  static std::mutex mu;
  static int blocks_per_sm[64] = {0};
  static int sms[64] = {0};
  constexpr int TPB = 32 * kPairWarpsPerCta;
  const size_t smem = (size_t)kPairWarpsPerCta * (2 * WS_HDR + 2 * (size_t)args.P_pad) * sizeof(double);
  auto kern = solve_pair_kernel<K, MODE, MIRROR>;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (device < 64 && blocks_per_sm[device] == 0) {
      int n = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, TPB, smem));
      if (n < 1) return fail(-5, "pair kernel does not fit on an SM (K=%d, smem=%zu)", K, smem);
      blocks_per_sm[device] = n;
      CUDA_TRY(cudaDeviceGetAttribute(&sms[device], cudaDevAttrMultiProcessorCount, device));
    }
  }
  int nb = 0, nsm = 0;
  if (device < 64) { nb = blocks_per_sm[device]; nsm = sms[device]; }
  if (nb == 0) {
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, TPB, smem));
    CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
  }
  // persistent grid: a multiple of the SM count; never more warps than pairs of sets
  long long grid = (long long)nsm * nb;
  const long long need = ((args.S + 1) / 2 + kPairWarpsPerCta - 1) / kPairWarpsPerCta;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, TPB, smem, stream>>>(args);
  count_launch();
  CUDA_TRY(cudaGetLastError());
  return 0;
}


}  // namespace

int launch_pair_kernel(int K, int mode, bool mirror, const KernelArgs& a, int device, cudaStream_t stream) {
#define GAB1_LAUNCH_PAIR(KK)                                                                             \
  case KK:                                                                                               \
    return mode == MODE_FAST_WHILE                                                                       \
               ? (mirror ? launch_pair<KK, MODE_FAST_WHILE, true>(a, device, stream)                     \
                         : launch_pair<KK, MODE_FAST_WHILE, false>(a, device, stream))                   \
               : (mirror ? launch_pair<KK, MODE_FAST_FOR, true>(a, device, stream)                       \
                         : launch_pair<KK, MODE_FAST_FOR, false>(a, device, stream));
  switch (K) {
    GAB1_LAUNCH_PAIR(1)
    GAB1_LAUNCH_PAIR(2)
    GAB1_LAUNCH_PAIR(4)
  }
#undef GAB1_LAUNCH_PAIR
  return fail(-6, "no pair kernel for K=%d", K);
}

}  // namespace gab1
