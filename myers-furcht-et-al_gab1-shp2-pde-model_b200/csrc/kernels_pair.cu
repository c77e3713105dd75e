// kernels_pair.cu — instantiations of the skewed fast kernels (pair_kernel.cuh) and their launchers; compiled twice:
// -DGAB1_GROUP_HW=16 (two sets per warp) and -DGAB1_GROUP_HW=32 (one set per warp).
#include <mutex>

#include "launch.h"
#include "pair_kernel.cuh"

namespace gab1 {
namespace {
template <int K, int MODE, bool MIRROR, int HW, bool SKEW, bool TOKEN>
int launch_pair(const KernelArgs& args, int device, cudaStream_t stream) {
  static std::mutex mu;
  static int blocks_per_sm[64] = {0};
  static int sms[64] = {0};
  constexpr int WPC = Shape<K, HW, SKEW, TOKEN>::warps;
  constexpr int TPB = 32 * WPC;
  const size_t smem = (size_t)WPC * ((32 / HW) * WS_HDR + 2 * (size_t)args.P_pad) * sizeof(double) + 64;   // + token flags
  auto kern = solve_pair_kernel<K, MODE, MIRROR, HW, SKEW, TOKEN>;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (device < 64 && blocks_per_sm[device] == 0) {
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
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
  const long long need = ((args.S + 32 / HW - 1) / (32 / HW) + WPC - 1) / WPC;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, TPB, smem, stream>>>(args);
  count_launch();
  CUDA_TRY(cudaGetLastError());
  return 0;
}


}  // namespace

#ifndef GAB1_GROUP_HW
#define GAB1_GROUP_HW 16
#endif

// variant: 0 = skewed loop, 1 = plain loop, 2 = plain loop with the interior token
#define GAB1_LAUNCH_V(KK, SKEW, TOKEN)                                                                            \
  (mode == MODE_FAST_WHILE                                                                                        \
       ? (mirror ? launch_pair<KK, MODE_FAST_WHILE, true, GAB1_GROUP_HW, SKEW, TOKEN>(a, device, stream)          \
                 : launch_pair<KK, MODE_FAST_WHILE, false, GAB1_GROUP_HW, SKEW, TOKEN>(a, device, stream))        \
       : (mirror ? launch_pair<KK, MODE_FAST_FOR, true, GAB1_GROUP_HW, SKEW, TOKEN>(a, device, stream)            \
                 : launch_pair<KK, MODE_FAST_FOR, false, GAB1_GROUP_HW, SKEW, TOKEN>(a, device, stream)))

#if GAB1_GROUP_HW == 16
// two parameter sets per warp: K nodes per lane of a 16-lane group
int launch_group16_kernel(int K, int variant, int mode, bool mirror, const KernelArgs& a, int device, cudaStream_t stream) {
  switch (K) {
    case 1: return GAB1_LAUNCH_V(1, true, false);
    case 2: return GAB1_LAUNCH_V(2, true, false);
#ifdef GAB1_BUILD_AB_VARIANTS
    case 4: return variant == 1 ? GAB1_LAUNCH_V(4, false, false) : GAB1_LAUNCH_V(4, true, false);
#else
    case 4: if (variant == 0) return GAB1_LAUNCH_V(4, true, false); break;
#endif
  }
  if (variant != 0) return fail(-6, "the plain / token loop variants are A/B builds (-DGAB1_BUILD_AB_VARIANTS)");
  return fail(-6, "no 16-lane kernel for K=%d", K);
}
#else
// one parameter set per warp: K nodes per lane
int launch_group32_kernel(int K, int variant, int mode, bool mirror, const KernelArgs& a, int device, cudaStream_t stream) {
  switch (K) {
#ifdef GAB1_BUILD_AB_VARIANTS
    case 2: return variant == 2 ? GAB1_LAUNCH_V(2, false, true) : variant == 1 ? GAB1_LAUNCH_V(2, false, false) : GAB1_LAUNCH_V(2, true, false);
#else
    case 2: if (variant == 0) return GAB1_LAUNCH_V(2, true, false); break;
#endif
    case 4: return GAB1_LAUNCH_V(4, true, false);
    case 8: return GAB1_LAUNCH_V(8, true, false);
  }
  if (variant != 0) return fail(-6, "the plain / token loop variants are A/B builds (-DGAB1_BUILD_AB_VARIANTS)");
  return fail(-6, "no 32-lane kernel for K=%d", K);
}
#endif
#undef GAB1_LAUNCH_V

}  // namespace gab1
