// kernels_gang.cu — instantiations of the several-sets-per-warp kernels (gang_kernel.cuh) and their launcher.
// Compiled once per gang width: -DGAB1_GANG_G=2|4|8|16|32 (so that the widths build in parallel).
#include <mutex>

#include "gang_kernel.cuh"
#include "launch.h"

#ifndef GAB1_GANG_G
#define GAB1_GANG_G 8
#endif

namespace gab1 {
namespace {
template <int G, int KN, int MODE>
int launch(const KernelArgs& args, int device, cudaStream_t stream) {
  static std::mutex mu;
  static int blocks_per_sm[64] = {0};
  static int sms[64] = {0};
  using L = GangLayout<G, KN>;
  const size_t smem = (size_t)L::doubles(args.P_pad) * sizeof(double);
  auto kern = gang_kernel<G, KN, MODE>;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (device < 64 && blocks_per_sm[device] == 0) {
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      int n = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, 32, smem));
      if (n < 1) return fail(-5, "gang kernel does not fit on an SM (G=%d, KN=%d, smem=%zu)", G, KN, smem);
      if (const char* e = getenv("GAB1_GANG_WARPS")) { const int v = atoi(e); if (v >= 1 && v < n) n = v; }
      blocks_per_sm[device] = n;
      CUDA_TRY(cudaDeviceGetAttribute(&sms[device], cudaDevAttrMultiProcessorCount, device));
    }
  }
  int nb = 0, nsm = 0;
  if (device < 64) { nb = blocks_per_sm[device]; nsm = sms[device]; }
  if (nb == 0) {
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 32, smem));
    CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
  }
  // persistent grid of one-warp CTAs; never more warps than groups of sets
  long long grid = (long long)nsm * nb;
  const long long need = (args.S + L::NS - 1) / L::NS;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, 32, smem, stream>>>(args);
  count_launch();
  CUDA_TRY(cudaGetLastError());
  return 0;
}

template <int G, int KN>
int launch_mode(int mode, const KernelArgs& a, int device, cudaStream_t stream) {
  return mode == MODE_FAST_WHILE ? launch<G, KN, MODE_FAST_WHILE>(a, device, stream) : launch<G, KN, MODE_FAST_FOR>(a, device, stream);
}
}  // namespace

#define GAB1_CAT2(a, b) a##b
#define GAB1_CAT(a, b) GAB1_CAT2(a, b)
// launch_gang_kernel_g<G>: KN in {7, 10, 13} (G = 32 also 8)
int GAB1_CAT(launch_gang_kernel_g, GAB1_GANG_G)(int KN, int mode, const KernelArgs& a, int device, cudaStream_t stream) {
  constexpr int G = GAB1_GANG_G;
  switch (KN) {
    case 7: return launch_mode<G, 7>(mode, a, device, stream);
    case 13: return launch_mode<G, 13>(mode, a, device, stream);
#if GAB1_GANG_G == 4
    case 10: return launch_mode<G, 10>(mode, a, device, stream);
#endif
#if GAB1_GANG_G == 32
    case 8: return launch_mode<G, 8>(mode, a, device, stream);
#endif
  }
  return fail(-6, "no gang kernel for G=%d, KN=%d", G, KN);
}

}  // namespace gab1
