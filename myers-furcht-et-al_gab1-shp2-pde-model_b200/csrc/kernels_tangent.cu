// kernels_tangent.cu — instantiations of the forward-mode kernels (tangent_kernel.cuh) and their launcher.
#include <mutex>

#include "launch.h"
#include "tangent_kernel.cuh"

namespace gab1 {
namespace {
template <int K, int NT>
int launch(const TangentArgs& ta, int device, cudaStream_t stream) {
  static std::mutex mu;
  static int blocks_per_sm[64] = {0};
  static int sms[64] = {0};
  constexpr int TPB = 32 * kWarpsPerCta;
  const size_t smem = (size_t)kWarpsPerCta * (TWS_HDR * (1 + NT) + 2 * (size_t)ta.a.P_pad) * sizeof(double);
  auto kern = tangent_kernel<K, NT>;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (device < 64 && blocks_per_sm[device] == 0) {
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      int n = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, TPB, smem));
      if (n < 1) return fail(-5, "tangent kernel does not fit on an SM (K=%d, NT=%d, smem=%zu)", K, NT, smem);
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
  long long grid = (long long)nsm * nb;
  const long long need = (ta.a.S * ta.groups + kWarpsPerCta - 1) / kWarpsPerCta;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, TPB, smem, stream>>>(ta);
  count_launch();
  CUDA_TRY(cudaGetLastError());
  return 0;
}
}  // namespace

int tangent_directions_per_item(int K, int n_dir) {
  // registers hold the state of NT + 1 components; (K, NT) = (2, 4) and (4, 2) spill the state itself and measured
  // 28x and 44x a primal solve — not built
  if (K == 1) return n_dir >= 4 ? 4 : (n_dir >= 2 ? 2 : 1);
  if (K == 2) return n_dir >= 2 ? 2 : 1;
  return 1;
}

int launch_tangent_kernel(int K, int NT, const TangentArgs& ta, int device, cudaStream_t stream) {
  switch (K * 10 + NT) {
    case 11: return launch<1, 1>(ta, device, stream);
    case 12: return launch<1, 2>(ta, device, stream);
    case 14: return launch<1, 4>(ta, device, stream);
    case 21: return launch<2, 1>(ta, device, stream);
    case 22: return launch<2, 2>(ta, device, stream);
    case 41: return launch<4, 1>(ta, device, stream);
  }
  return fail(-6, "no tangent kernel for K=%d, NT=%d", K, NT);
}

}  // namespace gab1
