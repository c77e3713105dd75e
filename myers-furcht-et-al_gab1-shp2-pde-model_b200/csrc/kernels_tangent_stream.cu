// kernels_tangent_stream.cu — instantiations of the streamed-partials forward-mode kernels (tangent_stream_kernel.cuh).
#include <mutex>

#include "launch.h"
#include "tangent_stream_kernel.cuh"

namespace gab1 {
namespace {
template <int K, int NT>
int launch(const TangentArgs& ta, int device, cudaStream_t stream) {
  static std::mutex mu;
  static int blocks_per_sm[64] = {0};
  static int sms[64] = {0};
  const size_t smem = ((size_t)TSLayout<K, NT>::ROWS + 2 * (size_t)ta.a.P_pad) * sizeof(double);
  auto kern = tangent_stream_kernel<K, NT>;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (device < 64 && blocks_per_sm[device] == 0) {
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      int n = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, 32, smem));
      if (n < 1) return fail(-5, "streamed tangent kernel does not fit on an SM (K=%d, NT=%d, smem=%zu)", K, NT, smem);
      if (n > 8) n = 8;            // two warps per scheduler: more only lengthens the tail of the queue
      blocks_per_sm[device] = n;
      CUDA_TRY(cudaDeviceGetAttribute(&sms[device], cudaDevAttrMultiProcessorCount, device));
    }
  }
  int nb = 0, nsm = 0;
  if (device < 64) { nb = blocks_per_sm[device]; nsm = sms[device]; }
  if (nb == 0) {
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 32, smem));
    if (nb > 8) nb = 8;
    CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
  }
  long long grid = (long long)nsm * nb;
  const long long need = ta.a.S * ta.groups;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, 32, smem, stream>>>(ta);
  count_launch();
  CUDA_TRY(cudaGetLastError());
  return 0;
}
}  // namespace

// 0 when no streamed kernel is built for (K, NT)
int launch_tangent_stream_kernel(int K, int NT, const TangentArgs& ta, int device, cudaStream_t stream) {
  switch (K * 10 + NT) {
    case 12: return launch<1, 2>(ta, device, stream);
    case 14: return launch<1, 4>(ta, device, stream);
    case 22: return launch<2, 2>(ta, device, stream);
    case 24: return launch<2, 4>(ta, device, stream);
    case 42: return launch<4, 2>(ta, device, stream);
    case 44: return launch<4, 4>(ta, device, stream);
  }
  return fail(-6, "no streamed tangent kernel for K=%d, NT=%d", K, NT);
}

}  // namespace gab1
