// kernels_stream.cu — instantiations of the shared-memory-resident solver for large grids (stream_kernel.cuh).
#include <mutex>

#include "launch.h"
#include "stream_kernel.cuh"

namespace gab1 {
namespace {
template <int K, int MODE, int MINB = 8>
int launch(const KernelArgs& args, int device, cudaStream_t stream) {
  static std::mutex mu;
  static int blocks_per_sm[64] = {0};
  static int sms[64] = {0};
  const size_t smem = ((size_t)SLayout<K>::ROWS + 2 * (size_t)args.P_pad) * sizeof(double);
  auto kern = stream_kernel<K, MODE, MINB>;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (device < 64 && blocks_per_sm[device] == 0) {
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      int n = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, 32, smem));
      if (n < 1) return fail(-5, "streamed kernel does not fit on an SM (K=%d, smem=%zu)", K, smem);
      if (const char* e = getenv("GAB1_STREAM_WARPS")) { const int v = atoi(e); if (v >= 1 && v < n) n = v; }
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
  long long grid = (long long)nsm * nb;
  if (grid > args.S) grid = args.S;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, 32, smem, stream>>>(args);
  count_launch();
  CUDA_TRY(cudaGetLastError());
  return 0;
}
}  // namespace

int launch_stream_kernel(int K, int mode, const KernelArgs& a, int device, cudaStream_t stream) {
  // A/B: GAB1_STREAM_MINB = 12 | 16 trades registers (168 | 128 per thread) for resident warps (fast `for` form, K <= 8)
  if (const char* e = getenv("GAB1_STREAM_MINB")) {
    const int v = atoi(e);
    if (mode == MODE_FAST_FOR && (v == 12 || v == 16)) {
      switch (K * 100 + v) {
        case 212: return launch<2, MODE_FAST_FOR, 12>(a, device, stream);
        case 216: return launch<2, MODE_FAST_FOR, 16>(a, device, stream);
        case 412: return launch<4, MODE_FAST_FOR, 12>(a, device, stream);
        case 416: return launch<4, MODE_FAST_FOR, 16>(a, device, stream);
        case 812: return launch<8, MODE_FAST_FOR, 12>(a, device, stream);
      }
    }
  }
#define GAB1_LAUNCH(KK)                                                                   \
  case KK:                                                                                \
    return mode == MODE_FAST_WHILE ? launch<KK, MODE_FAST_WHILE>(a, device, stream)       \
                                   : launch<KK, MODE_FAST_FOR>(a, device, stream);
  switch (K) {
    GAB1_LAUNCH(2)
    GAB1_LAUNCH(4)
    GAB1_LAUNCH(8)
    GAB1_LAUNCH(16)
  }
#undef GAB1_LAUNCH
  return fail(-6, "no streamed kernel for K=%d", K);
}

}  // namespace gab1
