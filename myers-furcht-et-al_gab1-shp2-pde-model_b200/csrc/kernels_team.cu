// kernels_team.cu — instantiations of the latency kernel (team_kernel.cuh): one CTA per parameter set.
#include <mutex>

#include "launch.h"
#include "team_kernel.cuh"
#include "team_tangent_kernel.cuh"

namespace gab1 {
namespace {
template <int MODE>
int launch(const KernelArgs& args, int device, cudaStream_t stream) {
  const int W = (args.o.Nr + 31) / 32;
  const int T = 32 * W;
  if (T > 256) return fail(-6, "team kernel: Nr = %d needs more than 8 warps", args.o.Nr);
  const size_t smem = ((size_t)2 * NCY * T + WS_HDR + 2 * (size_t)args.P_pad) * sizeof(double);
  auto kern = team_kernel<MODE>;
  static std::mutex mu;
  static bool attr_set[64] = {false};
  {
    std::lock_guard<std::mutex> lk(mu);
    if (device < 64 && !attr_set[device]) {
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set[device] = true;
    }
  }
  int nb = 0, nsm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, T, smem));
  CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
  if (nb < 1) return fail(-5, "team kernel does not fit on an SM (T=%d, smem=%zu)", T, smem);
  long long grid = (long long)nsm * nb;
  if (grid > args.S) grid = args.S;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, T, smem, stream>>>(args);
  count_launch();
  CUDA_TRY(cudaGetLastError());
  return 0;
}
}  // namespace

namespace {
template <int NT>
int launch_tt(const TangentArgs& ta, int device, cudaStream_t stream) {
  const int W = (ta.a.o.Nr + 31) / 32;
  const int T = 32 * W;
  if (T > 256) return fail(-6, "team tangent kernel: Nr = %d needs more than 8 warps", ta.a.o.Nr);
  constexpr int NC = 1 + NT;
  const size_t smem = ((size_t)2 * NC * NCY * T + 16 * NC + (size_t)ta.a.P_pad + ((C_N * NT + 3) & ~3) + (size_t)L_N * NT * 32) * sizeof(double);
  auto kern = team_tangent_kernel<NT>;
  static std::mutex mu;
  static bool attr_set[64] = {false};
  {
    std::lock_guard<std::mutex> lk(mu);
    if (device < 64 && !attr_set[device]) {
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set[device] = true;
    }
  }
  int nb = 0, nsm = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, T, smem));
  CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
  if (nb < 1) return fail(-5, "team tangent kernel does not fit on an SM (T=%d, smem=%zu)", T, smem);
  long long grid = (long long)nsm * nb;
  const long long items = ta.a.S * (long long)ta.groups;
  if (grid > items) grid = items;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, T, smem, stream>>>(ta);
  count_launch();
  CUDA_TRY(cudaGetLastError());
  return 0;
}
}  // namespace

// NT directions per CTA (1, 2 or 4); ta.groups must be ceil(n_dir / NT)
int launch_team_tangent_kernel(int NT, const TangentArgs& ta, int device, cudaStream_t stream) {
  if (NT == 4) return launch_tt<4>(ta, device, stream);
  if (NT == 2) return launch_tt<2>(ta, device, stream);
  return launch_tt<1>(ta, device, stream);
}

int launch_team_kernel(int mode, const KernelArgs& a, int device, cudaStream_t stream) {
  return mode == MODE_FAST_WHILE ? launch<MODE_FAST_WHILE>(a, device, stream) : launch<MODE_FAST_FOR>(a, device, stream);
}

}  // namespace gab1
