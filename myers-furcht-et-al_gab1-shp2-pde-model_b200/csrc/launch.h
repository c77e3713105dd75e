// launch.h — what the host translation unit (gab1pde.cu) and the kernel translation units share.
// The kernels are instantiated in their own .cu files so that the library builds in parallel.
#pragma once
#include <cuda_runtime.h>

#include "gab1pde.h"

namespace gab1 {

struct KernelArgs {
  gab1_opts o;
  long long S;
  const double* Co; long long Co_stride;
  const double* D; const double* k; const double* dt; const double* r;
  double* out; long long out_stride;
  int* status; int* n_saved; long long* n_steps; long long* n_bc;
  const int* order;          // sets in descending-work order, or nullptr
  unsigned int* counter;     // work queue head
  double R_pow3;             // R^3.0 (libm pow on the host; sapdesolver.jl:353)
  int P_pad;                 // doubles per staged row in shared memory
  const int* guard;          // when non-null the kernel runs only if *guard == guard_expect (see gab1pde.cu)
  int guard_expect;
  // hand-over between kernel families (gang_kernel.cuh): a several-sets-per-warp kernel appends the sets it gives up on
  // (every step runs into the iteration limit) to retry_list; the one-set-per-warp kernel enqueued behind it solves
  // exactly those, its queue ending at *dyn_count
  unsigned int* retry_count;
  int* retry_list;
  const unsigned int* dyn_count;
  unsigned int* duo_counter; // queue head of the latency lane (duo_kernel.cuh): its items are [0, *dyn_count)
  int* sm_resv;              // 4 reservation counters per SM (duo_kernel.cuh: isolation of the latency lane's sets)
  int isolate;               // 0: off, 1: a set of the latency lane reserves its SM, 2: its schedulers only
};

enum { MODE_FAST_FOR = 0, MODE_FAST_WHILE = 1, MODE_STRICT = 2 };

// thread-local error message of the C ABI (gab1_last_error); returns `code`
int fail(int code, const char* fmt, ...);
void count_launch();

#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t e_ = (expr);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return ::gab1::fail(-100 - (int)e_, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// one warp per parameter set, K in {1,2,4,8} nodes per lane (solver_kernel.cuh); mode is MODE_*
int launch_single_kernel(int K, int mode, const KernelArgs& args, int device, cudaStream_t stream);
// skewed fast kernels (pair_kernel.cuh), fast modes only: two parameter sets per warp with K in {1,2,4} nodes per
// lane of a 16-lane group, or one set per warp with K in {2,4,8}
// variant: 0 = skewed loop, 1 = plain loop, 2 = plain loop with the interior token (where instantiated)
int launch_group16_kernel(int K, int variant, int mode, bool mirror, const KernelArgs& args, int device, cudaStream_t stream);
int launch_group32_kernel(int K, int variant, int mode, bool mirror, const KernelArgs& args, int device, cudaStream_t stream);
// shared-memory-resident solver for large grids (stream_kernel.cuh): K in {2, 4, 8, 16} nodes per lane, fast modes only
int launch_stream_kernel(int K, int mode, const KernelArgs& args, int device, cudaStream_t stream);
// latency kernel (team_kernel.cuh): one CTA of ceil(Nr/32) warps per set, 32 < Nr <= 256, fast modes only
int launch_team_kernel(int mode, const KernelArgs& args, int device, cudaStream_t stream);
// the one-set-per-warp kernel with the latency lane in front (duo_kernel.cuh): two warps per set for items [0, *args.dyn_count)
// of the queue (head: args.duo_counter), one warp per set from *args.counter on; K in {2, 4}, fast modes only
int launch_duo_kernel(int K, int mode, const KernelArgs& args, int device, cudaStream_t stream);
// several sets per warp (gang_kernel.cuh): G lanes per set (one translation unit per G), KN node slots per lane
int launch_gang_kernel_g2(int KN, int mode, const KernelArgs& args, int device, cudaStream_t stream);
int launch_gang_kernel_g4(int KN, int mode, const KernelArgs& args, int device, cudaStream_t stream);
int launch_gang_kernel_g8(int KN, int mode, const KernelArgs& args, int device, cudaStream_t stream);
int launch_gang_kernel_g16(int KN, int mode, const KernelArgs& args, int device, cudaStream_t stream);
int launch_gang_kernel_g32(int KN, int mode, const KernelArgs& args, int device, cudaStream_t stream);
// order statistics across the sets of a FULL result (ensemble_stats.cu); all pointers but `p` are device pointers
size_t quantiles_workspace_bytes(long long S);
int ensemble_quantiles_device(const gab1_opts* o, int device, cudaStream_t stream, long long S, const double* out,
                              const int* status, unsigned matrices, int c0, int c1, int np, const double* p, double* q,
                              long long* n_valid, void* workspace);
// forward-mode kernels (tangent_kernel.cuh): K nodes per lane, NT directions per work item
struct TangentArgs {
  KernelArgs a;          // a.out_stride = doubles per component block; a.out holds (1 + n_dir) blocks per set
  const double* seeds;   // S x n_dir x GAB1_N_SEED: partials of [D(7); k(17); Co(5); dt] along each direction
  int n_dir;
  int groups;            // work items per set = ceil(n_dir / NT)
};
int tangent_directions_per_item(int K, int n_dir);
int launch_tangent_kernel(int K, int NT, const TangentArgs& ta, int device, cudaStream_t stream);
// latency path of forward mode (team_tangent_kernel.cuh): one CTA per (set, direction), 64 < Nr <= 256
int launch_team_tangent_kernel(int NT, const TangentArgs& ta, int device, cudaStream_t stream);
// streamed-partials kernels (tangent_stream_kernel.cuh): primal in registers, partials in shared memory; NT in {2, 4}
int launch_tangent_stream_kernel(int K, int NT, const TangentArgs& ta, int device, cudaStream_t stream);
// synthetic prior ensembles on the device (sampler.cu); mu, sigma are host pointers
int sample_prior_device(int device, cudaStream_t stream, long long S, unsigned long long seed, const double* mu,
                        const double* sigma, double EGF, double Kdd, double* D, double* k);
// diagnostics (kernels_single.cu)
void launch_recip_error_kernel(double lo, double hi, int n, double* out);

}  // namespace gab1
