// gab1pde.cu — C ABI of libgab1pde.so (include/gab1pde.h): option checking, work ordering, kernel launch,
// multi-GPU sharding of the host entry point.  No CPU implementation of the solver lives here.
#include <cuda_runtime.h>
#include <ctype.h>
#include <math.h>
#include <sched.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <cub/device/device_radix_sort.cuh>
#include <mutex>
#include <thread>
#include <vector>

#include "gab1pde.h"
#include "launch.h"

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace

namespace gab1 {
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
void count_launch() { g_launches.fetch_add(1); }
}  // namespace gab1


namespace {
using gab1::fail;

int popcount12(uint32_t m) { return __builtin_popcount(m & GAB1_MASK_ALL_MATRICES); }

int check_opts(const gab1_opts* o) {
  if (!o) return fail(-1, "opts is NULL");
  if (o->abi_version != GAB1_ABI_VERSION) return fail(-1, "abi_version %d != %d", o->abi_version, GAB1_ABI_VERSION);
  if (o->Nr < 2) return fail(-2, "Nr must be >= 2");
  if (o->Nts < 1) return fail(-2, "Nts must be >= 1");
  if (o->maxiters < 0) return fail(-2, "maxiters must be >= 0");
  if (o->geometry < 0 || o->geometry > 1 || o->sfk_mode < 0 || o->sfk_mode > 2 || o->bc_loop < 0 || o->bc_loop > 1 ||
      o->save_rule < 0 || o->save_rule > 1 || o->pg1tot_form < 0 || o->pg1tot_form > 1 || o->out_mode < 0 ||
      o->out_mode > 4 || o->arith < 0 || o->arith > 1)
    return fail(-3, "an enum field of gab1_opts is out of range");
  if (!(o->dr > 0.0) || !(o->R > 0.0)) return fail(-2, "R and dr must be positive");
  return 0;
}

// nodes per lane for a one-warp-per-set launch: nodes 1..Nr over 32 lanes
int pick_K(int Nr) {
  for (int K : {1, 2, 4, 8})
    if (Nr <= 32 * K) return K;
  return 0;
}
// Which fast kernel serves a grid (measured on B200):
//   Nr <= 32        two sets per warp, skewed loop (pair_kernel.cuh, HW = 16)        1.23x over the legacy kernel at Nr = 25
//   32 < Nr <= 128  the first-generation one-set-per-warp kernel (solver_kernel.cuh)  best at Nr = 40, 50, 100
//   128 < Nr <= 256 one set per warp, skewed loop (pair_kernel.cuh, HW = 32, K = 8)   1.45x over legacy at Nr = 200
// hw = lanes per parameter set (16: two sets per warp; 32: one; 0: legacy), K = nodes per lane, variant: 0 skewed loop,
// 1 plain loop, 2 plain loop with the interior token.  GAB1_KERNEL = legacy | group16 | group16p | group32 | group32p |
// group32t overrides the choice where the named family can hold the grid (A/B measurements).
struct FastPick { int hw, K, variant; };
FastPick pick_fast(int Nr) {
  auto fit = [&](int hw) -> int {
    for (int K : {1, 2, 4, 8}) {
      if (hw == 16 && K == 8) break;
      if (hw == 32 && K == 1) continue;
      if (Nr <= hw * K) return K;
    }
    return 0;
  };
  const char* e = getenv("GAB1_KERNEL");
  if (e && strcmp(e, "legacy") == 0) return {0, 0, 0};
  if (e && strcmp(e, "group16") == 0 && fit(16)) return {16, fit(16), 0};
  if (e && strcmp(e, "group16p") == 0 && fit(16) == 4) return {16, 4, 1};
  if (e && strcmp(e, "group32") == 0 && fit(32)) return {32, fit(32), 0};
  if (e && strcmp(e, "group32p") == 0 && fit(32) == 2) return {32, 2, 1};
  if (e && strcmp(e, "group32t") == 0 && fit(32) == 2) return {32, 2, 2};
  if (Nr <= 32) return {16, fit(16), 0};
  if (Nr <= 128) return {0, 0, 0};
  return {32, fit(32), 0};
}

// Shape of the several-sets-per-warp kernels for a grid: G lanes per set, KN node slots per lane, Nr <= G*KN; the
// narrowest gang that holds the grid (most sets per warp), then the shortest run.  GAB1_GANG="G,KN" overrides (A/B).
struct GangPick { int G, KN; };
GangPick pick_gang(int Nr) {
  if (const char* e = getenv("GAB1_GANG")) {
    int G = 0, KN = 0;
    if (sscanf(e, "%d,%d", &G, &KN) == 2 && G * KN >= Nr) return {G, KN};
  }
  static const GangPick table[] = {{2, 7}, {2, 13}, {4, 7}, {4, 10}, {4, 13}, {8, 7}, {8, 13}, {16, 13}, {32, 8}, {32, 13}};
  for (const GangPick& g : table)
    if (Nr <= g.G * g.KN) return g;
  return {0, 0};
}
int launch_gang(int G, int KN, int mode, const gab1::KernelArgs& a, int device, cudaStream_t stream) {
  switch (G) {
    case 2: return gab1::launch_gang_kernel_g2(KN, mode, a, device, stream);
    case 4: return gab1::launch_gang_kernel_g4(KN, mode, a, device, stream);
    case 8: return gab1::launch_gang_kernel_g8(KN, mode, a, device, stream);
    case 16: return gab1::launch_gang_kernel_g16(KN, mode, a, device, stream);
    case 32: return gab1::launch_gang_kernel_g32(KN, mode, a, device, stream);
  }
  return fail(-6, "no gang kernel for G = %d", G);
}

// ---- descending-work ordering: key = number of time steps, largest first -------------------------------------
// Thread 0 also sets the guard word the pair kernels check: their spherical stencil drops the zero-flux mirror term of
// node 1, which is exact only when 1 - dr/r[1] == 0, i.e. on the reference's own grid r = collect(0:dr:R).  For any
// other r the pair launch returns at once and the general kernel enqueued behind it does the work.
__global__ void work_keys_kernel(long long S, const double* dt, double tf, unsigned* keys, int* vals, const double* r,
                                 double dr, int spherical, int* guard) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i == 0) *guard = (spherical && 1.0 - dr / r[1] != 0.0) ? 1 : 0;
  if (i >= S) return;
  const double n = ceil(tf / dt[i]);
  keys[i] = (n >= 0.0 && n < 4.0e9) ? (unsigned)n : 0u;   // unusable dt: no work, goes last
  vals[i] = (int)i;
}

struct Workspace {   // carved out of the caller's workspace buffer
  unsigned* counter;
  unsigned *keys_in, *keys_out;
  int *vals_in, *vals_out;
  void* cub_tmp;
  size_t cub_bytes;
};

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

size_t cub_bytes_for(long long S) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, (const unsigned*)nullptr, (unsigned*)nullptr,
                                            (const int*)nullptr, (int*)nullptr, (int)S);
  return bytes;
}

constexpr size_t kCounterBytes = 8192;   // 64 queue words + 4 reservation words for each of up to 496 SMs

size_t carve(Workspace& w, void* base, long long S) {
  size_t off = 0;
  auto take = [&](size_t bytes) { void* p = base ? (char*)base + off : nullptr; off += align_up(bytes, 256); return p; };
  w.counter = (unsigned*)take(kCounterBytes);     // [0] queue head, [1] guard word, [2..9] hand-over / latency lane, [10] fused-or-plain guard, [64..] SM reservations
  w.keys_in = (unsigned*)take(sizeof(unsigned) * (size_t)S);
  w.keys_out = (unsigned*)take(sizeof(unsigned) * (size_t)S);
  w.vals_in = (int*)take(sizeof(int) * (size_t)S);
  w.vals_out = (int*)take(sizeof(int) * (size_t)S);
  w.cub_bytes = cub_bytes_for(S);
  w.cub_tmp = take(w.cub_bytes);
  return off;
}

// The several-sets-per-warp kernel, then the one-set-per-warp kernel over the sets the first one handed back (those
// that hit the iteration limit step after step: in a gang they would stall their warp's other sets).  Both launches are
// enqueued back to back; the second reads its queue length from device memory, so there is no host synchronisation.
int solve_with_gangs(const GangPick& gp, int mode, gab1::KernelArgs a, const Workspace& w, int device, cudaStream_t stream) {
  const bool retry = !getenv("GAB1_GANG_NO_RETRY");
  if (retry) { a.retry_count = w.counter + 2; a.retry_list = w.vals_in; }      // vals_in is free once the sort has run
  if (int rc = launch_gang(gp.G, gp.KN, mode, a, device, stream)) return rc;
  if (!retry) return 0;
  gab1::KernelArgs b = a;
  b.retry_count = nullptr; b.retry_list = nullptr;
  b.order = w.vals_in;
  b.counter = w.counter + 3;
  b.dyn_count = w.counter + 2;
  b.guard = nullptr;
  const int K = pick_K(a.o.Nr);
  return K ? gab1::launch_single_kernel(K, mode, b, device, stream)
           : gab1::launch_stream_kernel(16, mode, b, device, stream);
}

// ---- the latency lane (duo_kernel.cuh): which sets of a batch get two warps ------------------------------------------
// Decided on the device from the sorted step counts, so the host never waits: the first n sets of the descending-work
// queue get two warps each (queue [0, n), counter[5] / counter[4]) and the one-set-per-warp queue starts at n (counter[0]).  Both kernels give the same bits for a set, so n is free to follow the load:
//   S > warps        the batch fills the GPU: only sets whose step count exceeds 0.9 of the per-warp share of the whole batch
//                    (they would still be running when everything else has finished), at most one per two SMs.  Measured
//                    on shards of the 10^5-draw bench ensemble: longest set / share = 1.59 gains 12 %, 0.96 / 0.85 / 0.80 gain
//                    nothing (profiles/r2_latency_lane_probe.jsonl); round 2's first threshold of 0.5 sent 74 sets of a uniform
//                    two-sets-per-warp batch (2368 posterior rows at dr = 0.1) to the lane and cost it 25 % (356 -> 446 ms);
//   warps/4 < S      every set has a warp to itself and the GPU is at least half full: none.  Two warps per set execute more
//                    instructions per step (warp B recomputes node Nr-1), which costs more than the shorter chain saves once the
//                    schedulers are shared: 592 / 900 / 1184 posterior rows 30.7 / 31.5 / 37.8 ms one warp per set against
//                    35.9 / 36.4 / 41.0 ms with the lane (gpurun_out/r2n_duo_probe.jsonl);
//   S <= warps/4     every set (1 / 64 / 296 rows: 16.5 / 19.4 / 20.0 ms against 17.5 / 21.4 / 21.7 ms).
__global__ void __launch_bounds__(1024) duo_plan_kernel(long long S, const unsigned* keys, long long warps, int nsm, int all,
                                                        unsigned* counter) {
  __shared__ unsigned long long s_part[32];
  __shared__ unsigned long long s_total;
  __shared__ unsigned s_cnt;
  const int tid = threadIdx.x;
  unsigned long long acc = 0;
  for (long long i = tid; i < S; i += 1024) acc += keys[i];
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) s_part[tid >> 5] = acc;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  if (tid == 0) {
    unsigned long long tot = 0;
    for (int i = 0; i < 32; ++i) tot += s_part[i];
    s_total = tot;
  }
  __syncthreads();
  double thr;
  long long cap;
  if (all) { thr = -1.0; cap = S; }
  else if (S > warps) { thr = 0.9 * (double)s_total / (double)warps; cap = nsm / 2; }
  else if (4 * S > warps) { thr = 0.0; cap = 0; }
  else { thr = -1.0; cap = S; }
  const long long lim = S < cap ? S : cap;
  unsigned local = 0;
  for (long long i = tid; i < lim; i += 1024) local += ((double)keys[i] > thr) ? 1u : 0u;
  if (local) atomicAdd(&s_cnt, local);
  __syncthreads();
  if (tid == 0) { counter[0] = s_cnt; counter[4] = s_cnt; counter[5] = 0u; counter[10] = s_cnt ? 1u : 0u; }
}

// The plan, then the one-set-per-warp kernel with the latency lane in front (duo_kernel.cuh: duo_solve_kernel) — and, behind a
// guard word the plan writes, the plain one-set-per-warp kernel: exactly one of the two runs (the other's CTAs return at
// once).  The fused kernel's throughput lane carries 13 more register moves per step than solve_kernel (ptxas allocates
// for both lanes: 422 against 409 instructions, 120.5 vs 116.3 ms on 4736 posterior rows), so a batch that sends
// nothing to the latency lane — every full batch without a heavy tail — should not pay for it.
int solve_with_duo(int K, int mode, gab1::KernelArgs a, const Workspace& w, int device, cudaStream_t stream, bool all) {
  int nsm = 148;
  CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
  duo_plan_kernel<<<1, 1024, 0, stream>>>(a.S, w.keys_out, (long long)nsm * 8, nsm, all ? 1 : 0, w.counter);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  a.dyn_count = w.counter + 4;
  a.duo_counter = w.counter + 5;
  if (getenv("GAB1_DUO_TIMELINE")) CUDA_TRY(cudaMemsetAsync(w.counter + 8, 0xff, 8, stream));   // timing builds: earliest start
  // full batches: the few sets of the latency lane run isolated (duo_kernel.cuh); GAB1_ISOLATE = 0 | sm | sched overrides
  a.sm_resv = (int*)(w.counter + 64);
  a.isolate = (!all && a.S > (long long)nsm * 8 && (size_t)(64 + 4 * nsm) * sizeof(unsigned) <= kCounterBytes) ? 2 : 0;
  if (const char* e = getenv("GAB1_ISOLATE")) {
    if (a.isolate) a.isolate = strcmp(e, "sched") == 0 ? 2 : (strcmp(e, "sm") == 0 ? 1 : (e[0] == '0' ? 0 : a.isolate));
  }
  a.guard = (const int*)(w.counter + 10);     // written by the plan: 1 = the latency lane has sets
  a.guard_expect = 1;
  if (int rc = gab1::launch_duo_kernel(K, mode, a, device, stream)) return rc;
  if (all) return 0;
  a.guard_expect = 0;                 // no set for the latency lane: the plain kernel, its queue starting at counter[0] = 0
  a.dyn_count = nullptr; a.duo_counter = nullptr; a.sm_resv = nullptr; a.isolate = 0;
  return gab1::launch_single_kernel(K, mode, a, device, stream);
}

int solve_device(const gab1_opts* o, int device, cudaStream_t stream, long long S, const double* Co, long long Co_stride,
                 const double* D, const double* k, const double* dt, const double* r, double* out, int* status,
                 int* n_saved, long long* n_steps, long long* n_bc, void* workspace) {
  if (int rc = check_opts(o)) return rc;
  if (S < 0) return fail(-2, "S must be >= 0");
  if (S == 0) return 0;
  if (S > 2000000000LL) return fail(-2, "S too large for one call");
  if (Co_stride != 0 && Co_stride != GAB1_N_CO) return fail(-2, "Co_stride must be 0 or 5");
  if (!Co || !D || !k || !dt || !r || !out || !workspace) return fail(-2, "a required buffer is NULL");
  const int K = pick_K(o->Nr);
  if (o->Nr > 512) return fail(-6, "Nr = %d: grids with more than 512 nodes are not supported by this build", o->Nr);
  CUDA_TRY(cudaSetDevice(device));

  Workspace w;
  carve(w, workspace, S);
  gab1::KernelArgs a;
  memset(&a, 0, sizeof a);
  a.o = *o;
  a.o.device_ids = nullptr;
  a.S = S;
  a.Co = Co; a.Co_stride = Co_stride; a.D = D; a.k = k; a.dt = dt; a.r = r;
  a.out = out; a.out_stride = gab1_out_doubles_per_set(o);
  a.status = status; a.n_saved = n_saved; a.n_steps = n_steps; a.n_bc = n_bc;
  a.order = w.vals_out;
  a.counter = w.counter;
  a.R_pow3 = pow(o->R, 3.0);
  a.P_pad = (o->Nr + 1 + 3) & ~3;

  CUDA_TRY(cudaMemsetAsync(w.counter, 0, kCounterBytes, stream));
  {
    const int tb = 256;
    work_keys_kernel<<<(unsigned)((S + tb - 1) / tb), tb, 0, stream>>>(S, dt, o->tf, w.keys_in, w.vals_in, r, o->dr,
                                                                       o->geometry == GAB1_GEOM_SPHERICAL, (int*)(w.counter + 1));
    g_launches.fetch_add(1);
    CUDA_TRY(cudaGetLastError());
    size_t bytes = w.cub_bytes;
    CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(w.cub_tmp, bytes, w.keys_in, w.keys_out, w.vals_in, w.vals_out,
                                                       (int)S, 0, 32, stream));
  }
  // `maxiters = 0` with the for-loop form never runs the membrane block (degenerate but legal): the strict kernel
  // reproduces it exactly, the fast kernels assume at least one pass
  const bool degenerate = o->bc_loop == GAB1_BC_FOR_BREAK && o->maxiters == 0;
  const int mode = (o->arith == 1 || degenerate) ? gab1::MODE_STRICT
                   : (o->bc_loop == GAB1_BC_WHILE ? gab1::MODE_FAST_WHILE : gab1::MODE_FAST_FOR);
  if (mode == gab1::MODE_STRICT && K == 0)
    return fail(-6, "Nr = %d: the strict kernels (arith = 1, maxiters = 0) hold at most 256 nodes", o->Nr);
  // ---- large grids: the state lives in shared memory (stream_kernel.cuh).  Measured on B200:
  //      Nr = 200, 1184 sets: 1667 ms against 2808 ms for the register-resident K = 8 kernel (which spills its state);
  //      Nr = 100: 416 vs 392 ms, Nr = 50: 191 vs 117 ms — the register kernels keep the small grids.
  //      GAB1_KERNEL=stream forces this family for Nr > 32 (A/B measurements, parity tests).
  // ---- small batches: one CTA per set (team_kernel.cuh).  GAB1_TEAM_MAX_SETS overrides the measured threshold.
  if (mode != gab1::MODE_STRICT && o->Nr > 32 && o->Nr <= 256) {
    const char* e = getenv("GAB1_KERNEL");
    const bool named = e && e[0];
    // Measured crossover on B200 (tools/bench_small.py): teams of 2 warps (Nr <= 64) lose to one warp per
    // set even for a single solve (22 vs 15 ms); teams of 4 win 1.47x up to 148 sets and 1.17x at 296, lose at 592; teams of 7
    // win 2.75x for one set, 1.65x at 296, lose at 592.  Rule: all teams resident at <= 8 warps per SM (16 for 6+ warps).
    long long max_sets = 0;
    if (o->Nr > 64) {
      const int W = (o->Nr + 31) / 32;
      int nsm = 148;
      (void)cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
      max_sets = (long long)nsm * (W >= 6 ? 16 : 8) / W;
    }
    if (const char* m = getenv("GAB1_TEAM_MAX_SETS")) { if (m[0]) max_sets = atoll(m); }
    if ((named && strcmp(e, "team") == 0) || (!named && S <= max_sets))
      return gab1::launch_team_kernel(mode, a, device, stream);
  }
  // ---- large grids: several sets per warp, the state in shared memory (gang_kernel.cuh).  Measured on B200, full batches
  //      (DESIGN.md section 5): Nr = 200 310 ms against 348 ms streamed; Nr = 400 1123 ms against 2036 ms; at Nr <= 128
  //      the register-resident kernels below still win (Nr = 50: 393-467 ms against 424 ms; Nr = 25: 93 against 45 ms).
  //      GAB1_KERNEL=gang forces the family on any grid, GAB1_GANG="G,KN" the shape (A/B measurements, parity tests).
  if (mode != gab1::MODE_STRICT) {
    const char* e = getenv("GAB1_KERNEL");
    const bool named = e && e[0];
    if ((named && strcmp(e, "gang") == 0) || (!named && o->Nr > 128)) {
      const GangPick gp = pick_gang(o->Nr);
      if (!gp.G) return fail(-6, "Nr = %d: no gang kernel holds this grid", o->Nr);
      return solve_with_gangs(gp, mode, a, w, device, stream);
    }
  }
  if (mode != gab1::MODE_STRICT) {
    const char* e = getenv("GAB1_KERNEL");
    const bool forced = e && strcmp(e, "stream") == 0 && o->Nr > 32;
    const bool other = e && e[0] && strcmp(e, "stream") != 0 && o->Nr <= 256;     // an explicit request for another family
    if (forced || (o->Nr > 128 && !other))
      return gab1::launch_stream_kernel(o->Nr <= 64 ? 2 : (o->Nr <= 128 ? 4 : (o->Nr <= 256 ? 8 : 16)), mode, a, device, stream);
  }
  // ---- fast arithmetic: the skewed kernels of pair_kernel.cuh ----
  const FastPick fp = mode != gab1::MODE_STRICT ? pick_fast(o->Nr) : FastPick{0, 0, 0};
  if (fp.K) {
    const bool mirror = o->geometry != GAB1_GEOM_SPHERICAL;
    if (!mirror) { a.guard = (const int*)(w.counter + 1); a.guard_expect = 0; }
    const int rc = fp.hw == 16 ? gab1::launch_group16_kernel(fp.K, fp.variant, mode, mirror, a, device, stream)
                               : gab1::launch_group32_kernel(fp.K, fp.variant, mode, mirror, a, device, stream);
    if (rc || mirror) return rc;
    a.guard_expect = 1;          // the general kernel below runs only if the pair kernel declined the grid
  }
  // ---- the latency lane: two warps for the sets that would outlast the batch, or for all of a small batch (duo_kernel.cuh;
  //      bit-identical to the kernel below).  GAB1_KERNEL=duo sends every set there, GAB1_KERNEL=legacy or GAB1_DUO=0 none.
  if (mode != gab1::MODE_STRICT && !fp.K && (K == 2 || K == 4)) {
    const char* e = getenv("GAB1_KERNEL");
    const char* d = getenv("GAB1_DUO");
    const bool all = e && strcmp(e, "duo") == 0;
    const bool off = (e && e[0] && !all) || (d && d[0] == '0');
    if (!off) return solve_with_duo(K, mode, a, w, device, stream, all);
  }
  return gab1::launch_single_kernel(K, mode, a, device, stream);
}

// ---- forward-mode tangents (tangent_kernel.cuh) -------------------------------------------------------------------
int check_tangent_opts(const gab1_opts* o, int n_dir) {
  if (int rc = check_opts(o)) return rc;
  if (n_dir < 1 || n_dir > 64) return fail(-2, "n_dir must be in 1..64");
  if (o->bc_loop != GAB1_BC_FOR_BREAK) return fail(-6, "tangents: only the `for … break` membrane loop is supported");
  if (o->maxiters < 1) return fail(-6, "tangents: maxiters must be >= 1");
  if (o->save_rule != GAB1_SAVE_T_GE_TSAVE) return fail(-6, "tangents: only the t >= t_save snapshot rule is supported");
  if (o->t_prechase >= 0.0) return fail(-6, "tangents: pulse-chase is not supported");
  if (o->out_mode == GAB1_OUT_SIX) return fail(-6, "tangents: GAB1_OUT_SIX is piecewise constant in the parameters");
  if (o->Nr > 128) return fail(-6, "tangents: Nr = %d, grids with more than 128 nodes are not supported", o->Nr);
  return 0;
}

int solve_tangent_device(const gab1_opts* o, int device, cudaStream_t stream, long long S, int n_dir, const double* Co,
                         long long Co_stride, const double* D, const double* k, const double* dt, const double* seeds,
                         const double* r, double* out, int* status, int* n_saved, long long* n_steps, long long* n_bc,
                         void* workspace) {
  if (int rc = check_tangent_opts(o, n_dir)) return rc;
  if (S < 0) return fail(-2, "S must be >= 0");
  if (S == 0) return 0;
  if (S > 30000000LL) return fail(-2, "S too large for one call");
  if (Co_stride != 0 && Co_stride != GAB1_N_CO) return fail(-2, "Co_stride must be 0 or 5");
  if (!Co || !D || !k || !dt || !seeds || !r || !out || !workspace) return fail(-2, "a required buffer is NULL");
  const int K = pick_K(o->Nr);
  CUDA_TRY(cudaSetDevice(device));
  Workspace w;
  carve(w, workspace, S);
  gab1::TangentArgs ta;
  memset(&ta, 0, sizeof ta);
  gab1::KernelArgs& a = ta.a;
  a.o = *o;
  a.o.device_ids = nullptr;
  a.S = S;
  a.Co = Co; a.Co_stride = Co_stride; a.D = D; a.k = k; a.dt = dt; a.r = r;
  a.out = out; a.out_stride = gab1_out_doubles_per_set(o);
  a.status = status; a.n_saved = n_saved; a.n_steps = n_steps; a.n_bc = n_bc;
  a.order = w.vals_out;
  a.counter = w.counter;
  a.R_pow3 = pow(o->R, 3.0);
  a.P_pad = (o->Nr + 1 + 3) & ~3;
  ta.seeds = seeds;
  ta.n_dir = n_dir;
  // Kernel choice, measured on B200 with 4 partials per set:
  //   Nr <= 32        streamed kernel, 2 directions per item   (all five variants within 4 % of each other)
  //   32 < Nr <= 64   register kernel, 2 directions per item   (11.6x a primal solve; streamed with 4: 15.7x)
  //   64 < Nr <= 128  streamed kernel, 4 directions per item   (16.2x; register kernel with 1: 26.8x)
  //   one direction   register kernel
  // GAB1_TANGENT = reg | stream and GAB1_TANGENT_NT = 1 | 2 | 4 override for A/B measurements.
  //   small batches   register kernel, ONE direction per item: while every (set, direction) pair gets a warp of its own
  //                   with at most one warp per scheduler, latency is what counts — one loss-and-gradient evaluation with
  //                   4 partials: dr = 0.2 59 ms (2 per item: 88 ms), dr = 0.1 581 ms (streamed, 4 per item: 1202 ms);
  //                   a 101-point multistart population: 168 vs 223 ms and 651 vs 1304 ms (tools/tangent_latency.sh)
  bool streamed = n_dir >= 2 && K != 2;
  int NT = n_dir == 1 ? 1 : (K == 4 && n_dir >= 3 ? 4 : 2);
  {
    int nsm = 148;
    (void)cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    if (S * (long long)n_dir <= 4LL * nsm) { streamed = false; NT = 1; }
  }
  if (const char* e = getenv("GAB1_TANGENT")) {
    if (strcmp(e, "reg") == 0) { streamed = false; NT = gab1::tangent_directions_per_item(K, n_dir); }
    if (strcmp(e, "stream") == 0) { streamed = true; NT = n_dir >= 3 ? 4 : 2; }
  }
  if (const char* e = getenv("GAB1_TANGENT_NT")) {
    const int v = atoi(e);
    if (streamed ? (v == 2 || v == 4) : ((v == 1) || (v == 2 && K <= 2) || (v == 4 && K == 1))) NT = v;
  }
  // one CTA per (set, direction) while every one of them is resident (team_tangent_kernel.cuh; measured in DESIGN.md section 8)
  bool team = false;
  if (o->Nr > 32 && o->Nr <= 256) {
    int nsm = 148;
    (void)cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    // dr = 0.1, one / 37 / 101 evaluations with 4 partials: 127 / 138 / 309 ms against 577 / 628 / 654 ms for one warp per
    // (set, direction); two CTAs are resident per SM, so 8 * SMs items are four rounds of ~130 ms.
    // dr = 0.2 (teams of 2 warps): 1 / 37 / 101 / 296 evaluations 33 / 37 / 116 / 139 ms against 60 / 67 / 168 / 185 ms
    // full-size batches: dr = 0.1, 1184 sets x 4 partials 2.55 s (11.5x a primal solve) against 3.6 s streamed (16.2x): the
    // team kernel serves 64 < Nr <= 128 at every batch size; dr = 0.2, 4736 sets: 1.48 s against 1.19 s for the register
    // kernel with two directions per item, which keeps the large batches there
    team = o->Nr > 64 || S * (long long)n_dir <= 8LL * nsm;
    if (const char* e = getenv("GAB1_TANGENT")) { if (strcmp(e, "team") == 0) team = true; else if (e[0]) team = false; }
  }
  ta.groups = (n_dir + NT - 1) / NT;
  CUDA_TRY(cudaMemsetAsync(w.counter, 0, 4 * sizeof(unsigned), stream));
  const int tb = 256;
  work_keys_kernel<<<(unsigned)((S + tb - 1) / tb), tb, 0, stream>>>(S, dt, o->tf, w.keys_in, w.vals_in, r, o->dr, 0,
                                                                     (int*)(w.counter + 1));
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  size_t bytes = w.cub_bytes;
  CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(w.cub_tmp, bytes, w.keys_in, w.keys_out, w.vals_in, w.vals_out, (int)S,
                                                     0, 32, stream));
  if (team) {
    // directions per CTA: once the batch fills the GPU, up to four share one primal (coefficient partials in shared memory;
    // dr = 0.1, 1184 sets x 4 partials: 1.77 s with four, 1.96 s with two, 2.62 s with one); one per CTA is the latency optimum
    // (one gradient: 127 ms against 205 ms with two).  dr = 0.2: 1.18 / 1.24 s with two / four, the register kernel 1.19 s.
    int nsm = 148;
    (void)cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
    int NTt = (n_dir >= 2 && o->Nr > 64 && S * (long long)n_dir > 8LL * nsm) ? (n_dir >= 3 ? 4 : 2) : 1;
    if (const char* e = getenv("GAB1_TANGENT_NT")) { const int v = atoi(e); if (v == 1 || ((v == 2 || v == 4) && n_dir >= 2)) NTt = v; }
    ta.groups = (n_dir + NTt - 1) / NTt;
    return gab1::launch_team_tangent_kernel(NTt, ta, device, stream);
  }
  return streamed ? gab1::launch_tangent_stream_kernel(K, NT, ta, device, stream)
                  : gab1::launch_tangent_kernel(K, NT, ta, device, stream);
}

// ---- FP64 peak: register-resident DFMA chains ------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* sink, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456) sink[0] = s;
}

}  // namespace

extern "C" {

void gab1_opts_init(gab1_opts* o, double R, double dr, double tf, int32_t Nts) {
  memset(o, 0, sizeof *o);
  o->abi_version = GAB1_ABI_VERSION;
  o->geometry = GAB1_GEOM_SPHERICAL;
  o->sfk_mode = GAB1_SFK_DIFFUSIBLE;
  o->bc_loop = GAB1_BC_FOR_BREAK;
  o->save_rule = GAB1_SAVE_T_GE_TSAVE;
  o->pg1tot_form = GAB1_PG1TOT_VIA_STOT;
  o->out_mode = GAB1_OUT_FULL;
  o->matrix_mask = GAB1_MASK_ALL_MATRICES;
  o->maxiters = 100;                       // basepdesolver.jl:32
  o->tol = 1.0e-6;                         // basepdesolver.jl:33
  o->R = R; o->dr = dr; o->tf = tf;
  o->Nr = (int32_t)ceil(R / dr);           // basepdesolver.jl:71
  o->Nts = Nts;
  o->dt_save = tf / Nts;                   // basepdesolver.jl:31
  o->t_prechase = -1.0;
  o->pct_mul = 1.0; o->pct_div = 1.0;
  o->n_devices = 1;
}

int64_t gab1_out_doubles_per_set(const gab1_opts* o) {
  if (!o) return 0;
  const int64_t P = (int64_t)o->Nr + 1, C = (int64_t)o->Nts + 1;
  switch (o->out_mode) {
    case GAB1_OUT_FINAL4: return 4 * P;
    case GAB1_OUT_FULL: return (int64_t)popcount12(o->matrix_mask) * P * C + GAB1_N_VECTORS * C;
    case GAB1_OUT_SIX: return 6;
    case GAB1_OUT_PCT_BOUND: return 1;
    case GAB1_OUT_FINAL_STATE: return 10 * P + 8;
  }
  return 0;
}

int64_t gab1_full_matrix_offset(const gab1_opts* o, int32_t m) {
  if (!o || m < 0 || m >= GAB1_N_MATRICES || !((o->matrix_mask >> m) & 1u)) return -1;
  const int64_t P = (int64_t)o->Nr + 1, C = (int64_t)o->Nts + 1;
  return (int64_t)popcount12(o->matrix_mask & ((1u << m) - 1u)) * P * C;
}

int64_t gab1_full_vector_offset(const gab1_opts* o, int32_t v) {
  if (!o || v < 0 || v >= GAB1_N_VECTORS) return -1;
  const int64_t P = (int64_t)o->Nr + 1, C = (int64_t)o->Nts + 1;
  return (int64_t)popcount12(o->matrix_mask) * P * C + (int64_t)v * C;
}

int gab1_default_dt(int64_t S, const double* D, const double* k, double dr, double* dt) {
  if (!D || !k || !dt) return fail(-2, "a required buffer is NULL");
  for (int64_t i = 0; i < S; ++i) {
    double mx = D[i * GAB1_N_D];
    for (int q = 1; q < GAB1_N_D; ++q) mx = D[i * GAB1_N_D + q] > mx ? D[i * GAB1_N_D + q] : mx;
    double sk = 0.0;
    for (int q = 0; q < GAB1_N_K; ++q) sk += k[i * GAB1_N_K + q];
    dt[i] = 1.0 / (2.0 * (mx / (dr * dr) + sk / 4)) * 0.99;
  }
  return 0;
}

size_t gab1_workspace_bytes(int64_t S) {
  Workspace w;
  return carve(w, nullptr, S < 1 ? 1 : S);
}

int gab1_solve_batch_device(const gab1_opts* o, int32_t device, void* stream, int64_t S, const double* Co,
                            int64_t Co_stride, const double* D, const double* k, const double* dt, const double* r,
                            double* out, int32_t* status, int32_t* n_saved, int64_t* n_steps, int64_t* n_bc_iters,
                            void* workspace) {
  static_assert(sizeof(long long) == sizeof(int64_t), "int64_t must be long long");
  return solve_device(o, device, (cudaStream_t)stream, S, Co, Co_stride, D, k, dt, r, out, status, n_saved,
                      (long long*)n_steps, (long long*)n_bc_iters, workspace);
}

// Per-device arena of the host entry point: one slab of device memory and one stream per GPU, created on first use,
// grown when a call needs more and kept between calls (gab1_release_device_memory frees them), so that a call costs no
// cudaMalloc / cudaFree.  Calls that target the same device take turns; different devices run concurrently.
struct DeviceArena {
  std::mutex mu;
  cudaStream_t stream = nullptr;
  char* base = nullptr;
  size_t cap = 0;
};
static DeviceArena g_arena[64];
// per-set outputs up to this many doubles are sharded by dealing (gather / scatter on the host), larger ones contiguously
static const int64_t kDealtMaxDoubles = 8192;
static size_t& arena_prev_cap(int device) { static size_t prev[64] = {0}; return prev[device]; }


// Grow a device's arena to at least `need` bytes.  cudaFree + cudaMalloc cost 20-500 ms on a 180 GB device (measured), so
// the arena starts at 64 MB and grows by at least half / doubles while it is small — but never asks for more than the
// device can give: the request is capped by the free memory (after the old slab is released) and falls back to the exact
// size, so a shard that fits the budget never fails on the growth margin.
static int arena_reserve(DeviceArena& ar, int device, size_t need) {
  if (need <= ar.cap) return 0;
  if (ar.base) { CUDA_TRY(cudaStreamSynchronize(ar.stream)); cudaFree(ar.base); ar.base = nullptr; ar.cap = 0; }
  size_t want = need + need / 2;
  if (want < ((size_t)64 << 20)) want = (size_t)64 << 20;
  if (want < 2 * arena_prev_cap(device) && arena_prev_cap(device) <= ((size_t)4 << 30)) want = 2 * arena_prev_cap(device);
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
    const size_t room = free_b > ((size_t)256 << 20) ? free_b - ((size_t)256 << 20) : free_b;
    if (want > room) want = room;
  } else {
    (void)cudaGetLastError();
  }
  if (want < need) want = need;
  if (cudaMalloc((void**)&ar.base, want) != cudaSuccess) {
    (void)cudaGetLastError();
    ar.base = nullptr;
    want = need;
    CUDA_TRY(cudaMalloc((void**)&ar.base, want));
  }
  ar.cap = want;
  arena_prev_cap(device) = want;
  return 0;
}

// gab1_solve_ensemble_quantiles: the FULL result stays on the device and only order statistics come back
struct QReq {
  uint32_t matrices; int32_t c0, c1, np;
  const double* p;
  double* q;            // host: nm x np x (c1-c0) x (Nr+1)
  int64_t* n_valid;     // host
};

// bytes of device memory one shard may plan for: 85 % of the device, or GAB1_MAX_DEVICE_BYTES
static size_t device_budget_bytes(int device) {
  if (const char* e = getenv("GAB1_MAX_DEVICE_BYTES")) { const long long v = atoll(e); if (v > 0) return (size_t)v; }
  size_t free_b = 0, total_b = 0;
  if (cudaSetDevice(device) != cudaSuccess || cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  const size_t by_total = total_b / 100 * 85, by_free = free_b > ((size_t)1 << 30) ? free_b - ((size_t)1 << 30) : free_b / 2;
  size_t held = 0;       // this library's own slab on the device is reusable
  if (device >= 0 && device < 64) held = g_arena[device].cap;
  return by_total < by_free + held ? by_total : by_free + held;
}

// One shard of the host entry point: copy in, solve, copy out, on its device's stream.
static int run_shard(const gab1_opts* o, int device, int64_t lo, int64_t hi, const double* Co, int64_t Co_stride,
                     const double* D, const double* k, const double* dt, const double* r, double* out, int32_t* status,
                     int32_t* n_saved, int64_t* n_steps, int64_t* n_bc, const QReq* qr = nullptr) {
  const int64_t S = hi - lo;
  if (S <= 0) return 0;
  if (device < 0 || device >= 64) return fail(-7, "device ordinal %d out of range", device);
  CUDA_TRY(cudaSetDevice(device));
  const int64_t nout = gab1_out_doubles_per_set(o);
  // A shard whose staged output would not fit the device is solved in pieces, one after the other (0.5 MB per set at
  // Nr = 50: 3e5 full solutions fill a 180 GB B200; the order statistics need the whole ensemble resident and are
  // capped elsewhere).  GAB1_MAX_DEVICE_BYTES overrides the budget (tests).
  bool mapped_out = false;
  if (!qr) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, out + lo * nout) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
      mapped_out = true;              // the kernels write a pinned, mapped buffer directly: nothing is staged
    else
      (void)cudaGetLastError();
  }
  if (!qr && !mapped_out && S > 1) {
    size_t budget = device_budget_bytes(device);
    if (budget && (size_t)S * ((size_t)nout * sizeof(double) + 512) > budget) {
      const int64_t mid = lo + S / 2;
      if (int e = run_shard(o, device, lo, mid, Co, Co_stride, D, k, dt, r, out, status, n_saved, n_steps, n_bc)) return e;
      return run_shard(o, device, mid, hi, Co, Co_stride, D, k, dt, r, out, status, n_saved, n_steps, n_bc);
    }
  }
  DeviceArena& ar = g_arena[device];
  std::lock_guard<std::mutex> lk(ar.mu);
  if (!ar.stream) CUDA_TRY(cudaStreamCreateWithFlags(&ar.stream, cudaStreamNonBlocking));
  cudaStream_t st = ar.stream;
  const size_t P = (size_t)o->Nr + 1;
  // A pinned (mapped) caller buffer is written by the kernel directly: snapshot stores stream over PCIe while the
  // time loop runs, so there is no device copy of the 0.5 MB/set output and no D2H phase.  Pageable buffers are staged.
  double* out_direct = nullptr;
  if (!qr) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, out + lo * nout) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
      out_direct = (double*)attr.devicePointer;
    else
      (void)cudaGetLastError();
  }
  // carve the arena
  const size_t nCo = Co_stride ? (size_t)S * GAB1_N_CO : GAB1_N_CO;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t at = off; off += align_up(bytes, 256); return at; };
  const size_t oCo = take(nCo * sizeof(double)), oD = take((size_t)S * GAB1_N_D * sizeof(double)),
               oK = take((size_t)S * GAB1_N_K * sizeof(double)), oDt = take((size_t)S * sizeof(double)),
               oR = take(P * sizeof(double)), oSt = take((size_t)S * sizeof(int32_t)), oSv = take((size_t)S * sizeof(int32_t)),
               oNs = take((size_t)S * sizeof(int64_t)), oBc = take((size_t)S * sizeof(int64_t)),
               oWs = take(gab1_workspace_bytes(S)), oOut = take(out_direct ? 0 : (size_t)S * nout * sizeof(double));
  size_t nq = 0, oQ = 0, oQws = 0;
  if (qr) {
    nq = (size_t)__builtin_popcount(qr->matrices) * qr->np * (qr->c1 - qr->c0) * P;
    oQ = take(nq * sizeof(double));
    oQws = take(gab1::quantiles_workspace_bytes(S) + 256);
  }
  const bool dbg = getenv("GAB1_DEBUG_TIMING") != nullptr;
  auto now = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
  const double t_a = now();
  if (int e = arena_reserve(ar, device, off)) return e;
  const double t_b = now();
  if (dbg) fprintf(stderr, "[gab1] run_shard S=%lld arena %.1f ms (cap %zu)\n", (long long)S, 1e3 * (t_b - t_a), ar.cap);
  double *dCo = (double*)(ar.base + oCo), *dD = (double*)(ar.base + oD), *dk = (double*)(ar.base + oK),
         *ddt = (double*)(ar.base + oDt), *dr_ = (double*)(ar.base + oR), *dout = (double*)(ar.base + oOut);
  int32_t *dstatus = (int32_t*)(ar.base + oSt), *dsaved = (int32_t*)(ar.base + oSv);
  int64_t *dsteps = (int64_t*)(ar.base + oNs), *dbc = (int64_t*)(ar.base + oBc);
  void* ws = ar.base + oWs;
  CUDA_TRY(cudaMemcpyAsync(dCo, Co + (Co_stride ? lo * Co_stride : 0), nCo * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(dD, D + lo * GAB1_N_D, (size_t)S * GAB1_N_D * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(dk, k + lo * GAB1_N_K, (size_t)S * GAB1_N_K * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(ddt, dt + lo, (size_t)S * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(dr_, r, P * sizeof(double), cudaMemcpyHostToDevice, st));
  const double t_c = now();
  if (int e = solve_device(o, device, st, S, dCo, Co_stride, dD, dk, ddt, dr_, out_direct ? out_direct : dout, dstatus, dsaved,
                           (long long*)dsteps, (long long*)dbc, ws)) {
    cudaStreamSynchronize(st);
    return e;
  }
  if (dbg) { const double t_d = now(); cudaStreamSynchronize(st);
    fprintf(stderr, "[gab1] run_shard S=%lld enqueue %.1f ms, kernels %.1f ms\n", (long long)S, 1e3 * (t_d - t_c), 1e3 * (now() - t_d)); }
  if (qr) {
    double* dq = (double*)(ar.base + oQ);
    long long* dnv = (long long*)(ar.base + oQws);
    if (int e = gab1::ensemble_quantiles_device(o, device, st, S, dout, dstatus, qr->matrices, qr->c0, qr->c1, qr->np, qr->p, dq,
                                                dnv, ar.base + oQws + 256)) {
      cudaStreamSynchronize(st);
      return e;
    }
    CUDA_TRY(cudaMemcpyAsync(qr->q, dq, nq * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (qr->n_valid) CUDA_TRY(cudaMemcpyAsync(qr->n_valid, dnv, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  } else if (!out_direct)
    CUDA_TRY(cudaMemcpyAsync(out + lo * nout, dout, (size_t)S * nout * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (status) CUDA_TRY(cudaMemcpyAsync(status + lo, dstatus, (size_t)S * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (n_saved) CUDA_TRY(cudaMemcpyAsync(n_saved + lo, dsaved, (size_t)S * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (n_steps) CUDA_TRY(cudaMemcpyAsync(n_steps + lo, dsteps, (size_t)S * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  if (n_bc) CUDA_TRY(cudaMemcpyAsync(n_bc + lo, dbc, (size_t)S * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

size_t gab1_quantiles_workspace_bytes(int64_t S) { return gab1::quantiles_workspace_bytes(S); }

int gab1_ensemble_quantiles_device(const gab1_opts* o, int32_t device, void* stream, int64_t S, const double* out,
                                   const int32_t* status, uint32_t matrices, int32_t c0, int32_t c1, int32_t np,
                                   const double* p, double* q, int64_t* n_valid, void* workspace) {
  if (int rc = check_opts(o)) return rc;
  return gab1::ensemble_quantiles_device(o, device, (cudaStream_t)stream, S, out, status, matrices, c0, c1, np, p, q,
                                         (long long*)n_valid, workspace);
}

int gab1_solve_ensemble_quantiles(const gab1_opts* o, int64_t S, const double* Co, int64_t Co_stride, const double* D,
                                  const double* k, const double* dt, const double* r, uint32_t matrices, int32_t c0,
                                  int32_t c1, int32_t np, const double* p, double* q, int32_t* status, int32_t* n_saved,
                                  int64_t* n_steps, int64_t* n_bc_iters, int64_t* n_valid) {
  if (int rc = check_opts(o)) return rc;
  if (o->out_mode != GAB1_OUT_FULL) return fail(-2, "ensemble quantiles need out_mode = GAB1_OUT_FULL");
  if (S < 1) return fail(-2, "S must be >= 1");
  if (!Co || !D || !k || !dt || !r || !q || !p) return fail(-2, "a required buffer is NULL");
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1)
    return fail(-7, "no CUDA device is visible; this library has no CPU fallback");
  // the order statistics need every set of the ensemble on one device: a single GPU holds 3e5 full solutions at Nr = 50
  const int device = o->device_ids ? o->device_ids[0] : 0;
  QReq qr{matrices, c0, c1, np, p, q, n_valid};
  return run_shard(o, device, 0, S, Co, Co_stride, D, k, dt, r, nullptr, status, n_saved, n_steps, n_bc_iters, &qr);
}

// One shard of gab1_solve_tangent: copy in, solve, copy out, on its device's stream (arena shared with run_shard).
static int run_tangent_shard(const gab1_opts* o, int device, int64_t lo, int64_t hi, int n_dir, const double* Co,
                             int64_t Co_stride, const double* D, const double* k, const double* dt, const double* seeds,
                             const double* r, double* out, int32_t* status, int32_t* n_saved, int64_t* n_steps, int64_t* n_bc) {
  const int64_t S = hi - lo;
  if (S <= 0) return 0;
  if (device < 0 || device >= 64) return fail(-7, "device ordinal %d out of range", device);
  CUDA_TRY(cudaSetDevice(device));
  const int64_t nout = gab1_out_doubles_per_set(o) * (1 + n_dir);
  if (S > 1) {      // a shard whose staged output and seeds would not fit the device is solved in pieces (as run_shard does)
    const size_t budget = device_budget_bytes(device);
    const size_t per_set = (size_t)nout * sizeof(double) + (size_t)n_dir * GAB1_N_SEED * sizeof(double) + 512;
    if (budget && (size_t)S * per_set > budget) {
      const int64_t mid = lo + S / 2;
      if (int e = run_tangent_shard(o, device, lo, mid, n_dir, Co, Co_stride, D, k, dt, seeds, r, out, status, n_saved, n_steps, n_bc)) return e;
      return run_tangent_shard(o, device, mid, hi, n_dir, Co, Co_stride, D, k, dt, seeds, r, out, status, n_saved, n_steps, n_bc);
    }
  }
  DeviceArena& ar = g_arena[device];
  std::lock_guard<std::mutex> lk(ar.mu);
  if (!ar.stream) CUDA_TRY(cudaStreamCreateWithFlags(&ar.stream, cudaStreamNonBlocking));
  cudaStream_t st = ar.stream;
  const size_t P = (size_t)o->Nr + 1;
  const size_t nCo = Co_stride ? (size_t)S * GAB1_N_CO : GAB1_N_CO;
  const size_t nSeed = (size_t)S * n_dir * GAB1_N_SEED;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t at = off; off += align_up(bytes, 256); return at; };
  const size_t oCo = take(nCo * sizeof(double)), oD = take((size_t)S * GAB1_N_D * sizeof(double)),
               oK = take((size_t)S * GAB1_N_K * sizeof(double)), oDt = take((size_t)S * sizeof(double)),
               oSd = take(nSeed * sizeof(double)), oR = take(P * sizeof(double)), oSt = take((size_t)S * sizeof(int32_t)),
               oSv = take((size_t)S * sizeof(int32_t)), oNs = take((size_t)S * sizeof(int64_t)),
               oBc = take((size_t)S * sizeof(int64_t)), oWs = take(gab1_workspace_bytes(S)),
               oOut = take((size_t)S * nout * sizeof(double));
  if (int e = arena_reserve(ar, device, off)) return e;
  double *dCo = (double*)(ar.base + oCo), *dD = (double*)(ar.base + oD), *dk = (double*)(ar.base + oK),
         *ddt = (double*)(ar.base + oDt), *dsd = (double*)(ar.base + oSd), *dr_ = (double*)(ar.base + oR),
         *dout = (double*)(ar.base + oOut);
  int32_t *dstatus = (int32_t*)(ar.base + oSt), *dsaved = (int32_t*)(ar.base + oSv);
  int64_t *dsteps = (int64_t*)(ar.base + oNs), *dbc = (int64_t*)(ar.base + oBc);
  CUDA_TRY(cudaMemcpyAsync(dCo, Co + (Co_stride ? lo * Co_stride : 0), nCo * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(dD, D + lo * GAB1_N_D, (size_t)S * GAB1_N_D * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(dk, k + lo * GAB1_N_K, (size_t)S * GAB1_N_K * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(ddt, dt + lo, (size_t)S * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(dsd, seeds + lo * n_dir * GAB1_N_SEED, nSeed * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(dr_, r, P * sizeof(double), cudaMemcpyHostToDevice, st));
  if (int e = solve_tangent_device(o, device, st, S, n_dir, dCo, Co_stride, dD, dk, ddt, dsd, dr_, dout, dstatus, dsaved,
                                   (long long*)dsteps, (long long*)dbc, ar.base + oWs)) {
    cudaStreamSynchronize(st);
    return e;
  }
  CUDA_TRY(cudaMemcpyAsync(out + lo * nout, dout, (size_t)S * nout * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (status) CUDA_TRY(cudaMemcpyAsync(status + lo, dstatus, (size_t)S * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (n_saved) CUDA_TRY(cudaMemcpyAsync(n_saved + lo, dsaved, (size_t)S * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (n_steps) CUDA_TRY(cudaMemcpyAsync(n_steps + lo, dsteps, (size_t)S * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  if (n_bc) CUDA_TRY(cudaMemcpyAsync(n_bc + lo, dbc, (size_t)S * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

int gab1_solve_tangent_device(const gab1_opts* o, int32_t device, void* stream, int64_t S, int32_t n_dir, const double* Co,
                              int64_t Co_stride, const double* D, const double* k, const double* dt, const double* seeds,
                              const double* r, double* out, int32_t* status, int32_t* n_saved, int64_t* n_steps,
                              int64_t* n_bc_iters, void* workspace) {
  return solve_tangent_device(o, device, (cudaStream_t)stream, S, n_dir, Co, Co_stride, D, k, dt, seeds, r, out, status,
                              n_saved, (long long*)n_steps, (long long*)n_bc_iters, workspace);
}

int gab1_solve_tangent(const gab1_opts* o, int64_t S, int32_t n_dir, const double* Co, int64_t Co_stride, const double* D,
                       const double* k, const double* dt, const double* seeds, const double* r, double* out,
                       int32_t* status, int32_t* n_saved, int64_t* n_steps, int64_t* n_bc_iters) {
  if (int rc = check_tangent_opts(o, n_dir)) return rc;
  if (S < 0) return fail(-2, "S must be >= 0");
  if (S == 0) return 0;
  if (!Co || !D || !k || !dt || !seeds || !r || !out) return fail(-2, "a required buffer is NULL");
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1)
    return fail(-7, "no CUDA device is visible; this library has no CPU fallback");
  int nd = o->n_devices <= 0 ? visible : o->n_devices;
  if (nd > visible && !o->device_ids) return fail(-7, "n_devices = %d but only %d CUDA devices are visible", nd, visible);
  if (nd > S) nd = (int)S;
  std::vector<int> devs(nd);
  for (int i = 0; i < nd; ++i) devs[i] = o->device_ids ? o->device_ids[i] : i;
  std::vector<int64_t> bounds(nd + 1, 0);
  gab1_plan_shards(S, dt, o->tf, nd, bounds.data());
  if (nd == 1)
    return run_tangent_shard(o, devs[0], 0, S, n_dir, Co, Co_stride, D, k, dt, seeds, r, out, status, n_saved, n_steps, n_bc_iters);
  std::vector<int> rcs(nd, 0);
  std::vector<std::string> msgs(nd);
  std::vector<std::thread> th;
  for (int g = 0; g < nd; ++g)
    th.emplace_back([&, g]() {
      rcs[g] = run_tangent_shard(o, devs[g], bounds[g], bounds[g + 1], n_dir, Co, Co_stride, D, k, dt, seeds, r, out, status,
                                 n_saved, n_steps, n_bc_iters);
      if (rcs[g]) msgs[g] = g_err;
    });
  for (auto& t : th) t.join();
  for (int g = 0; g < nd; ++g)
    if (rcs[g]) return fail(rcs[g], "device %d: %s", devs[g], msgs[g].c_str());
  return 0;
}

// dt = 1.0/(2.0*(maximum(D)/(dr^2) + sum(k)/4))*0.99 on dual numbers (basepdesolver.jl:696): value and, in slot 29 of
// every seed row, its partial along that direction (sum rule over k, the partial of the largest D, quotient rule)
int gab1_default_dt_tangent(int64_t S, int32_t n_dir, const double* D, const double* k, double dr, double* dt, double* seeds) {
  if (!D || !k || !dt || !seeds || n_dir < 1) return fail(-2, "bad arguments to gab1_default_dt_tangent");
  for (int64_t i = 0; i < S; ++i) {
    int im = 0;
    for (int q = 1; q < GAB1_N_D; ++q) if (D[i * GAB1_N_D + q] > D[i * GAB1_N_D + im]) im = q;
    double sk = 0.0;
    for (int q = 0; q < GAB1_N_K; ++q) sk += k[i * GAB1_N_K + q];
    const double den = 2.0 * (D[i * GAB1_N_D + im] / (dr * dr) + sk / 4);
    const double inv = 1.0 / den;
    dt[i] = inv * 0.99;
    for (int d = 0; d < n_dir; ++d) {
      double* s = seeds + (i * n_dir + d) * GAB1_N_SEED;
      double dsk = 0.0;
      for (int q = 0; q < GAB1_N_K; ++q) dsk += s[GAB1_N_D + q];
      const double dden = 2.0 * (s[im] / (dr * dr) + dsk / 4);
      s[GAB1_N_SEED - 1] = (-(inv / den) * dden) * 0.99;
    }
  }
  return 0;
}

int gab1_sample_prior_device(int32_t device, void* stream, int64_t S, uint64_t seed, const double* mu, const double* sigma,
                             double EGF, double Kdd, double* D, double* k) {
  return gab1::sample_prior_device(device, (cudaStream_t)stream, S, seed, mu, sigma, EGF, Kdd, D, k);
}

int gab1_sample_prior(int64_t S, uint64_t seed, const double* mu, const double* sigma, double EGF, double Kdd, double* D,
                      double* k) {
  if (S < 0 || !mu || !sigma || !D || !k) return fail(-2, "bad arguments to gab1_sample_prior");
  if (S == 0) return 0;
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1)
    return fail(-7, "no CUDA device is visible; this library has no CPU fallback");
  CUDA_TRY(cudaSetDevice(0));
  double *dD = nullptr, *dk = nullptr;
  CUDA_TRY(cudaMalloc((void**)&dD, (size_t)S * GAB1_N_D * sizeof(double)));
  if (cudaMalloc((void**)&dk, (size_t)S * GAB1_N_K * sizeof(double)) != cudaSuccess) { cudaFree(dD); return fail(-8, "cudaMalloc failed"); }
  int rc = gab1::sample_prior_device(0, nullptr, S, seed, mu, sigma, EGF, Kdd, dD, dk);
  if (!rc && cudaMemcpy(D, dD, (size_t)S * GAB1_N_D * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(-9, "copy of D failed");
  if (!rc && cudaMemcpy(k, dk, (size_t)S * GAB1_N_K * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(-9, "copy of k failed");
  cudaFree(dD); cudaFree(dk);
  return rc;
}

void gab1_release_device_memory(void) {
  for (int d = 0; d < 64; ++d) {
    DeviceArena& ar = g_arena[d];
    std::lock_guard<std::mutex> lk(ar.mu);
    if (!ar.base && !ar.stream) continue;
    if (cudaSetDevice(d) != cudaSuccess) continue;
    if (ar.stream) { cudaStreamSynchronize(ar.stream); cudaStreamDestroy(ar.stream); ar.stream = nullptr; }
    if (ar.base) { cudaFree(ar.base); ar.base = nullptr; ar.cap = 0; }
    arena_prev_cap(d) = 0;
  }
}

int gab1_solve_batch(const gab1_opts* o, int64_t S, const double* Co, int64_t Co_stride, const double* D, const double* k,
                     const double* dt, const double* r, double* out, int32_t* status, int32_t* n_saved, int64_t* n_steps,
                     int64_t* n_bc_iters) {
  if (int rc = check_opts(o)) return rc;
  if (S < 0) return fail(-2, "S must be >= 0");
  if (S == 0) return 0;
  if (!Co || !D || !k || !dt || !r || !out) return fail(-2, "a required buffer is NULL");
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1)
    return fail(-7, "no CUDA device is visible; this library has no CPU fallback");
  int nd = o->n_devices <= 0 ? visible : o->n_devices;
  if (nd > visible && !o->device_ids) return fail(-7, "n_devices = %d but only %d CUDA devices are visible", nd, visible);
  if (nd > S) nd = (int)S;
  std::vector<int> devs(nd);
  for (int i = 0; i < nd; ++i) devs[i] = o->device_ids ? o->device_ids[i] : i;

  if (nd == 1) return run_shard(o, devs[0], 0, S, Co, Co_stride, D, k, dt, r, out, status, n_saved, n_steps, n_bc_iters);

  std::vector<int64_t> bounds(nd + 1, 0);
  std::vector<int> rcs(nd, 0);
  std::vector<std::string> msgs(nd);
  std::vector<std::thread> th;
  const int64_t nout = gab1_out_doubles_per_set(o);
  const char* plan_env = getenv("GAB1_SHARD_PLAN");       // contiguous | dealt (A/B measurements, tests)
  const bool dealt = plan_env && plan_env[0] ? strcmp(plan_env, "dealt") == 0 : nout <= kDealtMaxDoubles;
  std::vector<int64_t> perm;          // outlives the worker threads, which are joined below
  if (dealt) {
    // Small per-set outputs (final profiles, six scalars, one percentage): the sets are dealt to the devices from the
    // descending step-count order, so that every device receives the same mix of long and short solves (and of the
    // diverging ones, which are cheap) whatever the order of the caller's matrix; each device's sets are gathered into
    // contiguous staging rows, solved as one local batch, and scattered back to their columns.
    perm.resize((size_t)S);
    gab1_deal_shards(S, dt, o->tf, nd, perm.data(), bounds.data());
    for (int g = 0; g < nd; ++g)
      th.emplace_back([&, g]() {
        const int64_t lo = bounds[g], n = bounds[g + 1] - lo;
        if (n <= 0) return;
        const int64_t* idx = perm.data() + lo;
        std::vector<double> lD((size_t)n * GAB1_N_D), lk((size_t)n * GAB1_N_K), ldt((size_t)n), lCo(Co_stride ? (size_t)n * GAB1_N_CO : 0),
            lout((size_t)n * nout);
        std::vector<int32_t> lst((size_t)n), lsv((size_t)n);
        std::vector<int64_t> lns((size_t)n), lbc((size_t)n);
        for (int64_t i = 0; i < n; ++i) {
          const int64_t j = idx[i];
          memcpy(&lD[(size_t)i * GAB1_N_D], D + j * GAB1_N_D, GAB1_N_D * sizeof(double));
          memcpy(&lk[(size_t)i * GAB1_N_K], k + j * GAB1_N_K, GAB1_N_K * sizeof(double));
          ldt[(size_t)i] = dt[j];
          if (Co_stride) memcpy(&lCo[(size_t)i * GAB1_N_CO], Co + j * Co_stride, GAB1_N_CO * sizeof(double));
        }
        rcs[g] = run_shard(o, devs[g], 0, n, Co_stride ? lCo.data() : Co, Co_stride, lD.data(), lk.data(), ldt.data(), r,
                           lout.data(), lst.data(), lsv.data(), lns.data(), lbc.data());
        if (rcs[g]) { msgs[g] = g_err; return; }
        for (int64_t i = 0; i < n; ++i) {
          const int64_t j = idx[i];
          memcpy(out + j * nout, &lout[(size_t)i * nout], (size_t)nout * sizeof(double));
          if (status) status[j] = lst[(size_t)i];
          if (n_saved) n_saved[j] = lsv[(size_t)i];
          if (n_steps) n_steps[j] = lns[(size_t)i];
          if (n_bc_iters) n_bc_iters[j] = lbc[(size_t)i];
        }
      });
  } else {
    // Full snapshot output (0.5 MB per set and more): contiguous ranges balanced by step count, so that each device's
    // result lands in the caller's buffer as one block (or is written there by the kernels, when the buffer is mapped).
    gab1_plan_shards(S, dt, o->tf, nd, bounds.data());
    for (int g = 0; g < nd; ++g)
      th.emplace_back([&, g]() {
        rcs[g] = run_shard(o, devs[g], bounds[g], bounds[g + 1], Co, Co_stride, D, k, dt, r, out, status, n_saved, n_steps,
                           n_bc_iters);
        if (rcs[g]) msgs[g] = g_err;
      });
  }
  for (auto& t : th) t.join();
  for (int g = 0; g < nd; ++g)
    if (rcs[g]) return fail(rcs[g], "device %d: %s", devs[g], msgs[g].c_str());
  return 0;
}

// ---- certified solve (gab1pde.h) ---------------------------------------------------------------------------------
// The explicit scheme's dt ignores the second-order rate x concentration terms (basepdesolver.jl:30), so a few per mille of
// wide prior draws sit at the edge of stability: an alternating mode amplifies last-bit differences by 1e5..1e13 over the
// 4e4 steps without blowing up, and no arithmetic but the reference's own reproduces the reference there.  Those sets are
// found by the response of the final state to a one-ulp change of the initial concentrations (two fast solves) and
// re-solved with the strict kernels (arith = 1: bit-identical to the oracle), whose rows replace the fast ones.
int gab1_solve_batch_certified(const gab1_opts* o, int64_t S, const double* Co, int64_t Co_stride, const double* D,
                               const double* k, const double* dt, const double* r, double* out, int32_t* status,
                               int32_t* n_saved, int64_t* n_steps, int64_t* n_bc_iters, double response, int32_t* resolved,
                               int64_t* n_resolved) {
  if (n_resolved) *n_resolved = 0;
  if (int rc = check_opts(o)) return rc;
  if (S < 0) return fail(-2, "S must be >= 0");
  if (resolved) for (int64_t s = 0; s < S; ++s) resolved[s] = 0;
  if (S == 0) return 0;
  if (Co_stride != 0 && Co_stride != GAB1_N_CO) return fail(-2, "Co_stride must be 0 or 5");
  if (!(response > 0.0)) response = 1e-12;
  // the caller's solve; status and iteration counts are needed here whether or not the caller wants them
  std::vector<int32_t> st_own, sv_own;
  std::vector<int64_t> ns_own, bc_own;
  if (!status) { st_own.resize(S); status = st_own.data(); }
  if (!n_saved) { sv_own.resize(S); n_saved = sv_own.data(); }
  if (!n_steps) { ns_own.resize(S); n_steps = ns_own.data(); }
  if (!n_bc_iters) { bc_own.resize(S); n_bc_iters = bc_own.data(); }
  if (int rc = gab1_solve_batch(o, S, Co, Co_stride, D, k, dt, r, out, status, n_saved, n_steps, n_bc_iters)) return rc;
  if (o->arith == 1) return 0;                       // already the reference's arithmetic
  // final states at Co and at Co moved one ulp up
  gab1_opts fs = *o;
  fs.out_mode = GAB1_OUT_FINAL_STATE;
  const int64_t nf = gab1_out_doubles_per_set(&fs), nout = gab1_out_doubles_per_set(o);
  const int64_t nco = Co_stride ? S * GAB1_N_CO : GAB1_N_CO;
  std::vector<double> Co_up(nco), A, B((size_t)S * nf);
  for (int64_t i = 0; i < nco; ++i) Co_up[i] = nextafter(Co[i], INFINITY);
  std::vector<int32_t> sa_own, sb(S);
  std::vector<int64_t> ba_own, bb(S);
  const double* a = out;
  const int32_t* sa = status;
  const int64_t* ba = n_bc_iters;
  if (o->out_mode != GAB1_OUT_FINAL_STATE) {
    A.resize((size_t)S * nf); sa_own.resize(S); ba_own.resize(S);
    if (int rc = gab1_solve_batch(&fs, S, Co, Co_stride, D, k, dt, r, A.data(), sa_own.data(), nullptr, nullptr, ba_own.data())) return rc;
    a = A.data(); sa = sa_own.data(); ba = ba_own.data();
  }
  if (int rc = gab1_solve_batch(&fs, S, Co_up.data(), Co_stride, D, k, dt, r, B.data(), sb.data(), nullptr, nullptr, bb.data())) return rc;
  std::vector<int64_t> idx;
  for (int64_t s = 0; s < S; ++s) {
    const double* as = a + s * nf;
    const double* bs = B.data() + s * nf;
    double scale = 0.0;
    for (int64_t i = 0; i < nf; ++i) if (isfinite(as[i]) && fabs(as[i]) > scale) scale = fabs(as[i]);
    double e = 0.0;
    for (int64_t i = 0; i < nf; ++i) {
      if (isnan(as[i]) != isnan(bs[i])) { e = INFINITY; break; }
      if (!isfinite(as[i])) continue;
      const double den = fmax(fabs(as[i]), 1e-6 * scale);
      if (den > 0.0) { const double d = fabs(bs[i] - as[i]) / den; if (d > e || isnan(d)) e = isnan(d) ? INFINITY : d; }
    }
    const bool both_diverge = (sa[s] & GAB1_ST_NAN) && (sb[s] & GAB1_ST_NAN);     // NaN either way: dropped by every caller
    if ((e >= response || ba[s] != bb[s] || sa[s] != sb[s]) && !both_diverge) idx.push_back(s);
  }
  if (n_resolved) *n_resolved = (int64_t)idx.size();
  if (idx.empty()) return 0;
  // strict re-solve of the flagged sets, gathered into a dense batch
  const int64_t F = (int64_t)idx.size();
  std::vector<double> Dg((size_t)F * GAB1_N_D), kg((size_t)F * GAB1_N_K), dtg(F), Cog(Co_stride ? (size_t)F * GAB1_N_CO : 0),
      og((size_t)F * nout);
  std::vector<int32_t> sg(F), svg(F);
  std::vector<int64_t> nsg(F), bcg(F);
  for (int64_t j = 0; j < F; ++j) {
    const int64_t s = idx[j];
    memcpy(&Dg[j * GAB1_N_D], D + s * GAB1_N_D, sizeof(double) * GAB1_N_D);
    memcpy(&kg[j * GAB1_N_K], k + s * GAB1_N_K, sizeof(double) * GAB1_N_K);
    dtg[j] = dt[s];
    if (Co_stride) memcpy(&Cog[j * GAB1_N_CO], Co + s * GAB1_N_CO, sizeof(double) * GAB1_N_CO);
  }
  gab1_opts so = *o;
  so.arith = 1;
  if (int rc = gab1_solve_batch(&so, F, Co_stride ? Cog.data() : Co, Co_stride, Dg.data(), kg.data(), dtg.data(), r, og.data(),
                                sg.data(), svg.data(), nsg.data(), bcg.data())) return rc;
  for (int64_t j = 0; j < F; ++j) {
    const int64_t s = idx[j];
    memcpy(out + s * nout, &og[j * nout], sizeof(double) * nout);
    status[s] = sg[j]; n_saved[s] = svg[j]; n_steps[s] = nsg[j]; n_bc_iters[s] = bcg[j];
    if (resolved) resolved[s] = 1;
  }
  return 0;
}

// Dealt shards: the sets in descending step-count order go, one by one, to the shard with the least work so far (LPT),
// which keeps the shards' total step counts within one short solve of each other AND gives every shard the same
// distribution of solve lengths; perm[bounds[g] .. bounds[g+1]) lists shard g's set indices in ascending order.
int gab1_deal_shards(int64_t S, const double* dt, double tf, int32_t n_shards, int64_t* perm, int64_t* bounds) {
  if (S < 0 || n_shards < 1 || !dt || !perm || !bounds) return fail(-2, "bad arguments to gab1_deal_shards");
  std::vector<std::pair<double, int64_t>> w((size_t)S);
  for (int64_t i = 0; i < S; ++i) {
    const double n = ceil(tf / dt[i]);
    w[(size_t)i] = {(n > 0.0 && n < 4.0e9) ? n : 1.0, i};
  }
  std::stable_sort(w.begin(), w.end(), [](const std::pair<double, int64_t>& a, const std::pair<double, int64_t>& b) { return a.first > b.first; });
  std::vector<int32_t> shard_of((size_t)S);
  std::vector<int64_t> count((size_t)n_shards, 0);
  std::vector<double> load((size_t)n_shards, 0.0);
  for (int64_t p = 0; p < S; ++p) {
    // longest-processing-time rule: the next-longest solve goes to the least-loaded shard (ties: the fewest sets, then
    // the lowest ordinal); with thousands of sets per shard the totals end within one short solve of each other
    int32_t g = 0;
    for (int32_t c = 1; c < n_shards; ++c)
      if (load[(size_t)c] < load[(size_t)g] || (load[(size_t)c] == load[(size_t)g] && count[(size_t)c] < count[(size_t)g])) g = c;
    shard_of[(size_t)w[(size_t)p].second] = g;
    load[(size_t)g] += w[(size_t)p].first;
    ++count[(size_t)g];
  }
  bounds[0] = 0;
  for (int g = 0; g < n_shards; ++g) bounds[g + 1] = bounds[g] + count[(size_t)g];
  std::vector<int64_t> fill(bounds, bounds + n_shards);
  for (int64_t i = 0; i < S; ++i) perm[fill[(size_t)shard_of[(size_t)i]]++] = i;
  return 0;
}

// contiguous shards with (nearly) equal total step count, so that each device's gather is one copy
int gab1_plan_shards(int64_t S, const double* dt, double tf, int32_t n_shards, int64_t* bounds) {
  if (S < 0 || n_shards < 1 || !dt || !bounds) return fail(-2, "bad arguments to gab1_plan_shards");
  std::vector<double> work((size_t)S);
  double total = 0.0;
  for (int64_t i = 0; i < S; ++i) {
    const double n = ceil(tf / dt[i]);
    work[i] = (n > 0.0 && n < 4.0e9) ? n : 1.0;
    total += work[i];
  }
  bounds[0] = 0;
  double acc = 0.0;
  int g = 1;
  for (int64_t i = 0; i < S && g < n_shards; ++i) {
    acc += work[i];
    while (g < n_shards && acc >= total * g / n_shards) bounds[g++] = i + 1;
  }
  for (; g < n_shards; ++g) bounds[g] = S;
  bounds[n_shards] = S;
  return 0;
}

void* gab1_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
    fail(-8, "cudaHostAlloc(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}

// The NUMA node a GPU hangs off and that node's CPUs (sysfs); false when the platform does not say.
static bool cpus_near_device(int device, cpu_set_t* set, int* node_out) {
  char bus[32] = "";
  if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { (void)cudaGetLastError(); return false; }
  for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
  char path[256];
  snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE* f = fopen(path, "r");
  if (!f) return false;
  int node = -1;
  const int got = fscanf(f, "%d", &node);
  fclose(f);
  if (got != 1 || node < 0) return false;
  snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
  f = fopen(path, "r");
  if (!f) return false;
  char list[4096] = "";
  const bool ok = fgets(list, sizeof list, f) != nullptr;
  fclose(f);
  if (!ok) return false;
  CPU_ZERO(set);
  int n = 0;
  for (char* tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
    int a = 0, b = 0;
    const int k = sscanf(tok, "%d-%d", &a, &b);
    if (k == 1) b = a;
    if (k < 1) continue;
    for (int c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(c, set); ++n; }
  }
  if (node_out) *node_out = node;
  return n > 0;
}

// Pinned, mapped host memory whose pages live on the NUMA node of `device`: the calling thread is moved onto that node's
// CPUs while the pages are allocated and touched (Linux allocates on the toucher's node), then put back.  On a two-socket
// 8-GPU box every GPU then streams its snapshots into its own socket's DRAM instead of across the inter-socket link.
void* gab1_host_alloc_near(size_t bytes, int32_t device) {
  cpu_set_t old_set, near_set;
  const char* e = getenv("GAB1_NUMA");
  const bool want = !(e && e[0] == '0');
  bool moved = false;
  if (want && sched_getaffinity(0, sizeof old_set, &old_set) == 0 && cpus_near_device(device, &near_set, nullptr)) {
    cpu_set_t both;
    CPU_AND(&both, &old_set, &near_set);
    if (CPU_COUNT(&both) > 0) moved = sched_setaffinity(0, sizeof both, &both) == 0;
  }
  void* p = gab1_host_alloc(bytes);
  if (p && moved) {
    volatile char* c = (volatile char*)p;
    for (size_t i = 0; i < bytes; i += 4096) c[i] = 0;
  }
  if (moved) sched_setaffinity(0, sizeof old_set, &old_set);
  return p;
}

int gab1_device_numa_node(int32_t device) {
  cpu_set_t set;
  int node = -1;
  return cpus_near_device(device, &set, &node) ? node : -1;
}
void gab1_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

double gab1_measure_fp64_tflops(int32_t device, double seconds) {
  if (cudaSetDevice(device) != cudaSuccess) { fail(-7, "cudaSetDevice(%d) failed", device); return -1.0; }
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  double* sink = nullptr;
  if (cudaMalloc(&sink, 8) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = sms * 8, threads = 256;
  int iters = 2000;
  dfma_peak_kernel<<<blocks, threads>>>(sink, iters, 1.0000001, 1e-9);   // warm-up
  g_launches.fetch_add(1);
  cudaDeviceSynchronize();
  double best = 0.0, spent = 0.0;
  while (spent < seconds) {
    cudaEventRecord(e0);
    dfma_peak_kernel<<<blocks, threads>>>(sink, iters, 1.0000001, 1e-9);
    g_launches.fetch_add(1);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 8 * 16 * (double)iters * blocks * threads;
    const double tf = flop / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
    spent += ms * 1e-3;
    if (ms < 20.f) iters *= 2;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(sink);
  return best;
}

int64_t gab1_kernel_launches(void) { return g_launches.load(); }

int gab1_debug_recip_error(int32_t device, double lo, double hi, double* seed_err, double* recip_err) {
  CUDA_TRY(cudaSetDevice(device));
  double* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, 16));
  CUDA_TRY(cudaMemset(d, 0, 16));
  gab1::launch_recip_error_kernel(lo, hi, 1 << 22, d);
  double h[2] = {0, 0};
  CUDA_TRY(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
  cudaFree(d);
  if (seed_err) *seed_err = h[0];
  if (recip_err) *recip_err = h[1];
  return 0;
}

int gab1_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}
int gab1_version(void) { return GAB1_ABI_VERSION; }
const char* gab1_last_error(void) { return g_err; }

}  // extern "C"
