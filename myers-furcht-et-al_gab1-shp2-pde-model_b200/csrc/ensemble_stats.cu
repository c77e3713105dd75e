// ensemble_stats.cu — order statistics across the parameter sets of a FULL result, on the device.
//
// The reference's figure scripts stack the full solutions of an ensemble and take, node by node and snapshot by snapshot,
// `median(stack, dims=3)` and `quantile(stack[node, end, :], 0.5 -+ 0.341)` (run_base_model.jl:103-174).  Done here on the
// device, a caller receives a few surfaces (KBs) instead of 0.5 MB of snapshots per set.
//
// One CTA sorts, in shared memory, the values of all valid sets at NT neighbouring nodes of one snapshot column of one
// matrix (bitonic network, padded with +Inf to a power of two) and reads the requested order statistics off the sorted rows
// with the definitions of Julia's Statistics standard library (the reference's Manifest pins Julia 1.8.3; the stdlib source
// is not part of the reference tree, so this follows its published algorithm — parity unpinned at this boundary):
//   quantile(v, p) (alpha = beta = 1):  aleph = n*p + (1 + p*(1 - 1 - 1)), j = clamp(trunc(aleph), 1, n-1),
//                    g = clamp(aleph - j, 0, 1), q = v[j] + g*(v[j+1] - v[j])   (1-based; n = 1: q = v[1])
//   median(v):       v[(n+1)/2] for odd n, middle(a, b) = a/2 + b/2 of the two central values for even n
// Sets whose status carries GAB1_ST_NAN are left out, as run_ensemble drops them (get_param_posteriors.jl:155).
#include <cuda_runtime.h>
#include <math_constants.h>

#include "launch.h"

namespace gab1 {
namespace {

constexpr int QT = 1024;         // threads per CTA
constexpr int MAX_P = 8;         // probabilities per call
constexpr int MAX_M = 12;        // matrices per call

struct QArgs {
  const double* out; long long out_stride;
  long long moff[MAX_M]; int nm;          // offsets of the selected matrices inside a set's block
  int P, Cn, c0, ncols;
  const int* valid; const long long* n_valid;
  int N2, NT;
  int np; double p[MAX_P];                // p < 0: median rule
  double* q;                              // [matrix][p][column][node]
};

// compacted, ordered list of the sets without GAB1_ST_NAN (single CTA)
__global__ void __launch_bounds__(QT) valid_list_kernel(long long S, const int* status, int* valid, long long* n_valid) {
  __shared__ long long base[QT + 1];
  const long long chunk = (S + QT - 1) / QT, lo = threadIdx.x * chunk, hi = lo + chunk < S ? lo + chunk : S;
  long long cnt = 0;
  for (long long i = lo; i < hi; ++i) cnt += (status[i] & (int)GAB1_ST_NAN) == 0;
  base[threadIdx.x + 1] = cnt;
  if (threadIdx.x == 0) base[0] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int t = 1; t <= QT; ++t) base[t] += base[t - 1];
    *n_valid = base[QT];
  }
  __syncthreads();
  long long at = base[threadIdx.x];
  for (long long i = lo; i < hi; ++i)
    if ((status[i] & (int)GAB1_ST_NAN) == 0) valid[at++] = (int)i;
}

__global__ void __launch_bounds__(QT) quantile_kernel(const QArgs a) {
  extern __shared__ double rows[];                 // NT rows of N2 values
  const int tile = blockIdx.x, col = a.c0 + blockIdx.y, mi = blockIdx.z;
  const long long nv = *a.n_valid;
  const int N2 = a.N2, NT = a.NT, node0 = tile * NT;
  const double* src = a.out + a.moff[mi] + (long long)col * a.P + node0;
  for (long long v = threadIdx.x; v < N2; v += QT) {
    const bool in = v < nv;
    const double* s = in ? src + (long long)a.valid[v] * a.out_stride : nullptr;
    for (int n = 0; n < NT; ++n) rows[(long long)n * N2 + v] = (in && node0 + n < a.P) ? s[n] : CUDART_INF;
  }
  __syncthreads();
  // bitonic sorting network, all NT rows in the same passes
  const int half = N2 >> 1;
  for (int k = 2; k <= N2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < NT * half; t += QT) {
        const int row = t / half, i = t - row * half;
        const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1)), hi = lo | j;
        double* r = rows + (long long)row * N2;
        const double x = r[lo], y = r[hi];
        const bool up = (lo & k) == 0;
        if ((x > y) == up) { r[lo] = y; r[hi] = x; }
      }
      __syncthreads();
    }
  }
  for (int t = threadIdx.x; t < NT * a.np; t += QT) {
    const int n = t / a.np, pi = t - n * a.np;
    if (node0 + n >= a.P) continue;
    const double* r = rows + (long long)n * N2;
    double q;
    if (nv <= 0) q = CUDART_NAN;
    else if (a.p[pi] < 0.0) {                                   // median(v)
      q = (nv & 1) ? r[(nv - 1) / 2] : __dadd_rn(__ddiv_rn(r[nv / 2 - 1], 2.0), __ddiv_rn(r[nv / 2], 2.0));
    } else if (nv == 1) {
      q = r[0];
    } else {                                                    // quantile(v, p); j is 1-based as in the source
      const double pp = a.p[pi];
      const double aleph = __dadd_rn(__dmul_rn((double)nv, pp), __dadd_rn(1.0, __dmul_rn(pp, -1.0)));
      long long j = (long long)trunc(aleph);
      j = j < 1 ? 1 : (j > nv - 1 ? nv - 1 : j);
      double g = __dsub_rn(aleph, (double)j);
      g = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g);
      const double va = r[j - 1], vb = r[j];
      q = __dadd_rn(va, __dmul_rn(g, __dsub_rn(vb, va)));
    }
    a.q[(((long long)mi * a.np + pi) * a.ncols + (col - a.c0)) * a.P + node0 + n] = q;
  }
}

}  // namespace

size_t quantiles_workspace_bytes(long long S) { return 256 + sizeof(int) * (size_t)(S < 1 ? 1 : S); }

// All pointers are device pointers.  q: nm x np x (c1-c0) x (Nr+1) doubles.  Returns 0 or a negative code (gab1_last_error).
int ensemble_quantiles_device(const gab1_opts* o, int device, cudaStream_t stream, long long S, const double* out,
                              const int* status, unsigned matrices, int c0, int c1, int np, const double* p, double* q,
                              long long* n_valid, void* workspace) {
  if (!o || o->out_mode != GAB1_OUT_FULL) return fail(-2, "ensemble quantiles need a GAB1_OUT_FULL result");
  if (np < 1 || np > MAX_P) return fail(-2, "between 1 and %d probabilities per call", MAX_P);
  if (c0 < 0 || c1 > o->Nts + 1 || c0 >= c1) return fail(-2, "column range [%d, %d) outside 0..Nts+1", c0, c1);
  if (!out || !status || !q || !p || !workspace || !n_valid) return fail(-2, "a required buffer is NULL");
  if (matrices == 0 || (matrices & ~o->matrix_mask)) return fail(-2, "a requested matrix is not part of the result (matrix_mask)");
  if (S < 1) return fail(-2, "S must be >= 1");
  CUDA_TRY(cudaSetDevice(device));
  QArgs a;
  a.out = out;
  a.out_stride = gab1_out_doubles_per_set(o);
  a.P = o->Nr + 1; a.Cn = o->Nts + 1; a.c0 = c0; a.ncols = c1 - c0;
  a.nm = 0;
  for (int m = 0; m < GAB1_N_MATRICES; ++m)
    if ((matrices >> m) & 1u) a.moff[a.nm++] = gab1_full_matrix_offset(o, m);
  a.np = np;
  for (int i = 0; i < np; ++i) {
    if (p[i] > 1.0) return fail(-2, "probability %g outside [0, 1]", p[i]);
    a.p[i] = p[i];
  }
  int N2 = 2;
  while (N2 < S) N2 <<= 1;
  const size_t budget = 200 * 1024;
  if ((size_t)N2 * sizeof(double) > budget)
    return fail(-6, "ensemble quantiles sort one row of all sets in shared memory, padded to a power of two: at most 16384 sets per call "
                    "(%lld requested)", S);
  int NT = (int)(budget / ((size_t)N2 * sizeof(double)));
  if (NT > 4) NT = 4;
  a.N2 = N2; a.NT = NT;
  a.valid = (int*)((char*)workspace + 256);
  a.n_valid = (long long*)workspace;
  a.q = q;
  valid_list_kernel<<<1, QT, 0, stream>>>(S, status, (int*)a.valid, (long long*)a.n_valid);
  count_launch();
  const size_t smem = (size_t)NT * N2 * sizeof(double);
  static bool attr_set[64] = {false};
  if (device < 64 && !attr_set[device]) {
    CUDA_TRY(cudaFuncSetAttribute(quantile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set[device] = true;
  }
  dim3 grid((a.P + NT - 1) / NT, a.ncols, a.nm);
  quantile_kernel<<<grid, QT, smem, stream>>>(a);
  count_launch();
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(n_valid, a.n_valid, sizeof(long long), cudaMemcpyDeviceToDevice, stream));
  return 0;
}

}  // namespace gab1
