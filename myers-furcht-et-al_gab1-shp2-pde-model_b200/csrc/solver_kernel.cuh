// solver_kernel.cuh — the hot path: explicit finite-difference time stepping of the EGFR/GRB2/GAB1/SHP2/SFK
// reaction–diffusion model, one warp per parameter set, hand-written for sm_100a.
//
// Mapping.  Grid nodes 1..Nr (node 0 mirrors node 1 exactly because of the zero-flux condition, so it is never
// stored) are dealt to the lanes of ONE warp in contiguous runs of K nodes; all ten cytosolic species of a node live
// in that lane's registers for the whole solve.  Neighbour values cross lanes with two 64-bit shuffles per species per
// step.  The membrane fixed point (Robin closures + eight membrane ODEs) runs lane-parallel: lanes 0-9 each own one
// closure, lanes 10-17 one membrane species.  The time loop (3e4..2e6 steps) never touches global memory except for
// the requested snapshot columns, which are staged through shared memory and stored fully coalesced.
//
// Two arithmetic modes (gab1_opts.arith):
//   strict — every operation of the reference in source order with IEEE round-to-nearest intrinsics (never
//            contracted); bit-identical to the CPU oracle; used to pin the control flow on the GPU.
//   fast   — the product path: reciprocals hoisted, rate constants pre-scaled by dt, mass-action terms written as
//            net fluxes, FMA contraction; agrees with strict to ~1e-13 relative (bound asserted in tests: 1e-9).
//
// Reference: basepdesolver.jl:149-296 (time loop), :150-180 (interior), :183-192 (r=0), :197-242 (membrane loop),
// :265-295 (snapshots); basepdesolver_rect.jl:131-161; sapdesolver.jl:128-242; sapdesolver_memb-SFK.jl:175-222;
// pulsechase_solver.jl:156-158.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "gab1pde.h"
#include "launch.h"

namespace gab1 {

enum { iSFK, aSFK, GAB1, pGAB1, GRB2, G2G1, G2PG1, SHP2, PG1S, G2PG1S, NCY };   // basepdesolver.jl:199-202
enum { mE, mES, mESmES, E, EG2, EG2G1, EG2PG1, EG2PG1S, NMB };                  // basepdesolver.jl:203

constexpr unsigned FULL = 0xffffffffu;
constexpr int WS_HDR = 32;          // per-warp smem header: [0,16) interior-neighbour stage, [16,32) boundary stage
constexpr int ML = 10;              // first membrane lane
// Halo exchange area of the fast one-set-per-warp kernel with K >= 4 nodes per lane: the interior halo crosses lanes
// through shared memory (20 STS.64 + 20 LDS.64 per step) instead of 40 32-bit shuffles plus the register moves that
// re-pair their halves.  Measured on B200: K = 4 (dr = 0.1) 378.9 vs 388.9 ms; K = 2 (dr = 0.2) 126.8 vs 117.2 ms — with
// two nodes per lane there is too little independent work to cover store -> warp barrier -> load, so K <= 2 keeps shuffles.
// Round 2 tried the software-pipelined form (edge nodes stored behind the interior's barrier, i.e. inside the membrane block,
// and loaded at the top of the NEXT step with 128-bit accesses; boundary values and the old-time membrane values likewise):
// 12 % fewer instructions per step at K = 2 (278 against 317) and yet slower, 122.9 vs 116.2 ms / 432.7 vs 408.6 ms (5000
// rows / 20 000 prior draws): every halo value crosses the shared-memory data pipe twice (store + load; a shuffle moves it
// once), which takes that pipe from 46 % to ~75 % busy.  At K = 4 it measured equal to the form below (355.5 ms both, same
// session) — so it was removed again (profiles/r2x_halo_pipe_timings.txt, r2y_ab_*.txt).  Serving only the three old-time
// membrane values of the prologue from a shared x[lane] block (3 LDS.64 for 6 SHFL) does not pay either: ptxas answers with
// more moves than the shuffles had (413 against 409 instructions in the loop).
constexpr int WS_EX = 2 * 10 * 32;  // doubles per warp: [species][lane] of the last node, then of the first node


// ---------------------------------------------------------------------------------------------------------------
// strict arithmetic: a double whose operators are single IEEE operations that ptxas may not contract
struct sd {
  double v;
  __device__ __forceinline__ sd() {}
  __device__ __forceinline__ sd(double x) : v(x) {}
};
__device__ __forceinline__ sd operator+(sd a, sd b) { return sd(__dadd_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator-(sd a, sd b) { return sd(__dsub_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator*(sd a, sd b) { return sd(__dmul_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator/(sd a, sd b) { return sd(__ddiv_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator-(sd a) { return sd(-a.v); }

// 64-bit shuffles spelled as two 32-bit ones: with the built-in double overload ptxas lands the halves in swapped
// registers and repairs each shuffle with three XORs (ncu profile r1b: 60 LOP3 per step)
__device__ __forceinline__ double shfl(double x, int src) {
  const int lo = __shfl_sync(FULL, __double2loint(x), src), hi = __shfl_sync(FULL, __double2hiint(x), src);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_up1(double x) {
  const int lo = __shfl_up_sync(FULL, __double2loint(x), 1), hi = __shfl_up_sync(FULL, __double2hiint(x), 1);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_down1(double x) {
  const int lo = __shfl_down_sync(FULL, __double2loint(x), 1), hi = __shfl_down_sync(FULL, __double2hiint(x), 1);
  return __hiloint2double(hi, lo);
}
// shared memory through a 32-bit shared-window address: one register, no generic-address arithmetic in the time loop
__device__ __forceinline__ double lds(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts(unsigned addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void lds2(unsigned addr, double& v0, double& v1) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sts2(unsigned addr, double v0, double v1) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v0), "d"(v1) : "memory");
}
__device__ __forceinline__ bool is_special(double x) {   // zero, Inf or NaN — decided on the bit pattern, off the FP64 pipe
  const unsigned hi = (unsigned)__double2hiint(x) & 0x7fffffffu;
  const unsigned lo = (unsigned)__double2loint(x);
  return hi >= 0x7ff00000u || (hi | lo) == 0u;
}

// |1 - new/old| against tol exactly as the reference evaluates it (basepdesolver.jl:238-239).
// 0: <= tol, 1: > tol, 2: NaN (Julia's maximum would return NaN).
__device__ __forceinline__ int classify_exact(double oldv, double newv, double tol) {
  const double e = fabs(__dsub_rn(1.0, __ddiv_rn(newv, oldv)));
  return isnan(e) ? 2 : (e <= tol ? 0 : 1);
}

// Layout: the warp's 32*K slots are RIGHT-aligned on the grid, so that the boundary node Nr is always the last slot of
// lane G-1 (G = ceil(Nr/K)) and its inner neighbour the slot before it — compile-time register indices.  Slots left
// of node 1 are padding that stays zero.
template <int K>
struct Grid {          // per-lane constants of the radial grid, shared by every parameter set
  double a[K];         // strict: 1/(r_j*dr)                         (basepdesolver.jl:151)
  double cp[K], cm[K], c0[K];   // fast: lap = cp*u[j+1] + cm*u[j-1] + c0*u[j]; node 1 folds its mirror u[0]=u[1] into c0;
                                // all three are 0 on padding and on the boundary slot
  int node[K];         // node number: node n is Julia's index n+1; valid nodes are 1..Nr
  bool interior[K];    // 1 <= node <= Nr-1
  int G;               // lanes in use
};

// ---------------------------------------------------------------------------------------------------------------
// Output writers (rare path).  Rows are staged in shared memory so that global stores are unit-stride.
template <int K, typename F>
__device__ __forceinline__ void stage_row(double* row, int lane, const Grid<K>& g, int Nr, F val) {
  (void)lane;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    if (g.node[i] >= 1 && g.node[i] <= Nr) row[g.node[i]] = val(i);
    if (g.node[i] == 1) row[0] = val(i);   // node 0 == node 1 (basepdesolver.jl:183-192)
  }
  __syncwarp();
}
// 16-byte stores: a lane writes two consecutive nodes with one st.global.v2.f64 (a 51-node column is 26 lanes' worth instead of
// two passes of 8-byte stores); a column that starts on an odd double (every other one when P is odd) sheds its first node
// as a scalar so that the pairs are aligned.  The same code serves device memory and mapped host memory.
__device__ __forceinline__ void stg2(double* p, double v0, double v1) {
  asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v0), "d"(v1) : "memory");
}
template <typename F>
__device__ __forceinline__ bool store_row_v2(double* dst, int P, int lane, F val) {
  bool nan_seen = false;
  const int head = (int)((reinterpret_cast<unsigned long long>(dst) >> 3) & 1ull);
  const int npairs = (P - head) >> 1;
  for (int j = lane; j < npairs; j += 32) {
    const int n = head + 2 * j;
    const double v0 = val(n), v1 = val(n + 1);
    nan_seen |= isnan(v0) || isnan(v1);
    stg2(dst + n, v0, v1);
  }
  if (lane == 31) {
    if (head) { const double v = val(0); nan_seen |= isnan(v); dst[0] = v; }
    const int tail = head + 2 * npairs;
    if (tail < P) { const double v = val(tail); nan_seen |= isnan(v); dst[tail] = v; }
  }
  return __any_sync(FULL, nan_seen);
}
__device__ __forceinline__ bool flush_row(double* dst, const double* row, int P, int lane) {
  const bool nan_seen = store_row_v2(dst, P, lane, [&](int n) { return row[n]; });
  __syncwarp();
  return nan_seen;
}

template <int K>
__device__ __forceinline__ double species_at(const double (&u)[NCY][K], int q, int i) {
  // q is a compile-time constant at every call site after unrolling
  return u[q][i];
}

template <int K>
__device__ __forceinline__ double derived_stot(const double (&u)[NCY][K], int i) {
  return __dadd_rn(u[PG1S][i], u[G2PG1S][i]);                              // basepdesolver.jl:299
}
template <int K>
__device__ __forceinline__ double derived_ptot(const double (&u)[NCY][K], int i, int form) {
  if (form == GAB1_PG1TOT_VIA_STOT)                                        // basepdesolver.jl:300
    return __dadd_rn(__dadd_rn(u[G2PG1][i], u[pGAB1][i]), derived_stot<K>(u, i));
  return __dadd_rn(__dadd_rn(__dadd_rn(u[G2PG1][i], u[pGAB1][i]), u[PG1S][i]), u[G2PG1S][i]);   // basepdesolver_rect.jl:261
}

// one snapshot column of GAB1_OUT_FULL (basepdesolver.jl:268-294)
template <int K>
__device__ void write_full_column(const KernelArgs& a, double* oset, int c, const double (&u)[NCY][K],
                                  const double (&m)[NMB], double t, double CoEGFR, int lane, const Grid<K>& g,
                                  double* row, unsigned& status) {
  const int Nr = a.o.Nr, P = Nr + 1;
  const long long Cn = a.o.Nts + 1;
  const unsigned mask = a.o.matrix_mask;
  long long off = 0;
  constexpr int kSpecies[10] = {iSFK, aSFK, GRB2, GAB1, SHP2, G2G1, G2PG1, G2PG1S, pGAB1, PG1S};   // basepdesolver.jl:271-280
#pragma unroll
  for (int mi = 0; mi < 12; ++mi) {
    if (!((mask >> mi) & 1u)) continue;
    if (mi < 10) {
      const int q = kSpecies[mi];
      stage_row<K>(row, lane, g, Nr, [&](int i) { return u[q][i]; });
    } else if (mi == GAB1_M_PG1tot) {
      stage_row<K>(row, lane, g, Nr, [&](int i) { return derived_ptot<K>(u, i, a.o.pg1tot_form); });
    } else {
      stage_row<K>(row, lane, g, Nr, [&](int i) { return derived_stot<K>(u, i); });
    }
    const bool nan_seen = flush_row(oset + off + (long long)c * P, row, P, lane);
    if (mi == GAB1_M_PG1S && nan_seen) status |= GAB1_ST_NAN;
    off += (long long)P * Cn;
  }
  if (!((mask >> GAB1_M_PG1S) & 1u)) {       // the NaN filter looks at PG1S whether or not it is materialised
    bool ns = false;
#pragma unroll
    for (int i = 0; i < K; ++i) ns |= (g.node[i] >= 1 && g.node[i] <= Nr) && isnan(u[PG1S][i]);
    if (__any_sync(FULL, ns)) status |= GAB1_ST_NAN;
  }
  if (lane == 0) {
    double* v = oset + off;
    const double Etot = __dmul_rn(2.0, __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(m[E], m[EG2]), m[EG2G1]), m[EG2PG1]), m[EG2PG1S]));  // :263
    v[GAB1_V_pE * Cn + c] = __ddiv_rn(__dmul_rn(Etot, 100.0), CoEGFR);                     // :287
    v[GAB1_V_mE * Cn + c] = m[mE];
    v[GAB1_V_mES * Cn + c] = m[mES];
    v[GAB1_V_mESmES * Cn + c] = m[mESmES];
    v[GAB1_V_E * Cn + c] = m[E];
    v[GAB1_V_EG2 * Cn + c] = m[EG2];
    v[GAB1_V_EG2G1 * Cn + c] = m[EG2G1];
    v[GAB1_V_EG2PG1 * Cn + c] = m[EG2PG1];
    v[GAB1_V_EG2PG1S * Cn + c] = m[EG2PG1S];
    v[GAB1_V_EGFR_SHP2 * Cn + c] = __ddiv_rn(__dmul_rn(m[EG2PG1S], 100.0), CoEGFR);        // basepdesolver_rect.jl:264
    v[GAB1_V_t_out * Cn + c] = t;
  }
}

// NumericalIntegration.integrate(r, y .* r.^2): trapezoid, left-to-right, on a staged row (every lane computes it)
__device__ __forceinline__ double trapz_r2(const double* r, const double* y, int P) {
  double acc = 0.0;
  double yi = __dmul_rn(y[0], __dmul_rn(r[0], r[0]));
  for (int i = 0; i + 1 < P; ++i) {
    const double yn = __dmul_rn(y[i + 1], __dmul_rn(r[i + 1], r[i + 1]));
    acc = __dadd_rn(acc, __dmul_rn(__dsub_rn(r[i + 1], r[i]), __dadd_rn(yi, yn)));
    yi = yn;
  }
  return __dmul_rn(0.5, acc);
}
// R - minimum(r[y .>= f*maximum(y)])  (sapdesolver.jl:344-347)
__device__ __forceinline__ double length_scale(const double* r, const double* y, int P, double f, double R, bool& threw) {
  double mx = y[0];
  bool nan_seen = isnan(y[0]);
  for (int i = 1; i < P; ++i) {
    if (isnan(y[i])) nan_seen = true;
    else if (!(mx >= y[i])) mx = y[i];
  }
  if (nan_seen) mx = CUDART_NAN;
  const double thr = __dmul_rn(f, mx);
  for (int i = 0; i < P; ++i)
    if (y[i] >= thr) return __dsub_rn(R, r[i]);
  threw = true;
  return 0.0;
}

// final-time outputs: FINAL4 (sapdesolver.jl:245-279), SIX (sapdesolver.jl:343-356), FINAL_STATE
template <int K>
__device__ void write_final(const KernelArgs& a, double* oset, const double (&u)[NCY][K], const double (&m)[NMB],
                            int lane, const Grid<K>& g, double* rowA, double* rowB, unsigned& status) {
  const int Nr = a.o.Nr, P = Nr + 1;
  if (a.o.out_mode == GAB1_OUT_FINAL4) {
    bool ns = false;
    stage_row<K>(rowA, lane, g, Nr, [&](int i) { return u[iSFK][i]; });
    ns |= flush_row(oset, rowA, P, lane);
    stage_row<K>(rowA, lane, g, Nr, [&](int i) { return u[aSFK][i]; });
    ns |= flush_row(oset + P, rowA, P, lane);
    stage_row<K>(rowA, lane, g, Nr, [&](int i) { return derived_ptot<K>(u, i, a.o.pg1tot_form); });
    ns |= flush_row(oset + 2 * P, rowA, P, lane);
    stage_row<K>(rowA, lane, g, Nr, [&](int i) { return derived_stot<K>(u, i); });
    ns |= flush_row(oset + 3 * P, rowA, P, lane);
    if (ns) status |= GAB1_ST_NAN;
  } else if (a.o.out_mode == GAB1_OUT_FINAL_STATE) {
    bool ns = false;
#pragma unroll
    for (int q = 0; q < NCY; ++q) {
      stage_row<K>(rowA, lane, g, Nr, [&](int i) { return u[q][i]; });
      ns |= flush_row(oset + (long long)q * P, rowA, P, lane);
    }
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < NMB; ++j) { oset[(long long)NCY * P + j] = m[j]; ns |= isnan(m[j]); }
    }
    if (__any_sync(FULL, ns)) status |= GAB1_ST_NAN;
  } else if (a.o.out_mode == GAB1_OUT_SIX) {
    stage_row<K>(rowA, lane, g, Nr, [&](int i) { return u[aSFK][i]; });
    stage_row<K>(rowB, lane, g, Nr, [&](int i) { return derived_stot<K>(u, i); });
    bool threw = false;
    double six[6];
    const double R = a.o.R;
    six[0] = length_scale(a.r, rowA, P, 0.5, R, threw);
    six[1] = length_scale(a.r, rowA, P, 0.1, R, threw);
    six[2] = length_scale(a.r, rowB, P, 0.5, R, threw);
    six[3] = length_scale(a.r, rowB, P, 0.1, R, threw);
    six[4] = __ddiv_rn(rowB[0], rowB[P - 1]);
    six[5] = __ddiv_rn(__dmul_rn(trapz_r2(a.r, rowB, P), 3.0), a.R_pow3);
    __syncwarp();
    if (threw) status |= GAB1_ST_THROW;
    bool ns = false;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double v = threw ? 0.0 : six[i];
      ns |= isnan(v);
      if (lane == 0) oset[i] = v;
    }
    if (ns) status |= GAB1_ST_NAN;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Interior update, strict: the reference's expressions, term by term (basepdesolver.jl:151-179, rect :132-160)
struct Rates {   // basepdesolver.jl:52-68
  double kS2f, kS2r, kG1f, kG1r, kG2f, kG2r, kG1p, kG1dp, kSa, kSi, kp, kdp, kEGFf, kEGFr, EGF, kdf, kdr;
};

__device__ __forceinline__ void kinetics_strict(const sd (&L)[NCY], const double (&c)[NCY], const Rates& k, sd dt, double (&w)[NCY]) {
  const sd Si = c[iSFK], Sa = c[aSFK], G1 = c[GAB1], pG1 = c[pGAB1], G2 = c[GRB2], g2g1 = c[G2G1], g2pg1 = c[G2PG1],
           S2 = c[SHP2], pg1s = c[PG1S], g2pg1s = c[G2PG1S];
  const sd kS2f = k.kS2f, kS2r = k.kS2r, kG1f = k.kG1f, kG1r = k.kG1r, kG1p = k.kG1p, kG1dp = k.kG1dp, kSi = k.kSi;
  w[iSFK] = ((L[iSFK] + kSi * Sa) * dt + Si).v;
  w[aSFK] = ((L[aSFK] - kSi * Sa) * dt + Sa).v;
  w[GAB1] = ((L[GAB1] - kG1f * G1 * G2 + kG1r * g2g1 - kG1p * Sa * G1 + kG1dp * pG1) * dt + G1).v;
  w[pGAB1] = ((L[pGAB1] - kG1f * pG1 * G2 + kG1r * g2pg1 + kG1p * Sa * G1 - kG1dp * pG1 - kS2f * S2 * pG1 + kS2r * pg1s) * dt + pG1).v;
  w[GRB2] = ((L[GRB2] - kG1f * G1 * G2 + kG1r * g2g1 - kG1f * pG1 * G2 + kG1r * g2pg1 - kG1f * G2 * pg1s + kG1r * g2pg1s) * dt + G2).v;
  w[G2G1] = ((L[G2G1] + kG1f * G1 * G2 - kG1r * g2g1 - kG1p * Sa * g2g1 + kG1dp * g2pg1) * dt + g2g1).v;
  w[G2PG1] = ((L[G2PG1] + kG1f * pG1 * G2 - kG1r * g2pg1 + kG1p * Sa * g2g1 - kG1dp * g2pg1 - kS2f * S2 * g2pg1 + kS2r * g2pg1s) * dt + g2pg1).v;
  w[SHP2] = ((L[SHP2] - kS2f * S2 * pG1 + kS2r * pg1s - kS2f * S2 * g2pg1 + kS2r * g2pg1s) * dt + S2).v;
  w[PG1S] = ((L[PG1S] + kS2f * S2 * pG1 - kS2r * pg1s - kG1f * G2 * pg1s + kG1r * g2pg1s) * dt + pg1s).v;
  w[G2PG1S] = ((L[G2PG1S] + kG1f * G2 * pg1s - kG1r * g2pg1s + kS2f * S2 * g2pg1 - kS2r * g2pg1s) * dt + g2pg1s).v;
}

// Membrane fixed point, strict, evaluated identically by every lane (basepdesolver.jl:197-242).
// b[] holds the boundary values u[Nr+1,2]; m1/m2 the membrane columns [1]/[2]. Returns the iteration count.
__device__ inline int membrane_strict(const KernelArgs& a, const Rates& k, double kp_now, const double (&Dsp)[NCY],
                               const double (&In)[NCY], double (&b)[NCY], const double (&m1)[NMB], double (&m2)[NMB],
                               double dt_, unsigned& status, bool& unconverged) {
  const sd dr = a.o.dr, dt = dt_, one = 1.0, two = 2.0;
  const sd kS2f = k.kS2f, kS2r = k.kS2r, kG1f = k.kG1f, kG1r = k.kG1r, kG2f = k.kG2f, kG2r = k.kG2r, kSa = k.kSa,
           kp = kp_now, kdp = k.kdp, kEGFf = k.kEGFf, kEGFr = k.kEGFr, EGF = k.EGF, kdf = k.kdf, kdr = k.kdr;
  const sd D_Si = Dsp[iSFK], D_Sa = Dsp[aSFK], D_G1 = Dsp[GAB1], D_G2 = Dsp[GRB2], D_G2G1 = Dsp[G2G1], D_S2 = Dsp[SHP2],
           D_G1S2 = Dsp[PG1S], D_G2G1S2 = Dsp[G2PG1S];
  const double tol = a.o.tol;
  double err = __dmul_rn(tol, 2.0);
  int it = 0;
  unconverged = false;
  for (;;) {
    if (a.o.bc_loop == GAB1_BC_FOR_BREAK) { if (it >= a.o.maxiters) { unconverged = true; break; } }
    else { if (!(err > tol)) break; if (it >= a.o.maxiters) { status |= GAB1_ST_ITER_CAP; break; } }
    ++it;
    double cold[NCY], mold[NMB];
#pragma unroll
    for (int q = 0; q < NCY; ++q) cold[q] = b[q];
#pragma unroll
    for (int j = 0; j < NMB; ++j) mold[j] = m2[j];
    const sd mE2 = m2[E], mEG2 = m2[EG2], mEG2G1 = m2[EG2G1], mEG2PG1 = m2[EG2PG1], mEG2PG1S = m2[EG2PG1S];
    const sd Etot = two * (mE2 + mEG2 + mEG2G1 + mEG2PG1 + mEG2PG1S);                                  // :205
    const sd bi = sd(In[iSFK]) / (one + kSa * Etot * dr / D_Si);                                      // :206
    const sd ba = sd(In[aSFK]) + kSa * bi * Etot * dr / D_Sa;                                         // :207
    const sd bG1 = (kG1r * mEG2G1 * dr / D_G1 + sd(In[GAB1])) / (one + kG1f * mEG2 * dr / D_G1);      // :208
    const sd bpG1 = (kG1r * mEG2PG1 * dr / D_G1 + sd(In[pGAB1])) / (one + kG1f * mEG2 * dr / D_G1);   // :209
    const sd bG2 = (kG2r * mEG2 * dr / D_G2 + sd(In[GRB2])) / (one + kG2f * mE2 * dr / D_G2);         // :210
    const sd bg2g1 = (kG2r * mEG2G1 * dr / D_G2G1 + sd(In[G2G1])) / (one + kG2f * mE2 * dr / D_G2G1); // :211
    const sd bg2pg1 = (kG2r * mEG2PG1 * dr / D_G2G1 + sd(In[G2PG1])) / (one + kG2f * mE2 * dr / D_G2G1);            // :212
    const sd bS2 = (kS2r * mEG2PG1S * dr / D_S2 + sd(In[SHP2])) / (one + kS2f * mEG2PG1 * dr / D_S2);               // :213
    const sd bpg1s = (kG1r * mEG2PG1S * dr / D_G1S2 + sd(In[PG1S])) / (one + kG1f * mEG2 * dr / D_G1S2);            // :214
    const sd bg2pg1s = (kG2r * mEG2PG1S * dr / D_G2G1S2 + sd(In[G2PG1S])) / (one + kG2f * mE2 * dr / D_G2G1S2);     // :215
    b[iSFK] = bi.v; b[aSFK] = ba.v; b[GAB1] = bG1.v; b[pGAB1] = bpG1.v; b[GRB2] = bG2.v; b[G2G1] = bg2g1.v;
    b[G2PG1] = bg2pg1.v; b[SHP2] = bS2.v; b[PG1S] = bpg1s.v; b[G2PG1S] = bg2pg1s.v;
    const sd o_mE = m1[mE], o_mES = m1[mES], o_mm = m1[mESmES], o_E = m1[E], o_EG2 = m1[EG2], o_EG2G1 = m1[EG2G1],
             o_EG2PG1 = m1[EG2PG1], o_EG2PG1S = m1[EG2PG1S];
    m2[mE] = ((-kEGFf * EGF * o_mE + kEGFr * o_mES) * dt + o_mE).v;                                                  // :220
    m2[mES] = ((kEGFf * EGF * o_mE - kEGFr * o_mES - two * kdf * o_mES * o_mES + two * kdr * o_mm) * dt + o_mES).v;  // :221
    m2[mESmES] = ((kdf * o_mES * o_mES - kdr * o_mm - kp * o_mm + kdp * o_E) * dt + o_mm).v;                         // :222
    m2[E] = ((kp * o_mm - kdp * o_E - kG2f * o_E * bG2 + kG2r * o_EG2 - kG2f * o_E * bg2g1 + kG2r * o_EG2G1
              - kG2f * o_E * bg2pg1 + kG2r * o_EG2PG1 - kG2f * o_E * bg2pg1s + kG2r * o_EG2PG1S) * dt + o_E).v;      // :223-224
    m2[EG2] = ((kG2f * bG2 * o_E - kG2r * o_EG2 - kG1f * bG1 * o_EG2 + kG1r * o_EG2G1 - kG1f * bpG1 * o_EG2
                + kG1r * o_EG2PG1 - kG1f * bpg1s * o_EG2 + kG1r * o_EG2PG1S) * dt + o_EG2).v;                        // :225-226
    m2[EG2G1] = ((kG2f * bg2g1 * o_E - kG2r * o_EG2G1 + kG1f * bG1 * o_EG2 - kG1r * o_EG2G1) * dt + o_EG2G1).v;      // :227
    m2[EG2PG1] = ((kG2f * bg2pg1 * o_E - kG2r * o_EG2PG1 + kG1f * bpG1 * o_EG2 - kG1r * o_EG2PG1
                   - kS2f * bS2 * o_EG2PG1 + kS2r * o_EG2PG1S) * dt + o_EG2PG1).v;                                   // :228-229
    m2[EG2PG1S] = ((kS2f * bS2 * o_EG2PG1 - kS2r * o_EG2PG1S + kG1f * bpg1s * o_EG2 - kG1r * o_EG2PG1S
                    + kG2f * bg2pg1s * o_E - kG2r * o_EG2PG1S) * dt + o_EG2PG1S).v;                                  // :230-231
    // error = maximum(abs.(1 .- new./old)) with NaN propagation (:238)
    double mx = -CUDART_INF;
    bool nan_seen = false;
#pragma unroll
    for (int q = 0; q < NCY; ++q) {
      const double e = fabs(__dsub_rn(1.0, __ddiv_rn(b[q], cold[q])));
      if (isnan(e)) nan_seen = true; else if (e > mx) mx = e;
    }
#pragma unroll
    for (int j = 0; j < NMB; ++j) {
      const double e = fabs(__dsub_rn(1.0, __ddiv_rn(m2[j], mold[j])));
      if (isnan(e)) nan_seen = true; else if (e > mx) mx = e;
    }
    err = nan_seen ? CUDART_NAN : mx;
    if (a.o.bc_loop == GAB1_BC_FOR_BREAK && err <= tol) break;
  }
  return it;
}

// ---------------------------------------------------------------------------------------------------------------
// Per-set driver.  u[q][i]: species q at the lane's i-th node.
// MODE 0: fast arithmetic, `for ... break` membrane loop; 1: fast arithmetic, `while error > tol` loop; 2: strict.

// 1/b for the Robin closures: hardware seed (MUFU.RCP64H, ~2^-20) and one cubic step r*(1 + e + e^2), e = 1 - b*r:
// relative error ~2^-60 before the final rounding, three dependent FP64 operations, no slow-path branch.
// (The quotient num*r then carries <= 2 ulp; measured by gab1_debug_recip_error / tests.)
__device__ __forceinline__ double fast_recip(double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  const double e = fma(-b, r, 1.0);
  return fma(r, fma(e, e, e), r);
}

template <int K, int MODE>
__device__ void solve_set(const KernelArgs& a, long long set, int lane, double* ws, const Grid<K>& g) {
  constexpr bool STRICT = MODE == MODE_STRICT;
  const int Nr = a.o.Nr, P = Nr + 1, Nts = a.o.Nts, Cn = Nts + 1;
  double* rowA = ws + WS_HDR;
  double* rowB = rowA + a.P_pad;
  double* oset = a.out + set * a.out_stride;
  unsigned status = 0;

  // ---- parameters of this set (uniform loads) ----
  const double* Co = a.Co + set * a.Co_stride;
  const double* Dv = a.D + set * GAB1_N_D;
  const double* kv = a.k + set * GAB1_N_K;
  const double dt = a.dt[set];
  const double CoSFK = Co[0], CoG2 = Co[1], CoG1 = Co[2], CoS2 = Co[3], CoEGFR = Co[4];
  Rates k;
  k.kS2f = kv[0]; k.kS2r = kv[1]; k.kG1f = kv[2]; k.kG1r = kv[3]; k.kG2f = kv[4]; k.kG2r = kv[5]; k.kG1p = kv[6];
  k.kG1dp = kv[7]; k.kSa = kv[8]; k.kSi = kv[9]; k.kp = kv[10]; k.kdp = kv[11]; k.kEGFf = kv[12]; k.kEGFr = kv[13];
  k.EGF = kv[14]; k.kdf = kv[15]; k.kdr = kv[16];
  double D_Si = Dv[0], D_Sa = Dv[0];
  if (a.o.sfk_mode == GAB1_SFK_MEMBRANE) D_Sa = 1e-32;                                  // basepdesolver.jl:366
  if (a.o.sfk_mode == GAB1_SFK_BOTH_FROZEN) { D_Si = 1e-32; D_Sa = 1e-32; }              // basepdesolver_rect.jl:305-306
  const double Dsp[NCY] = {D_Si, D_Sa, Dv[4], Dv[4], Dv[1], Dv[2], Dv[2], Dv[6], Dv[5], Dv[3]};   // basepdesolver.jl:43-49

  const bool track_t = (a.o.out_mode == GAB1_OUT_FULL || a.o.out_mode == GAB1_OUT_PCT_BOUND);
  const long long nout = a.out_stride;

  // the kernel writes every element of the set's block itself (snapshot columns that never become due are zero-filled
  // at the end, like the reference's zeros(...) arrays), so `out` needs no clearing and may be mapped host memory

  // Nt = Int64(ceil(tf/dt)) (basepdesolver.jl:72)
  const double nt_f = ceil(__ddiv_rn(a.o.tf, dt));
  if (!(nt_f >= 0.0 && nt_f < 9.0e18)) {
    for (long long i = lane; i < nout; i += 32) oset[i] = 0.0;
    if (lane == 0) {
      if (a.status) a.status[set] = GAB1_ST_THROW;
      if (a.n_saved) a.n_saved[set] = 0;
      if (a.n_steps) a.n_steps[set] = 0;
      if (a.n_bc) a.n_bc[set] = 0;
    }
    return;
  }
  const long long Nt = (long long)nt_f;

  // ---- state ----
  double u[NCY][K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const bool on = g.node[i] >= 1 && g.node[i] <= Nr;
#pragma unroll
    for (int q = 0; q < NCY; ++q) u[q][i] = 0.0;
    u[iSFK][i] = on ? CoSFK : 0.0;      // basepdesolver.jl:137-140
    u[GAB1][i] = on ? CoG1 : 0.0;
    u[GRB2][i] = on ? CoG2 : 0.0;
    u[SHP2][i] = on ? CoS2 : 0.0;
  }
  // where the boundary node Nr and its inner neighbour Nr-1 live
  const int lane_b = g.G - 1;
  constexpr int idx_b = K - 1;
  const int lane_i = K >= 2 ? g.G - 1 : g.G - 2;
  constexpr int idx_i = K >= 2 ? K - 2 : 0;

  // initial column of the FULL output (basepdesolver.jl:94-97,111)
  if (a.o.out_mode == GAB1_OUT_FULL) {
    const unsigned mask = a.o.matrix_mask;
    long long off = 0;
    for (int mi = 0; mi < 12; ++mi) {
      if (!((mask >> mi) & 1u)) continue;
      const double v0 = mi == GAB1_M_iSFK ? CoSFK : mi == GAB1_M_GRB2 ? CoG2 : mi == GAB1_M_SHP2 ? CoS2 : mi == GAB1_M_GAB1 ? CoG1 : 0.0;
      for (int n = lane; n < P; n += 32) oset[off + n] = v0;
      off += (long long)P * Cn;
    }
    if (lane < GAB1_N_VECTORS) oset[off + (long long)lane * Cn] = lane == GAB1_V_mE ? CoEGFR : 0.0;
  }

  double t = 0.0, t_save = a.o.dt_save;
  int nts = 1;
  const double modulus_step = (a.o.save_rule == GAB1_SAVE_MODULUS) ? rint(__ddiv_rn((double)Nt, (double)Nts)) : 0.0;
  long long bc_total = 0;
  double kp_now = k.kp;
  double pct_ave = 0.0, pct_memb = 0.0;      // PCT_BOUND: from snapshot column Nts+1 (zeros if never written)
  bool dead = false;                          // every state value is NaN: nothing can change any more
  long long step = 1;

  if constexpr (STRICT) {
    // =========================================================================================== strict path
    double b[NCY], m1[NMB], m2[NMB];
#pragma unroll
    for (int q = 0; q < NCY; ++q) b[q] = 0.0;             // u[Nr+1,2] starts at zero (basepdesolver.jl:115-124)
#pragma unroll
    for (int j = 0; j < NMB; ++j) { m1[j] = 0.0; m2[j] = 0.0; }
    m1[mE] = CoEGFR;
    const sd dr2 = __dmul_rn(a.o.dr, a.o.dr);
    const sd sdt = dt;
    for (; step <= Nt && !dead; ++step) {
      if (a.o.t_prechase >= 0.0) {                        // pulsechase_solver.jl:156-158
        if (__dadd_rn(a.o.t_prechase, dt) > t && t >= a.o.t_prechase) kp_now = 0.0;
      }
      double w[NCY][K];
#pragma unroll
      for (int i = 0; i < K; ++i) {
        sd L[NCY];
        double c[NCY], wn[NCY];
#pragma unroll
        for (int q = 0; q < NCY; ++q) {
          const double left = shfl(u[q][K - 1], lane - 1 < 0 ? 0 : lane - 1);
          const double right = shfl(u[q][0], lane + 1 > 31 ? 31 : lane + 1);
          const sd um = g.node[i] == 1 ? u[q][i] : (i > 0 ? u[q][i - 1] : left);
          const sd up = i < K - 1 ? u[q][i + 1] : right;
          const sd uc = u[q][i];
          c[q] = uc.v;
          if (a.o.geometry == GAB1_GEOM_SPHERICAL)
            L[q] = sd(Dsp[q]) * (sd(g.a[i]) * (up - um) + (up - sd(2.0) * uc + um) / dr2);
          else
            L[q] = sd(Dsp[q]) * (up - sd(2.0) * uc + um) / dr2;
        }
        kinetics_strict(L, c, k, sdt, wn);
#pragma unroll
        for (int q = 0; q < NCY; ++q) w[q][i] = g.interior[i] ? wn[q] : u[q][i];
      }
      // inner-neighbour values of the boundary node, broadcast to every lane
      double In[NCY];
#pragma unroll
      for (int q = 0; q < NCY; ++q) {
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < K; ++i) if (i == idx_i) v = w[q][i];
        In[q] = shfl(v, lane_i);
      }
      bool unconverged;
      const int it = membrane_strict(a, k, kp_now, Dsp, In, b, m1, m2, dt, status, unconverged);
      bc_total += it;
#pragma unroll
      for (int q = 0; q < NCY; ++q) {
#pragma unroll
        for (int i = 0; i < K; ++i) u[q][i] = (lane == lane_b && i == idx_b) ? b[q] : w[q][i];
      }
#pragma unroll
      for (int j = 0; j < NMB; ++j) m1[j] = m2[j];
      if (unconverged || isnan(m2[mE])) {
        bool all_nan = true;
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
          for (int q = 0; q < NCY; ++q) all_nan &= !(g.node[i] >= 1 && g.node[i] <= Nr) || isnan(u[q][i]);   // slots outside 1..Nr are padding
#pragma unroll
        for (int j = 0; j < NMB; ++j) all_nan &= isnan(m2[j]);
        dead = __all_sync(FULL, all_nan);
      }
      if (track_t) {
        t = __dadd_rn(t, dt);
        const bool save = a.o.save_rule == GAB1_SAVE_T_GE_TSAVE ? (t >= t_save) : (fmod((double)step, modulus_step) == 0.0);
        if (save) {
          if (nts >= Cn) status |= GAB1_ST_OVERFLOW;
          else {
            const int c = nts++;
            if (a.o.out_mode == GAB1_OUT_FULL) write_full_column<K>(a, oset, c, u, m2, t, CoEGFR, lane, g, rowA, status);
            else if (c == Cn - 1) {
              stage_row<K>(rowA, lane, g, Nr, [&](int i) { return derived_stot<K>(u, i); });
              pct_ave = trapz_r2(a.r, rowA, P);
              pct_memb = m2[EG2PG1S];
              __syncwarp();
            }
          }
          if (a.o.save_rule == GAB1_SAVE_T_GE_TSAVE) t_save = __dadd_rn(t_save, a.o.dt_save);
        }
      }
    }
    // ---- all-NaN state: only the clock and the snapshot schedule still evolve ----
    for (; step <= Nt; ++step) {
      bc_total += (a.o.bc_loop == GAB1_BC_FOR_BREAK) ? a.o.maxiters : 1;
      if (!track_t) { bc_total += (Nt - step) * (long long)((a.o.bc_loop == GAB1_BC_FOR_BREAK) ? a.o.maxiters : 1); break; }
      t = __dadd_rn(t, dt);
      const bool save = a.o.save_rule == GAB1_SAVE_T_GE_TSAVE ? (t >= t_save) : (fmod((double)step, modulus_step) == 0.0);
      if (save) {
        if (nts >= Cn) status |= GAB1_ST_OVERFLOW;
        else {
          const int c = nts++;
          if (a.o.out_mode == GAB1_OUT_FULL) write_full_column<K>(a, oset, c, u, m2, t, CoEGFR, lane, g, rowA, status);
          else if (c == Cn - 1) { pct_ave = CUDART_NAN; pct_memb = CUDART_NAN; }
        }
        if (a.o.save_rule == GAB1_SAVE_T_GE_TSAVE) t_save = __dadd_rn(t_save, a.o.dt_save);
      }
    }
    if (Nt == 0) {      // no step taken: column 2 of every work array is still zero (sapdesolver.jl:245)
#pragma unroll
      for (int q = 0; q < NCY; ++q)
#pragma unroll
        for (int i = 0; i < K; ++i) u[q][i] = 0.0;
    }
    write_final<K>(a, oset, u, m2, lane, g, rowA, rowB, status);
  } else {
    // ============================================================================================= fast path
    constexpr bool WHILE = MODE == MODE_FAST_WHILE;
    // ---- interior constants: rate constants and diffusivities pre-scaled by dt ----
    const double kS2f_t = k.kS2f * dt, kS2r_t = k.kS2r * dt, kG1f_t = k.kG1f * dt, kG1r_t = k.kG1r * dt,
                 kG1p_t = k.kG1p * dt, kG1dp_t = k.kG1dp * dt, kSi_t = k.kSi * dt;
    const double Dt_Si = D_Si * dt, Dt_Sa = D_Sa * dt, Dt_G1 = Dv[4] * dt, Dt_G2 = Dv[1] * dt, Dt_G2G1 = Dv[2] * dt,
                 Dt_S2 = Dv[6] * dt, Dt_G1S2 = Dv[5] * dt, Dt_G2G1S2 = Dv[3] * dt;

    // ---- membrane block: lane roles ----
    // lanes 0..9   Robin closure of cytosolic species `lane`:  b = (cr*M_num + I)/(1 + cf*M_den)   (basepdesolver.jl:206-215)
    // lanes 10..17 membrane species `lane-10`;  lane 18 carries Etot = 2(E+EG2+EG2G1+EG2PG1+EG2PG1S) as a pseudo-species
    // lane 31      always holds zeros: the source of every unused shuffle
    constexpr int LZ = 31, LE = ML + NMB;
    double kf = 0.0, kr = 0.0, Dq = 1.0;
    int src_num = LZ, src_den = LZ;
    switch (lane) {
      case iSFK:   kf = k.kSa;  Dq = D_Si; src_den = LE; break;                     // I/(1 + kSa*Etot*dr/D_S)
      case aSFK:   kf = k.kSa;  Dq = D_Si; src_num = LE; src_den = LE; break;       // rewritten over the same denominator, see cr_a
      case GAB1:   kf = k.kG1f; kr = k.kG1r; Dq = Dv[4]; src_num = ML + EG2G1;   src_den = ML + EG2;    break;
      case pGAB1:  kf = k.kG1f; kr = k.kG1r; Dq = Dv[4]; src_num = ML + EG2PG1;  src_den = ML + EG2;    break;
      case GRB2:   kf = k.kG2f; kr = k.kG2r; Dq = Dv[1]; src_num = ML + EG2;     src_den = ML + E;      break;
      case G2G1:   kf = k.kG2f; kr = k.kG2r; Dq = Dv[2]; src_num = ML + EG2G1;   src_den = ML + E;      break;
      case G2PG1:  kf = k.kG2f; kr = k.kG2r; Dq = Dv[2]; src_num = ML + EG2PG1;  src_den = ML + E;      break;
      case SHP2:   kf = k.kS2f; kr = k.kS2r; Dq = Dv[6]; src_num = ML + EG2PG1S; src_den = ML + EG2PG1; break;
      case PG1S:   kf = k.kG1f; kr = k.kG1r; Dq = Dv[5]; src_num = ML + EG2PG1S; src_den = ML + EG2;    break;
      case G2PG1S: kf = k.kG2f; kr = k.kG2r; Dq = Dv[3]; src_num = ML + EG2PG1S; src_den = ML + E;      break;
      default: break;
    }
    const double drD = a.o.dr / Dq;
    const double cf = kf * drD;
    const double cr_fixed = kr * drD;
    const double ca = k.kSa * (a.o.dr / D_Sa);            // aSFK closure coefficient; a true division (D_Sa may be 1e-32)
    const bool is_flux = lane >= GAB1 && lane <= G2PG1S;   // this closure's net binding flux feeds the membrane ODEs
    const double kf_t = is_flux ? kf * dt : 0.0, kr_t = is_flux ? kr * dt : 0.0;
    // membrane species j: new = base + sum_i sg_i * F[fs_i], F = dt*(kf*M_den*b - kr*M_num) held by the closure lanes
    // (basepdesolver.jl:220-231 regrouped by reaction: every binding term appears once with each sign)
    int fs0 = LZ, fs1 = LZ, fs2 = LZ, fs3 = LZ;
    double sg0 = 0.0, sg1 = 0.0, sg2 = 0.0, sg3 = 0.0;
    switch (lane - ML) {
      case E:       fs0 = GRB2;   fs1 = G2G1;  fs2 = G2PG1; fs3 = G2PG1S; sg0 = -1.0; sg1 = -1.0; sg2 = -1.0; sg3 = -1.0; break;
      case EG2:     fs0 = GRB2;   fs1 = GAB1;  fs2 = pGAB1; fs3 = PG1S;   sg0 = 1.0;  sg1 = -1.0; sg2 = -1.0; sg3 = -1.0; break;
      case EG2G1:   fs0 = G2G1;   fs1 = GAB1;  sg0 = 1.0; sg1 = 1.0; break;
      case EG2PG1:  fs0 = G2PG1;  fs1 = pGAB1; fs2 = SHP2; sg0 = 1.0; sg1 = 1.0; sg2 = -1.0; break;
      case EG2PG1S: fs0 = G2PG1S; fs1 = PG1S;  fs2 = SHP2; sg0 = 1.0; sg1 = 1.0; sg2 = 1.0; break;
      default: break;
    }
    // membrane-only reactions, old-time values: f = m*(alpha + alpha2*m) - beta*m_next on lanes 10,11,12
    //   lane 10: kEGFf*EGF*mE - kEGFr*mES     lane 11: kdf*mES^2 - kdr*mESmES     lane 12: kp*mESmES - kdp*E
    // and their contribution to d(m)/dt: dm = s_own*f + s_src*f[f_src]   (lane 18: dEtot/dt = 2*f3, all bindings cancel)
    double alpha = 0.0, alpha2 = 0.0, beta = 0.0, s_own = 0.0, s_src = 0.0;
    int f_src = LZ;
    switch (lane - ML) {
      case mE:     alpha = k.kEGFf * k.EGF; beta = k.kEGFr; s_own = -1.0; break;
      case mES:    alpha2 = k.kdf;          beta = k.kdr;   s_own = -2.0; s_src = 1.0; f_src = ML + mE; break;
      case mESmES: alpha = kp_now;          beta = k.kdp;   s_own = -1.0; s_src = 1.0; f_src = ML + mES; break;
      case E:      s_src = 1.0; f_src = ML + mESmES; break;
      case NMB:    s_src = 2.0; f_src = ML + mESmES; break;
      default: break;
    }
    const double tol = a.o.tol;
    const bool untracked = lane >= LE;
    const int iq_idx = lane < NCY ? lane : 10;             // ws[10..15] stay zero
    const bool pulse = a.o.t_prechase >= 0.0;
    const int maxiters = a.o.maxiters;
    const unsigned ws_s = (unsigned)__cvta_generic_to_shared(ws);
    // x: the value this lane tracks across iterations and steps — boundary value u[Nr+1] of species `lane`
    // (lanes 0..9), membrane species lane-10 (lanes 10..17), Etot (lane 18), zero elsewhere
    double x = (lane == ML + mE) ? CoEGFR : 0.0;

    // ---- time loop.  Every rare event (snapshot due, pulse-chase switch, last step, dead state) hides behind one integer
    //      countdown, so the common step carries a single predictable branch besides the fixed-point loop.  `plan`
    //      returns a conservative count of steps that cannot contain an event; the exact floating-point tests of the
    //      reference (t >= t_save, t_prechase + dt > t >= t_prechase) are then applied step by step around the event.
    bool pulse_pending = pulse;
    if (pulse_pending && a.o.t_prechase + dt > t && t >= a.o.t_prechase) {      // pulsechase_solver.jl:156-158 at step 1
      kp_now = 0.0; if (lane == ML + mESmES) alpha = 0.0; pulse_pending = false;
    }
    auto plan = [&]() -> int {
      long long n = Nt - step + 1;
      auto bound = [&](double t_event) {
        const double q = floor((t_event - t) / dt) - 1.0;                      // accumulated t is within ulps of step*dt
        if (!(q >= 1.0)) n = 1;
        else if (q < (double)n) n = (long long)q;
      };
      if (track_t) {
        if (a.o.save_rule == GAB1_SAVE_T_GE_TSAVE) bound(t_save); else n = 1;
      }
      if (pulse_pending) bound(a.o.t_prechase);
      return (int)(n > 1000000000LL ? 1000000000LL : n);
    };
    int countdown = plan();
    if (Nt >= 1)
    for (;;) {
      // ---- membrane block prologue: everything that depends only on old-time values.  It is independent of the
      //      interior update below, so the two instruction streams interleave and hide each other's latency ----
      const double m_old = x;                                        // lanes >= 10: value at the old time level
      const double m_next = shfl_down1(m_old);
      const double f = fma(m_old, fma(alpha2, m_old, alpha), -(beta * m_next));
      const double base = fma(dt, fma(s_own, f, s_src * shfl(f, f_src)), m_old);
      // the first iterate of the membrane column is the old-time column, so these two shuffles serve both the
      // old-time flux coefficients and the first pass of the fixed point
      const double Md1 = shfl(m_old, src_den), Mn1 = shfl(m_old, src_num);
      const double A_t = kf_t * Md1;                                 // F = dt*(kf*M_den*b - kr*M_num), old-time M
      const double B_t = kr_t * Mn1;
      const double rden1 = fast_recip(fma(cf, Md1, 1.0));            // 1/(1 + cf*M_den) of the first pass

      // ---- interior: D*lap + kinetics, explicit Euler, updated in place (basepdesolver.jl:150-180) ----
      {
        double hl[NCY], hr[NCY];
        if constexpr (K >= 4) {
          const unsigned ex = ws_s + 8u * (unsigned)(WS_HDR + 2 * a.P_pad);
#pragma unroll
          for (int q = 0; q < NCY; ++q) {
            sts(ex + 8u * (unsigned)(q * 32 + lane), u[q][K - 1]);
            sts(ex + 8u * (unsigned)((NCY + q) * 32 + lane), u[q][0]);
          }
          __syncwarp();
          const int ll = lane > 0 ? lane - 1 : 0, lr = lane < 31 ? lane + 1 : 31;
#pragma unroll
          for (int q = 0; q < NCY; ++q) {
            hl[q] = lds(ex + 8u * (unsigned)(q * 32 + ll));
            hr[q] = lds(ex + 8u * (unsigned)((NCY + q) * 32 + lr));
          }
        } else {
#pragma unroll
          for (int q = 0; q < NCY; ++q) {
            hl[q] = shfl_up1(u[q][K - 1]);
            hr[q] = shfl_down1(u[q][0]);
          }
        }
        double L[2][NCY];     // Laplacians of the node being updated and of the next one (which still needs old values)
#pragma unroll
        for (int q = 0; q < NCY; ++q) {
          const double up = K > 1 ? u[q][1] : hr[q];
          L[0][q] = fma(g.cp[0], up, fma(g.cm[0], hl[q], g.c0[0] * u[q][0]));
        }
#pragma unroll
        for (int i = 0; i < K; ++i) {
          if (i + 1 < K) {
#pragma unroll
            for (int q = 0; q < NCY; ++q) {
              const double up = i + 2 < K ? u[q][i + 2] : hr[q];
              L[(i + 1) & 1][q] = fma(g.cp[i + 1], up, fma(g.cm[i + 1], u[q][i], g.c0[i + 1] * u[q][i + 1]));
            }
          }
          const double(&lap)[NCY] = L[i & 1];
          const double Si = u[iSFK][i], Sa = u[aSFK][i], G1 = u[GAB1][i], pG1 = u[pGAB1][i], G2 = u[GRB2][i],
                       g2g1 = u[G2G1][i], g2pg1 = u[G2PG1][i], S2 = u[SHP2][i], pg1s = u[PG1S][i], g2pg1s = u[G2PG1S][i];
          const double gb = kG1f_t * G2, ph = kG1p_t * Sa, sb = kS2f_t * S2;
          const double v1 = fma(gb, G1, -(kG1r_t * g2g1));        // GRB2 + GAB1   <-> G2G1
          const double v3 = fma(gb, pG1, -(kG1r_t * g2pg1));      // GRB2 + pGAB1  <-> G2PG1
          const double v5 = fma(gb, pg1s, -(kG1r_t * g2pg1s));    // GRB2 + PG1S   <-> G2PG1S
          const double v2 = fma(ph, G1, -(kG1dp_t * pG1));        // GAB1  <-> pGAB1 (aSFK / phosphatase)
          const double v6 = fma(ph, g2g1, -(kG1dp_t * g2pg1));    // G2G1  <-> G2PG1
          const double v4 = fma(sb, pG1, -(kS2r_t * pg1s));       // SHP2 + pGAB1  <-> PG1S
          const double v7 = fma(sb, g2pg1, -(kS2r_t * g2pg1s));   // SHP2 + G2PG1  <-> G2PG1S
          u[iSFK][i] = fma(Dt_Si, lap[iSFK], fma(kSi_t, Sa, Si));             // aSFK -> iSFK
          u[aSFK][i] = fma(Dt_Sa, lap[aSFK], fma(-kSi_t, Sa, Sa));
          u[GAB1][i] = fma(Dt_G1, lap[GAB1], G1 - v1 - v2);
          u[pGAB1][i] = fma(Dt_G1, lap[pGAB1], pG1 - v3 + v2 - v4);
          u[GRB2][i] = fma(Dt_G2, lap[GRB2], G2 - v1 - v3 - v5);
          u[G2G1][i] = fma(Dt_G2G1, lap[G2G1], g2g1 + v1 - v6);
          u[G2PG1][i] = fma(Dt_G2G1, lap[G2PG1], g2pg1 + v3 + v6 - v7);
          u[SHP2][i] = fma(Dt_S2, lap[SHP2], S2 - v4 - v7);
          u[PG1S][i] = fma(Dt_G1S2, lap[PG1S], pg1s + v4 - v5);
          u[G2PG1S][i] = fma(Dt_G2G1S2, lap[G2PG1S], g2pg1s + v5 + v7);
          if (i == idx_i) {
            // hand the inner-neighbour values u+[Nr-1] to the closure lanes as soon as they exist; the loads that
            // pick them up sit behind the rest of the interior work
            if (lane == lane_i) {
#pragma unroll
              for (int q = 0; q < NCY; ++q) sts(ws_s + 8 * q, u[q][i]);
            }
          }
        }
      }
      __syncwarp();
      const double Iq = lds(ws_s + 8 * iq_idx);
      // aSFK: I_a + ca*Etot*I_i/(1 + cf*Etot) = (I_a + (cf*I_a + ca*I_i)*Etot)/(1 + cf*Etot)   (basepdesolver.jl:206-207)
      const double cr = lane == aSFK ? fma(cf, Iq, ca * lds(ws_s + 8 * iSFK)) : cr_fixed;

      // ---- fixed-point iterations (basepdesolver.jl:197-242); the first pass is peeled: its reciprocal is ready ----
      int it = 1;               // the host routes `maxiters = 0` to the strict kernel: at least one pass runs here
      bool unconverged = false, nan_exit = false;
      {
        // everything after the closure value qv of one pass; returns true when another pass is needed
        auto finish_pass = [&](double qv) -> bool {
          const double F = fma(A_t, qv, -B_t);
          const double F0 = shfl(F, fs0), F1 = shfl(F, fs1), F2 = shfl(F, fs2), F3 = shfl(F, fs3);
          const double mnew = fma(sg0, F0, sg1 * F1) + fma(sg2, F2, fma(sg3, F3, base));     // depth 3 instead of 4
          const double xnew = lane < NCY ? qv : mnew;
          if constexpr (!WHILE) {
            // |1 - new/old| <= tol  <=>  |old - new| <= tol*|old|; the strict `<` also rejects old = new = 0 (0/0 = NaN
            // in the reference) and old = +-Inf, so no special cases remain; NaN operands compare false
            const bool ok = (fabs(x - xnew) < tol * fabs(x)) || untracked;
            x = xnew;
            if (__all_sync(FULL, ok)) return false;
            if (it >= maxiters) { unconverged = true; return false; }
            return true;
          } else {
            // `while error > tol`: a NaN error leaves the loop, so NaN has to be told apart exactly
            int cls;
            const bool special = !untracked && (is_special(x) || is_special(xnew));
            if (__any_sync(FULL, special)) cls = untracked ? 0 : classify_exact(x, xnew, tol);
            else cls = (!untracked && !(fabs(x - xnew) <= tol * fabs(x))) ? 1 : 0;
            x = xnew;
            const bool any_nan = __any_sync(FULL, cls == 2);
            const bool all_ok = __all_sync(FULL, cls == 0);
            if (any_nan || all_ok) { nan_exit = any_nan; return false; }
            if (it >= maxiters) { status |= GAB1_ST_ITER_CAP; return false; }
            return true;
          }
        };
        bool more = finish_pass(fma(cr, Mn1, Iq) * rden1);
        while (more) {
          ++it;
          const double Mn = shfl(x, src_num);
          const double Md = shfl(x, src_den);
          more = finish_pass(fma(cr, Mn, Iq) * fast_recip(fma(cf, Md, 1.0)));
        }
      }
      bc_total += it;
      // ---- boundary values back to the lane that owns node Nr ----
      if (lane < NCY) sts(ws_s + 8 * (16 + lane), x);
      __syncwarp();
      if (lane == lane_b) {
#pragma unroll
        for (int q = 0; q < NCY; ++q) u[q][idx_b] = lds(ws_s + 8 * (16 + q));
      }
      if (unconverged || nan_exit) {
        bool all_nan = (lane < ML || lane >= LE) || isnan(x);
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
          for (int q = 0; q < NCY; ++q) all_nan &= !(g.node[i] >= 1 && g.node[i] <= Nr) || isnan(u[q][i]);   // slots outside 1..Nr are padding
        dead = __all_sync(FULL, all_nan);
        if (dead) countdown = 1;
      }
      t = t + dt;                                                   // basepdesolver.jl:265
      if (--countdown > 0) { ++step; continue; }
      // ---- rare path: exact event tests for the step just taken ----
      if (track_t) {
        const bool save = a.o.save_rule == GAB1_SAVE_T_GE_TSAVE ? (t >= t_save) : (fmod((double)step, modulus_step) == 0.0);
        if (save) {
          if (nts >= Cn) status |= GAB1_ST_OVERFLOW;
          else {
            const int c = nts++;
            double m[NMB];
#pragma unroll
            for (int j = 0; j < NMB; ++j) m[j] = shfl(x, ML + j);
            if (a.o.out_mode == GAB1_OUT_FULL) write_full_column<K>(a, oset, c, u, m, t, CoEGFR, lane, g, rowA, status);
            else if (c == Cn - 1) {
              stage_row<K>(rowA, lane, g, Nr, [&](int i) { return derived_stot<K>(u, i); });
              pct_ave = trapz_r2(a.r, rowA, P);
              pct_memb = m[EG2PG1S];
              __syncwarp();
            }
          }
          if (a.o.save_rule == GAB1_SAVE_T_GE_TSAVE) t_save = t_save + a.o.dt_save;
        }
      }
      if (pulse_pending) {                                          // the test the next step would make at its start
        if (a.o.t_prechase + dt > t && t >= a.o.t_prechase) { kp_now = 0.0; if (lane == ML + mESmES) alpha = 0.0; pulse_pending = false; }
        else if (t >= a.o.t_prechase + dt) pulse_pending = false;   // the window was stepped over: the reference never switches
      }
      ++step;
      if (dead || step > Nt) break;
      countdown = plan();
    }
    double m[NMB];
#pragma unroll
    for (int j = 0; j < NMB; ++j) m[j] = shfl(x, ML + j);
    // ---- all-NaN state: only the clock and the snapshot schedule still evolve ----
    for (; step <= Nt; ++step) {
      const long long per = WHILE ? 1 : maxiters;
      if (!track_t) { bc_total += (Nt - step + 1) * per; break; }
      bc_total += per;
      t = t + dt;
      const bool save = a.o.save_rule == GAB1_SAVE_T_GE_TSAVE ? (t >= t_save) : (fmod((double)step, modulus_step) == 0.0);
      if (save) {
        if (nts >= Cn) status |= GAB1_ST_OVERFLOW;
        else {
          const int c = nts++;
          if (a.o.out_mode == GAB1_OUT_FULL) write_full_column<K>(a, oset, c, u, m, t, CoEGFR, lane, g, rowA, status);
          else if (c == Cn - 1) { pct_ave = CUDART_NAN; pct_memb = CUDART_NAN; }
        }
        if (a.o.save_rule == GAB1_SAVE_T_GE_TSAVE) t_save = t_save + a.o.dt_save;
      }
    }
    if (Nt == 0) {
#pragma unroll
      for (int q = 0; q < NCY; ++q)
#pragma unroll
        for (int i = 0; i < K; ++i) u[q][i] = 0.0;
#pragma unroll
      for (int j = 0; j < NMB; ++j) m[j] = 0.0;
    }
    write_final<K>(a, oset, u, m, lane, g, rowA, rowB, status);
  }

  if (a.o.out_mode == GAB1_OUT_PCT_BOUND) {      // run_base_model.jl:272-276
    const double R = a.o.R;
    const double ave = __ddiv_rn(__dmul_rn(pct_ave, 3.0), __dmul_rn(__dmul_rn(R, R), R));
    const double mem = __ddiv_rn(__dmul_rn(pct_memb, a.o.pct_mul), a.o.pct_div);
    const double pct = __dmul_rn(__ddiv_rn(__dadd_rn(ave, mem), CoG1), 100.0);
    if (isnan(pct)) status |= GAB1_ST_NAN;
    if (lane == 0) oset[0] = pct;
  }
  if (track_t && nts < Cn) {
    status |= GAB1_ST_SHORT;
    if (a.o.out_mode == GAB1_OUT_FULL) {           // columns nts..Nts were never due: they stay zero in the reference
      long long off = 0;
      for (int mi = 0; mi < 12; ++mi) {
        if (!((a.o.matrix_mask >> mi) & 1u)) continue;
        for (long long i = (long long)nts * P + lane; i < (long long)Cn * P; i += 32) oset[off + i] = 0.0;
        off += (long long)P * Cn;
      }
      for (int v = 0; v < GAB1_N_VECTORS; ++v)
        for (int c = nts + lane; c < Cn; c += 32) oset[off + (long long)v * Cn + c] = 0.0;
    }
  }
  if (lane == 0) {
    if (a.status) a.status[set] = (int)status;
    if (a.n_saved) a.n_saved[set] = track_t ? nts : 0;
    if (a.n_steps) a.n_steps[set] = Nt;
    if (a.n_bc) a.n_bc[set] = bc_total;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent kernel: every warp pulls parameter sets from a queue ordered by descending work.
// Launch shape of the product kernels (K <= 2, fast arithmetic): GAB1_WARPS warps per CTA, at least GAB1_MINB CTAs per
// SM (which caps registers per thread).  Chosen from measurements on B200, see DESIGN.md.
#ifndef GAB1_WARPS
#define GAB1_WARPS 4
#endif
#ifndef GAB1_MINB
#define GAB1_MINB 2
#endif
constexpr int kWarpsPerCta = GAB1_WARPS;
template <int K, int MODE>
__global__ void __launch_bounds__(32 * GAB1_WARPS, (MODE == MODE_STRICT || K > 2) ? 1 : GAB1_MINB)
solve_kernel(const KernelArgs a) {
  if (a.guard && *a.guard != a.guard_expect) return;
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* ws = smem + (size_t)warp * (WS_HDR + 2 * a.P_pad + WS_EX);
  const int Nr = a.o.Nr;
  ws[lane] = 0.0;
  __syncwarp();

  Grid<K> g;
  {
    const double dr = a.o.dr;
    const double inv_dr2 = 1.0 / (dr * dr);
    g.G = (Nr + K - 1) / K;
    const int off = Nr - g.G * K;                      // <= 0
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const int n = lane * K + i + 1 + off;
      g.node[i] = n;
      g.interior[i] = n >= 1 && n <= Nr - 1;
      const double r = (n >= 1 && n <= Nr) ? a.r[n] : 1.0;
      g.a[i] = __ddiv_rn(1.0, __dmul_rn(r, dr));
      const double aj = (a.o.geometry == GAB1_GEOM_SPHERICAL) ? 1.0 / (r * dr) : 0.0;
      double cp = inv_dr2 + aj, cm = inv_dr2 - aj, c0 = -2.0 * inv_dr2;
      if (n == 1) { c0 += cm; cm = 0.0; }             // u[0] = u[1]: the mirror term joins the centre coefficient
      g.cp[i] = g.interior[i] ? cp : 0.0;
      g.cm[i] = g.interior[i] ? cm : 0.0;
      g.c0[i] = g.interior[i] ? c0 : 0.0;
    }
  }
  for (;;) {
    unsigned item = 0;
    if (lane == 0) item = atomicAdd(a.counter, 1u);
    item = __shfl_sync(FULL, item, 0);
    if ((long long)item >= a.S || (a.dyn_count && item >= *a.dyn_count)) break;
    const long long set = a.order ? (long long)a.order[item] : (long long)item;
    solve_set<K, MODE>(a, set, lane, ws, g);
    __syncwarp();
  }
}

#ifdef GAB1_WITH_DIAGNOSTICS
// max relative error of the seed and of fast_recip over n log-spaced operands in [lo, hi] (diagnostic)
__global__ void recip_error_kernel(double lo, double hi, int n, double* out) {
  double worst_seed = 0.0, worst = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double b = lo * exp(log(hi / lo) * ((double)i / (double)(n - 1)));
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(b));
    const double exact = __drcp_rn(b);
    worst_seed = fmax(worst_seed, fabs(r0 - exact) / exact);
    worst = fmax(worst, fabs(fast_recip(b) - exact) / exact);
  }
  for (int o = 16; o; o >>= 1) {
    worst_seed = fmax(worst_seed, __shfl_xor_sync(FULL, worst_seed, o));
    worst = fmax(worst, __shfl_xor_sync(FULL, worst, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax((unsigned long long*)&out[0], (unsigned long long)__double_as_longlong(worst_seed));
    atomicMax((unsigned long long*)&out[1], (unsigned long long)__double_as_longlong(worst));
  }
}
#endif

}  // namespace gab1
