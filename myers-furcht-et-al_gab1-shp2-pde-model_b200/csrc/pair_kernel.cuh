// pair_kernel.cuh — the product kernels (fast arithmetic): one parameter set per GROUP of HW lanes, HW = 16 (two
// sets per warp, grids of up to 64 nodes) or HW = 32 (one set per warp, up to 256 nodes).  Below, "half" means
// such a group; with HW = 32 there is a single one and every per-half vote degenerates to a warp vote.
//
// Why pairs.  The one-set-per-warp kernel (solver_kernel.cuh) spends, per time step of one set, ~190 FP64 warp
// instructions, ~62 shuffles and a serial membrane fixed point whose lanes are mostly idle; ncu (profiles/r1_*) shows
// the FP64 pipe, the LSU/shuffle pipe and the issue slots each about half busy with two resident warps per scheduler
// and no room for a third (244 registers).  Packing two sets into one warp with K = 4 nodes per lane
//   * halves the halo shuffles per node (two per species per FOUR nodes instead of two),
//   * runs the membrane fixed point of both sets in the same instructions (it costs one set's worth per pair),
//   * gives every warp 4 nodes x 10 species of independent work per lane to cover its own latencies,
// and the stencil is regrouped so that one species-node costs 4 FP64 instructions instead of 6:
//     u+ = lam_q*(cp_i*u[i+1] + cm_i*u[i-1]) + (c_q*u[i] + kinetics),   lam_q = D_q*dt/dr^2, c_q = 1 - 2*lam_q,
//     cp_i = 1 + dr/r_i, cm_i = 1 - dr/r_i   (spherical; both 1 for the planar Laplacian)
// (algebraically the reference's D*(1/(r*dr)*(u[i+1]-u[i-1]) + (u[i+1]-2u[i]+u[i-1])/dr^2)*dt + u, basepdesolver.jl:151).
//
// Skew.  The membrane fixed point of step n needs only u[Nr-1](n), and of everything the interior computes for step
// n+1 only u[Nr-1](n+1) needs its result b(n) — through the single term lam_q*cp*b(n).  So one loop iteration runs
//     interior(n+1) for every node, with the boundary contribution to node Nr-1 left out,
//     membrane(n) — two passes in straight-line code, further passes in a (rare) loop,
//     u[Nr-1](n+1) += lam_q*cp*b(n)
// in ONE instruction stream: the long dependent chain of the fixed point hides behind the interior's FP64 work instead
// of stalling the warp.  Whenever a half has an event the pipeline is drained (membrane(n+1) runs alone) so that
// outputs see one consistent time level; the order of the arithmetic of a set does not depend on its partner.
//
// The two halves run the same instruction stream on their own parameter set, clock, snapshot schedule and
// fixed-point iteration count; everything rare (snapshot due, pulse-chase switch, last step, diverged state) is
// resolved per half behind one warp-wide countdown with the reference's exact floating-point tests.
//
// Lane roles inside a half (hl = lane & 15):
//   grid        nodes 1..Nr right-aligned on the half's 16*K slots (node Nr = last slot of lane G-1)
//   closures    hl 0..9: Robin closure of cytosolic species hl                       (basepdesolver.jl:206-215)
//   membrane    hl 0..7: membrane species hl; hl 8: Etot pseudo-species; hl 15: zeros (basepdesolver.jl:220-231)
// Reference: basepdesolver.jl:149-296, basepdesolver_rect.jl:131-161, sapdesolver.jl:128-242,
// sapdesolver_memb-SFK.jl:175-222, pulsechase_solver.jl:156-158.
#pragma once
#ifdef GAB1_PHASE_TIMING
#include <cstdio>
#endif
#include "solver_kernel.cuh"

namespace gab1 {


template <int K>
struct PGrid {
  double cp[K], cm[K];   // 1 +- dr/r_i on interior nodes, 0 elsewhere (padding, boundary slot)
  double m1[K];          // MIRROR only: coefficient of u[1] standing in for u[0] = u[1] (zero-flux), on node 1
  int node[K];
  int G;
};

template <int HW>
__device__ __forceinline__ double shflg_down1(double x) {
  const int lo = __shfl_down_sync(FULL, __double2loint(x), 1, HW), hi = __shfl_down_sync(FULL, __double2hiint(x), 1, HW);
  return __hiloint2double(hi, lo);
}
template <int HW>
__device__ __forceinline__ double shflg_up1(double x) {
  const int lo = __shfl_up_sync(FULL, __double2loint(x), 1, HW), hi = __shfl_up_sync(FULL, __double2hiint(x), 1, HW);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ long long bcast_ll(long long v, int src) {
  return (long long)__shfl_sync(FULL, (unsigned long long)v, src);
}
__device__ __forceinline__ double* bcast_ptr(double* v, int src) {
  return reinterpret_cast<double*>(__shfl_sync(FULL, reinterpret_cast<unsigned long long>(v), src));
}

__device__ __forceinline__ double2 lds2(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}
// sts2: solver_kernel.cuh

// ---------------------------------------------------------------------------------------------------------------
// Interior token.  A scheduler holds two of these warps (255 registers each).  Left alone they drift into phase:
// both run their FP64-dense interior update at once and share the pipe, then both sit in the latency-bound membrane
// fixed point and leave it idle (ncu: pipe ~50% busy, per-warp period I + M with I doubled by the sharing).  The
// token makes the two warps of a scheduler take turns at the interior, so that each warp's fixed point runs under
// its partner's interior: period max(I_a + I_b, I + M) with I at its stand-alone speed.
// Two named barriers per pair carry the token (bar.arrive = hand over, bar.sync = wait for it); a flag word per warp,
// published before each hand-over, tells the partner when a warp has run out of parameter sets, so that the pair
// stops shaking hands at the same round.
struct Token {
  int bar_wait, bar_post;
  volatile int* mine;
  volatile int* peer;
  bool leader;            // the warp of the pair that takes the first turn
  bool published;         // what this warp last told its partner about being out of work
  bool peer_done;         // what the partner last told this warp

  __device__ __forceinline__ void acquire() {
    asm volatile("bar.sync %0, 64;" ::"r"(bar_wait) : "memory");
    peer_done = *peer != 0;
  }
  __device__ __forceinline__ void release(bool out_of_work) {
    *mine = out_of_work ? 1 : 0;
    published = out_of_work;
    __threadfence_block();
    asm volatile("bar.arrive %0, 64;" ::"r"(bar_post) : "memory");
  }
  // after the last parameter set: keep taking (empty) turns until the partner has run out of work as well
  __device__ void drain() {
    for (;;) {
      const bool was_published = published;
      acquire();
      if (!leader && peer_done && was_published) return;       // the leader leaves after this round: so do we
      release(true);
      if (leader && peer_done) return;                          // both out of work as of this round
    }
  }
};

// lanes of the owning half stage one row into shared memory; the whole warp then stores it unit-stride
template <int K, typename F>
__device__ __forceinline__ void pstage_row(double* row, bool mine, const PGrid<K>& g, int Nr, F val) {
  if (mine) {
#pragma unroll
    for (int i = 0; i < K; ++i) {
      if (g.node[i] >= 1 && g.node[i] <= Nr) row[g.node[i]] = val(i);
      if (g.node[i] == 1) row[0] = val(i);   // node 0 == node 1 (basepdesolver.jl:183-192)
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
template <int K, int MODE, bool MIRROR, int HW, bool SKEW, bool TOKEN>
__device__ void solve_pair(const KernelArgs& a, long long item, int lane, double* ws, const PGrid<K>& g, Token& tok) {
  constexpr bool WHILE = MODE == MODE_FAST_WHILE;
  constexpr int NH = 32 / HW;                                   // parameter sets per warp
  constexpr unsigned HMASK = HW == 32 ? 0xffffffffu : 0xffffu;  // the lanes of one set, shifted down to bit 0
  const int hl = lane & (HW - 1), hbase = lane & (32 - HW), half = lane / HW;
  const int Nr = a.o.Nr, P = Nr + 1, Cn = a.o.Nts + 1;
  double* rowA = ws + NH * WS_HDR;
  double* rowB = rowA + a.P_pad;
  const unsigned wsh = (unsigned)__cvta_generic_to_shared(ws) + (unsigned)(half * WS_HDR * 8);

  // ---- which parameter set this half solves (an odd tail leaves the second half without one: it shadows its
  //      partner's parameters and writes nothing) ----
  const long long idx = NH * item + half;
  const bool have = idx < a.S;
  const long long pidx = have ? idx : NH * item;
  const long long set = a.order ? (long long)a.order[pidx] : pidx;
  double* oset = a.out + set * a.out_stride;
  unsigned status = 0;

  const double* Co = a.Co + set * a.Co_stride;
  const double* Dv = a.D + set * GAB1_N_D;
  const double* kv = a.k + set * GAB1_N_K;
  const double dt = a.dt[set];
  const double CoSFK = Co[0], CoG2 = Co[1], CoG1 = Co[2], CoS2 = Co[3], CoEGFR = Co[4];
  // k = [kS2f,kS2r,kG1f,kG1r,kG2f,kG2r,kG1p,kG1dp,kSa,kSi,kp,kdp,kEGFf,kEGFr,EGF,kdf,kdr]   (basepdesolver.jl:52-68)
  const double kS2f = kv[0], kS2r = kv[1], kG1f = kv[2], kG1r = kv[3], kG2f = kv[4], kG2r = kv[5], kG1p = kv[6],
               kG1dp = kv[7], kSa = kv[8], kSi = kv[9], kp = kv[10], kdp = kv[11], kEGFf = kv[12], kEGFr = kv[13],
               EGF = kv[14], kdf = kv[15], kdr = kv[16];
  double D_Si = Dv[0], D_Sa = Dv[0];
  if (a.o.sfk_mode == GAB1_SFK_MEMBRANE) D_Sa = 1e-32;                                  // basepdesolver.jl:366
  if (a.o.sfk_mode == GAB1_SFK_BOTH_FROZEN) { D_Si = 1e-32; D_Sa = 1e-32; }              // basepdesolver_rect.jl:305-306

  const bool track_t = (a.o.out_mode == GAB1_OUT_FULL || a.o.out_mode == GAB1_OUT_PCT_BOUND);
  const long long nout = a.out_stride;

  // Nt = Int64(ceil(tf/dt)) (basepdesolver.jl:72)
  const double nt_f = ceil(__ddiv_rn(a.o.tf, dt));
  const bool usable = nt_f >= 0.0 && nt_f < 9.0e18;
  const long long Nt = usable ? (long long)nt_f : 0;
  bool done = !have;
  if (have && !usable) {                     // the reference throws InexactError: block of zeros, status only
    for (long long i = hl; i < nout; i += HW) oset[i] = 0.0;
    if (hl == 0) {
      if (a.status) a.status[set] = GAB1_ST_THROW;
      if (a.n_saved) a.n_saved[set] = 0;
      if (a.n_steps) a.n_steps[set] = 0;
      if (a.n_bc) a.n_bc[set] = 0;
    }
    done = true;
  }

  // ---- state: u[q][i] = species q at this lane's i-th node ----
  double u[NCY][K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    // Nt == 0: column 2 is still zero (sapdesolver.jl:245).  Node Nr is not kept on the grid while the loop runs:
    // its slot holds zero and its value lives in the closure lanes (xc)
    const bool on = g.node[i] >= 1 && g.node[i] <= (SKEW ? Nr - 1 : Nr) && Nt > 0;
#pragma unroll
    for (int q = 0; q < NCY; ++q) u[q][i] = 0.0;
    u[iSFK][i] = on ? CoSFK : 0.0;      // basepdesolver.jl:137-140
    u[GAB1][i] = on ? CoG1 : 0.0;
    u[GRB2][i] = on ? CoG2 : 0.0;
    u[SHP2][i] = on ? CoS2 : 0.0;
  }
  const int lane_b = g.G - 1;                                  // (half-relative) lane of the boundary node Nr ...
  constexpr int idx_b = K - 1;                                 // ... always its last slot
  const int lane_i = K >= 2 ? g.G - 1 : g.G - 2;               // inner neighbour Nr-1
  constexpr int idx_i = K >= 2 ? K - 2 : 0;

  // initial column of the FULL output (basepdesolver.jl:94-97,111)
  if (!done && a.o.out_mode == GAB1_OUT_FULL) {
    const unsigned mask = a.o.matrix_mask;
    long long off = 0;
    for (int mi = 0; mi < 12; ++mi) {
      if (!((mask >> mi) & 1u)) continue;
      const double v0 = mi == GAB1_M_iSFK ? CoSFK : mi == GAB1_M_GRB2 ? CoG2 : mi == GAB1_M_SHP2 ? CoS2 : mi == GAB1_M_GAB1 ? CoG1 : 0.0;
      for (int n = hl; n < P; n += HW) oset[off + n] = v0;
      off += (long long)P * Cn;
    }
    if (hl < GAB1_N_VECTORS) oset[off + (long long)hl * Cn] = hl == GAB1_V_mE ? CoEGFR : 0.0;
  }

  double t = 0.0, t_save = a.o.dt_save;
  int nts = 1;
  const double modulus_step = (a.o.save_rule == GAB1_SAVE_MODULUS) ? rint(__ddiv_rn((double)Nt, (double)a.o.Nts)) : 0.0;
  long long bc_total = 0;
  double pct_ave = 0.0, pct_memb = 0.0;      // PCT_BOUND: from snapshot column Nts+1 (zeros if never written)
  bool dead = false;                          // every state value of this half is NaN: nothing can change any more

  // ---- interior constants (per half) ----
  const double inv_dr2 = 1.0 / (a.o.dr * a.o.dr);
  const double kS2f_t = kS2f * dt, kS2r_t = kS2r * dt, kG1f_t = kG1f * dt, kG1r_t = kG1r * dt, kG1p_t = kG1p * dt,
               kG1dp_t = kG1dp * dt, kSi_t = kSi * dt;
  const double l_Si = D_Si * dt * inv_dr2, l_Sa = D_Sa * dt * inv_dr2, l_G1 = Dv[4] * dt * inv_dr2, l_G2 = Dv[1] * dt * inv_dr2,
               l_G2G1 = Dv[2] * dt * inv_dr2, l_S2 = Dv[6] * dt * inv_dr2, l_G1S2 = Dv[5] * dt * inv_dr2,
               l_G2G1S2 = Dv[3] * dt * inv_dr2;
  const double c_Si = fma(-2.0, l_Si, 1.0), c_Sa = fma(-2.0, l_Sa, 1.0) - kSi_t, c_G1 = fma(-2.0, l_G1, 1.0),
               c_G2 = fma(-2.0, l_G2, 1.0), c_G2G1 = fma(-2.0, l_G2G1, 1.0), c_S2 = fma(-2.0, l_S2, 1.0),
               c_G1S2 = fma(-2.0, l_G1S2, 1.0), c_G2G1S2 = fma(-2.0, l_G2G1S2, 1.0);

  // ---- membrane block: lane roles (half-relative lane numbers; LZ holds zeros in both roles) ----
  constexpr int LZ = HW - 1, LE = NMB;                       // LE: Etot pseudo-species
  // closure role, hl = q < 10:  b = (cr*M_num + I)/(1 + cf*M_den)
  double kf = 0.0, kr = 0.0, Dq = 1.0;
  int src_num = LZ, src_den = LZ;
  switch (hl) {
    case iSFK:   kf = kSa;  Dq = D_Si; src_den = LE; break;                     // I/(1 + kSa*Etot*dr/D_S)
    case aSFK:   kf = kSa;  Dq = D_Si; src_num = LE; src_den = LE; break;       // rewritten over the same denominator, see cr
    case GAB1:   kf = kG1f; kr = kG1r; Dq = Dv[4]; src_num = EG2G1;   src_den = EG2;    break;
    case pGAB1:  kf = kG1f; kr = kG1r; Dq = Dv[4]; src_num = EG2PG1;  src_den = EG2;    break;
    case GRB2:   kf = kG2f; kr = kG2r; Dq = Dv[1]; src_num = EG2;     src_den = E;      break;
    case G2G1:   kf = kG2f; kr = kG2r; Dq = Dv[2]; src_num = EG2G1;   src_den = E;      break;
    case G2PG1:  kf = kG2f; kr = kG2r; Dq = Dv[2]; src_num = EG2PG1;  src_den = E;      break;
    case SHP2:   kf = kS2f; kr = kS2r; Dq = Dv[6]; src_num = EG2PG1S; src_den = EG2PG1; break;
    case PG1S:   kf = kG1f; kr = kG1r; Dq = Dv[5]; src_num = EG2PG1S; src_den = EG2;    break;
    case G2PG1S: kf = kG2f; kr = kG2r; Dq = Dv[3]; src_num = EG2PG1S; src_den = E;      break;
    default: break;
  }
  src_num += hbase; src_den += hbase;
  const double drD = a.o.dr / Dq;
  const double cf = kf * drD;
  // aSFK: I_a + ca*Etot*I_i/(1 + cf*Etot) = (I_a + (cf*I_a + ca*I_i)*Etot)/(1 + cf*Etot)   (basepdesolver.jl:206-207);
  // the aSFK lane keeps ca = kSa*dr/D_Sa (a true division: D_Sa may be 1e-32) where the others keep cr = kr*dr/D
  const bool is_a = hl == aSFK;
  const double cr_fixed = is_a ? kSa * (a.o.dr / D_Sa) : kr * drD;
  const bool is_flux = hl >= GAB1 && hl <= G2PG1S;        // this closure's net binding flux feeds the membrane ODEs
  const double kf_t = is_flux ? kf * dt : 0.0, kr_t = is_flux ? kr * dt : 0.0;
  // membrane role, hl = j < 8 (+ Etot on hl 8):  new = base + sa*(F[fs0] + sb*((F[fs1] + F[fs2]) + F[fs3])),
  // F = dt*(kf*M_den*b - kr*M_num) held by the closure lanes (basepdesolver.jl:220-231 regrouped by reaction)
  int fs0 = LZ, fs1 = LZ, fs2 = LZ, fs3 = LZ;
  double sa = 0.0, sb = 0.0;
  switch (hl) {
    case E:       fs0 = GRB2;   fs1 = G2G1;  fs2 = G2PG1; fs3 = G2PG1S; sa = -1.0; sb = 1.0;  break;   // -(all four)
    case EG2:     fs0 = GRB2;   fs1 = GAB1;  fs2 = pGAB1; fs3 = PG1S;   sa = 1.0;  sb = -1.0; break;
    case EG2G1:   fs0 = G2G1;   fs1 = GAB1;  sa = 1.0; sb = 1.0; break;
    case EG2PG1:  fs0 = SHP2;   fs1 = G2PG1; fs2 = pGAB1; sa = -1.0; sb = -1.0; break;                   // -F_S2 + F_G2PG1 + F_pG1
    case EG2PG1S: fs0 = G2PG1S; fs1 = PG1S;  fs2 = SHP2; sa = 1.0; sb = 1.0; break;
    default: break;
  }
  fs0 += hbase; fs1 += hbase; fs2 += hbase; fs3 += hbase;
  // membrane-only reactions, old-time values: f = m*(alpha + alpha2*m) - beta*m_next on hl 0,1,2
  //   hl 0: kEGFf*EGF*mE - kEGFr*mES     hl 1: kdf*mES^2 - kdr*mESmES     hl 2: kp*mESmES - kdp*E
  // and their contribution dm = dt*(s_own*f + s_src*f[f_src])   (Etot: 2*f of hl 2, all bindings cancel)
  double alpha = 0.0, alpha2 = 0.0, beta = 0.0, s_own_t = 0.0, s_src_t = 0.0;
  int f_src = LZ;
  switch (hl) {
    case mE:     alpha = kEGFf * EGF; beta = kEGFr; s_own_t = -dt; break;
    case mES:    alpha2 = kdf;        beta = kdr;   s_own_t = -2.0 * dt; s_src_t = dt; f_src = mE; break;
    case mESmES: alpha = kp;          beta = kdp;   s_own_t = -dt; s_src_t = dt; f_src = mES; break;
    case E:      s_src_t = dt; f_src = mESmES; break;
    case LE:     s_src_t = 2.0 * dt; f_src = mESmES; break;
    default: break;
  }
  f_src += hbase;
  // what b contributes to node Nr-1 in the next interior update: lam_q * cp(Nr-1) * b
  double lcp = 0.0, b0 = 0.0;          // b0: u[Nr](0) = Co, the initial column (basepdesolver.jl:137-140)
  {
    const double cpb = a.o.geometry == GAB1_GEOM_SPHERICAL ? 1.0 + a.o.dr / a.r[Nr - 1] : 1.0;
    switch (hl) {
      case iSFK: lcp = l_Si; b0 = CoSFK; break;
      case aSFK: lcp = l_Sa; break;
      case GAB1: lcp = l_G1; b0 = CoG1; break;
      case pGAB1: lcp = l_G1; break;
      case GRB2: lcp = l_G2; b0 = CoG2; break;
      case G2G1: case G2PG1: lcp = l_G2G1; break;
      case SHP2: lcp = l_S2; b0 = CoS2; break;
      case PG1S: lcp = l_G1S2; break;
      case G2PG1S: lcp = l_G2G1S2; break;
      default: break;
    }
    lcp *= cpb;
    if (Nr < 2) lcp = 0.0;
  }
  const double tol = a.o.tol;
  const bool untracked_c = hl >= NCY, untracked_m = hl >= NMB;
  const unsigned iq_addr = wsh + 8u * (unsigned)(hl < NCY ? hl : 10);       // stage slots 10..15 stay zero
  const int maxiters = a.o.maxiters;
  // xc: boundary value u[Nr] of species hl (closure role); xm: membrane species hl / Etot (membrane role)
  double xc = 0.0;
  double xm = (hl == mE && Nt > 0) ? CoEGFR : 0.0;

  // ---- events.  `ev_step`: the step after which this half must run the reference's exact tests again ----
  bool pulse_pending = a.o.t_prechase >= 0.0;
  if (pulse_pending && a.o.t_prechase + dt > t && t >= a.o.t_prechase) {      // pulsechase_solver.jl:156-158 at step 1
    if (hl == mESmES) alpha = 0.0;
    pulse_pending = false;
  }
  long long step = 1;                       // warp-uniform: both halves start together
  auto plan = [&]() -> long long {
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
    return step + n - 1;
  };
  constexpr long long NEVER = 0x7fffffffffffffffLL;

  // ---- output helpers (rare path; every call site is warp-uniform, `h` names the half being served) ----
  auto membrane_of = [&](int h, double (&m)[NMB]) {
#pragma unroll
    for (int j = 0; j < NMB; ++j) m[j] = shfl(xm, HW * h + j);
  };
  auto write_column = [&](int h, int c) {                       // basepdesolver.jl:268-294
    const bool mine = half == h;
    const int src = HW * h;
    double* o_h = bcast_ptr(oset, src);
    const double t_h = shfl(t, src), CoE_h = shfl(CoEGFR, src);
    double m[NMB];
    membrane_of(h, m);
    const unsigned mask = a.o.matrix_mask;
    long long off = 0;
    constexpr int kSpecies[10] = {iSFK, aSFK, GRB2, GAB1, SHP2, G2G1, G2PG1, G2PG1S, pGAB1, PG1S};   // basepdesolver.jl:271-280
#pragma unroll
    for (int mi = 0; mi < 12; ++mi) {
      if (!((mask >> mi) & 1u)) continue;
      if (mi < 10) {
        const int q = kSpecies[mi];
        pstage_row<K>(rowA, mine, g, Nr, [&](int i) { return u[q][i]; });
      } else if (mi == GAB1_M_PG1tot) {
        pstage_row<K>(rowA, mine, g, Nr, [&](int i) { return derived_ptot<K>(u, i, a.o.pg1tot_form); });
      } else {
        pstage_row<K>(rowA, mine, g, Nr, [&](int i) { return derived_stot<K>(u, i); });
      }
      const bool nan_seen = flush_row(o_h + off + (long long)c * P, rowA, P, lane);
      if (mi == GAB1_M_PG1S && nan_seen && mine) status |= GAB1_ST_NAN;
      off += (long long)P * Cn;
    }
    if (!((mask >> GAB1_M_PG1S) & 1u)) {       // the NaN filter looks at PG1S whether or not it is materialised
      bool ns = false;
#pragma unroll
      for (int i = 0; i < K; ++i) ns |= (g.node[i] >= 1 && g.node[i] <= Nr) && isnan(u[PG1S][i]);
      const unsigned b = __ballot_sync(FULL, ns);
      if (mine && ((b >> hbase) & HMASK)) status |= GAB1_ST_NAN;
    }
    if (lane == src) {
      double* v = o_h + off;
      const double Etot = __dmul_rn(2.0, __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(m[E], m[EG2]), m[EG2G1]), m[EG2PG1]), m[EG2PG1S]));  // :263
      v[GAB1_V_pE * Cn + c] = __ddiv_rn(__dmul_rn(Etot, 100.0), CoE_h);                     // :287
      v[GAB1_V_mE * Cn + c] = m[mE];
      v[GAB1_V_mES * Cn + c] = m[mES];
      v[GAB1_V_mESmES * Cn + c] = m[mESmES];
      v[GAB1_V_E * Cn + c] = m[E];
      v[GAB1_V_EG2 * Cn + c] = m[EG2];
      v[GAB1_V_EG2G1 * Cn + c] = m[EG2G1];
      v[GAB1_V_EG2PG1 * Cn + c] = m[EG2PG1];
      v[GAB1_V_EG2PG1S * Cn + c] = m[EG2PG1S];
      v[GAB1_V_EGFR_SHP2 * Cn + c] = __ddiv_rn(__dmul_rn(m[EG2PG1S], 100.0), CoE_h);        // basepdesolver_rect.jl:264
      v[GAB1_V_t_out * Cn + c] = t_h;
    }
  };
  auto pct_column = [&](int h) {                                // run_base_model.jl:272-273 on the last column
    const bool mine = half == h;
    pstage_row<K>(rowA, mine, g, Nr, [&](int i) { return derived_stot<K>(u, i); });
    const double ave = trapz_r2(a.r, rowA, P);
    const double memb = shfl(xm, HW * h + EG2PG1S);
    if (mine) { pct_ave = ave; pct_memb = memb; }
    __syncwarp();
  };
  // final-time outputs (sapdesolver.jl:245-279, :343-356), PCT epilogue, never-due columns, per-set diagnostics
  auto finalize = [&](int h) {
    const bool mine = half == h;
    const int src = HW * h;
    double* o_h = bcast_ptr(oset, src);
    if (a.o.out_mode == GAB1_OUT_FINAL4) {
      bool ns = false;
      pstage_row<K>(rowA, mine, g, Nr, [&](int i) { return u[iSFK][i]; });
      ns |= flush_row(o_h, rowA, P, lane);
      pstage_row<K>(rowA, mine, g, Nr, [&](int i) { return u[aSFK][i]; });
      ns |= flush_row(o_h + P, rowA, P, lane);
      pstage_row<K>(rowA, mine, g, Nr, [&](int i) { return derived_ptot<K>(u, i, a.o.pg1tot_form); });
      ns |= flush_row(o_h + 2 * P, rowA, P, lane);
      pstage_row<K>(rowA, mine, g, Nr, [&](int i) { return derived_stot<K>(u, i); });
      ns |= flush_row(o_h + 3 * P, rowA, P, lane);
      if (ns && mine) status |= GAB1_ST_NAN;
    } else if (a.o.out_mode == GAB1_OUT_FINAL_STATE) {
      bool ns = false;
#pragma unroll
      for (int q = 0; q < NCY; ++q) {
        pstage_row<K>(rowA, mine, g, Nr, [&](int i) { return u[q][i]; });
        ns |= flush_row(o_h + (long long)q * P, rowA, P, lane);
      }
      double m[NMB];
      membrane_of(h, m);
#pragma unroll
      for (int j = 0; j < NMB; ++j) { if (lane == src) o_h[(long long)NCY * P + j] = m[j]; ns |= isnan(m[j]); }
      if (ns && mine) status |= GAB1_ST_NAN;
    } else if (a.o.out_mode == GAB1_OUT_SIX) {
      pstage_row<K>(rowA, mine, g, Nr, [&](int i) { return u[aSFK][i]; });
      pstage_row<K>(rowB, mine, g, Nr, [&](int i) { return derived_stot<K>(u, i); });
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
      if (threw && mine) status |= GAB1_ST_THROW;
      bool ns = false;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const double v = threw ? 0.0 : six[i];
        ns |= isnan(v);
        if (lane == src) o_h[i] = v;
      }
      if (ns && mine) status |= GAB1_ST_NAN;
    }
    if (mine) {
      if (a.o.out_mode == GAB1_OUT_PCT_BOUND) {      // run_base_model.jl:272-276
        const double R = a.o.R;
        const double ave = __ddiv_rn(__dmul_rn(pct_ave, 3.0), __dmul_rn(__dmul_rn(R, R), R));
        const double mem = __ddiv_rn(__dmul_rn(pct_memb, a.o.pct_mul), a.o.pct_div);
        const double pct = __dmul_rn(__ddiv_rn(__dadd_rn(ave, mem), CoG1), 100.0);
        if (isnan(pct)) status |= GAB1_ST_NAN;
        if (hl == 0) oset[0] = pct;
      }
      if (track_t && nts < Cn) {
        status |= GAB1_ST_SHORT;
        if (a.o.out_mode == GAB1_OUT_FULL) {           // columns nts..Nts were never due: they stay zero in the reference
          long long off = 0;
          for (int mi = 0; mi < 12; ++mi) {
            if (!((a.o.matrix_mask >> mi) & 1u)) continue;
            for (long long i = (long long)nts * P + hl; i < (long long)Cn * P; i += HW) oset[off + i] = 0.0;
            off += (long long)P * Cn;
          }
          for (int v = 0; v < GAB1_N_VECTORS; ++v)
            for (int c = nts + hl; c < Cn; c += HW) oset[off + (long long)v * Cn + c] = 0.0;
        }
      }
      if (hl == 0) {
        if (a.status) a.status[set] = (int)status;
        if (a.n_saved) a.n_saved[set] = track_t ? nts : 0;
        if (a.n_steps) a.n_steps[set] = Nt;
        if (a.n_bc) a.n_bc[set] = bc_total;
      }
      done = true;
    }
    __syncwarp();
  };

  // a set with Nt == 0 takes no step: its outputs are the untouched work arrays
  {
    const unsigned b0 = __ballot_sync(FULL, !done && Nt == 0);
#pragma unroll 1
    for (int h = 0; h < NH; ++h)
      if ((b0 >> (HW * h)) & 1u) finalize(h);
  }

  // ---- the membrane block in pieces, so that the loop can weave it into the interior update ----
  struct Pro { double base, A_t, B_t, rden1, Mn1; };
  // everything that depends only on old-time membrane values (basepdesolver.jl:220-231, the terms without b)
  auto prologue = [&]() -> Pro {
    Pro p;
    const double m_old = xm;
    const double m_next = shflg_down1<HW>(m_old);
    const double f = fma(m_old, fma(alpha2, m_old, alpha), -__dmul_rn(beta, m_next));
    p.base = fma(s_own_t, f, fma(s_src_t, shfl(f, f_src), m_old));
    // the first iterate of the membrane column is the old-time column: these two shuffles serve both the old-time
    // flux coefficients and the first pass of the fixed point
    const double Md1 = shfl(m_old, src_den);
    p.Mn1 = shfl(m_old, src_num);
    p.A_t = __dmul_rn(kf_t, Md1);                                 // F = dt*(kf*M_den*b - kr*M_num), old-time M
    p.B_t = __dmul_rn(kr_t, p.Mn1);
    p.rden1 = fast_recip(fma(cf, Md1, 1.0));                      // 1/(1 + cf*M_den) of the first pass
    return p;
  };
  // Robin closure from the current membrane iterate (basepdesolver.jl:206-215)
  auto closure = [&](double Iq, double cr) -> double {
    const double Mn = shfl(xm, src_num);
    const double Md = shfl(xm, src_den);
    return __dmul_rn(fma(cr, Mn, Iq), fast_recip(fma(cf, Md, 1.0)));
  };
  // membrane species from the closure values of this pass (basepdesolver.jl:220-231)
  auto membrane = [&](const Pro& p, double qv) -> double {
    const double F = fma(p.A_t, qv, -p.B_t);
    const double F0 = shfl(F, fs0), F1 = shfl(F, fs1), F2 = shfl(F, fs2), F3 = shfl(F, fs3);
    return fma(sa, fma(sb, __dadd_rn(__dadd_rn(F1, F2), F3), F0), p.base);
  };
  // this lane's verdict on one pass: `go` = its value asks for another pass, `nanl` = the reference's error is NaN
  auto judge = [&](double qv, double mnew, bool& go, bool& nanl) {
    nanl = false;
    if constexpr (!WHILE) {
      // |1 - new/old| <= tol  <=>  |old - new| <= tol*|old|; the strict `<` also rejects old = new = 0 (0/0 = NaN
      // in the reference) and old = +-Inf, so no special cases remain; NaN operands compare false
      const bool okc = (fabs(__dsub_rn(xc, qv)) < __dmul_rn(tol, fabs(xc))) || untracked_c;
      const bool okm = (fabs(__dsub_rn(xm, mnew)) < __dmul_rn(tol, fabs(xm))) || untracked_m;
      go = !(okc && okm);
    } else {
      // `while error > tol`: a NaN error leaves the loop, so NaN has to be told apart exactly
      const bool spc = !untracked_c && (is_special(xc) || is_special(qv));
      const bool spm = !untracked_m && (is_special(xm) || is_special(mnew));
      int clc, clm;
      if (__any_sync(FULL, spc || spm)) {
        clc = untracked_c ? 0 : classify_exact(xc, qv, tol);
        clm = untracked_m ? 0 : classify_exact(xm, mnew, tol);
      } else {
        clc = (!untracked_c && !(fabs(__dsub_rn(xc, qv)) <= __dmul_rn(tol, fabs(xc)))) ? 1 : 0;
        clm = (!untracked_m && !(fabs(__dsub_rn(xm, mnew)) <= __dmul_rn(tol, fabs(xm)))) ? 1 : 0;
      }
      go = clc == 1 || clm == 1;
      nanl = clc == 2 || clm == 2;
    }
  };
  bool unconv = false, nan_exit = false;
  // does this half need another pass after pass number `pass`?  (`act`: it was still iterating in that pass)
  auto decide = [&](bool act, bool go, bool nanl, int pass) -> bool {
    bool more;
    if constexpr (NH == 1) {
      more = act && __any_sync(FULL, go);                 // `act` is warp-uniform: one set per warp
      if constexpr (WHILE) {
        if (act && __any_sync(FULL, nanl)) { nan_exit = true; more = false; }
      }
    } else {
      const unsigned bgo = __ballot_sync(FULL, act && go);
      more = act && ((bgo >> hbase) & HMASK) != 0u;
      if constexpr (WHILE) {
        const unsigned bnan = __ballot_sync(FULL, act && nanl);
        if (act && ((bnan >> hbase) & HMASK)) { nan_exit = true; more = false; }
      }
    }
    if (more && pass >= maxiters) {
      if constexpr (WHILE) status |= GAB1_ST_ITER_CAP; else unconv = true;
      more = false;
    }
    return more;
  };
  auto any_set = [&](bool p) -> bool {                     // p is uniform within a set
    if constexpr (NH == 1) return p; else return __any_sync(FULL, p);
  };
  const unsigned st_I = wsh, st_b = wsh + 8u * 16u;     // stage: [0,16) inner-neighbour values, [16,32) boundary terms
  // closure lanes publish lam*cp*b for node Nr-1; the lane that owns node Nr-1 adds it and hands the finished
  // u[Nr-1] back as the inner-neighbour value of the next membrane block
  auto fixup = [&](bool publish) {
    if (publish && hl < NCY) sts(st_b + 8u * (unsigned)hl, __dmul_rn(lcp, xc));
    __syncwarp();
    if (hl == lane_i) {
#pragma unroll
      for (int q = 0; q < NCY; q += 2) {
        const double2 c = lds2(st_b + 8u * (unsigned)q);
        u[q][idx_i] = __dadd_rn(u[q][idx_i], c.x);
        u[q + 1][idx_i] = __dadd_rn(u[q + 1][idx_i], c.y);
        sts2(st_I + 8u * (unsigned)q, u[q][idx_i], u[q + 1][idx_i]);
      }
    }
    __syncwarp();
  };
  // dead state: every value of the half is NaN; only the clock and the snapshot schedule still evolve.
  // `behind`: membrane steps still to be accounted for
  auto dead_check = [&](long long behind) -> bool {
    bool newly = false;
    if (any_set(unconv || nan_exit)) {
      bool all_nan = (untracked_c || isnan(xc)) && (hl > LE || isnan(xm));
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int q = 0; q < NCY; ++q) all_nan &= !(g.node[i] >= 1 && g.node[i] <= Nr - 1) || isnan(u[q][i]);
      const unsigned bn = __ballot_sync(FULL, all_nan);
      newly = !(done || dead) && ((bn >> hbase) & HMASK) == HMASK;
      if (newly) {
        dead = true;
        // a NaN error never passes `<= tol` (all maxiters passes run) and leaves `while error > tol` at once
        bc_total += behind * (long long)(WHILE ? 1 : maxiters);
      }
    }
    unconv = false; nan_exit = false;
    return any_set(newly);
  };
  // the whole membrane block on its own (pipeline drain)
  auto membrane_passes = [&](const Pro& p) {
    const double Iq = lds(iq_addr);
    const double cr = is_a ? fma(cf, Iq, __dmul_rn(cr_fixed, lds(st_I + 8u * iSFK))) : cr_fixed;
    bool act = !(done || dead);
    int pass = 1, it_mine = 0;
    {
      // first pass: the membrane iterate is still the old-time column, whose shuffles and reciprocal the prologue holds
      const double qv = __dmul_rn(fma(cr, p.Mn1, Iq), p.rden1);
      const double mnew = membrane(p, qv);
      bool go, nanl;
      judge(qv, mnew, go, nanl);
      if (act) { xc = qv; xm = mnew; it_mine = 1; }
      act = decide(act, go, nanl, 1);
    }
    while (any_set(act)) {
      ++pass;
      const double qv = closure(Iq, cr);
      const double mnew = membrane(p, qv);
      bool go, nanl;
      judge(qv, mnew, go, nanl);
      if (act) { xc = qv; xm = mnew; it_mine = pass; }
      act = decide(act, go, nanl, pass);
    }
    bc_total += it_mine;
  };
  auto membrane_alone = [&]() {
    const Pro p = prologue();
    membrane_passes(p);
  };
  // node Nr on / off the grid around output code
  auto boundary_to_grid = [&]() {
    if (hl < NCY) sts(st_b + 8u * (unsigned)hl, xc);
    __syncwarp();
    if (hl == lane_b) {
#pragma unroll
      for (int q = 0; q < NCY; ++q) u[q][idx_b] = lds(st_b + 8u * (unsigned)q);
    }
    __syncwarp();
  };
  auto boundary_off_grid = [&]() {
    if (hl == lane_b) {
#pragma unroll
      for (int q = 0; q < NCY; ++q) u[q][idx_b] = 0.0;
    }
  };

  long long ev_step = done ? NEVER : plan();
  bool compute;                 // warp-uniform: at least one half still integrates
  int countdown;
  auto arm = [&]() -> bool {    // returns false when both halves are finished
    if (!any_set(!done)) return false;
    compute = any_set(!done && !dead);
    long long e = ev_step;
    if constexpr (NH == 2) {
      const long long eo = bcast_ll(e, lane ^ HW);
      if (eo < e) e = eo;
    }
    e = e - step + 1;
    countdown = (int)(e > 1000000000LL ? 1000000000LL : e);
    return true;
  };
  if (!arm()) return;

  // pipeline state (warp-uniform): `fresh` = the membrane is already at the time level of the interior, i.e. the
  // membrane block of this iteration has nothing to do.  True at the start (b(0) = Co) and after every drain.
  bool fresh = true;
  if (SKEW && hl < NCY) sts(st_b + 8u * (unsigned)hl, __dmul_rn(lcp, b0));
  __syncwarp();

  // plain interior update of every node this lane holds (the skewed loop below carries its own copy, woven with
  // the membrane block); basepdesolver.jl:150-180, in place, left to right; `carry` = cm_i*u[i-1] of the old level
  auto interior_update = [&]() {
    double hr[NCY], carry[NCY];
#pragma unroll
    for (int q = 0; q < NCY; ++q) {
      carry[q] = g.cm[0] * shflg_up1<HW>(u[q][K - 1]);
      hr[q] = shflg_down1<HW>(u[q][0]);
    }
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const double Si = u[iSFK][i], Sa = u[aSFK][i], G1 = u[GAB1][i], pG1 = u[pGAB1][i], G2 = u[GRB2][i],
                   g2g1 = u[G2G1][i], g2pg1 = u[G2PG1][i], S2 = u[SHP2][i], pg1s = u[PG1S][i], g2pg1s = u[G2PG1S][i];
      const double gb = kG1f_t * G2, ph = kG1p_t * Sa, sbd = kS2f_t * S2;
      const double v1 = fma(gb, G1, -(kG1r_t * g2g1));        // GRB2 + GAB1   <-> G2G1
      const double v3 = fma(gb, pG1, -(kG1r_t * g2pg1));      // GRB2 + pGAB1  <-> G2PG1
      const double v5 = fma(gb, pg1s, -(kG1r_t * g2pg1s));    // GRB2 + PG1S   <-> G2PG1S
      const double v2 = fma(ph, G1, -(kG1dp_t * pG1));        // GAB1  <-> pGAB1 (aSFK / phosphatase)
      const double v6 = fma(ph, g2g1, -(kG1dp_t * g2pg1));    // G2G1  <-> G2PG1
      const double v4 = fma(sbd, pG1, -(kS2r_t * pg1s));      // SHP2 + pGAB1  <-> PG1S
      const double v7 = fma(sbd, g2pg1, -(kS2r_t * g2pg1s));  // SHP2 + G2PG1  <-> G2PG1S
      double ks[NCY];                                          // c_q*u + kinetics
      ks[iSFK] = fma(c_Si, Si, kSi_t * Sa);                    // aSFK -> iSFK
      ks[aSFK] = c_Sa * Sa;                                    // (the decay -kSi*aSFK is folded into c_Sa)
      ks[GAB1] = fma(c_G1, G1, -(v1 + v2));
      ks[pGAB1] = fma(c_G1, pG1, (v2 - v3) - v4);
      ks[GRB2] = fma(c_G2, G2, -((v1 + v3) + v5));
      ks[G2G1] = fma(c_G2G1, g2g1, v1 - v6);
      ks[G2PG1] = fma(c_G2G1, g2pg1, (v3 + v6) - v7);
      ks[SHP2] = fma(c_S2, S2, -(v4 + v7));
      ks[PG1S] = fma(c_G1S2, pg1s, v4 - v5);
      ks[G2PG1S] = fma(c_G2G1S2, g2pg1s, v5 + v7);
      const double lam[NCY] = {l_Si, l_Sa, l_G1, l_G1, l_G2, l_G2G1, l_G2G1, l_S2, l_G1S2, l_G2G1S2};
#pragma unroll
      for (int q = 0; q < NCY; ++q) {
        const double up = i + 1 < K ? u[q][i + 1] : hr[q];
        double nb = fma(g.cp[i], up, carry[q]);
        if constexpr (MIRROR) nb = fma(g.m1[i], u[q][i], nb);
        if (i + 1 < K) carry[q] = g.cm[i + 1] * u[q][i];
        u[q][i] = fma(lam[q], nb, ks[q]);
      }
    }
  };

#ifdef GAB1_PHASE_TIMING
  long long ph_wait = 0, ph_int = 0, ph_mem = 0, ph_n = 0;
  const long long ph_t0 = clock64();
#endif
  for (;;) {
    if constexpr (!SKEW) {
      // ===== plain order: interior(step), then membrane(step) on its own.  Leaner in registers and instructions; the
      //       fixed point's latency is covered by the other resident warps instead of by this warp's own interior =====
      if (compute) {
#ifdef GAB1_PHASE_TIMING
        const long long c0 = clock64();
#endif
        // old-time part of the membrane block first: its shuffles and the first reciprocal are in flight while the
        // interior runs (and while this warp waits for its turn)
        const Pro p = prologue();
        if constexpr (TOKEN) tok.acquire();
#ifdef GAB1_PHASE_TIMING
        const long long c1 = clock64();
#endif
        interior_update();
#ifdef GAB1_PHASE_TIMING
        const long long c2 = clock64();
#endif
        if constexpr (TOKEN) tok.release(false);
        if (hl == lane_i) {
#pragma unroll
          for (int q = 0; q < NCY; q += 2) sts2(st_I + 8u * (unsigned)q, u[q][idx_i], u[q + 1][idx_i]);
        }
        __syncwarp();
        membrane_passes(p);
        boundary_to_grid();
        if (dead_check(Nt - step)) countdown = 1;
#ifdef GAB1_PHASE_TIMING
        const long long c3 = clock64();
        ph_wait += c1 - c0; ph_int += c2 - c1; ph_mem += c3 - c2; ++ph_n;
#endif
      }
    } else if (compute) {
      // ===== one iteration: interior(step) woven with membrane(step-1), then the boundary term of node Nr-1 =====
      const bool act0 = !(done || dead) && !fresh;
      const double Iq = lds(iq_addr);
      const double cr = is_a ? fma(cf, Iq, __dmul_rn(cr_fixed, lds(st_I + 8u * iSFK))) : cr_fixed;
      const Pro p = prologue();
      const double qv1 = __dmul_rn(fma(cr, p.Mn1, Iq), p.rden1);            // pass 1: closures
      double m1 = 0.0, qv2 = 0.0, m2 = 0.0;
      bool go1 = false, nan1 = false, go2 = false, nan2 = false, more1 = false, more2 = false;
      int it_mine = 0;

      double hr[NCY], carry[NCY];
#pragma unroll
      for (int q = 0; q < NCY; ++q) {
        carry[q] = g.cm[0] * shflg_up1<HW>(u[q][K - 1]);
        hr[q] = shflg_down1<HW>(u[q][0]);
      }
#pragma unroll
      for (int i = 0; i < K; ++i) {
        // ---- a slice of the membrane block between two nodes ----
        if (i == (K >= 4 ? 1 : 0)) {
          m1 = membrane(p, qv1);                                            // pass 1: membrane species
          judge(qv1, m1, go1, nan1);
          if (act0) { xc = qv1; xm = m1; it_mine = 1; }
          qv2 = closure(Iq, cr);                                            // pass 2: closures (used only if needed)
        }
        if (i == (K >= 4 ? 2 : K - 1)) {
          more1 = decide(act0, go1, nan1, 1);
          m2 = membrane(p, qv2);                                            // pass 2: membrane species
          judge(qv2, m2, go2, nan2);
          if (more1) { xc = qv2; xm = m2; it_mine = 2; }
        }
        // ---- interior node i (basepdesolver.jl:150-180), in place, left to right; `carry` = cm_i*u[i-1] of the old level ----
        const double Si = u[iSFK][i], Sa = u[aSFK][i], G1 = u[GAB1][i], pG1 = u[pGAB1][i], G2 = u[GRB2][i],
                     g2g1 = u[G2G1][i], g2pg1 = u[G2PG1][i], S2 = u[SHP2][i], pg1s = u[PG1S][i], g2pg1s = u[G2PG1S][i];
        const double gb = kG1f_t * G2, ph = kG1p_t * Sa, sbd = kS2f_t * S2;
        const double v1 = fma(gb, G1, -(kG1r_t * g2g1));        // GRB2 + GAB1   <-> G2G1
        const double v3 = fma(gb, pG1, -(kG1r_t * g2pg1));      // GRB2 + pGAB1  <-> G2PG1
        const double v5 = fma(gb, pg1s, -(kG1r_t * g2pg1s));    // GRB2 + PG1S   <-> G2PG1S
        const double v2 = fma(ph, G1, -(kG1dp_t * pG1));        // GAB1  <-> pGAB1 (aSFK / phosphatase)
        const double v6 = fma(ph, g2g1, -(kG1dp_t * g2pg1));    // G2G1  <-> G2PG1
        const double v4 = fma(sbd, pG1, -(kS2r_t * pg1s));      // SHP2 + pGAB1  <-> PG1S
        const double v7 = fma(sbd, g2pg1, -(kS2r_t * g2pg1s));  // SHP2 + G2PG1  <-> G2PG1S
        double ks[NCY];                                          // c_q*u + kinetics
        ks[iSFK] = fma(c_Si, Si, kSi_t * Sa);                    // aSFK -> iSFK
        ks[aSFK] = c_Sa * Sa;                                    // (the decay -kSi*aSFK is folded into c_Sa)
        ks[GAB1] = fma(c_G1, G1, -(v1 + v2));
        ks[pGAB1] = fma(c_G1, pG1, (v2 - v3) - v4);
        ks[GRB2] = fma(c_G2, G2, -((v1 + v3) + v5));
        ks[G2G1] = fma(c_G2G1, g2g1, v1 - v6);
        ks[G2PG1] = fma(c_G2G1, g2pg1, (v3 + v6) - v7);
        ks[SHP2] = fma(c_S2, S2, -(v4 + v7));
        ks[PG1S] = fma(c_G1S2, pg1s, v4 - v5);
        ks[G2PG1S] = fma(c_G2G1S2, g2pg1s, v5 + v7);
        const double lam[NCY] = {l_Si, l_Sa, l_G1, l_G1, l_G2, l_G2G1, l_G2G1, l_S2, l_G1S2, l_G2G1S2};
#pragma unroll
        for (int q = 0; q < NCY; ++q) {
          const double up = i + 1 < K ? u[q][i + 1] : hr[q];
          double nb = fma(g.cp[i], up, carry[q]);
          if constexpr (MIRROR) nb = fma(g.m1[i], u[q][i], nb);
          if (i + 1 < K) carry[q] = g.cm[i + 1] * u[q][i];
          u[q][i] = fma(lam[q], nb, ks[q]);
        }
      }
      // ---- further passes (rare: the first steps of a solve, stiff corners of parameter space) ----
      more2 = decide(more1, go2, nan2, 2);
      if (any_set(more2)) {
        bool act = more2;
        int pass = 2;
        do {
          ++pass;
          const double qv = closure(Iq, cr);
          const double mnew = membrane(p, qv);
          bool go, nanl;
          judge(qv, mnew, go, nanl);
          if (act) { xc = qv; xm = mnew; it_mine = pass; }
          act = decide(act, go, nanl, pass);
        } while (any_set(act));
      }
      bc_total += it_mine;
      fixup(!fresh);             // after a drain (and at the start) the boundary terms are already published
      fresh = false;
      // membrane steps `step`..Nt are still ahead of a half that turns out to be dead now
      if (dead_check(Nt - step + 1)) countdown = 1;             // re-arm: maybe nothing is left to integrate
    }
    t = t + dt;                                                   // basepdesolver.jl:265
    if (--countdown > 0) { ++step; continue; }

    // ---- rare path.  Drain the pipeline: membrane(step) on its own, so that both time levels agree ----
    if (SKEW && compute) {
      if (!fresh) {
        membrane_alone();
        if (hl < NCY) sts(st_b + 8u * (unsigned)hl, __dmul_rn(lcp, xc));     // what the next fixup() adds to node Nr-1
        __syncwarp();
        fresh = true;
        (void)dead_check(Nt - step);
      }
    }
    // exact event tests for the step just taken, per half
    const bool my_ev = !done && step >= ev_step;
    const bool save = track_t && my_ev &&
                      (a.o.save_rule == GAB1_SAVE_T_GE_TSAVE ? (t >= t_save) : (fmod((double)step, modulus_step) == 0.0));
    const unsigned bs = __ballot_sync(FULL, save);
    const unsigned bf = __ballot_sync(FULL, my_ev && step == Nt);
    if (bs | bf) {
      // output code reads node Nr from the grid; the published boundary terms are rewritten below
      if constexpr (SKEW) boundary_to_grid();
#pragma unroll 1
      for (int h = 0; h < NH; ++h) {
        if ((bs >> (HW * h)) & 1u) {
          const int c = __shfl_sync(FULL, nts, HW * h);
          if (c >= Cn) { if (half == h) status |= GAB1_ST_OVERFLOW; }
          else {
            if (half == h) ++nts;
            if (a.o.out_mode == GAB1_OUT_FULL) write_column(h, c);
            else if (c == Cn - 1) pct_column(h);
          }
        }
      }
      if (save && a.o.save_rule == GAB1_SAVE_T_GE_TSAVE) t_save = t_save + a.o.dt_save;
#pragma unroll 1
      for (int h = 0; h < NH; ++h)
        if ((bf >> (HW * h)) & 1u) finalize(h);
      if constexpr (SKEW) {
        boundary_off_grid();
        if (hl < NCY) sts(st_b + 8u * (unsigned)hl, __dmul_rn(lcp, xc));
        __syncwarp();
      }
    }
    if (my_ev && pulse_pending) {                                 // the test the next step would make at its start
      if (a.o.t_prechase + dt > t && t >= a.o.t_prechase) { if (hl == mESmES) alpha = 0.0; pulse_pending = false; }
      else if (t >= a.o.t_prechase + dt) pulse_pending = false;   // the window was stepped over: the reference never switches
    }
    ++step;
    if (done) ev_step = NEVER;
    else if (my_ev) ev_step = plan();
    if (!arm()) break;
  }
#ifdef GAB1_PHASE_TIMING
  if (lane == 0 && blockIdx.x == 3 && threadIdx.x < 64 && item < 3000)
    printf("item %lld warp %d: steps %lld, per step: wait %.0f interior %.0f membrane+rest %.0f, whole loop %.0f cycles\n", item,
           (int)(threadIdx.x >> 5), ph_n, (double)ph_wait / ph_n, (double)ph_int / ph_n, (double)ph_mem / ph_n,
           (double)(clock64() - ph_t0) / ph_n);
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent kernel: every warp pulls PAIRS of parameter sets from a queue ordered by descending work, so the two
// halves of a warp finish within a few steps of each other.
#ifndef GAB1_PAIR_WARPS
#define GAB1_PAIR_WARPS 4
#endif
#ifndef GAB1_PAIR_MINB
#define GAB1_PAIR_MINB 2
#endif
constexpr int kPairWarpsPerCta = GAB1_PAIR_WARPS;

// launch shape of the plain (non-skewed) kernels with at most 2 nodes per lane and set: warps per CTA and CTAs per SM
#ifndef GAB1_PLAIN_WARPS
#define GAB1_PLAIN_WARPS 4
#endif
#ifndef GAB1_PLAIN_MINB
#define GAB1_PLAIN_MINB 2
#endif
template <int K, int HW, bool SKEW, bool TOKEN>
struct Shape {
  static constexpr bool small = !SKEW && K * (32 / HW) <= 2;
  // token kernels: one CTA of eight warps per SM, warps w and w+4 share a scheduler and a token
#ifdef GAB1_STATIC_QUEUE
  // experiment: one warp per CTA and a static, CTA-index-derived assignment of parameter sets, so that ptxas can
  // prove the per-set constants warp-uniform and keep them in uniform registers
  static constexpr int warps = 1;
  static constexpr int minb = GAB1_STATIC_QUEUE;
#else
  static constexpr int warps = TOKEN ? 8 : small ? GAB1_PLAIN_WARPS : GAB1_PAIR_WARPS;
  static constexpr int minb = TOKEN ? 1 : (HW == 32 && K > 2) ? 1 : small ? GAB1_PLAIN_MINB : GAB1_PAIR_MINB;
#endif
};
template <int K, int MODE, bool MIRROR, int HW, bool SKEW, bool TOKEN>
__global__ void __launch_bounds__(32 * Shape<K, HW, SKEW, TOKEN>::warps, Shape<K, HW, SKEW, TOKEN>::minb)
solve_pair_kernel(const KernelArgs a) {
  static_assert(!(TOKEN && SKEW), "the token belongs to the plain loop");
  // the spherical stencil without the mirror term relies on cm = 1 - dr/r[1] being exactly zero; the host side
  // routes any other grid to the general kernel through this flag (set by work_keys_kernel)
  if (a.guard && *a.guard != a.guard_expect) return;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NH = 32 / HW;
  double* ws = smem + (size_t)warp * (NH * WS_HDR + 2 * a.P_pad);
  const int Nr = a.o.Nr;
#pragma unroll
  for (int h = 0; h < NH; ++h) ws[32 * h + lane] = 0.0;
  __syncwarp();

  PGrid<K> g;
  {
    const int hl = lane & (HW - 1);
    g.G = (Nr + K - 1) / K;
    const int off = Nr - g.G * K;                      // <= 0
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const int n = hl * K + i + 1 + off;
      g.node[i] = n;
      const bool interior = n >= 1 && n <= Nr - 1;
      const double r = (n >= 1 && n <= Nr) ? a.r[n] : 1.0;
      const double e = (a.o.geometry == GAB1_GEOM_SPHERICAL) ? a.o.dr / r : 0.0;
      g.cp[i] = interior ? 1.0 + e : 0.0;
      g.cm[i] = (interior && n != 1) ? 1.0 - e : 0.0;   // node 1: its left neighbour is the mirror image (m1) and the slot before it is padding
      g.m1[i] = (interior && n == 1) ? 1.0 - e : 0.0;   // u[0] = u[1]: the mirror term joins the centre (basepdesolver.jl:183-192)
    }
  }
  Token tok;
  if constexpr (TOKEN) {
    constexpr int W = Shape<K, HW, SKEW, TOKEN>::warps;
    int* flags = reinterpret_cast<int*>(smem + (size_t)W * (NH * WS_HDR + 2 * a.P_pad));
    if (threadIdx.x < W) flags[threadIdx.x] = 0;
    __syncthreads();
    const int pair = warp & 3;                      // warps w and w+4 sit on the same scheduler
    tok.leader = warp < 4;
    const int P = 1 + 2 * pair, Q = 2 + 2 * pair;   // P: leader -> follower, Q: follower -> leader
    tok.bar_wait = tok.leader ? Q : P;
    tok.bar_post = tok.leader ? P : Q;
    tok.mine = flags + warp;
    tok.peer = flags + (warp ^ 4);
    tok.published = false;
    tok.peer_done = false;
    if (!tok.leader) asm volatile("bar.arrive %0, 64;" ::"r"(Q) : "memory");    // the leader takes the first turn
  }
  const long long n_items = (a.S + NH - 1) / NH;
#ifdef GAB1_STATIC_QUEUE
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
#else
  for (;;) {
    unsigned item = 0;
    if (lane == 0) item = atomicAdd(a.counter, 1u);
    item = __shfl_sync(FULL, item, 0);
    if ((long long)item >= n_items) break;
#endif
    solve_pair<K, MODE, MIRROR, HW, SKEW, TOKEN>(a, (long long)item, lane, ws, g, tok);
    __syncwarp();
  }
  if constexpr (TOKEN) tok.drain();
}

}  // namespace gab1
