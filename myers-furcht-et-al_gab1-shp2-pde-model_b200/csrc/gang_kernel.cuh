// gang_kernel.cuh — the throughput kernels: SEVERAL parameter sets per warp.
//
// Mapping.  A warp carries NS = 32/G parameter sets at once; the G lanes of a "gang" share one set.  The nodes 1..Nr of
// the set are dealt to the gang in contiguous runs of KN slots per lane (slot k = g*KN + i holds node k+1; slot Nr-1 is
// the boundary node, written by the membrane block; slots beyond are padding).  The state of all NS sets lives in
// shared memory as state[i][species][lane] — every access of the time loop is a compile-time offset from one per-lane
// base address and touches 32 consecutive doubles across the warp: no bank conflicts, no address arithmetic.
//
// Why it is faster than one set per warp (solver_kernel.cuh), measured on B200:
//   * no halo shuffles: a lane walks its KN nodes with a rolling three-node window (left / centre / right, ten species
//     each); the neighbours' edge nodes are read from shared memory like any other node — 10 loads + 10 stores per
//     node-step instead of 62 shuffles + 40 register-pairing moves per step;
//   * the grid fills its slots: 49 interior nodes on 8 x 7 = 56 slots (87.5 %) instead of 64 (76.6 %);
//   * the membrane fixed point — a latency-bound chain of ~230 cycles per pass — is lane-split inside each gang and
//     runs for all NS sets of the warp at once, so its cost per set drops by NS;
//   * per-step bookkeeping (clock, event countdown, loop control) is paid once per NS sets.
// Arithmetic forms are those of the other fast kernels (regrouped stencil of pair_kernel.cuh, hardware-seeded
// reciprocal, net fluxes), so the family is held to the same parity bar: 1e-9 against the oracle (observed ~1e-13)
// with identical step counts, snapshot schedules and membrane-iteration counts.
//
// Reference: basepdesolver.jl:149-296 (time loop), :150-180 (interior), :183-192 (r = 0), :197-242 (membrane loop),
// :265-295 (snapshots); basepdesolver_rect.jl:131-161; sapdesolver.jl:128-242; sapdesolver_memb-SFK.jl:175-222;
// pulsechase_solver.jl:156-158.
#pragma once
#include "solver_kernel.cuh"

namespace gab1 {

template <int G, int KN>
struct GangLayout {
  static constexpr int NS = 32 / G;                    // sets per warp
  static constexpr int SLOTS = G * KN;                 // node slots per set (nodes 1..Nr need Nr <= SLOTS)
  static constexpr int NC = (NCY + G - 1) / G;         // closure tasks per lane
  static constexpr int NO = (NMB + 1 + G - 1) / G;     // membrane tasks per lane (8 species + the Etot pseudo-species)
  // exchange area of one gang: X[0..9] boundary iterate b (= u[Nr+1,2]), X[10..17] membrane iterate, X[18] Etot,
  // F[19..28] binding fluxes, [29..31] zeros (the source of every unused operand), [32..37] per-set constants of the
  // membrane-only reactions.  The gangs' areas are XS doubles apart, XS = 16/NS mod 16, so that the same entry of
  // different gangs never shares a bank
  static constexpr int X_B = 0, X_M = 10, X_ETOT = 18, X_F = 19, X_ZERO = 29, X_PAR = 32;
  static constexpr int XS = 48 + (NS >= 16 ? 1 : 16 / NS);
  // per-lane constants of the membrane tasks, tab[k][lane]: 6 per closure task, 6 per membrane task
  static constexpr int T_CF = 0, T_CRF = NC, T_CAQ = 2 * NC, T_CFA = 3 * NC, T_KFT = 4 * NC, T_KRT = 5 * NC;
  static constexpr int T_SG = 6 * NC, T_WA = 6 * NC + 4 * NO, T_WB = 6 * NC + 5 * NO, NT = 6 * NC + 6 * NO;
  // shared memory of one warp, in doubles
  static constexpr int OFF_STATE = 0;
  static constexpr int OFF_CTAB = KN * NCY * 32;       // c+ / c- per slot: [i][2][G]
  static constexpr int OFF_XCH = OFF_CTAB + 2 * KN * G;
  static constexpr int OFF_TAB = OFF_XCH + NS * XS;
  // scratch: [species][lane] lines where a slot update goes when it must not land (time loop), and the two staged rows of
  // the output writers (rare path, epilogue) — never live at the same time, so they share the space
  static constexpr int OFF_DUMMY = OFF_TAB + NT * 32;
  static constexpr int OFF_ROWS = OFF_DUMMY;
  static __host__ __device__ constexpr int doubles(int P_pad) { return OFF_DUMMY + (NCY * 32 > 2 * P_pad ? NCY * 32 : 2 * P_pad); }
};

// value of species q at node n (0..Nr; node 0 mirrors node 1) of gang s
template <int G, int KN>
__device__ __forceinline__ double gang_node(const double* state, int s, int q, int n) {
  const int k = (n < 1 ? 1 : n) - 1;
  return state[((k % KN) * NCY + q) * 32 + s * G + k / KN];
}

template <int G, int KN, typename F>
__device__ __forceinline__ bool gang_write_row(double* dst, int P, int lane, F val) {
  return store_row_v2(dst, P, lane, val);     // 16-byte stores (solver_kernel.cuh)
}

template <int G, int KN>
__device__ __forceinline__ double gang_stot(const double* st, int s, int n) {
  return __dadd_rn(gang_node<G, KN>(st, s, PG1S, n), gang_node<G, KN>(st, s, G2PG1S, n));           // basepdesolver.jl:299
}
template <int G, int KN>
__device__ __forceinline__ double gang_ptot(const double* st, int s, int n, int form) {
  const double a = __dadd_rn(gang_node<G, KN>(st, s, G2PG1, n), gang_node<G, KN>(st, s, pGAB1, n));
  if (form == GAB1_PG1TOT_VIA_STOT) return __dadd_rn(a, gang_stot<G, KN>(st, s, n));                  // basepdesolver.jl:300
  return __dadd_rn(__dadd_rn(a, gang_node<G, KN>(st, s, PG1S, n)), gang_node<G, KN>(st, s, G2PG1S, n));   // basepdesolver_rect.jl:261
}

// one snapshot column of GAB1_OUT_FULL for gang s, written by the whole warp (basepdesolver.jl:268-294)
template <int G, int KN>
__device__ void gang_write_full_column(const KernelArgs& a, double* oset, int c, const double* st, const double* X, int s,
                                       double t, double CoEGFR, int lane, unsigned& status_bits) {
  const int Nr = a.o.Nr, P = Nr + 1;
  const long long Cn = a.o.Nts + 1;
  const unsigned mask = a.o.matrix_mask;
  long long off = 0;
  constexpr int kSpecies[10] = {iSFK, aSFK, GRB2, GAB1, SHP2, G2G1, G2PG1, G2PG1S, pGAB1, PG1S};   // basepdesolver.jl:271-280
  bool pg1s_nan = false;
#pragma unroll 1
  for (int mi = 0; mi < 12; ++mi) {
    if (!((mask >> mi) & 1u)) continue;
    bool ns;
    double* dst = oset + off + (long long)c * P;
    if (mi < 10) {
      const int q = kSpecies[mi];
      ns = gang_write_row<G, KN>(dst, P, lane, [&](int n) { return gang_node<G, KN>(st, s, q, n); });
    } else if (mi == GAB1_M_PG1tot) {
      ns = gang_write_row<G, KN>(dst, P, lane, [&](int n) { return gang_ptot<G, KN>(st, s, n, a.o.pg1tot_form); });
    } else {
      ns = gang_write_row<G, KN>(dst, P, lane, [&](int n) { return gang_stot<G, KN>(st, s, n); });
    }
    if (mi == GAB1_M_PG1S) pg1s_nan = ns;
    off += (long long)P * Cn;
  }
  if (!((mask >> GAB1_M_PG1S) & 1u)) {       // the NaN filter looks at PG1S whether or not it is materialised
    bool ns = false;
    for (int n = lane; n < P; n += 32) ns |= isnan(gang_node<G, KN>(st, s, PG1S, n));
    pg1s_nan = __any_sync(FULL, ns);
  }
  if (pg1s_nan) status_bits |= GAB1_ST_NAN;
  if (lane == 0) {
    const double* m = X + GangLayout<G, KN>::X_M;
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

// final-time outputs of gang s, written by the whole warp: FINAL4 (sapdesolver.jl:245-279), SIX (:343-356), FINAL_STATE
template <int G, int KN>
__device__ void gang_write_final(const KernelArgs& a, double* oset, const double* st, const double* X, int s, int lane,
                                 double* rowA, double* rowB, unsigned& status_bits, bool zero_state) {
  const int Nr = a.o.Nr, P = Nr + 1;
  auto node = [&](int q, int n) { return zero_state ? 0.0 : gang_node<G, KN>(st, s, q, n); };
  auto stot = [&](int n) { return zero_state ? 0.0 : gang_stot<G, KN>(st, s, n); };
  auto ptot = [&](int n) { return zero_state ? 0.0 : gang_ptot<G, KN>(st, s, n, a.o.pg1tot_form); };
  if (a.o.out_mode == GAB1_OUT_FINAL4) {
    bool ns = false;
    ns |= gang_write_row<G, KN>(oset, P, lane, [&](int n) { return node(iSFK, n); });
    ns |= gang_write_row<G, KN>(oset + P, P, lane, [&](int n) { return node(aSFK, n); });
    ns |= gang_write_row<G, KN>(oset + 2 * P, P, lane, [&](int n) { return ptot(n); });
    ns |= gang_write_row<G, KN>(oset + 3 * P, P, lane, [&](int n) { return stot(n); });
    if (ns) status_bits |= GAB1_ST_NAN;
  } else if (a.o.out_mode == GAB1_OUT_FINAL_STATE) {
    bool ns = false;
#pragma unroll 1
    for (int q = 0; q < NCY; ++q) ns |= gang_write_row<G, KN>(oset + (long long)q * P, P, lane, [&](int n) { return node(q, n); });
    bool mn = false;
    if (lane < NMB) {
      const double v = zero_state ? 0.0 : X[GangLayout<G, KN>::X_M + lane];
      oset[(long long)NCY * P + lane] = v;
      mn = isnan(v);
    }
    if (__any_sync(FULL, mn) || ns) status_bits |= GAB1_ST_NAN;
  } else if (a.o.out_mode == GAB1_OUT_SIX) {
    for (int n = lane; n < P; n += 32) { rowA[n] = node(aSFK, n); rowB[n] = stot(n); }
    __syncwarp();
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
    if (threw) status_bits |= GAB1_ST_THROW;
    bool ns = false;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double v = threw ? 0.0 : six[i];
      ns |= isnan(v);
      if (lane == 0) oset[i] = v;
    }
    if (ns) status_bits |= GAB1_ST_NAN;
  }
}

// ---------------------------------------------------------------------------------------------------------------
template <int G, int KN, int MODE>
__global__ void __launch_bounds__(32, 1) gang_kernel(const KernelArgs a) {
  static_assert(MODE == MODE_FAST_FOR || MODE == MODE_FAST_WHILE, "the gang kernels are fast-arithmetic kernels");
  using L = GangLayout<G, KN>;
  constexpr bool WHILE = MODE == MODE_FAST_WHILE;
  constexpr int NS = L::NS, NC = L::NC, NO = L::NO;
  extern __shared__ double smem[];
  const int lane = threadIdx.x;
  const int s = lane / G, g = lane % G;
  const unsigned gang_mask = (G == 32 ? FULL : ((1u << (G & 31)) - 1u)) << (s * G);
  double* const state = smem + L::OFF_STATE;
  double* const ctab = smem + L::OFF_CTAB;
  double* const X = smem + L::OFF_XCH + s * L::XS;           // this gang's exchange area
  double* const par = X + L::X_PAR;
  double* const tab = smem + L::OFF_TAB + lane;              // tab[k*32]
  double* const dummy = smem + L::OFF_DUMMY + lane;
  double* const rowA = smem + L::OFF_ROWS;
  double* const rowB = rowA + a.P_pad;
  const int Nr = a.o.Nr, P = Nr + 1, Nts = a.o.Nts, Cn = Nts + 1;
  const bool track_t = (a.o.out_mode == GAB1_OUT_FULL || a.o.out_mode == GAB1_OUT_PCT_BOUND);
  const long long nout = a.out_stride;
  const double tol = a.o.tol;
  const int maxiters = a.o.maxiters;
  const bool pulse = a.o.t_prechase >= 0.0;

  // ---- grid coefficients of the regrouped stencil, shared by every set: c+ = 1 + dr/r, c- = 1 - dr/r (planar: 1, 1) ----
  for (int e = lane; e < KN * G; e += 32) {
    const int i = e / G, gg = e % G;
    const int node = gg * KN + i + 1;
    double cp = 0.0, cm = 0.0;
    if (node <= Nr - 1) {
      const double q = (a.o.geometry == GAB1_GEOM_SPHERICAL) ? a.o.dr / a.r[node] : 0.0;
      cp = 1.0 + q;
      cm = 1.0 - q;
    }
    ctab[(i * 2 + 0) * G + gg] = cp;
    ctab[(i * 2 + 1) * G + gg] = cm;
  }
  __syncwarp();

  // per-lane addresses of the time loop
  double* const st = state + lane;                                           // st[(i*NCY + q)*32]
  // left halo of slot 0: the last slot of the lane to the left; the gang's first lane reads its own slot 0 instead,
  // which is the zero-flux mirror u[0] = u[1] (basepdesolver.jl:183-192)
  const double* const hl = (g == 0) ? st : state + ((KN - 1) * NCY) * 32 + (lane - 1);
  // right halo of the last slot: slot 0 of the lane to the right (never used by the gang's last lane: its last slot is
  // the boundary node or padding)
  const double* const hr_p = state + (lane < 31 ? lane + 1 : 31);
  const double* const cb = ctab + g;
  const int n_valid = max(0, min(KN, (Nr - 1) - g * KN));                    // my slots 0..n_valid-1 are interior nodes
  // where the boundary node Nr and its inner neighbour Nr-1 live (same for every gang, shifted by the gang's lanes)
  const int kb = Nr - 1, ki = Nr - 2;
  double* const st_b = state + ((kb % KN) * NCY) * 32 + s * G + kb / KN;     // st_b[q*32]: boundary value of species q
  const double* const st_i = state + ((ki % KN) * NCY) * 32 + s * G + ki / KN;

  for (;;) {
    // ---------------------------------------------------------------------------------------------- next NS sets
    unsigned base_item = 0;
    if (lane == 0) base_item = atomicAdd(a.counter, (unsigned)NS);
    base_item = __shfl_sync(FULL, base_item, 0);
    if ((long long)base_item >= a.S) break;
    const long long item = (long long)base_item + s;
    const bool has = item < a.S;
    // a gang without a set shadows the warp's first set (finite arithmetic, nothing stored)
    const long long set = a.order ? (long long)a.order[has ? item : base_item] : (has ? item : (long long)base_item);

    const double* Co = a.Co + set * a.Co_stride;
    const double* Dv = a.D + set * GAB1_N_D;
    const double* kv = a.k + set * GAB1_N_K;
    const double dt = a.dt[set];
    const double CoSFK = Co[0], CoG2 = Co[1], CoG1 = Co[2], CoS2 = Co[3], CoEGFR = Co[4];
    const double kS2f = kv[0], kS2r = kv[1], kG1f = kv[2], kG1r = kv[3], kG2f = kv[4], kG2r = kv[5], kG1p = kv[6],
                 kG1dp = kv[7], kSa = kv[8], kSi = kv[9], kp = kv[10], kdp = kv[11], kEGFf = kv[12], kEGFr = kv[13],
                 EGF = kv[14], kdf = kv[15], kdr = kv[16];
    double D_Si = Dv[0], D_Sa = Dv[0];
    if (a.o.sfk_mode == GAB1_SFK_MEMBRANE) D_Sa = 1e-32;                                  // basepdesolver.jl:366
    if (a.o.sfk_mode == GAB1_SFK_BOTH_FROZEN) { D_Si = 1e-32; D_Sa = 1e-32; }              // basepdesolver_rect.jl:305-306

    unsigned status = 0;
    bool alive = has;
    // Nt = Int64(ceil(tf/dt)) (basepdesolver.jl:72)
    const double nt_f = ceil(__ddiv_rn(a.o.tf, dt));
    const bool bad_nt = !(nt_f >= 0.0 && nt_f < 9.0e18);
    const long long Nt = bad_nt ? 0 : (long long)nt_f;

    // ---- interior constants: rate constants pre-scaled by dt; lam_q = D_q dt / dr^2, c_q = 1 - 2 lam_q ----
    const double kS2f_t = kS2f * dt, kS2r_t = kS2r * dt, kG1f_t = kG1f * dt, kG1r_t = kG1r * dt, kG1p_t = kG1p * dt,
                 kG1dp_t = kG1dp * dt, kSi_t = kSi * dt;
    const double inv_dr2 = 1.0 / (a.o.dr * a.o.dr);
    const double l_Si = D_Si * dt * inv_dr2, l_Sa = D_Sa * dt * inv_dr2, l_G1 = Dv[4] * dt * inv_dr2, l_G2 = Dv[1] * dt * inv_dr2,
                 l_G2G1 = Dv[2] * dt * inv_dr2, l_S2 = Dv[6] * dt * inv_dr2, l_G1S2 = Dv[5] * dt * inv_dr2,
                 l_G2G1S2 = Dv[3] * dt * inv_dr2;
    const double c_Si = fma(-2.0, l_Si, 1.0), c_Sa = fma(-2.0, l_Sa, 1.0), c_G1 = fma(-2.0, l_G1, 1.0), c_G2 = fma(-2.0, l_G2, 1.0),
                 c_G2G1 = fma(-2.0, l_G2G1, 1.0), c_S2 = fma(-2.0, l_S2, 1.0), c_G1S2 = fma(-2.0, l_G1S2, 1.0),
                 c_G2G1S2 = fma(-2.0, l_G2G1S2, 1.0);
    // (ptxas otherwise re-derives c_q from lam_q at every use to save registers: +10 DFMA per node-step)
    auto pin = [](const double& v) { double w = v; asm volatile("" : "+d"(w)); return w; };
    const double p_Si = pin(c_Si), p_Sa = pin(c_Sa), p_G1 = pin(c_G1), p_G2 = pin(c_G2), p_G2G1 = pin(c_G2G1), p_S2 = pin(c_S2),
                 p_G1S2 = pin(c_G1S2), p_G2G1S2 = pin(c_G2G1S2);

    // ---- membrane block: the gang's lanes share the 10 Robin closures and the 9 membrane updates ----
    // closure task c of lane g: species q = g + c*G:  b = (cr*M_num + I)/(1 + cf*M_den)      (basepdesolver.jl:206-215)
    // The tasks' constants live in shared memory (tab[k][lane]) and their iterates in the exchange area, so that the
    // interior update has the register file to itself; only the small operand offsets stay in registers.
    int cq_[NC], c_on[NC], c_od[NC];
    bool c_ok[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int q = g + c * G;
      cq_[c] = q < NCY ? q : 0;
      c_ok[c] = q < NCY;
      double kf = 0.0, kr = 0.0, Dq = 1.0;
      int on = L::X_ZERO, od = L::X_ZERO;
      switch (q) {
        case iSFK:   kf = kSa;  Dq = D_Si; od = L::X_ETOT; break;                                  // I/(1 + kSa*Etot*dr/D_S)
        case aSFK:   kf = kSa;  Dq = D_Si; on = L::X_ETOT; od = L::X_ETOT; break;                  // over the same denominator
        case GAB1:   kf = kG1f; kr = kG1r; Dq = Dv[4]; on = L::X_M + EG2G1;   od = L::X_M + EG2;    break;
        case pGAB1:  kf = kG1f; kr = kG1r; Dq = Dv[4]; on = L::X_M + EG2PG1;  od = L::X_M + EG2;    break;
        case GRB2:   kf = kG2f; kr = kG2r; Dq = Dv[1]; on = L::X_M + EG2;     od = L::X_M + E;      break;
        case G2G1:   kf = kG2f; kr = kG2r; Dq = Dv[2]; on = L::X_M + EG2G1;   od = L::X_M + E;      break;
        case G2PG1:  kf = kG2f; kr = kG2r; Dq = Dv[2]; on = L::X_M + EG2PG1;  od = L::X_M + E;      break;
        case SHP2:   kf = kS2f; kr = kS2r; Dq = Dv[6]; on = L::X_M + EG2PG1S; od = L::X_M + EG2PG1; break;
        case PG1S:   kf = kG1f; kr = kG1r; Dq = Dv[5]; on = L::X_M + EG2PG1S; od = L::X_M + EG2;    break;
        case G2PG1S: kf = kG2f; kr = kG2r; Dq = Dv[3]; on = L::X_M + EG2PG1S; od = L::X_M + E;      break;
        default: break;
      }
      const double drD = a.o.dr / Dq;
      const double cf_ = kf * drD;
      tab[(L::T_CF + c) * 32] = cf_;
      tab[(L::T_CRF + c) * 32] = q == aSFK ? 0.0 : kr * drD;
      // aSFK: I_a + ca*Etot*I_i/(1 + cf*Etot) = (I_a + (cf*I_a + ca*I_i)*Etot)/(1 + cf*Etot)   (basepdesolver.jl:206-207)
      tab[(L::T_CAQ + c) * 32] = q == aSFK ? kSa * (a.o.dr / D_Sa) : 0.0;          // a true division (D_Sa may be 1e-32)
      tab[(L::T_CFA + c) * 32] = q == aSFK ? cf_ : 0.0;
      const bool is_flux = q >= GAB1 && q <= G2PG1S;
      tab[(L::T_KFT + c) * 32] = is_flux ? kf * dt : 0.0;
      tab[(L::T_KRT + c) * 32] = is_flux ? kr * dt : 0.0;
      c_on[c] = on;
      c_od[c] = od;
    }
    // membrane task d of lane g: species j = g + d*G (j = 8: Etot):  new = base + sum_i sg_i * F[fs_i]
    // (basepdesolver.jl:220-231 regrouped by reaction: every binding term appears once with each sign)
    int o_j[NO], fs0[NO], fs1[NO], fs2[NO], fs3[NO], ia[NO], ib[NO];
    bool o_ok[NO], o_tracked[NO];
#pragma unroll
    for (int d = 0; d < NO; ++d) {
      const int j = g + d * G;
      o_ok[d] = j <= NMB;
      o_tracked[d] = j < NMB;                                     // Etot rides along but is not part of the error
      o_j[d] = j <= NMB ? j : NMB;
      int f0 = L::X_ZERO, f1 = L::X_ZERO, f2 = L::X_ZERO, f3 = L::X_ZERO;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      switch (j) {
        case E:       f0 = L::X_F + GRB2;   f1 = L::X_F + G2G1;  f2 = L::X_F + G2PG1; f3 = L::X_F + G2PG1S; s0 = -1.0; s1 = -1.0; s2 = -1.0; s3 = -1.0; break;
        case EG2:     f0 = L::X_F + GRB2;   f1 = L::X_F + GAB1;  f2 = L::X_F + pGAB1; f3 = L::X_F + PG1S;   s0 = 1.0;  s1 = -1.0; s2 = -1.0; s3 = -1.0; break;
        case EG2G1:   f0 = L::X_F + G2G1;   f1 = L::X_F + GAB1;  s0 = 1.0; s1 = 1.0; break;
        case EG2PG1:  f0 = L::X_F + G2PG1;  f1 = L::X_F + pGAB1; f2 = L::X_F + SHP2; s0 = 1.0; s1 = 1.0; s2 = -1.0; break;
        case EG2PG1S: f0 = L::X_F + G2PG1S; f1 = L::X_F + PG1S;  f2 = L::X_F + SHP2; s0 = 1.0; s1 = 1.0; s2 = 1.0; break;
        default: break;
      }
      fs0[d] = f0; fs1[d] = f1; fs2[d] = f2; fs3[d] = f3;
      tab[(L::T_SG + 4 * d + 0) * 32] = s0; tab[(L::T_SG + 4 * d + 1) * 32] = s1;
      tab[(L::T_SG + 4 * d + 2) * 32] = s2; tab[(L::T_SG + 4 * d + 3) * 32] = s3;
      // membrane-only reactions at the old time level: f0 = kEGFf*EGF*mE - kEGFr*mES, f1 = kdf*mES^2 - kdr*mESmES,
      // f2 = kp*mESmES - kdp*E;  base_j = m_j + dt*(wa*f[ia] + wb*f[ib])
      double a_ = 0.0, b_ = 0.0;
      int ia_ = -1, ib_ = -1;                                       // -1: no such term (its operand is a literal zero)
      switch (j) {
        case mE:     a_ = -1.0; ia_ = 0; break;
        case mES:    a_ = -2.0; ia_ = 1; b_ = 1.0; ib_ = 0; break;
        case mESmES: a_ = -1.0; ia_ = 2; b_ = 1.0; ib_ = 1; break;
        case E:      b_ = 1.0; ib_ = 2; break;
        case NMB:    b_ = 2.0; ib_ = 2; break;
        default: break;
      }
      tab[(L::T_WA + d) * 32] = a_; tab[(L::T_WB + d) * 32] = b_; ia[d] = ia_; ib[d] = ib_;
    }

    // ---- state ----
#pragma unroll
    for (int i = 0; i < KN; ++i) {
      const bool on = g * KN + i + 1 <= Nr;
#pragma unroll
      for (int q = 0; q < NCY; ++q) st[(i * NCY + q) * 32] = 0.0;
      st[(i * NCY + iSFK) * 32] = on ? CoSFK : 0.0;      // basepdesolver.jl:137-140
      st[(i * NCY + GAB1) * 32] = on ? CoG1 : 0.0;
      st[(i * NCY + GRB2) * 32] = on ? CoG2 : 0.0;
      st[(i * NCY + SHP2) * 32] = on ? CoS2 : 0.0;
    }
    // the iterates: u[Nr+1,2] and the membrane column start at zero but for mE (basepdesolver.jl:115-130)
    for (int e = g; e < L::X_PAR; e += G) X[e] = (e == L::X_M + mE) ? CoEGFR : 0.0;
    double kp_now = kp;
    if (g == 0) {
      par[0] = kEGFf * EGF; par[1] = kEGFr; par[2] = kdf; par[3] = kdr; par[4] = kp_now; par[5] = kdp;
    }
    __syncwarp();

    // ---- a set whose step count is unusable: zeros and the THROW status, as the strict kernels report it ----
    if (__any_sync(FULL, has && bad_nt)) {
#pragma unroll 1
      for (int ss = 0; ss < NS; ++ss) {
        if (!__shfl_sync(FULL, (int)(has && bad_nt), ss * G)) continue;
        const long long set_ss = __shfl_sync(FULL, set, ss * G);
        double* o2 = a.out + set_ss * nout;
        for (long long i = lane; i < nout; i += 32) o2[i] = 0.0;
        if (lane == 0) {
          if (a.status) a.status[set_ss] = GAB1_ST_THROW;
          if (a.n_saved) a.n_saved[set_ss] = 0;
          if (a.n_steps) a.n_steps[set_ss] = 0;
          if (a.n_bc) a.n_bc[set_ss] = 0;
        }
      }
      if (bad_nt) alive = false;
    }
    bool reported = has && !bad_nt;                // this gang writes its set's outputs at the end
    int slow = 0;                                  // consecutive steps that ran into the iteration limit

    // ---- initial column of the FULL output (basepdesolver.jl:94-97,111) ----
    if (a.o.out_mode == GAB1_OUT_FULL) {
#pragma unroll 1
      for (int ss = 0; ss < NS; ++ss) {
        if (!__shfl_sync(FULL, (int)reported, ss * G)) continue;
        const long long set_ss = __shfl_sync(FULL, set, ss * G);
        const double* Co2 = a.Co + set_ss * a.Co_stride;
        double* o2 = a.out + set_ss * nout;
        const unsigned mask = a.o.matrix_mask;
        long long off = 0;
        for (int mi = 0; mi < 12; ++mi) {
          if (!((mask >> mi) & 1u)) continue;
          const double v0 = mi == GAB1_M_iSFK ? Co2[0] : mi == GAB1_M_GRB2 ? Co2[1] : mi == GAB1_M_SHP2 ? Co2[3] : mi == GAB1_M_GAB1 ? Co2[2] : 0.0;
          for (int n = lane; n < P; n += 32) o2[off + n] = v0;
          off += (long long)P * Cn;
        }
        if (lane < GAB1_N_VECTORS) o2[off + (long long)lane * Cn] = lane == GAB1_V_mE ? Co2[4] : 0.0;
      }
    }

    double t = 0.0, t_save = a.o.dt_save;
    int nts = 1;
    const double modulus_step = (a.o.save_rule == GAB1_SAVE_MODULUS) ? rint(__ddiv_rn((double)Nt, (double)Nts)) : 0.0;
    long long bc_total = 0;
    double pct_ave = 0.0, pct_memb = 0.0;      // PCT_BOUND: from snapshot column Nts+1 (zeros if never written)
    bool dead = false;                          // every state value is NaN: nothing can change any more
    long long step = 1;
    bool pulse_pending = pulse;
    if (pulse_pending && a.o.t_prechase + dt > t && t >= a.o.t_prechase) {      // pulsechase_solver.jl:156-158 at step 1
      kp_now = 0.0; pulse_pending = false;
      if (g == 0) par[4] = 0.0;
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
    if (Nt < 1) alive = false;
    __syncwarp();

    // ================================================================================================ time loop
    while (__any_sync(FULL, alive)) {
      const int nv = alive ? n_valid : 0;            // a finished gang keeps computing, but stores nothing
      // ---- interior: D*lap + kinetics, explicit Euler, in place (basepdesolver.jl:150-180) ----
      {
        double left[NCY], cur[NCY], nxt[NCY], hr[NCY];
#pragma unroll
        for (int q = 0; q < NCY; ++q) {
          hr[q] = hr_p[q * 32];
          left[q] = hl[q * 32];
          cur[q] = st[q * 32];
        }
        __syncwarp();                                  // every halo value is in registers before any slot is overwritten
#pragma unroll
        for (int i = 0; i < KN; ++i) {
#pragma unroll
          for (int q = 0; q < NCY; ++q) nxt[q] = (i + 1 < KN) ? st[((i + 1) * NCY + q) * 32] : hr[q];
          const double cp = cb[(i * 2 + 0) * G], cm = cb[(i * 2 + 1) * G];
          const double Si = cur[iSFK], Sa = cur[aSFK], G1 = cur[GAB1], pG1 = cur[pGAB1], G2 = cur[GRB2], g2g1 = cur[G2G1],
                       g2pg1 = cur[G2PG1], S2 = cur[SHP2], pg1s = cur[PG1S], g2pg1s = cur[G2PG1S];
          const double gb = kG1f_t * G2, ph = kG1p_t * Sa, sb = kS2f_t * S2;
          const double v1 = fma(gb, G1, -(kG1r_t * g2g1));        // GRB2 + GAB1   <-> G2G1
          const double v3 = fma(gb, pG1, -(kG1r_t * g2pg1));      // GRB2 + pGAB1  <-> G2PG1
          const double v5 = fma(gb, pg1s, -(kG1r_t * g2pg1s));    // GRB2 + PG1S   <-> G2PG1S
          const double v2 = fma(ph, G1, -(kG1dp_t * pG1));        // GAB1  <-> pGAB1 (aSFK / phosphatase)
          const double v6 = fma(ph, g2g1, -(kG1dp_t * g2pg1));    // G2G1  <-> G2PG1
          const double v4 = fma(sb, pG1, -(kS2r_t * pg1s));       // SHP2 + pGAB1  <-> PG1S
          const double v7 = fma(sb, g2pg1, -(kS2r_t * g2pg1s));   // SHP2 + G2PG1  <-> G2PG1S
          double nw[NCY];
          nw[iSFK] = fma(l_Si, fma(cm, left[iSFK], cp * nxt[iSFK]), fma(p_Si, Si, kSi_t * Sa));          // aSFK -> iSFK
          nw[aSFK] = fma(l_Sa, fma(cm, left[aSFK], cp * nxt[aSFK]), fma(p_Sa, Sa, -(kSi_t * Sa)));
          nw[GAB1] = fma(l_G1, fma(cm, left[GAB1], cp * nxt[GAB1]), fma(p_G1, G1, -(v1 + v2)));
          nw[pGAB1] = fma(l_G1, fma(cm, left[pGAB1], cp * nxt[pGAB1]), fma(p_G1, pG1, (v2 - v3) - v4));
          nw[GRB2] = fma(l_G2, fma(cm, left[GRB2], cp * nxt[GRB2]), fma(p_G2, G2, -((v1 + v3) + v5)));
          nw[G2G1] = fma(l_G2G1, fma(cm, left[G2G1], cp * nxt[G2G1]), fma(p_G2G1, g2g1, v1 - v6));
          nw[G2PG1] = fma(l_G2G1, fma(cm, left[G2PG1], cp * nxt[G2PG1]), fma(p_G2G1, g2pg1, (v3 + v6) - v7));
          nw[SHP2] = fma(l_S2, fma(cm, left[SHP2], cp * nxt[SHP2]), fma(p_S2, S2, -(v4 + v7)));
          nw[PG1S] = fma(l_G1S2, fma(cm, left[PG1S], cp * nxt[PG1S]), fma(p_G1S2, pg1s, v4 - v5));
          nw[G2PG1S] = fma(l_G2G1S2, fma(cm, left[G2PG1S], cp * nxt[G2PG1S]), fma(p_G2G1S2, g2pg1s, v5 + v7));
          // every lane computes every slot (straight-line code: the scheduler overlaps the loads, the kinetics and the
          // stencil of neighbouring slots); what is not an interior node of a running set is stored to a scratch line
          double* const dst = (i < nv) ? st + i * NCY * 32 : dummy;
#pragma unroll
          for (int q = 0; q < NCY; ++q) dst[q * 32] = nw[q];
#pragma unroll
          for (int q = 0; q < NCY; ++q) { left[q] = cur[q]; cur[q] = nxt[q]; }
        }
      }
      __syncwarp();

      // ---- membrane block.  Old-time quantities first: the membrane-only reactions, each task's base value, the flux
      //      coefficients and the first pass's reciprocal (the first iterate of the membrane column is the old column) ----
      double base[NO], A_t[NC], B_t[NC], rden1[NC], Mn1[NC], Iq[NC], cr[NC];
      {
        const double m_E = X[L::X_M + mE], m_ES = X[L::X_M + mES], m_mm = X[L::X_M + mESmES], m_e = X[L::X_M + E];
        const double f0 = fma(m_E, par[0], -(par[1] * m_ES));
        const double f1 = fma(m_ES, par[2] * m_ES, -(par[3] * m_mm));
        const double f2 = fma(m_mm, par[4], -(par[5] * m_e));
#pragma unroll
        for (int d = 0; d < NO; ++d) {
          const double fa = ia[d] == 0 ? f0 : (ia[d] == 1 ? f1 : (ia[d] == 2 ? f2 : 0.0));
          const double fb = ib[d] == 0 ? f0 : (ib[d] == 1 ? f1 : (ib[d] == 2 ? f2 : 0.0));
          base[d] = fma(dt, fma(tab[(L::T_WA + d) * 32], fa, tab[(L::T_WB + d) * 32] * fb), X[L::X_M + o_j[d]]);
        }
        const double Ii = st_i[iSFK * 32];                                 // inner-neighbour values u+[Nr-1]
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const double Md1 = X[c_od[c]];
          const double cf_ = tab[(L::T_CF + c) * 32];
          Mn1[c] = X[c_on[c]];
          A_t[c] = tab[(L::T_KFT + c) * 32] * Md1;                        // F = dt*(kf*M_den*b - kr*M_num), old-time M
          B_t[c] = tab[(L::T_KRT + c) * 32] * Mn1[c];
          rden1[c] = fast_recip(fma(cf_, Md1, 1.0));                      // 1/(1 + cf*M_den) of the first pass
          Iq[c] = st_i[cq_[c] * 32];
          cr[c] = fma(tab[(L::T_CFA + c) * 32], Iq[c], fma(tab[(L::T_CAQ + c) * 32], Ii, tab[(L::T_CRF + c) * 32]));
        }
      }

      // ---- fixed-point passes (basepdesolver.jl:197-242); the first pass is peeled: its reciprocal is ready ----
      int it = 1;               // the host routes `maxiters = 0` to the strict kernel: at least one pass runs here
      bool unconverged = false, nan_exit = false;
      {
        bool more = alive;      // this gang still iterates
        bool first = true;
        for (;;) {
          // closures of this pass
          double qv[NC], xb[NC];
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const double Mn = first ? Mn1[c] : X[c_on[c]];
            const double rd = first ? rden1[c] : fast_recip(fma(tab[(L::T_CF + c) * 32], X[c_od[c]], 1.0));
            xb[c] = X[L::X_B + cq_[c]];                                    // my previous iterate
            qv[c] = fma(cr[c], Mn, Iq[c]) * rd;
            if (more && c_ok[c]) X[L::X_F + cq_[c]] = fma(A_t[c], qv[c], -B_t[c]);
          }
          __syncwarp();
          double mnew[NO], xm[NO];
#pragma unroll
          for (int d = 0; d < NO; ++d) {
            const double F0 = X[fs0[d]], F1 = X[fs1[d]], F2 = X[fs2[d]], F3 = X[fs3[d]];
            const double s0 = tab[(L::T_SG + 4 * d + 0) * 32], s1 = tab[(L::T_SG + 4 * d + 1) * 32],
                         s2 = tab[(L::T_SG + 4 * d + 2) * 32], s3 = tab[(L::T_SG + 4 * d + 3) * 32];
            xm[d] = X[L::X_M + o_j[d]];
            mnew[d] = fma(s0, F0, s1 * F1) + fma(s2, F2, fma(s3, F3, base[d]));     // depth 3 instead of 4
          }
          // convergence of my tracked values
          bool my_more = false;
          if constexpr (!WHILE) {
            // |1 - new/old| <= tol  <=>  |old - new| <= tol*|old|; the strict `<` also rejects old = new = 0 (0/0 = NaN
            // in the reference) and old = +-Inf, so no special cases remain; NaN operands compare false
            bool ok = true;
#pragma unroll
            for (int c = 0; c < NC; ++c) ok &= !c_ok[c] || (fabs(xb[c] - qv[c]) < tol * fabs(xb[c]));
#pragma unroll
            for (int d = 0; d < NO; ++d) ok &= !o_tracked[d] || (fabs(xm[d] - mnew[d]) < tol * fabs(xm[d]));
            const unsigned okb = __ballot_sync(FULL, ok);
            const bool all_ok = (okb & gang_mask) == gang_mask;
            if (more) {
              if (all_ok) my_more = false;
              else if (it >= maxiters) { unconverged = true; my_more = false; }
              else my_more = true;
            }
          } else {
            // `while error > tol`: a NaN error leaves the loop, so NaN has to be told apart exactly
            bool special = false;
#pragma unroll
            for (int c = 0; c < NC; ++c) special |= c_ok[c] && (is_special(xb[c]) || is_special(qv[c]));
#pragma unroll
            for (int d = 0; d < NO; ++d) special |= o_tracked[d] && (is_special(xm[d]) || is_special(mnew[d]));
            const bool gang_special = (__ballot_sync(FULL, special) & gang_mask) != 0u;
            bool any_bad = false, any_nan = false;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              if (!c_ok[c]) continue;
              const int cls = gang_special ? classify_exact(xb[c], qv[c], tol) : (!(fabs(xb[c] - qv[c]) <= tol * fabs(xb[c])) ? 1 : 0);
              any_bad |= cls == 1; any_nan |= cls == 2;
            }
#pragma unroll
            for (int d = 0; d < NO; ++d) {
              if (!o_tracked[d]) continue;
              const int cls = gang_special ? classify_exact(xm[d], mnew[d], tol) : (!(fabs(xm[d] - mnew[d]) <= tol * fabs(xm[d])) ? 1 : 0);
              any_bad |= cls == 1; any_nan |= cls == 2;
            }
            const bool g_nan = (__ballot_sync(FULL, any_nan) & gang_mask) != 0u;
            const bool g_bad = (__ballot_sync(FULL, any_bad) & gang_mask) != 0u;
            if (more) {
              if (g_nan) { nan_exit = true; my_more = false; }
              else if (!g_bad) my_more = false;
              else if (it >= maxiters) { status |= GAB1_ST_ITER_CAP; my_more = false; }
              else my_more = true;
            }
          }
          // accept the pass: the iterates go back to the exchange area
          if (more) {
#pragma unroll
            for (int c = 0; c < NC; ++c)
              if (c_ok[c]) X[L::X_B + cq_[c]] = qv[c];
#pragma unroll
            for (int d = 0; d < NO; ++d)
              if (o_ok[d]) X[L::X_M + o_j[d]] = mnew[d];
          }
          more = my_more;
          __syncwarp();
          if (!__any_sync(FULL, more)) break;
          if (more) ++it;
          first = false;
        }
      }
      if (alive) bc_total += it;
      // ---- a set that runs into the iteration limit step after step (it is blowing up, or its fixed point never meets
      //      tol) makes every set of its warp wait for maxiters passes per step: it is handed to the one-set-per-warp
      //      kernel enqueued behind this one, which solves it from the start (the first ~5 steps of every solve hit the
      //      limit too: 0/0 errors on the zero-initialised column, basepdesolver.jl:115-124,238) ----
      if (a.retry_list) {
        const bool struggling = alive && (WHILE ? (status & GAB1_ST_ITER_CAP) != 0u : (unconverged && maxiters >= 4));
        slow = struggling ? slow + 1 : 0;
        if (slow >= (WHILE ? 1 : 12)) {
          if (g == 0) a.retry_list[atomicAdd(a.retry_count, 1u)] = (int)set;
          alive = false;
          reported = false;
          slow = 0;
        }
      }
      // ---- boundary values into the node slot of node Nr ----
      if (alive) {
#pragma unroll
        for (int c = 0; c < NC; ++c)
          if (c_ok[c]) st_b[cq_[c] * 32] = X[L::X_B + cq_[c]];
      }
      __syncwarp();
      if (__any_sync(FULL, alive && (unconverged || nan_exit))) {
        // has the whole state of a gang turned NaN?  (rare: a diverging set)
        bool all_nan = true;
#pragma unroll
        for (int d = 0; d < NO; ++d) all_nan &= !o_tracked[d] || isnan(X[L::X_M + o_j[d]]);
#pragma unroll 1
        for (int i = 0; i < KN; ++i) {
          if (g * KN + i + 1 > Nr) break;
#pragma unroll 1
          for (int q = 0; q < NCY; ++q) all_nan &= isnan(st[(i * NCY + q) * 32]);
        }
        const bool gang_all = (__ballot_sync(FULL, all_nan) & gang_mask) == gang_mask;
        if (alive && (unconverged || nan_exit) && gang_all) { dead = true; countdown = 1; }
      }
      if (alive) { t = t + dt; --countdown; }                                   // basepdesolver.jl:265
      if (!__any_sync(FULL, alive && countdown <= 0)) { if (alive) ++step; continue; }

      // ------------------------------------------------------------------------------------------ rare path
      // exact event tests for the step just taken, gang by gang; the whole warp writes a gang's outputs
      if (alive && countdown > 0) ++step;
#pragma unroll 1
      for (int ss = 0; ss < NS; ++ss) {
        const bool ev = alive && countdown <= 0;
        if (!__shfl_sync(FULL, (int)ev, ss * G)) continue;
        const bool mine = s == ss;
        const long long set_ss = __shfl_sync(FULL, set, ss * G);
        double* const o2 = a.out + set_ss * nout;
        const double* const X2 = smem + L::OFF_XCH + ss * L::XS;
        const double t_ss = __shfl_sync(FULL, t, ss * G);
        if (track_t) {
          const long long step_ss = __shfl_sync(FULL, step, ss * G);
          const double tsave_ss = __shfl_sync(FULL, t_save, ss * G);
          const double mstep_ss = __shfl_sync(FULL, modulus_step, ss * G);
          const bool save = a.o.save_rule == GAB1_SAVE_T_GE_TSAVE ? (t_ss >= tsave_ss) : (fmod((double)step_ss, mstep_ss) == 0.0);
          if (save) {
            const int nts_ss = __shfl_sync(FULL, nts, ss * G);
            unsigned st_bits = 0;
            if (nts_ss >= Cn) st_bits |= GAB1_ST_OVERFLOW;
            else {
              const int c = nts_ss;
              if (mine) ++nts;
              if (a.o.out_mode == GAB1_OUT_FULL) {
                const double CoE = __shfl_sync(FULL, CoEGFR, ss * G);
                gang_write_full_column<G, KN>(a, o2, c, state, X2, ss, t_ss, CoE, lane, st_bits);
              } else if (c == Cn - 1) {
                for (int n = lane; n < P; n += 32) rowA[n] = gang_stot<G, KN>(state, ss, n);
                __syncwarp();
                const double pa = trapz_r2(a.r, rowA, P);
                if (mine) { pct_ave = pa; pct_memb = X2[L::X_M + EG2PG1S]; }
                __syncwarp();
              }
            }
            if (mine) {
              status |= st_bits;
              if (a.o.save_rule == GAB1_SAVE_T_GE_TSAVE) t_save = t_save + a.o.dt_save;
            }
          }
        }
        if (mine && pulse_pending) {                                  // the test the next step would make at its start
          if (a.o.t_prechase + dt > t && t >= a.o.t_prechase) { kp_now = 0.0; pulse_pending = false; if (g == 0) par[4] = 0.0; }
          else if (t >= a.o.t_prechase + dt) pulse_pending = false;   // the window was stepped over: the reference never switches
        }
        if (mine) {
          ++step;
          if (dead || step > Nt) alive = false;
          else countdown = plan();
        }
        __syncwarp();
      }
    }

    // ================================================================================================ epilogue, gang by gang
#pragma unroll 1
    for (int ss = 0; ss < NS; ++ss) {
      if (!__shfl_sync(FULL, (int)reported, ss * G)) continue;
      const bool mine = s == ss;
      const long long set_ss = __shfl_sync(FULL, set, ss * G);
      double* const o2 = a.out + set_ss * nout;
      const double* const X2 = smem + L::OFF_XCH + ss * L::XS;
      const long long Nt_ss = __shfl_sync(FULL, Nt, ss * G);
      const double dt_ss = __shfl_sync(FULL, dt, ss * G);
      const double CoE = __shfl_sync(FULL, CoEGFR, ss * G);
      const double mstep_ss = __shfl_sync(FULL, modulus_step, ss * G);
      long long step_ss = __shfl_sync(FULL, step, ss * G);
      double t_ss = __shfl_sync(FULL, t, ss * G);
      double tsave_ss = __shfl_sync(FULL, t_save, ss * G);
      int nts_ss = __shfl_sync(FULL, nts, ss * G);
      long long bc_ss = __shfl_sync(FULL, bc_total, ss * G);
      unsigned st_ss = __shfl_sync(FULL, status, ss * G);
      double pa_ss = __shfl_sync(FULL, pct_ave, ss * G), pm_ss = __shfl_sync(FULL, pct_memb, ss * G);
      // ---- all-NaN state: only the clock and the snapshot schedule still evolve ----
      for (; step_ss <= Nt_ss; ++step_ss) {
        const long long per = WHILE ? 1 : maxiters;
        if (!track_t) { bc_ss += (Nt_ss - step_ss + 1) * per; break; }
        bc_ss += per;
        t_ss = t_ss + dt_ss;
        const bool save = a.o.save_rule == GAB1_SAVE_T_GE_TSAVE ? (t_ss >= tsave_ss) : (fmod((double)step_ss, mstep_ss) == 0.0);
        if (save) {
          if (nts_ss >= Cn) st_ss |= GAB1_ST_OVERFLOW;
          else {
            const int c = nts_ss++;
            if (a.o.out_mode == GAB1_OUT_FULL) gang_write_full_column<G, KN>(a, o2, c, state, X2, ss, t_ss, CoE, lane, st_ss);
            else if (c == Cn - 1) { pa_ss = CUDART_NAN; pm_ss = CUDART_NAN; }
          }
          if (a.o.save_rule == GAB1_SAVE_T_GE_TSAVE) tsave_ss = tsave_ss + a.o.dt_save;
        }
      }
      gang_write_final<G, KN>(a, o2, state, X2, ss, lane, rowA, rowB, st_ss, Nt_ss == 0);
      if (a.o.out_mode == GAB1_OUT_PCT_BOUND) {      // run_base_model.jl:272-276
        const double R = a.o.R;
        const double CoG1_ss = __shfl_sync(FULL, CoG1, ss * G);
        const double ave = __ddiv_rn(__dmul_rn(pa_ss, 3.0), __dmul_rn(__dmul_rn(R, R), R));
        const double mem = __ddiv_rn(__dmul_rn(pm_ss, a.o.pct_mul), a.o.pct_div);
        const double pct = __dmul_rn(__ddiv_rn(__dadd_rn(ave, mem), CoG1_ss), 100.0);
        if (isnan(pct)) st_ss |= GAB1_ST_NAN;
        if (lane == 0) o2[0] = pct;
      }
      if (track_t && nts_ss < Cn) {
        st_ss |= GAB1_ST_SHORT;
        if (a.o.out_mode == GAB1_OUT_FULL) {           // columns nts..Nts were never due: they stay zero in the reference
          long long off = 0;
          for (int mi = 0; mi < 12; ++mi) {
            if (!((a.o.matrix_mask >> mi) & 1u)) continue;
            for (long long i = (long long)nts_ss * P + lane; i < (long long)Cn * P; i += 32) o2[off + i] = 0.0;
            off += (long long)P * Cn;
          }
          for (int v = 0; v < GAB1_N_VECTORS; ++v)
            for (int c = nts_ss + lane; c < Cn; c += 32) o2[off + (long long)v * Cn + c] = 0.0;
        }
      }
      if (lane == 0) {
        if (a.status) a.status[set_ss] = (int)st_ss;
        if (a.n_saved) a.n_saved[set_ss] = track_t ? nts_ss : 0;
        if (a.n_steps) a.n_steps[set_ss] = Nt_ss;
        if (a.n_bc) a.n_bc[set_ss] = bc_ss;
      }
      (void)mine;
      __syncwarp();
    }
    __syncwarp();
  }
}

}  // namespace gab1
