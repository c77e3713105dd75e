// stream_kernel.cuh — primal solver for LARGE radial grids (128 < Nr <= 512): one warp per parameter set with the
// state of the set in the warp's slice of SHARED MEMORY, streamed through the registers node by node.
//
// Why: with K = 8 nodes per lane the register-resident kernels hold 80 doubles of state per thread; ptxas spills the state
// into local memory inside the time loop and the K = 8 kernel runs at 24.6 % of the FP64 peak where K = 4 reaches 59 %
// (DESIGN.md section 5), and K = 16 (dr = 0.025, the reference's finest grid, SURVEY §8a4 "stretch") does not fit at all.
// B200 has 227 KB of shared memory per SM: 10 species x 256 slots are 20 KB.  The interior update walks the lane's K
// nodes with a rolling three-node window (left / current / right, 10 species each), reads every value once and writes it
// once per step, in place; the halo crosses lanes by shuffle as before.  Everything else — arithmetic forms, the
// lane-parallel membrane fixed point, the event countdown, outputs — is the fast path of solver_kernel.cuh, so results
// are held to the same bar (1e-9, identical control flow) by the same tests.
//
// Reference: basepdesolver.jl:149-296, basepdesolver_rect.jl:131-161 (config "rectangular geometry at 4x radial grid
// refinement": dr = 0.05, Nr = 200; dr = 0.025, Nr = 400), sapdesolver*.jl time loops, pulsechase_solver.jl:156-158.
#pragma once
#include "solver_kernel.cuh"

namespace gab1 {

template <int K>
struct SLayout {      // doubles, per warp
  static constexpr int HDR = 0;                       // WS_HDR: [0,16) inner neighbour, [16,32) boundary values
  static constexpr int A = WS_HDR;                    // K * 32: 1/(r*dr) of the lane's nodes (0 for the planar Laplacian)
  static constexpr int U = A + K * 32;                // NCY * K * 32: the state
  static constexpr int ROWS = U + NCY * K * 32;       // rowA, rowB (2 * P_pad) follow
};

// the staged-row writers of solver_kernel.cuh take the state as a register array; here a lane's values come out of smem
template <int K, typename F>
__device__ __forceinline__ void stage_row_s(double* row, int lane, int off, int Nr, F val) {
#pragma unroll 1
  for (int i = 0; i < K; ++i) {
    const int n = lane * K + i + 1 + off;
    if (n >= 1 && n <= Nr) { const double v = val(i); row[n] = v; if (n == 1) row[0] = v; }   // node 0 == node 1
  }
  __syncwarp();
}

template <int K, int MODE>
__device__ void solve_set_streamed(const KernelArgs& a, long long set, int lane, double* ws) {
  constexpr bool WHILE = MODE == MODE_FAST_WHILE;
  typedef SLayout<K> LY;
  const int Nr = a.o.Nr, P = Nr + 1, Nts = a.o.Nts, Cn = Nts + 1;
  double* rowA = ws + LY::ROWS;
  double* rowB = rowA + a.P_pad;
  double* oset = a.out + set * a.out_stride;
  unsigned status = 0;
  const unsigned ws_s = (unsigned)__cvta_generic_to_shared(ws);
  // The state and the grid coefficient are touched by their own lane only: plain accesses (not the volatile lds/sts of
  // the exchange header), so that ptxas may hoist the next node's loads over this node's arithmetic and stores.
  double* const U_ = ws + LY::U + lane;
  double* const A_ = ws + LY::A + lane;
  auto u_at = [&](int q, int i) -> double& { return U_[(q * K + i) * 32]; };
  auto a_at = [&](int i) -> double& { return A_[i * 32]; };
  const int G = (Nr + K - 1) / K;                     // lanes in use; slots are right-aligned on the grid
  const int off = Nr - G * K;                         // node of (lane, i) = lane*K + i + 1 + off

  // ---- parameters of this set (uniform loads) ----
  const double* Co = a.Co + set * a.Co_stride;
  const double* Dv = a.D + set * GAB1_N_D;
  const double* kv = a.k + set * GAB1_N_K;
  const double dt = a.dt[set];
  const double CoSFK = Co[0], CoG2 = Co[1], CoG1 = Co[2], CoS2 = Co[3], CoEGFR = Co[4];
  double D_Si = Dv[0], D_Sa = Dv[0];
  if (a.o.sfk_mode == GAB1_SFK_MEMBRANE) D_Sa = 1e-32;                                  // basepdesolver.jl:366
  if (a.o.sfk_mode == GAB1_SFK_BOTH_FROZEN) { D_Si = 1e-32; D_Sa = 1e-32; }              // basepdesolver_rect.jl:305-306

  const bool track_t = (a.o.out_mode == GAB1_OUT_FULL || a.o.out_mode == GAB1_OUT_PCT_BOUND);
  const long long nout = a.out_stride;

  const double nt_f = ceil(__ddiv_rn(a.o.tf, dt));                                      // basepdesolver.jl:72
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

  // ---- state and grid coefficients into shared memory ----
  {
    const double dr = a.o.dr;
#pragma unroll 1
    for (int i = 0; i < K; ++i) {
      const int n = lane * K + i + 1 + off;
      const bool on = n >= 1 && n <= Nr;
      const double r = on ? a.r[n] : 1.0;
      a_at(i) = (a.o.geometry == GAB1_GEOM_SPHERICAL) ? 1.0 / (r * dr) : 0.0;
#pragma unroll
      for (int q = 0; q < NCY; ++q) u_at(q, i) = 0.0;
      u_at(iSFK, i) = on ? CoSFK : 0.0;          // basepdesolver.jl:137-140
      u_at(GAB1, i) = on ? CoG1 : 0.0;
      u_at(GRB2, i) = on ? CoG2 : 0.0;
      u_at(SHP2, i) = on ? CoS2 : 0.0;
    }
  }
  const double inv_dr2 = 1.0 / (a.o.dr * a.o.dr);
  const int lane_b = G - 1;
  constexpr int idx_b = K - 1;
  const int lane_i = G - 1;                          // K >= 2: node Nr-1 sits next to node Nr in the same lane
  constexpr int idx_i = K - 2;

  // output helpers: one species (or a derived profile) of the whole grid staged into a row
  auto stage_species = [&](double* row, int q) { stage_row_s<K>(row, lane, off, Nr, [&](int i) { return u_at(q, i); }); };
  auto stage_stot = [&](double* row) {
    stage_row_s<K>(row, lane, off, Nr, [&](int i) { return __dadd_rn(u_at(PG1S, i), u_at(G2PG1S, i)); });   // basepdesolver.jl:299
  };
  auto stage_ptot = [&](double* row) {
    stage_row_s<K>(row, lane, off, Nr, [&](int i) {
      const double g2pg1 = u_at(G2PG1, i), pg1 = u_at(pGAB1, i), pg1s = u_at(PG1S, i), g2pg1s = u_at(G2PG1S, i);
      if (a.o.pg1tot_form == GAB1_PG1TOT_VIA_STOT) return __dadd_rn(__dadd_rn(g2pg1, pg1), __dadd_rn(pg1s, g2pg1s));   // :300
      return __dadd_rn(__dadd_rn(__dadd_rn(g2pg1, pg1), pg1s), g2pg1s);                                               // basepdesolver_rect.jl:261
    });
  };
  // one snapshot column of GAB1_OUT_FULL (basepdesolver.jl:268-294)
  auto write_column = [&](int c, const double (&m)[NMB], double t) {
    const unsigned mask = a.o.matrix_mask;
    long long o2 = 0;
    constexpr int kSpecies[10] = {iSFK, aSFK, GRB2, GAB1, SHP2, G2G1, G2PG1, G2PG1S, pGAB1, PG1S};
    bool pg1s_nan = false;
#pragma unroll 1
    for (int mi = 0; mi < 12; ++mi) {
      if (!((mask >> mi) & 1u)) continue;
      if (mi < 10) stage_species(rowA, kSpecies[mi]);
      else if (mi == GAB1_M_PG1tot) stage_ptot(rowA);
      else stage_stot(rowA);
      const bool nan_seen = flush_row(oset + o2 + (long long)c * P, rowA, P, lane);
      if (mi == GAB1_M_PG1S) pg1s_nan = nan_seen;
      o2 += (long long)P * Cn;
    }
    if (!((mask >> GAB1_M_PG1S) & 1u)) {
      bool ns = false;
#pragma unroll 1
      for (int i = 0; i < K; ++i) { const int n = lane * K + i + 1 + off; ns |= (n >= 1 && n <= Nr) && isnan(u_at(PG1S, i)); }
      pg1s_nan = __any_sync(FULL, ns);
    }
    if (pg1s_nan) status |= GAB1_ST_NAN;
    if (lane == 0) {
      double* v = oset + o2;
      const double Etot = __dmul_rn(2.0, __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(m[E], m[EG2]), m[EG2G1]), m[EG2PG1]), m[EG2PG1S]));
      v[GAB1_V_pE * (long long)Cn + c] = __ddiv_rn(__dmul_rn(Etot, 100.0), CoEGFR);
      v[GAB1_V_mE * (long long)Cn + c] = m[mE];
      v[GAB1_V_mES * (long long)Cn + c] = m[mES];
      v[GAB1_V_mESmES * (long long)Cn + c] = m[mESmES];
      v[GAB1_V_E * (long long)Cn + c] = m[E];
      v[GAB1_V_EG2 * (long long)Cn + c] = m[EG2];
      v[GAB1_V_EG2G1 * (long long)Cn + c] = m[EG2G1];
      v[GAB1_V_EG2PG1 * (long long)Cn + c] = m[EG2PG1];
      v[GAB1_V_EG2PG1S * (long long)Cn + c] = m[EG2PG1S];
      v[GAB1_V_EGFR_SHP2 * (long long)Cn + c] = __ddiv_rn(__dmul_rn(m[EG2PG1S], 100.0), CoEGFR);
      v[GAB1_V_t_out * (long long)Cn + c] = t;
    }
  };

  // initial column of the FULL output (basepdesolver.jl:94-97,111)
  if (a.o.out_mode == GAB1_OUT_FULL) {
    const unsigned mask = a.o.matrix_mask;
    long long o2 = 0;
    for (int mi = 0; mi < 12; ++mi) {
      if (!((mask >> mi) & 1u)) continue;
      const double v0 = mi == GAB1_M_iSFK ? CoSFK : mi == GAB1_M_GRB2 ? CoG2 : mi == GAB1_M_SHP2 ? CoS2 : mi == GAB1_M_GAB1 ? CoG1 : 0.0;
      for (int n = lane; n < P; n += 32) oset[o2 + n] = v0;
      o2 += (long long)P * Cn;
    }
    if (lane < GAB1_N_VECTORS) oset[o2 + (long long)lane * Cn] = lane == GAB1_V_mE ? CoEGFR : 0.0;
  }

  double t = 0.0, t_save = a.o.dt_save;
  int nts = 1;
  const double modulus_step = (a.o.save_rule == GAB1_SAVE_MODULUS) ? rint(__ddiv_rn((double)Nt, (double)Nts)) : 0.0;
  long long bc_total = 0;
  double kp_now = kv[10];
  double pct_ave = 0.0, pct_memb = 0.0;
  bool dead = false;
  long long step = 1;

  // ---- interior constants: rate constants and diffusivities pre-scaled by dt ----
  const double kS2f_t = kv[0] * dt, kS2r_t = kv[1] * dt, kG1f_t = kv[2] * dt, kG1r_t = kv[3] * dt,
               kG1p_t = kv[6] * dt, kG1dp_t = kv[7] * dt, kSi_t = kv[9] * dt;
  const double Dt_Si = D_Si * dt, Dt_Sa = D_Sa * dt, Dt_G1 = Dv[4] * dt, Dt_G2 = Dv[1] * dt, Dt_G2G1 = Dv[2] * dt,
               Dt_S2 = Dv[6] * dt, Dt_G1S2 = Dv[5] * dt, Dt_G2G1S2 = Dv[3] * dt;

  // ---- membrane block: lane roles (solver_kernel.cuh) ----
  constexpr int LZ = 31, LE = ML + NMB;
  double kf = 0.0, kr = 0.0, Dq = 1.0;
  int src_num = LZ, src_den = LZ;
  switch (lane) {
    case iSFK:   kf = kv[8]; Dq = D_Si; src_den = LE; break;
    case aSFK:   kf = kv[8]; Dq = D_Si; src_num = LE; src_den = LE; break;
    case GAB1:   kf = kv[2]; kr = kv[3]; Dq = Dv[4]; src_num = ML + EG2G1;   src_den = ML + EG2;    break;
    case pGAB1:  kf = kv[2]; kr = kv[3]; Dq = Dv[4]; src_num = ML + EG2PG1;  src_den = ML + EG2;    break;
    case GRB2:   kf = kv[4]; kr = kv[5]; Dq = Dv[1]; src_num = ML + EG2;     src_den = ML + E;      break;
    case G2G1:   kf = kv[4]; kr = kv[5]; Dq = Dv[2]; src_num = ML + EG2G1;   src_den = ML + E;      break;
    case G2PG1:  kf = kv[4]; kr = kv[5]; Dq = Dv[2]; src_num = ML + EG2PG1;  src_den = ML + E;      break;
    case SHP2:   kf = kv[0]; kr = kv[1]; Dq = Dv[6]; src_num = ML + EG2PG1S; src_den = ML + EG2PG1; break;
    case PG1S:   kf = kv[2]; kr = kv[3]; Dq = Dv[5]; src_num = ML + EG2PG1S; src_den = ML + EG2;    break;
    case G2PG1S: kf = kv[4]; kr = kv[5]; Dq = Dv[3]; src_num = ML + EG2PG1S; src_den = ML + E;      break;
    default: break;
  }
  const double drD = a.o.dr / Dq;
  const double cf = kf * drD;
  const double cr_fixed = kr * drD;
  const double ca = kv[8] * (a.o.dr / D_Sa);            // a true division (D_Sa may be 1e-32)
  const bool is_flux = lane >= GAB1 && lane <= G2PG1S;
  const double kf_t = is_flux ? kf * dt : 0.0, kr_t = is_flux ? kr * dt : 0.0;
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
  double alpha = 0.0, alpha2 = 0.0, beta = 0.0, s_own = 0.0, s_src = 0.0;
  int f_src = LZ;
  switch (lane - ML) {
    case mE:     alpha = kv[12] * kv[14]; beta = kv[13]; s_own = -1.0; break;
    case mES:    alpha2 = kv[15];         beta = kv[16]; s_own = -2.0; s_src = 1.0; f_src = ML + mE; break;
    case mESmES: alpha = kp_now;          beta = kv[11]; s_own = -1.0; s_src = 1.0; f_src = ML + mES; break;
    case E:      s_src = 1.0; f_src = ML + mESmES; break;
    case NMB:    s_src = 2.0; f_src = ML + mESmES; break;
    default: break;
  }
  const double tol = a.o.tol;
  const bool untracked = lane >= LE;
  const int iq_idx = lane < NCY ? lane : 10;
  const bool pulse = a.o.t_prechase >= 0.0;
  const int maxiters = a.o.maxiters;
  double x = (lane == ML + mE) ? CoEGFR : 0.0;

  bool pulse_pending = pulse;
  if (pulse_pending && a.o.t_prechase + dt > t && t >= a.o.t_prechase) {      // pulsechase_solver.jl:156-158 at step 1
    kp_now = 0.0; if (lane == ML + mESmES) alpha = 0.0; pulse_pending = false;
  }
  auto plan = [&]() -> int {
    long long n = Nt - step + 1;
    auto bound = [&](double t_event) {
      const double q = floor((t_event - t) / dt) - 1.0;
      if (!(q >= 1.0)) n = 1;
      else if (q < (double)n) n = (long long)q;
    };
    if (track_t) {
      if (a.o.save_rule == GAB1_SAVE_T_GE_TSAVE) bound(t_save); else n = 1;
    }
    if (pulse_pending) bound(a.o.t_prechase);
    return (int)(n > 1000000000LL ? 1000000000LL : n);
  };
  auto all_state_nan = [&]() -> bool {
    bool all_nan = (lane < ML || lane >= LE) || isnan(x);
#pragma unroll 1
    for (int i = 0; i < K; ++i) {
      const int n = lane * K + i + 1 + off;
      if (n >= 1 && n <= Nr) {
#pragma unroll
        for (int q = 0; q < NCY; ++q) all_nan &= isnan(u_at(q, i));
      }
    }
    return __all_sync(FULL, all_nan);
  };
  int countdown = plan();
  if (Nt >= 1)
  for (;;) {
    // ---- membrane block prologue: everything that depends only on old-time values ----
    const double m_old = x;
    const double m_next = shfl_down1(m_old);
    const double f = fma(m_old, fma(alpha2, m_old, alpha), -(beta * m_next));
    const double base = fma(dt, fma(s_own, f, s_src * shfl(f, f_src)), m_old);
    const double Md1 = shfl(m_old, src_den), Mn1 = shfl(m_old, src_num);
    const double A_t = kf_t * Md1;
    const double B_t = kr_t * Mn1;
    const double rden1 = fast_recip(fma(cf, Md1, 1.0));

    // ---- interior: rolling three-node window over the lane's K nodes, in place (basepdesolver.jl:150-180) ----
    {
      double left[NCY], cur[NCY], hr[NCY];
#pragma unroll
      for (int q = 0; q < NCY; ++q) cur[q] = u_at(q, 0);
#pragma unroll
      for (int q = 0; q < NCY; ++q) left[q] = u_at(q, K - 1);
#pragma unroll
      for (int q = 0; q < NCY; ++q) { left[q] = shfl_up1(left[q]); hr[q] = shfl_down1(cur[q]); }
      const int n0 = lane * K + 1 + off;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        double right[NCY];
        if (i + 1 < K) {
#pragma unroll
          for (int q = 0; q < NCY; ++q) right[q] = u_at(q, i + 1);
        } else {
#pragma unroll
          for (int q = 0; q < NCY; ++q) right[q] = hr[q];
        }
        const double aj = a_at(i);
        const int n = n0 + i;
        const bool interior = n >= 1 && n <= Nr - 1;
        // lap = cp*u[j+1] + cm*u[j-1] + c0*u[j]; node 1 folds its mirror u[0] = u[1] into c0; zero outside the interior
        const double cm_full = inv_dr2 - aj;
        const double cpc = interior ? inv_dr2 + aj : 0.0;
        const double cmc = (interior && n != 1) ? cm_full : 0.0;
        const double c0c = interior ? (n == 1 ? fma(-2.0, inv_dr2, cm_full) : -2.0 * inv_dr2) : 0.0;
        const double Si = cur[iSFK], Sa = cur[aSFK], G1 = cur[GAB1], pG1 = cur[pGAB1], G2 = cur[GRB2],
                     g2g1 = cur[G2G1], g2pg1 = cur[G2PG1], S2 = cur[SHP2], pg1s = cur[PG1S], g2pg1s = cur[G2PG1S];
        const double gb = kG1f_t * G2, ph = kG1p_t * Sa, sb = kS2f_t * S2;
        const double v1 = fma(gb, G1, -(kG1r_t * g2g1));
        const double v3 = fma(gb, pG1, -(kG1r_t * g2pg1));
        const double v5 = fma(gb, pg1s, -(kG1r_t * g2pg1s));
        const double v2 = fma(ph, G1, -(kG1dp_t * pG1));
        const double v6 = fma(ph, g2g1, -(kG1dp_t * g2pg1));
        const double v4 = fma(sb, pG1, -(kS2r_t * pg1s));
        const double v7 = fma(sb, g2pg1, -(kS2r_t * g2pg1s));
        double nw[NCY];
        auto lap = [&](int q) { return fma(cpc, right[q], fma(cmc, left[q], c0c * cur[q])); };
        nw[iSFK] = fma(Dt_Si, lap(iSFK), fma(kSi_t, Sa, Si));
        nw[aSFK] = fma(Dt_Sa, lap(aSFK), fma(-kSi_t, Sa, Sa));
        nw[GAB1] = fma(Dt_G1, lap(GAB1), G1 - v1 - v2);
        nw[pGAB1] = fma(Dt_G1, lap(pGAB1), pG1 - v3 + v2 - v4);
        nw[GRB2] = fma(Dt_G2, lap(GRB2), G2 - v1 - v3 - v5);
        nw[G2G1] = fma(Dt_G2G1, lap(G2G1), g2g1 + v1 - v6);
        nw[G2PG1] = fma(Dt_G2G1, lap(G2PG1), g2pg1 + v3 + v6 - v7);
        nw[SHP2] = fma(Dt_S2, lap(SHP2), S2 - v4 - v7);
        nw[PG1S] = fma(Dt_G1S2, lap(PG1S), pg1s + v4 - v5);
        nw[G2PG1S] = fma(Dt_G2G1S2, lap(G2PG1S), g2pg1s + v5 + v7);
#pragma unroll
        for (int q = 0; q < NCY; ++q) {
          u_at(q, i) = nw[q];
          if (i == idx_i && lane == lane_i) sts(ws_s + 8 * q, nw[q]);      // u+[Nr-1] for the closure lanes
          left[q] = cur[q];
          cur[q] = right[q];
        }
      }
    }
    __syncwarp();
    const double Iq = lds(ws_s + 8 * iq_idx);
    const double cr = lane == aSFK ? fma(cf, Iq, ca * lds(ws_s + 8 * iSFK)) : cr_fixed;

    // ---- fixed-point iterations (basepdesolver.jl:197-242), as solver_kernel.cuh ----
    int it = 1;
    bool unconverged = false, nan_exit = false;
    {
      auto finish_pass = [&](double qv) -> bool {
        const double F = fma(A_t, qv, -B_t);
        const double F0 = shfl(F, fs0), F1 = shfl(F, fs1), F2 = shfl(F, fs2), F3 = shfl(F, fs3);
        const double mnew = fma(sg0, F0, sg1 * F1) + fma(sg2, F2, fma(sg3, F3, base));
        const double xnew = lane < NCY ? qv : mnew;
        if constexpr (!WHILE) {
          const bool ok = (fabs(x - xnew) < tol * fabs(x)) || untracked;
          x = xnew;
          if (__all_sync(FULL, ok)) return false;
          if (it >= maxiters) { unconverged = true; return false; }
          return true;
        } else {
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
      for (int q = 0; q < NCY; ++q) u_at(q, idx_b) = lds(ws_s + 8 * (16 + q));
    }
    if (unconverged || nan_exit) {
      dead = all_state_nan();
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
          if (a.o.out_mode == GAB1_OUT_FULL) write_column(c, m, t);
          else if (c == Cn - 1) {
            stage_stot(rowA);
            pct_ave = trapz_r2(a.r, rowA, P);
            pct_memb = m[EG2PG1S];
            __syncwarp();
          }
        }
        if (a.o.save_rule == GAB1_SAVE_T_GE_TSAVE) t_save = t_save + a.o.dt_save;
      }
    }
    if (pulse_pending) {
      if (a.o.t_prechase + dt > t && t >= a.o.t_prechase) { kp_now = 0.0; if (lane == ML + mESmES) alpha = 0.0; pulse_pending = false; }
      else if (t >= a.o.t_prechase + dt) pulse_pending = false;
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
        if (a.o.out_mode == GAB1_OUT_FULL) write_column(c, m, t);
        else if (c == Cn - 1) { pct_ave = CUDART_NAN; pct_memb = CUDART_NAN; }
      }
      if (a.o.save_rule == GAB1_SAVE_T_GE_TSAVE) t_save = t_save + a.o.dt_save;
    }
  }
  if (Nt == 0) {
#pragma unroll 1
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int q = 0; q < NCY; ++q) u_at(q, i) = 0.0;
#pragma unroll
    for (int j = 0; j < NMB; ++j) m[j] = 0.0;
  }
  // ---- final-time outputs: FINAL4 (sapdesolver.jl:245-279), SIX (:343-356), FINAL_STATE ----
  if (a.o.out_mode == GAB1_OUT_FINAL4) {
    bool ns = false;
    stage_species(rowA, iSFK); ns |= flush_row(oset, rowA, P, lane);
    stage_species(rowA, aSFK); ns |= flush_row(oset + P, rowA, P, lane);
    stage_ptot(rowA); ns |= flush_row(oset + 2 * P, rowA, P, lane);
    stage_stot(rowA); ns |= flush_row(oset + 3 * P, rowA, P, lane);
    if (ns) status |= GAB1_ST_NAN;
  } else if (a.o.out_mode == GAB1_OUT_FINAL_STATE) {
    bool ns = false;
#pragma unroll 1
    for (int q = 0; q < NCY; ++q) { stage_species(rowA, q); ns |= flush_row(oset + (long long)q * P, rowA, P, lane); }
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < NMB; ++j) { oset[(long long)NCY * P + j] = m[j]; ns |= isnan(m[j]); }
    }
    if (__any_sync(FULL, ns)) status |= GAB1_ST_NAN;
  } else if (a.o.out_mode == GAB1_OUT_SIX) {
    stage_species(rowA, aSFK);
    stage_stot(rowB);
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
    if (a.o.out_mode == GAB1_OUT_FULL) {
      long long o2 = 0;
      for (int mi = 0; mi < 12; ++mi) {
        if (!((a.o.matrix_mask >> mi) & 1u)) continue;
        for (long long i = (long long)nts * P + lane; i < (long long)Cn * P; i += 32) oset[o2 + i] = 0.0;
        o2 += (long long)P * Cn;
      }
      for (int v = 0; v < GAB1_N_VECTORS; ++v)
        for (int c = nts + lane; c < Cn; c += 32) oset[o2 + (long long)v * Cn + c] = 0.0;
    }
  }
  if (lane == 0) {
    if (a.status) a.status[set] = (int)status;
    if (a.n_saved) a.n_saved[set] = track_t ? nts : 0;
    if (a.n_steps) a.n_steps[set] = Nt;
    if (a.n_bc) a.n_bc[set] = bc_total;
  }
}

// Persistent kernel, one warp per CTA (a warp's smem slice is 24-54 KB).
// MINB = CTAs (= warps) per SM the register allocation must allow: 8 -> 255 registers, 12 -> 168, 16 -> 128
template <int K, int MODE, int MINB = 8>
__global__ void __launch_bounds__(32, MINB)
stream_kernel(const KernelArgs a) {
  if (a.guard && *a.guard != a.guard_expect) return;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  double* ws = smem;
  ws[lane] = 0.0;
  __syncwarp();
  for (;;) {
    unsigned item = 0;
    if (lane == 0) item = atomicAdd(a.counter, 1u);
    item = __shfl_sync(FULL, item, 0);
    if ((long long)item >= a.S || (a.dyn_count && item >= *a.dyn_count)) break;
    const long long set = a.order ? (long long)a.order[item] : (long long)item;
    solve_set_streamed<K, MODE>(a, set, lane, ws);
    __syncwarp();
  }
}

}  // namespace gab1
