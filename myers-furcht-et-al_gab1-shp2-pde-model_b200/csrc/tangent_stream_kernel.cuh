// tangent_stream_kernel.cuh — forward-mode kernel, second generation: the PRIMAL state of a set lives in registers
// (as in solver_kernel.cuh) while the partials of all NT directions live in the warp's slice of shared memory and are
// streamed through the registers one direction at a time.
//
// Why: tangent_kernel.cuh keeps value + NT partials of all ten species in registers; beyond NT = 1 (K = 2) ptxas spills
// the state into local memory inside the time loop (measured on B200: 4 partials at dr = 0.2 cost 11.6x a primal solve
// with NT = 2, 28x with NT = 4; at dr = 0.1 27x with NT = 1).  B200 has 227 KB of shared memory per SM: 4 partials of a
// 64-slot grid are 20 KB per warp.  With the partials there, ONE work item carries all directions of a set, the primal
// is computed once instead of once per direction group, and a direction's working set (10 species of a 3-node window)
// fits the register file next to the primal state.
//
// Per time step and lane:
//   P0  primal Laplacians L_q(u) of the lane's K nodes (halo by shuffle)           -> smem LV (read by every direction)
//   T_n for each direction n: rolling 3-node window over the lane's nodes; partials of the 7 net fluxes by the product
//       rule against the OLD primal values in registers; u'_q += kin'_q + Dt'_q L_q(u) + Dt_q L_q(u')   (in place, smem)
//   P1  primal update in place (registers), as solver_kernel.cuh's fast path
//   M   membrane fixed point, lane-parallel on duals (as tangent_kernel.cuh); the partials of the per-lane coefficients
//       are read from shared memory at the start of the block, when the interior's working set is dead
// What it stands in for, arithmetic and tolerances: see tangent_kernel.cuh.
#pragma once
#include "tangent_kernel.cuh"

namespace gab1 {

// indices of the interior constants whose partials live in smem (CP): rate constants and diffusivities pre-scaled by dt
enum { C_kS2f, C_kS2r, C_kG1f, C_kG1r, C_kG1p, C_kG1dp, C_kSi, C_DSi, C_DSa, C_DG1, C_DG2, C_DG2G1, C_DS2, C_DG1S2,
       C_DG2G1S2, C_dt, C_ca, C_N };
// per-lane membrane coefficients whose partials live in smem (LP)
enum { L_cf, L_cr, L_kft, L_krt, L_alpha, L_alpha2, L_beta, L_N };

template <int K, int NT>
struct TSLayout {     // doubles, per warp
  static constexpr int HDR = 0;                                   // TWS_HDR * (1 + NT): inner-neighbour / boundary exchange
  static constexpr int CP = HDR + TWS_HDR * (1 + NT);              // C_N * NT, uniform
  static constexpr int LP = CP + ((C_N * NT + 3) & ~3);            // L_N * NT * 32, lane-indexed
  static constexpr int SP = LP + L_N * NT * 32;                    // NT * NCY * K * 32: the partials of the state
  static constexpr int LV = SP + NT * NCY * K * 32;                // NCY * K * 32: primal Laplacians of this step
  static constexpr int ROWS = LV + NCY * K * 32;                   // rowA, rowB (2 * P_pad) follow
};

template <int K, int NT>
__device__ void solve_set_tangent_stream(const TangentArgs& ta, long long set, int group, int lane, double* ws, const Grid<K>& g) {
  typedef Dn<NT> T;
  typedef TSLayout<K, NT> LY;
  const KernelArgs& a = ta.a;
  const int Nr = a.o.Nr, P = Nr + 1, Nts = a.o.Nts, Cn = Nts + 1;
  const long long nout = a.out_stride;
  double* rowA = ws + LY::ROWS;
  double* rowB = rowA + a.P_pad;
  double* oset = a.out + set * nout * (1 + ta.n_dir);
  const bool lead = group == 0;
  unsigned status = 0, scratch_status = 0;
  const unsigned ws_s = (unsigned)__cvta_generic_to_shared(ws);
  // The partials, this step's primal Laplacians and the coefficient partials are read and written by their own lane only
  // (the uniform ones are written once by lane 0 behind a __syncwarp): plain accesses, not the volatile lds/sts of the
  // exchange header, so that ptxas may hoist a node's loads over the previous node's arithmetic and stores.
  double* const SP_ = ws + LY::SP + lane;
  double* const LV_ = ws + LY::LV + lane;
  double* const CP_ = ws + LY::CP;
  double* const LP_ = ws + LY::LP + lane;
  auto sp_at = [&](int n, int q, int i) -> double& { return SP_[((n * NCY + q) * K + i) * 32]; };
  auto lv_at = [&](int q, int i) -> double& { return LV_[(q * K + i) * 32]; };
  auto cp_at = [&](int c, int n) -> double& { return CP_[c * NT + n]; };
  auto lp_at = [&](int c, int n) -> double& { return LP_[(c * NT + n) * 32]; };

  auto block_of = [&](int c) -> double* {
    if (c == 0) return lead ? oset : nullptr;
    const int d = group * NT + (c - 1);
    return d < ta.n_dir ? oset + (long long)(1 + d) * nout : nullptr;
  };

  // ---- parameters and their partials (uniform loads) ----
  const double* Cov = a.Co + set * a.Co_stride;
  const double* Dv = a.D + set * GAB1_N_D;
  const double* kv = a.k + set * GAB1_N_K;
  const double* sd[NT];
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const int d = group * NT + n;
    sd[n] = d < ta.n_dir ? ta.seeds + (set * ta.n_dir + d) * GAB1_N_SEED : nullptr;
  }
  auto seed = [&](int n, int i) -> double { return sd[n] ? sd[n][i] : 0.0; };
  auto Dd = [&](int i) { T r; r.v = Dv[i];
#pragma unroll
    for (int n = 0; n < NT; ++n) r.p[n] = seed(n, i); return r; };
  auto kd = [&](int i) { T r; r.v = kv[i];
#pragma unroll
    for (int n = 0; n < NT; ++n) r.p[n] = seed(n, GAB1_N_D + i); return r; };
  auto Cod = [&](int i) { T r; r.v = Cov[i];
#pragma unroll
    for (int n = 0; n < NT; ++n) r.p[n] = seed(n, GAB1_N_D + GAB1_N_K + i); return r; };
  T dt; dt.v = a.dt[set];
#pragma unroll
  for (int n = 0; n < NT; ++n) dt.p[n] = seed(n, GAB1_N_SEED - 1);
  const double dtv = dt.v;

  const bool track_t = (a.o.out_mode == GAB1_OUT_FULL || a.o.out_mode == GAB1_OUT_PCT_BOUND);

  const double nt_f = ceil(__ddiv_rn(a.o.tf, dtv));
  if (!(nt_f >= 0.0 && nt_f < 9.0e18)) {
    for (int c = 0; c <= NT; ++c) {
      double* ob = block_of(c);
      if (ob) for (long long i = lane; i < nout; i += 32) ob[i] = 0.0;
    }
    if (lead && lane == 0) {
      if (a.status) a.status[set] = GAB1_ST_THROW;
      if (a.n_saved) a.n_saved[set] = 0;
      if (a.n_steps) a.n_steps[set] = 0;
      if (a.n_bc) a.n_bc[set] = 0;
    }
    return;
  }
  const long long Nt = (long long)nt_f;

  // ---- state: values in registers, partials in shared memory ----
  double uv[NCY][K];
  double CoG1v, CoEGFRv;
  double CoG1p[NT], CoEGFRp[NT];
  {
    const T CoSFK = Cod(0), CoG2 = Cod(1), CoG1 = Cod(2), CoS2 = Cod(3), CoEGFR = Cod(4);
    CoG1v = CoG1.v; CoEGFRv = CoEGFR.v;
#pragma unroll
    for (int n = 0; n < NT; ++n) { CoG1p[n] = CoG1.p[n]; CoEGFRp[n] = CoEGFR.p[n]; }
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const bool on = g.node[i] >= 1 && g.node[i] <= Nr;
#pragma unroll
      for (int q = 0; q < NCY; ++q) {
        uv[q][i] = 0.0;
#pragma unroll
        for (int n = 0; n < NT; ++n) sp_at(n, q, i) = 0.0;
      }
      uv[iSFK][i] = on ? CoSFK.v : 0.0;      // basepdesolver.jl:776-779
      uv[GAB1][i] = on ? CoG1.v : 0.0;
      uv[GRB2][i] = on ? CoG2.v : 0.0;
      uv[SHP2][i] = on ? CoS2.v : 0.0;
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        sp_at(n, iSFK, i) = on ? CoSFK.p[n] : 0.0;
        sp_at(n, GAB1, i) = on ? CoG1.p[n] : 0.0;
        sp_at(n, GRB2, i) = on ? CoG2.p[n] : 0.0;
        sp_at(n, SHP2, i) = on ? CoS2.p[n] : 0.0;
      }
    }
    // initial column of the FULL output, every component
    if (a.o.out_mode == GAB1_OUT_FULL) {
#pragma unroll
      for (int c = 0; c <= NT; ++c) {
        double* ob = block_of(c);
        if (!ob) continue;
        auto comp = [&](const T& x) { return c == 0 ? x.v : x.p[c > 0 ? c - 1 : 0]; };
        long long off = 0;
        for (int mi = 0; mi < 12; ++mi) {
          if (!((a.o.matrix_mask >> mi) & 1u)) continue;
          const double v0 = mi == GAB1_M_iSFK ? comp(CoSFK) : mi == GAB1_M_GRB2 ? comp(CoG2) : mi == GAB1_M_SHP2 ? comp(CoS2)
                            : mi == GAB1_M_GAB1 ? comp(CoG1) : 0.0;
          for (int nn = lane; nn < P; nn += 32) ob[off + nn] = v0;
          off += (long long)P * Cn;
        }
        if (lane < GAB1_N_VECTORS) ob[off + (long long)lane * Cn] = lane == GAB1_V_mE ? comp(CoEGFR) : 0.0;
      }
    }
  }

  const int lane_b = g.G - 1;
  constexpr int idx_b = K - 1;
  const int lane_i = K >= 2 ? g.G - 1 : g.G - 2;
  constexpr int idx_i = K >= 2 ? K - 2 : 0;

  // ---- interior constants: values in registers, partials in smem ----
  double cv[C_ca + 1];
  // ---- membrane block: lane roles (solver_kernel.cuh, fast path) ----
  constexpr int LZ = 31, LE = ML + NMB;
  int src_num = LZ, src_den = LZ;
  double lv_[L_N];
  double cav;
  {
    T D_Si = Dd(0), D_Sa = Dd(0);
    if (a.o.sfk_mode == GAB1_SFK_MEMBRANE) D_Sa = dconst<NT>(1e-32);                                   // basepdesolver.jl:366
    if (a.o.sfk_mode == GAB1_SFK_BOTH_FROZEN) { D_Si = dconst<NT>(1e-32); D_Sa = dconst<NT>(1e-32); }  // basepdesolver_rect.jl:305-306
    auto put_c = [&](int c, const T& x) { cv[c] = x.v;
      if (lane == 0) {
#pragma unroll
        for (int n = 0; n < NT; ++n) cp_at(c, n) = x.p[n];
      } };
    put_c(C_kS2f, kd(0) * dt); put_c(C_kS2r, kd(1) * dt); put_c(C_kG1f, kd(2) * dt); put_c(C_kG1r, kd(3) * dt);
    put_c(C_kG1p, kd(6) * dt); put_c(C_kG1dp, kd(7) * dt); put_c(C_kSi, kd(9) * dt);
    put_c(C_DSi, D_Si * dt); put_c(C_DSa, D_Sa * dt); put_c(C_DG1, Dd(4) * dt); put_c(C_DG2, Dd(1) * dt);
    put_c(C_DG2G1, Dd(2) * dt); put_c(C_DS2, Dd(6) * dt); put_c(C_DG1S2, Dd(5) * dt); put_c(C_DG2G1S2, Dd(3) * dt);
    put_c(C_dt, dt);
    const T ca = kd(8) * drdiv<NT>(a.o.dr, D_Sa);          // aSFK closure coefficient; a true division (D_Sa may be 1e-32)
    put_c(C_ca, ca);
    cav = ca.v;

    T kf = dconst<NT>(0.0), kr = dconst<NT>(0.0), Dq = dconst<NT>(1.0);
    switch (lane) {
      case iSFK:   kf = kd(8); Dq = D_Si; src_den = LE; break;
      case aSFK:   kf = kd(8); Dq = D_Si; src_num = LE; src_den = LE; break;
      case GAB1:   kf = kd(2); kr = kd(3); Dq = Dd(4); src_num = ML + EG2G1;   src_den = ML + EG2;    break;
      case pGAB1:  kf = kd(2); kr = kd(3); Dq = Dd(4); src_num = ML + EG2PG1;  src_den = ML + EG2;    break;
      case GRB2:   kf = kd(4); kr = kd(5); Dq = Dd(1); src_num = ML + EG2;     src_den = ML + E;      break;
      case G2G1:   kf = kd(4); kr = kd(5); Dq = Dd(2); src_num = ML + EG2G1;   src_den = ML + E;      break;
      case G2PG1:  kf = kd(4); kr = kd(5); Dq = Dd(2); src_num = ML + EG2PG1;  src_den = ML + E;      break;
      case SHP2:   kf = kd(0); kr = kd(1); Dq = Dd(6); src_num = ML + EG2PG1S; src_den = ML + EG2PG1; break;
      case PG1S:   kf = kd(2); kr = kd(3); Dq = Dd(5); src_num = ML + EG2PG1S; src_den = ML + EG2;    break;
      case G2PG1S: kf = kd(4); kr = kd(5); Dq = Dd(3); src_num = ML + EG2PG1S; src_den = ML + E;      break;
      default: break;
    }
    const T drD = drdiv<NT>(a.o.dr, Dq);
    const bool is_flux = lane >= GAB1 && lane <= G2PG1S;
    T alpha = dconst<NT>(0.0), alpha2 = dconst<NT>(0.0), beta = dconst<NT>(0.0);
    switch (lane - ML) {
      case mE:     alpha = kd(12) * kd(14); beta = kd(13); break;
      case mES:    alpha2 = kd(15);         beta = kd(16); break;
      case mESmES: alpha = kd(10);          beta = kd(11); break;
      default: break;
    }
    auto put_l = [&](int c, const T& x) { lv_[c] = x.v;
#pragma unroll
      for (int n = 0; n < NT; ++n) lp_at(c, n) = x.p[n]; };
    put_l(L_cf, kf * drD);
    put_l(L_cr, kr * drD);
    put_l(L_kft, is_flux ? kf * dt : dconst<NT>(0.0));
    put_l(L_krt, is_flux ? kr * dt : dconst<NT>(0.0));
    put_l(L_alpha, alpha); put_l(L_alpha2, alpha2); put_l(L_beta, beta);
  }
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
  double s_own = 0.0, s_src = 0.0;
  int f_src = LZ;
  switch (lane - ML) {
    case mE:     s_own = -1.0; break;
    case mES:    s_own = -2.0; s_src = 1.0; f_src = ML + mE; break;
    case mESmES: s_own = -1.0; s_src = 1.0; f_src = ML + mES; break;
    case E:      s_src = 1.0; f_src = ML + mESmES; break;
    case NMB:    s_src = 2.0; f_src = ML + mESmES; break;
    default: break;
  }
  __syncwarp();
  const double tol = a.o.tol;
  const bool untracked = lane >= LE;
  const int iq_idx = lane < NCY ? lane : 10;
  const int maxiters = a.o.maxiters;
  auto hdr_ld = [&](int slot) { T r; r.v = lds(ws_s + 8 * slot);
#pragma unroll
    for (int n = 0; n < NT; ++n) r.p[n] = lds(ws_s + 8 * (TWS_HDR * (1 + n) + slot)); return r; };
  auto hdr_st = [&](int slot, const T& x) { sts(ws_s + 8 * slot, x.v);
#pragma unroll
    for (int n = 0; n < NT; ++n) sts(ws_s + 8 * (TWS_HDR * (1 + n) + slot), x.p[n]); };
  auto lane_const = [&](int c) { T r; r.v = lv_[c];
#pragma unroll
    for (int n = 0; n < NT; ++n) r.p[n] = lp_at(c, n); return r; };
  auto warp_const = [&](int c, double v) { T r; r.v = v;
#pragma unroll
    for (int n = 0; n < NT; ++n) r.p[n] = cp_at(c, n); return r; };

  T x = dconst<NT>(0.0);
  if (lane == ML + mE) { x.v = CoEGFRv;
#pragma unroll
    for (int n = 0; n < NT; ++n) x.p[n] = CoEGFRp[n]; }

  T t = dconst<NT>(0.0);
  double t_save = a.o.dt_save;
  int nts = 1;
  long long bc_total = 0;
  T pct_ave = dconst<NT>(0.0), pct_memb = dconst<NT>(0.0);

  auto gather_m = [&](int c, double (&m)[NMB]) {
#pragma unroll
    for (int j = 0; j < NMB; ++j) {
      double val = x.v;
#pragma unroll
      for (int n = 0; n < NT; ++n) if (c == n + 1) val = x.p[n];
      m[j] = shfl(val, ML + j);
    }
  };
  auto load_partials = [&](int n, double (&w)[NCY][K]) {
#pragma unroll
    for (int q = 0; q < NCY; ++q)
#pragma unroll
      for (int i = 0; i < K; ++i) w[q][i] = sp_at(n, q, i);
  };
  auto snapshot = [&](int col) {
    double mv[NMB];
    gather_m(0, mv);
    if (a.o.out_mode == GAB1_OUT_FULL) {
      if (lead) write_full_column<K>(a, oset, col, uv, mv, t.v, CoEGFRv, lane, g, rowA, status);
      const double Etot_v = 2.0 * (mv[E] + mv[EG2] + mv[EG2G1] + mv[EG2PG1] + mv[EG2PG1S]);
#pragma unroll 1
      for (int n = 0; n < NT; ++n) {
        double* ob = block_of(1 + n);
        if (!ob) continue;
        double mp[NMB], w[NCY][K];
        gather_m(1 + n, mp);
        load_partials(n, w);
        double tp = 0.0, cop = 0.0;
#pragma unroll
        for (int nn = 0; nn < NT; ++nn) if (nn == n) { tp = t.p[nn]; cop = CoEGFRp[nn]; }
        write_full_column<K>(a, ob, col, w, mp, tp, CoEGFRv, lane, g, rowA, scratch_status);
        if (lane == 0) {      // the two outputs that divide by CoEGFR (basepdesolver.jl:287; basepdesolver_rect.jl:264): quotient rule
          double* v = ob + (long long)__popc(a.o.matrix_mask & GAB1_MASK_ALL_MATRICES) * P * Cn;
          const double Etot_p = 2.0 * (mp[E] + mp[EG2] + mp[EG2G1] + mp[EG2PG1] + mp[EG2PG1S]);
          const double wq = cop / CoEGFRv;
          v[GAB1_V_pE * (long long)Cn + col] = (Etot_p * 100.0) / CoEGFRv - (Etot_v * 100.0 / CoEGFRv) * wq;
          v[GAB1_V_EGFR_SHP2 * (long long)Cn + col] = (mp[EG2PG1S] * 100.0) / CoEGFRv - (mv[EG2PG1S] * 100.0 / CoEGFRv) * wq;
        }
      }
    } else if (col == Cn - 1) {       // PCT_BOUND: trapezoid of (PG1S + G2PG1S) r^2 is linear in the profile
      stage_row<K>(rowA, lane, g, Nr, [&](int i) { return derived_stot<K>(uv, i); });
      pct_ave.v = trapz_r2(a.r, rowA, P);
      pct_memb.v = mv[EG2PG1S];
      __syncwarp();
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        double mp[NMB];
        gather_m(1 + n, mp);
        stage_row<K>(rowA, lane, g, Nr, [&](int i) { return sp_at(n, PG1S, i) + sp_at(n, G2PG1S, i); });
        pct_ave.p[n] = trapz_r2(a.r, rowA, P);
        pct_memb.p[n] = mp[EG2PG1S];
        __syncwarp();
      }
    }
  };

  for (long long step = 1; step <= Nt; ++step) {
    // ---- P0: primal Laplacians of the lane's nodes -> smem (every direction reads them: Dt' * L(u)) ----
    {
      double hl[NCY], hr[NCY];
#pragma unroll
      for (int q = 0; q < NCY; ++q) { hl[q] = shfl_up1(uv[q][K - 1]); hr[q] = shfl_down1(uv[q][0]); }
#pragma unroll
      for (int q = 0; q < NCY; ++q) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const double um = i > 0 ? uv[q][i - 1] : hl[q], upn = i + 1 < K ? uv[q][i + 1] : hr[q];
          lv_at(q, i) = fma(g.cp[i], upn, fma(g.cm[i], um, g.c0[i] * uv[q][i]));
        }
      }
    }
    // ---- T_n: partials of direction n, rolling window over the lane's nodes, against the OLD primal values ----
#pragma unroll 1
    for (int n = 0; n < NT; ++n) {
      double cp_[C_DG2G1S2 + 1];
#pragma unroll
      for (int c = 0; c <= C_DG2G1S2; ++c) cp_[c] = cp_at(c, n);
      double left[NCY], cur[NCY], hr[NCY];
#pragma unroll
      for (int q = 0; q < NCY; ++q) cur[q] = sp_at(n, q, 0);
#pragma unroll
      for (int q = 0; q < NCY; ++q) left[q] = K > 1 ? sp_at(n, q, K - 1) : cur[q];
#pragma unroll
      for (int q = 0; q < NCY; ++q) { left[q] = shfl_up1(left[q]); hr[q] = shfl_down1(cur[q]); }
#pragma unroll
      for (int i = 0; i < K; ++i) {
        // every load of this node first (they are in flight while the fluxes are formed), every store last
        double right[NCY], Lv[NCY];
#pragma unroll
        for (int q = 0; q < NCY; ++q) right[q] = i + 1 < K ? sp_at(n, q, i + 1) : hr[q];
#pragma unroll
        for (int q = 0; q < NCY; ++q) Lv[q] = lv_at(q, i);
        const double Sa = uv[aSFK][i], G1 = uv[GAB1][i], pG1 = uv[pGAB1][i], G2 = uv[GRB2][i], g2g1 = uv[G2G1][i],
                     g2pg1 = uv[G2PG1][i], S2 = uv[SHP2][i], pg1s = uv[PG1S][i], g2pg1s = uv[G2PG1S][i];
        const double gbv = cv[C_kG1f] * G2, phv = cv[C_kG1p] * Sa, sbv = cv[C_kS2f] * S2;
        const double gbp = fma(cp_[C_kG1f], G2, cv[C_kG1f] * cur[GRB2]);
        const double php = fma(cp_[C_kG1p], Sa, cv[C_kG1p] * cur[aSFK]);
        const double sbp = fma(cp_[C_kS2f], S2, cv[C_kS2f] * cur[SHP2]);
        // d(a*b - kr*c) = a'*b + a*b' - kr'*c - kr*c'
        auto flux = [&](double ap, double av, double bv, double bp, int ckr, double cvv, double cpp) {
          return fma(ap, bv, fma(av, bp, -fma(cp_[ckr], cvv, cv[ckr] * cpp)));
        };
        const double v1 = flux(gbp, gbv, G1, cur[GAB1], C_kG1r, g2g1, cur[G2G1]);
        const double v3 = flux(gbp, gbv, pG1, cur[pGAB1], C_kG1r, g2pg1, cur[G2PG1]);
        const double v5 = flux(gbp, gbv, pg1s, cur[PG1S], C_kG1r, g2pg1s, cur[G2PG1S]);
        const double v2 = flux(php, phv, G1, cur[GAB1], C_kG1dp, pG1, cur[pGAB1]);
        const double v6 = flux(php, phv, g2g1, cur[G2G1], C_kG1dp, g2pg1, cur[G2PG1]);
        const double v4 = flux(sbp, sbv, pG1, cur[pGAB1], C_kS2r, pg1s, cur[PG1S]);
        const double v7 = flux(sbp, sbv, g2pg1, cur[G2PG1], C_kS2r, g2pg1s, cur[G2PG1S]);
        const double sk = fma(cp_[C_kSi], Sa, cv[C_kSi] * cur[aSFK]);
        double kin[NCY];
        kin[iSFK] = cur[iSFK] + sk;
        kin[aSFK] = cur[aSFK] - sk;
        kin[GAB1] = cur[GAB1] - v1 - v2;
        kin[pGAB1] = cur[pGAB1] - v3 + v2 - v4;
        kin[GRB2] = cur[GRB2] - v1 - v3 - v5;
        kin[G2G1] = cur[G2G1] + v1 - v6;
        kin[G2PG1] = cur[G2PG1] + v3 + v6 - v7;
        kin[SHP2] = cur[SHP2] - v4 - v7;
        kin[PG1S] = cur[PG1S] + v4 - v5;
        kin[G2PG1S] = cur[G2PG1S] + v5 + v7;
        constexpr int cD[NCY] = {C_DSi, C_DSa, C_DG1, C_DG1, C_DG2, C_DG2G1, C_DG2G1, C_DS2, C_DG1S2, C_DG2G1S2};
#pragma unroll
        for (int q = 0; q < NCY; ++q) {
          const double Lp = fma(g.cp[i], right[q], fma(g.cm[i], left[q], g.c0[i] * cur[q]));
          kin[q] = fma(cp_[cD[q]], Lv[q], fma(cv[cD[q]], Lp, kin[q]));
          left[q] = cur[q];
          cur[q] = right[q];
        }
#pragma unroll
        for (int q = 0; q < NCY; ++q) {
          sp_at(n, q, i) = kin[q];
          if (i == idx_i && lane == lane_i) sts(ws_s + 8 * (TWS_HDR * (1 + n) + q), kin[q]);     // u'+[Nr-1] for the closure lanes
        }
      }
    }
    // ---- P1: primal update in place (solver_kernel.cuh fast path) ----
    {
      double v1[K], v2[K], v3[K], v4[K], v5[K], v6[K], v7[K], sk[K];
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const double Sa = uv[aSFK][i], G1 = uv[GAB1][i], pG1 = uv[pGAB1][i], G2 = uv[GRB2][i], g2g1 = uv[G2G1][i],
                     g2pg1 = uv[G2PG1][i], S2 = uv[SHP2][i], pg1s = uv[PG1S][i], g2pg1s = uv[G2PG1S][i];
        const double gb = cv[C_kG1f] * G2, ph = cv[C_kG1p] * Sa, sb = cv[C_kS2f] * S2;
        v1[i] = fma(gb, G1, -(cv[C_kG1r] * g2g1));
        v3[i] = fma(gb, pG1, -(cv[C_kG1r] * g2pg1));
        v5[i] = fma(gb, pg1s, -(cv[C_kG1r] * g2pg1s));
        v2[i] = fma(ph, G1, -(cv[C_kG1dp] * pG1));
        v6[i] = fma(ph, g2g1, -(cv[C_kG1dp] * g2pg1));
        v4[i] = fma(sb, pG1, -(cv[C_kS2r] * pg1s));
        v7[i] = fma(sb, g2pg1, -(cv[C_kS2r] * g2pg1s));
        sk[i] = cv[C_kSi] * Sa;
      }
#pragma unroll
      for (int i = 0; i < K; ++i) {
        double Lv[NCY];
#pragma unroll
        for (int q = 0; q < NCY; ++q) Lv[q] = lv_at(q, i);
        const double Si = uv[iSFK][i], Sa = uv[aSFK][i], G1 = uv[GAB1][i], pG1 = uv[pGAB1][i], G2 = uv[GRB2][i],
                     g2g1 = uv[G2G1][i], g2pg1 = uv[G2PG1][i], S2 = uv[SHP2][i], pg1s = uv[PG1S][i], g2pg1s = uv[G2PG1S][i];
        uv[iSFK][i] = fma(cv[C_DSi], Lv[iSFK], Si + sk[i]);
        uv[aSFK][i] = fma(cv[C_DSa], Lv[aSFK], Sa - sk[i]);
        uv[GAB1][i] = fma(cv[C_DG1], Lv[GAB1], G1 - v1[i] - v2[i]);
        uv[pGAB1][i] = fma(cv[C_DG1], Lv[pGAB1], pG1 - v3[i] + v2[i] - v4[i]);
        uv[GRB2][i] = fma(cv[C_DG2], Lv[GRB2], G2 - v1[i] - v3[i] - v5[i]);
        uv[G2G1][i] = fma(cv[C_DG2G1], Lv[G2G1], g2g1 + v1[i] - v6[i]);
        uv[G2PG1][i] = fma(cv[C_DG2G1], Lv[G2PG1], g2pg1 + v3[i] + v6[i] - v7[i]);
        uv[SHP2][i] = fma(cv[C_DS2], Lv[SHP2], S2 - v4[i] - v7[i]);
        uv[PG1S][i] = fma(cv[C_DG1S2], Lv[PG1S], pg1s + v4[i] - v5[i]);
        uv[G2PG1S][i] = fma(cv[C_DG2G1S2], Lv[G2PG1S], g2pg1s + v5[i] + v7[i]);
      }
    }
    if (lane == lane_i) {
#pragma unroll
      for (int q = 0; q < NCY; ++q) sts(ws_s + 8 * q, uv[q][idx_i]);
    }
    __syncwarp();

    // ---- M: membrane block on duals; the per-lane coefficients' partials come out of smem here ----
    int it = 0;
    {
      const T cf = lane_const(L_cf), cr_fixed = lane_const(L_cr), kf_t = lane_const(L_kft), kr_t = lane_const(L_krt),
              alpha = lane_const(L_alpha), alpha2 = lane_const(L_alpha2), beta = lane_const(L_beta);
      const T dtd = warp_const(C_dt, dtv);
      const T m_old = x;
      const T m_next = dshfl_down1<NT>(m_old);
      const T f = dfms<NT>(m_old, dfma<NT>(alpha2, m_old, alpha), beta * m_next);
      const T fsrc = dshfl<NT>(f, f_src);
      T dm;
      dm.v = fma(s_own, f.v, s_src * fsrc.v);
#pragma unroll
      for (int n = 0; n < NT; ++n) dm.p[n] = fma(s_own, f.p[n], s_src * fsrc.p[n]);
      const T base = dfma<NT>(dtd, dm, m_old);
      const T Md1 = dshfl<NT>(m_old, src_den), Mn1 = dshfl<NT>(m_old, src_num);
      const T A_t = kf_t * Md1;
      const T B_t = kr_t * Mn1;
      const T Iq = hdr_ld(iq_idx);
      // aSFK: I_a + ca*Etot*I_i/(1 + cf*Etot) = (I_a + (cf*I_a + ca*I_i)*Etot)/(1 + cf*Etot)   (basepdesolver.jl:853-854)
      T cr = cr_fixed;
      if (lane == aSFK) cr = dfma<NT>(cf, Iq, warp_const(C_ca, cav) * hdr_ld(iSFK));
      T Mn = Mn1, Md = Md1;
      for (;;) {
        ++it;
        const T num = dfma<NT>(cr, Mn, Iq);
        T den = cf * Md;
        den.v += 1.0;
        const double rden = fast_recip(den.v);
        T qv;
        qv.v = num.v * rden;
#pragma unroll
        for (int n = 0; n < NT; ++n) qv.p[n] = fma(-qv.v, den.p[n], num.p[n]) * rden;      // quotient rule
        const T F = dfms<NT>(A_t, qv, B_t);
        const T F0 = dshfl<NT>(F, fs0), F1 = dshfl<NT>(F, fs1), F2 = dshfl<NT>(F, fs2), F3 = dshfl<NT>(F, fs3);
        T mnew;
        mnew.v = fma(sg0, F0.v, sg1 * F1.v) + fma(sg2, F2.v, fma(sg3, F3.v, base.v));
#pragma unroll
        for (int n = 0; n < NT; ++n) mnew.p[n] = fma(sg0, F0.p[n], sg1 * F1.p[n]) + fma(sg2, F2.p[n], fma(sg3, F3.p[n], base.p[n]));
        const T xnew = lane < NCY ? qv : mnew;
        const bool ok = (fabs(x.v - xnew.v) < tol * fabs(x.v)) || untracked;      // values decide (solver_kernel.cuh finish_pass)
        x = xnew;
        if (__all_sync(FULL, ok)) break;
        if (it >= maxiters) break;
        Mn = dshfl<NT>(x, src_num);
        Md = dshfl<NT>(x, src_den);
      }
      t = t + dtd;                                                  // basepdesolver.jl:908
    }
    bc_total += it;
    // ---- boundary values back to the lane that owns node Nr ----
    if (lane < NCY) hdr_st(16 + lane, x);
    __syncwarp();
    if (lane == lane_b) {
#pragma unroll
      for (int q = 0; q < NCY; ++q) {
        const T b = hdr_ld(16 + q);
        uv[q][idx_b] = b.v;
#pragma unroll
        for (int n = 0; n < NT; ++n) sp_at(n, q, idx_b) = b.p[n];
      }
    }
    if (track_t && t.v >= t_save) {                                 // basepdesolver.jl:912
      if (nts >= Cn) status |= GAB1_ST_OVERFLOW;
      else snapshot(nts++);
      t_save = t_save + a.o.dt_save;
    }
  }

  // ---- final-time outputs ----
  if (a.o.out_mode == GAB1_OUT_FINAL4 || a.o.out_mode == GAB1_OUT_FINAL_STATE) {
    if (Nt == 0) {
#pragma unroll
      for (int q = 0; q < NCY; ++q)
#pragma unroll
        for (int i = 0; i < K; ++i) { uv[q][i] = 0.0;
#pragma unroll
          for (int n = 0; n < NT; ++n) sp_at(n, q, i) = 0.0; }
      x = dconst<NT>(0.0);
    }
    double m[NMB];
    gather_m(0, m);
    if (lead) write_final<K>(a, oset, uv, m, lane, g, rowA, rowB, status);
#pragma unroll 1
    for (int n = 0; n < NT; ++n) {
      double* ob = block_of(1 + n);
      if (!ob) continue;
      double w[NCY][K];
      gather_m(1 + n, m);
      load_partials(n, w);
      write_final<K>(a, ob, w, m, lane, g, rowA, rowB, scratch_status);
    }
  }
  if (a.o.out_mode == GAB1_OUT_PCT_BOUND) {      // param_fitting+inference_finitediff.jl:211-216
    const double R = a.o.R, R3 = R * R * R;
    const double ave_v = pct_ave.v * 3.0 / R3, mem_v = pct_memb.v * a.o.pct_mul / a.o.pct_div;
    const double pct_v = (ave_v + mem_v) / CoG1v * 100.0;
    if (isnan(pct_v)) status |= GAB1_ST_NAN;
    if (lead && lane == 0) oset[0] = pct_v;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      double* ob = block_of(1 + n);
      if (!ob) continue;
      const double ave_p = pct_ave.p[n] * 3.0 / R3, mem_p = pct_memb.p[n] * a.o.pct_mul / a.o.pct_div;
      if (lane == 0) ob[0] = ((ave_p + mem_p) / CoG1v - ((ave_v + mem_v) / CoG1v) * (CoG1p[n] / CoG1v)) * 100.0;
    }
  }
  if (track_t && nts < Cn) {
    status |= GAB1_ST_SHORT;
    if (a.o.out_mode == GAB1_OUT_FULL) {
      for (int c = 0; c <= NT; ++c) {
        double* ob = block_of(c);
        if (!ob) continue;
        long long off = 0;
        for (int mi = 0; mi < 12; ++mi) {
          if (!((a.o.matrix_mask >> mi) & 1u)) continue;
          for (long long i = (long long)nts * P + lane; i < (long long)Cn * P; i += 32) ob[off + i] = 0.0;
          off += (long long)P * Cn;
        }
        for (int v = 0; v < GAB1_N_VECTORS; ++v)
          for (int cc = nts + lane; cc < Cn; cc += 32) ob[off + (long long)v * Cn + cc] = 0.0;
      }
    }
  }
  if (lead && lane == 0) {
    if (a.status) a.status[set] = (int)status;
    if (a.n_saved) a.n_saved[set] = track_t ? nts : 0;
    if (a.n_steps) a.n_steps[set] = Nt;
    if (a.n_bc) a.n_bc[set] = bc_total;
  }
}

// Persistent kernel, ONE warp per CTA (the warp's smem slice is 25-62 KB: 3 to 8 warps fit an SM).
template <int K, int NT>
__global__ void __launch_bounds__(32)
tangent_stream_kernel(const TangentArgs ta) {
  extern __shared__ double smem[];
  const KernelArgs& a = ta.a;
  const int lane = threadIdx.x & 31;
  double* ws = smem;
  const int Nr = a.o.Nr;
  for (int i = lane; i < TSLayout<K, NT>::LP; i += 32) ws[i] = 0.0;       // exchange header and uniform constants
  __syncwarp();

  Grid<K> g;
  {
    const double dr = a.o.dr;
    const double inv_dr2 = 1.0 / (dr * dr);
    g.G = (Nr + K - 1) / K;
    const int off = Nr - g.G * K;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const int n = lane * K + i + 1 + off;
      g.node[i] = n;
      g.interior[i] = n >= 1 && n <= Nr - 1;
      const double r = (n >= 1 && n <= Nr) ? a.r[n] : 1.0;
      g.a[i] = 0.0;
      const double aj = (a.o.geometry == GAB1_GEOM_SPHERICAL) ? 1.0 / (r * dr) : 0.0;
      double cp = inv_dr2 + aj, cm = inv_dr2 - aj, c0 = -2.0 * inv_dr2;
      if (n == 1) { c0 += cm; cm = 0.0; }             // u[0] = u[1] (basepdesolver.jl:830-839)
      g.cp[i] = g.interior[i] ? cp : 0.0;
      g.cm[i] = g.interior[i] ? cm : 0.0;
      g.c0[i] = g.interior[i] ? c0 : 0.0;
    }
  }
  const long long items = a.S * ta.groups;
  for (;;) {
    unsigned item = 0;
    if (lane == 0) item = atomicAdd(a.counter, 1u);
    item = __shfl_sync(FULL, item, 0);
    if ((long long)item >= items) break;
    const long long si = item / ta.groups;
    const int group = (int)(item - si * ta.groups);
    const long long set = a.order ? (long long)a.order[si] : si;
    solve_set_tangent_stream<K, NT>(ta, set, group, lane, ws, g);
    __syncwarp();
  }
}

}  // namespace gab1
