// team_tangent_kernel.cuh — LATENCY path of forward mode: one CTA per (parameter set, direction), one thread per node.
//
// LBFGS and NUTS ask for one loss-and-gradient evaluation at a time (param_fitting+inference_finitediff.jl:238-270,
// 308-370), and the final optimisation stage runs at dr = 0.1: with one warp per (set, direction) that is 140 116
// steps of a lone warp that carries four nodes per lane and spills (572 ms per gradient).  Here the team layout of
// team_kernel.cuh carries a dual number per node: value and ONE partial of all ten species as a ping-pong pair in
// shared memory, thread t = node t, the last warp runs the lane-parallel fixed point on duals (tangent_kernel.cuh),
// one __syncthreads() per step.  The four directions of a gradient are four CTAs.
// Measured: one gradient (4 partials) 127 ms at dr = 0.1, 33 ms at dr = 0.2 (one warp per pair: 577 and 60 ms).
// Same restrictions as gab1_solve_tangent; arithmetic forms as tangent_kernel.cuh.
#pragma once
#include "tangent_stream_kernel.cuh"     // the coefficient enums C_*, L_*

namespace gab1 {

template <int NT>
__global__ void __launch_bounds__(256)
team_tangent_kernel(const TangentArgs ta) {
  typedef Dn<NT> D1;
  constexpr int NC = 1 + NT;                           // components: value, NT partials
  const KernelArgs& a = ta.a;
  extern __shared__ double smem[];
  __shared__ unsigned s_item;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = blockDim.x, W = T >> 5;
  const bool mwarp = warp == W - 1;
  const int Nr = a.o.Nr, P = Nr + 1, Nts = a.o.Nts, Cn = Nts + 1;
  const int off = Nr - (T - 1);
  const int node = tid + off;
  double* U = smem;                                    // [2 buffers][NC components][NCY][T]
  double* hdr = smem + 2 * NC * NCY * T;               // [16 c, 16 c + 8): membrane values of component c
  double* row = hdr + 16 * NC;                         // P_pad doubles
  // partials of the coefficients: C_N uniform ones (x NT), L_N + 2 per-lane ones of the membrane warp (x NT x 32) — in
  // registers they spill as soon as a CTA carries more than one direction
  const unsigned cps_s = (unsigned)__cvta_generic_to_shared(row + a.P_pad);
  const unsigned lps_s = cps_s + 8u * (unsigned)(((C_N * NT + 3) & ~3));
  auto ldc = [](unsigned addr) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v; };
  auto u_at = [&](int buf, int c, int q, int slot) -> double& { return U[((buf * NC + c) * NCY + q) * T + slot]; };

  const bool interior = node >= 1 && node <= Nr - 1;
  double cpc = 0.0, cmc = 0.0, c0c = 0.0;
  {
    const double dr = a.o.dr, inv_dr2 = 1.0 / (dr * dr);
    const double r = (node >= 1 && node <= Nr) ? a.r[node] : 1.0;
    const double aj = (a.o.geometry == GAB1_GEOM_SPHERICAL) ? 1.0 / (r * dr) : 0.0;
    double cp = inv_dr2 + aj, cm = inv_dr2 - aj, c0 = -2.0 * inv_dr2;
    if (node == 1) { c0 += cm; cm = 0.0; }
    if (interior) { cpc = cp; cmc = cm; c0c = c0; }
  }
  const int sl = tid > 0 ? tid - 1 : 0, sr = tid < T - 1 ? tid + 1 : T - 1;
  const long long items = a.S * (long long)ta.groups;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_item = atomicAdd(a.counter, 1u);
    __syncthreads();
    const unsigned item = s_item;
    if ((long long)item >= items) break;
    const long long si = item / ta.groups;
    const int group = (int)(item - si * ta.groups);
    const long long set = a.order ? (long long)a.order[si] : si;
    const bool lead = group == 0;                        // owns the value block and the diagnostics

    const long long nout = a.out_stride;
    double* oset = a.out + set * nout * (1 + ta.n_dir);
    double* ob[NC];
    const double* sd[NT];
    ob[0] = lead ? oset : nullptr;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const int d = group * NT + n;
      ob[1 + n] = d < ta.n_dir ? oset + (long long)(1 + d) * nout : nullptr;
      sd[n] = d < ta.n_dir ? ta.seeds + (set * ta.n_dir + d) * GAB1_N_SEED : nullptr;
    }
    auto seed = [&](int n, int i) -> double { return sd[n] ? sd[n][i] : 0.0; };
    unsigned status = 0;
    const double* Cov = a.Co + set * a.Co_stride;
    const double* Dv = a.D + set * GAB1_N_D;
    const double* kv = a.k + set * GAB1_N_K;
    auto Dd = [&](int i) { D1 r; r.v = Dv[i];
#pragma unroll
      for (int n = 0; n < NT; ++n) r.p[n] = seed(n, i); return r; };
    auto kd = [&](int i) { D1 r; r.v = kv[i];
#pragma unroll
      for (int n = 0; n < NT; ++n) r.p[n] = seed(n, GAB1_N_D + i); return r; };
    auto Cod = [&](int i) { D1 r; r.v = Cov[i];
#pragma unroll
      for (int n = 0; n < NT; ++n) r.p[n] = seed(n, GAB1_N_D + GAB1_N_K + i); return r; };
    D1 dt; dt.v = a.dt[set];
#pragma unroll
    for (int n = 0; n < NT; ++n) dt.p[n] = seed(n, GAB1_N_SEED - 1);
    auto comp_of = [&](const D1& x, int c) { double v = x.v;
#pragma unroll
      for (int n = 0; n < NT; ++n) if (c == n + 1) v = x.p[n]; return v; };
    const D1 CoSFK = Cod(0), CoG2 = Cod(1), CoG1 = Cod(2), CoS2 = Cod(3), CoEGFR = Cod(4);
    D1 D_Si = Dd(0), D_Sa = Dd(0);
    if (a.o.sfk_mode == GAB1_SFK_MEMBRANE) D_Sa = dconst<NT>(1e-32);
    if (a.o.sfk_mode == GAB1_SFK_BOTH_FROZEN) { D_Si = dconst<NT>(1e-32); D_Sa = dconst<NT>(1e-32); }
    const bool track_t = (a.o.out_mode == GAB1_OUT_FULL || a.o.out_mode == GAB1_OUT_PCT_BOUND);

    const double nt_f = ceil(__ddiv_rn(a.o.tf, dt.v));
    if (!(nt_f >= 0.0 && nt_f < 9.0e18)) {
      for (int c = 0; c < NC; ++c)
        if (ob[c]) for (long long i = tid; i < nout; i += T) ob[c][i] = 0.0;
      if (lead && tid == 0) {
        if (a.status) a.status[set] = GAB1_ST_THROW;
        if (a.n_saved) a.n_saved[set] = 0;
        if (a.n_steps) a.n_steps[set] = 0;
        if (a.n_bc) a.n_bc[set] = 0;
      }
      continue;
    }
    const long long Nt = (long long)nt_f;

    {
      const bool on = node >= 1 && node <= Nr;
#pragma unroll
      for (int q = 0; q < NCY; ++q)
#pragma unroll
        for (int c = 0; c < NC; ++c) { u_at(0, c, q, tid) = 0.0; u_at(1, c, q, tid) = 0.0; }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        u_at(0, c, iSFK, tid) = on ? comp_of(CoSFK, c) : 0.0;
        u_at(0, c, GAB1, tid) = on ? comp_of(CoG1, c) : 0.0;
        u_at(0, c, GRB2, tid) = on ? comp_of(CoG2, c) : 0.0;
        u_at(0, c, SHP2, tid) = on ? comp_of(CoS2, c) : 0.0;
      }
      if (tid < 16 * NC) hdr[tid] = 0.0;
    }
    __syncthreads();                                     // the initial state is visible to the neighbours
    int cur = 0;
    auto at_node = [&](int b, int c, int q, int n) -> double { return u_at(b, c, q, (n < 1 ? 1 : n) - off); };
    auto stot_at = [&](int b, int c, int n) { return __dadd_rn(at_node(b, c, PG1S, n), at_node(b, c, G2PG1S, n)); };
    auto ptot_at = [&](int b, int c, int n) {
      const double g2pg1 = at_node(b, c, G2PG1, n), pg1 = at_node(b, c, pGAB1, n), pg1s = at_node(b, c, PG1S, n), g2pg1s = at_node(b, c, G2PG1S, n);
      if (a.o.pg1tot_form == GAB1_PG1TOT_VIA_STOT) return __dadd_rn(__dadd_rn(g2pg1, pg1), __dadd_rn(pg1s, g2pg1s));
      return __dadd_rn(__dadd_rn(__dadd_rn(g2pg1, pg1), pg1s), g2pg1s);
    };

    if (a.o.out_mode == GAB1_OUT_FULL) {
      for (int c = 0; c < NC; ++c) {
        if (!ob[c]) continue;
        auto comp = [&](const D1& x) { return comp_of(x, c); };
        long long o2 = 0;
        for (int mi = 0; mi < 12; ++mi) {
          if (!((a.o.matrix_mask >> mi) & 1u)) continue;
          const double v0 = mi == GAB1_M_iSFK ? comp(CoSFK) : mi == GAB1_M_GRB2 ? comp(CoG2) : mi == GAB1_M_SHP2 ? comp(CoS2)
                            : mi == GAB1_M_GAB1 ? comp(CoG1) : 0.0;
          for (int n = tid; n < P; n += T) ob[c][o2 + n] = v0;
          o2 += (long long)P * Cn;
        }
        if (tid < GAB1_N_VECTORS) ob[c][o2 + (long long)tid * Cn] = tid == GAB1_V_mE ? comp(CoEGFR) : 0.0;
      }
    }

    D1 t = dconst<NT>(0.0);
    double t_save = a.o.dt_save;
    int nts = 1;
    long long bc_total = 0;
    D1 pct_ave = dconst<NT>(0.0), pct_memb = dconst<NT>(0.0);

    double cv[C_N];                 // values of the uniform coefficients; their partials: cps_s
    double lv[L_N];                 // values of this lane's membrane coefficients (last warp); partials: lps_s
    constexpr int LZ = 31, LE = ML + NMB;
    int src_num = LZ, src_den = LZ;
    {
      auto put_c = [&](int c, const D1& v) { cv[c] = v.v;
        if (tid == 0) {
#pragma unroll
          for (int n = 0; n < NT; ++n) sts(cps_s + 8u * (unsigned)(c * NT + n), v.p[n]);
        } };
      put_c(C_kS2f, kd(0) * dt); put_c(C_kS2r, kd(1) * dt); put_c(C_kG1f, kd(2) * dt); put_c(C_kG1r, kd(3) * dt);
      put_c(C_kG1p, kd(6) * dt); put_c(C_kG1dp, kd(7) * dt); put_c(C_kSi, kd(9) * dt);
      put_c(C_DSi, D_Si * dt); put_c(C_DSa, D_Sa * dt); put_c(C_DG1, Dd(4) * dt); put_c(C_DG2, Dd(1) * dt);
      put_c(C_DG2G1, Dd(2) * dt); put_c(C_DS2, Dd(6) * dt); put_c(C_DG1S2, Dd(5) * dt); put_c(C_DG2G1S2, Dd(3) * dt);
      put_c(C_dt, dt);
      put_c(C_ca, kd(8) * drdiv<NT>(a.o.dr, D_Sa));
      D1 kf = dconst<NT>(0.0), kr = dconst<NT>(0.0), Dq = dconst<NT>(1.0);
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
      const D1 drD = drdiv<NT>(a.o.dr, Dq);
      const bool is_flux = lane >= GAB1 && lane <= G2PG1S;
      D1 alpha = dconst<NT>(0.0), alpha2 = dconst<NT>(0.0), beta = dconst<NT>(0.0);
      switch (lane - ML) {
        case mE:     alpha = kd(12) * kd(14); beta = kd(13); break;
        case mES:    alpha2 = kd(15);         beta = kd(16); break;
        case mESmES: alpha = kd(10);          beta = kd(11); break;
        default: break;
      }
      auto put_l = [&](int c, const D1& v) { lv[c] = v.v;
        if (mwarp) {
#pragma unroll
          for (int n = 0; n < NT; ++n) sts(lps_s + 8u * (unsigned)((c * NT + n) * 32 + lane), v.p[n]);
        } };
      put_l(L_cf, kf * drD);
      put_l(L_cr, kr * drD);
      put_l(L_kft, is_flux ? kf * dt : dconst<NT>(0.0));
      put_l(L_krt, is_flux ? kr * dt : dconst<NT>(0.0));
      put_l(L_alpha, alpha); put_l(L_alpha2, alpha2); put_l(L_beta, beta);
    }
    auto KC = [&](int c) { D1 r; r.v = cv[c];
#pragma unroll
      for (int n = 0; n < NT; ++n) r.p[n] = ldc(cps_s + 8u * (unsigned)(c * NT + n)); return r; };
    auto LC = [&](int c) { D1 r; r.v = lv[c];
#pragma unroll
      for (int n = 0; n < NT; ++n) r.p[n] = ldc(lps_s + 8u * (unsigned)((c * NT + n) * 32 + lane)); return r; };
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
    __syncthreads();                                     // coefficient partials visible
    const double tol = a.o.tol;
    const bool untracked = lane >= LE;
    const int maxiters = a.o.maxiters;
    D1 x = (lane == ML + mE) ? CoEGFR : dconst<NT>(0.0);

    auto publish_m = [&]() {
      if (mwarp && lane >= ML && lane < LE) {
#pragma unroll
        for (int c = 0; c < NC; ++c) hdr[16 * c + lane - ML] = comp_of(x, c);
      }
    };
    // one snapshot column, both components; hdr holds the membrane values (after publish_m + __syncthreads)
    auto write_column = [&](int col, int b) {
      constexpr int kSpecies[10] = {iSFK, aSFK, GRB2, GAB1, SHP2, G2G1, G2PG1, G2PG1S, pGAB1, PG1S};
      for (int c = 0; c < NC; ++c) {
        if (!ob[c]) continue;
        long long o2 = 0;
        for (int mi = 0; mi < 12; ++mi) {
          if (!((a.o.matrix_mask >> mi) & 1u)) continue;
          for (int n = tid; n < P; n += T)
            ob[c][o2 + (long long)col * P + n] = mi < 10 ? at_node(b, c, kSpecies[mi], n) : (mi == GAB1_M_PG1tot ? ptot_at(b, c, n) : stot_at(b, c, n));
          o2 += (long long)P * Cn;
        }
        if (tid == 0) {
          const double* m = hdr + 16 * c;
          double* v = ob[c] + o2;
          const double Etot = 2.0 * (m[E] + m[EG2] + m[EG2G1] + m[EG2PG1] + m[EG2PG1S]);
          if (c == 0) {
            v[GAB1_V_pE * (long long)Cn + col] = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(m[E], m[EG2]), m[EG2G1]), m[EG2PG1]), m[EG2PG1S])), 100.0), CoEGFR.v);
            v[GAB1_V_EGFR_SHP2 * (long long)Cn + col] = __ddiv_rn(__dmul_rn(m[EG2PG1S], 100.0), CoEGFR.v);
            v[GAB1_V_t_out * (long long)Cn + col] = t.v;
          } else {                                             // quotient rule for the two outputs that divide by CoEGFR
            const double* mv = hdr;
            const double Etot_v = 2.0 * (mv[E] + mv[EG2] + mv[EG2G1] + mv[EG2PG1] + mv[EG2PG1S]);
            const double w = comp_of(CoEGFR, c) / CoEGFR.v;
            v[GAB1_V_pE * (long long)Cn + col] = (Etot * 100.0) / CoEGFR.v - (Etot_v * 100.0 / CoEGFR.v) * w;
            v[GAB1_V_EGFR_SHP2 * (long long)Cn + col] = (m[EG2PG1S] * 100.0) / CoEGFR.v - (mv[EG2PG1S] * 100.0 / CoEGFR.v) * w;
            v[GAB1_V_t_out * (long long)Cn + col] = comp_of(t, c);
          }
          v[GAB1_V_mE * (long long)Cn + col] = m[mE];
          v[GAB1_V_mES * (long long)Cn + col] = m[mES];
          v[GAB1_V_mESmES * (long long)Cn + col] = m[mESmES];
          v[GAB1_V_E * (long long)Cn + col] = m[E];
          v[GAB1_V_EG2 * (long long)Cn + col] = m[EG2];
          v[GAB1_V_EG2G1 * (long long)Cn + col] = m[EG2G1];
          v[GAB1_V_EG2PG1 * (long long)Cn + col] = m[EG2PG1];
          v[GAB1_V_EG2PG1S * (long long)Cn + col] = m[EG2PG1S];
        }
      }
      bool ns = false;
      for (int n = tid; n < P; n += T) ns |= isnan(at_node(b, 0, PG1S, n));
      if (ns) status |= GAB1_ST_NAN;
    };

    for (long long step = 1; step <= Nt; ++step) {
      const int nxt = cur ^ 1;
      // ---- interior on duals: thread = node, old buffer -> new buffer ----
      {
        auto ldq = [&](int q, int slot) { D1 r; r.v = u_at(cur, 0, q, slot);
#pragma unroll
          for (int n = 0; n < NT; ++n) r.p[n] = u_at(cur, 1 + n, q, slot); return r; };
        const D1 Si = ldq(iSFK, tid), Sa = ldq(aSFK, tid), G1 = ldq(GAB1, tid), pG1 = ldq(pGAB1, tid), G2 = ldq(GRB2, tid),
                 g2g1 = ldq(G2G1, tid), g2pg1 = ldq(G2PG1, tid), S2 = ldq(SHP2, tid), pg1s = ldq(PG1S, tid), g2pg1s = ldq(G2PG1S, tid);
        auto lap = [&](int q, const D1& uc) {
          const D1 up = ldq(q, sr), um = ldq(q, sl);
          D1 r;
          r.v = fma(cpc, up.v, fma(cmc, um.v, c0c * uc.v));
#pragma unroll
          for (int n = 0; n < NT; ++n) r.p[n] = fma(cpc, up.p[n], fma(cmc, um.p[n], c0c * uc.p[n]));
          return r;
        };
        auto stq = [&](int q, const D1& val) { u_at(nxt, 0, q, tid) = val.v;
#pragma unroll
          for (int n = 0; n < NT; ++n) u_at(nxt, 1 + n, q, tid) = val.p[n]; };
        const D1 kg1r = KC(C_kG1r), kg1dp = KC(C_kG1dp), ks2r = KC(C_kS2r);
        const D1 gb = KC(C_kG1f) * G2, ph = KC(C_kG1p) * Sa, sb = KC(C_kS2f) * S2;
        const D1 v1 = dfms<NT>(gb, G1, kg1r * g2g1);
        const D1 v3 = dfms<NT>(gb, pG1, kg1r * g2pg1);
        const D1 v5 = dfms<NT>(gb, pg1s, kg1r * g2pg1s);
        const D1 v2 = dfms<NT>(ph, G1, kg1dp * pG1);
        const D1 v6 = dfms<NT>(ph, g2g1, kg1dp * g2pg1);
        const D1 v4 = dfms<NT>(sb, pG1, ks2r * pg1s);
        const D1 v7 = dfms<NT>(sb, g2pg1, ks2r * g2pg1s);
        const D1 sk = KC(C_kSi) * Sa;
        stq(iSFK, dfma<NT>(KC(C_DSi), lap(iSFK, Si), Si + sk));
        stq(aSFK, dfma<NT>(KC(C_DSa), lap(aSFK, Sa), Sa - sk));
        stq(GAB1, dfma<NT>(KC(C_DG1), lap(GAB1, G1), G1 - v1 - v2));
        stq(pGAB1, dfma<NT>(KC(C_DG1), lap(pGAB1, pG1), pG1 - v3 + v2 - v4));
        stq(GRB2, dfma<NT>(KC(C_DG2), lap(GRB2, G2), G2 - v1 - v3 - v5));
        stq(G2G1, dfma<NT>(KC(C_DG2G1), lap(G2G1, g2g1), g2g1 + v1 - v6));
        stq(G2PG1, dfma<NT>(KC(C_DG2G1), lap(G2PG1, g2pg1), g2pg1 + v3 + v6 - v7));
        stq(SHP2, dfma<NT>(KC(C_DS2), lap(SHP2, S2), S2 - v4 - v7));
        stq(PG1S, dfma<NT>(KC(C_DG1S2), lap(PG1S, pg1s), pg1s + v4 - v5));
        stq(G2PG1S, dfma<NT>(KC(C_DG2G1S2), lap(G2PG1S, g2pg1s), g2pg1s + v5 + v7));
      }
      if (mwarp) {
        // ---- membrane fixed point on duals (tangent_kernel.cuh); exit decided by the values ----
        const D1 cf = LC(L_cf), cr_fixed = LC(L_cr), kf_t = LC(L_kft), kr_t = LC(L_krt), alpha = LC(L_alpha), alpha2 = LC(L_alpha2),
                 beta = LC(L_beta), dtd = KC(C_dt);
        const D1 m_old = x;
        const D1 m_next = dshfl_down1<NT>(m_old);
        const D1 f = dfms<NT>(m_old, dfma<NT>(alpha2, m_old, alpha), beta * m_next);
        const D1 fsrc = dshfl<NT>(f, f_src);
        D1 dm;
        dm.v = fma(s_own, f.v, s_src * fsrc.v);
#pragma unroll
        for (int n = 0; n < NT; ++n) dm.p[n] = fma(s_own, f.p[n], s_src * fsrc.p[n]);
        const D1 base = dfma<NT>(dtd, dm, m_old);
        const D1 Md1 = dshfl<NT>(m_old, src_den), Mn1 = dshfl<NT>(m_old, src_num);
        const D1 A_t = kf_t * Md1;
        const D1 B_t = kr_t * Mn1;
        __syncwarp();
        D1 Iq = dconst<NT>(0.0), Ii;
        if (lane < NCY) { Iq.v = u_at(nxt, 0, lane, T - 2);
#pragma unroll
          for (int n = 0; n < NT; ++n) Iq.p[n] = u_at(nxt, 1 + n, lane, T - 2); }
        Ii.v = u_at(nxt, 0, iSFK, T - 2);
#pragma unroll
        for (int n = 0; n < NT; ++n) Ii.p[n] = u_at(nxt, 1 + n, iSFK, T - 2);
        D1 cr = cr_fixed;
        if (lane == aSFK) cr = dfma<NT>(cf, Iq, KC(C_ca) * Ii);
        int it = 0;
        D1 Mn = Mn1, Md = Md1;
        for (;;) {
          ++it;
          const D1 num = dfma<NT>(cr, Mn, Iq);
          D1 den = cf * Md;
          den.v += 1.0;
          const double rden = fast_recip(den.v);
          D1 qv;
          qv.v = num.v * rden;
#pragma unroll
          for (int n = 0; n < NT; ++n) qv.p[n] = fma(-qv.v, den.p[n], num.p[n]) * rden;
          const D1 F = dfms<NT>(A_t, qv, B_t);
          const D1 F0 = dshfl<NT>(F, fs0), F1 = dshfl<NT>(F, fs1), F2 = dshfl<NT>(F, fs2), F3 = dshfl<NT>(F, fs3);
          D1 mnew;
          mnew.v = fma(sg0, F0.v, sg1 * F1.v) + fma(sg2, F2.v, fma(sg3, F3.v, base.v));
#pragma unroll
          for (int n = 0; n < NT; ++n) mnew.p[n] = fma(sg0, F0.p[n], sg1 * F1.p[n]) + fma(sg2, F2.p[n], fma(sg3, F3.p[n], base.p[n]));
          const D1 xnew = lane < NCY ? qv : mnew;
          const bool ok = (fabs(x.v - xnew.v) < tol * fabs(x.v)) || untracked;
          x = xnew;
          if (__all_sync(FULL, ok)) break;
          if (it >= maxiters) break;
          Mn = dshfl<NT>(x, src_num);
          Md = dshfl<NT>(x, src_den);
        }
        bc_total += it;
        if (lane < NCY) { u_at(nxt, 0, lane, T - 1) = x.v;
#pragma unroll
          for (int n = 0; n < NT; ++n) u_at(nxt, 1 + n, lane, T - 1) = x.p[n]; }
      }
      __syncthreads();
      cur = nxt;
      t = t + KC(C_dt);
      if (track_t && t.v >= t_save) {
        if (nts >= Cn) status |= GAB1_ST_OVERFLOW;
        else {
          const int col = nts++;
          publish_m();
          __syncthreads();
          if (a.o.out_mode == GAB1_OUT_FULL) write_column(col, cur);
          else if (col == Cn - 1) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              __syncthreads();
              for (int n = tid; n < P; n += T) row[n] = stot_at(cur, c, n);
              __syncthreads();
              const double tr = trapz_r2(a.r, row, P);
              if (c == 0) { pct_ave.v = tr; pct_memb.v = hdr[EG2PG1S]; } else { pct_ave.p[c > 0 ? c - 1 : 0] = tr; pct_memb.p[c > 0 ? c - 1 : 0] = hdr[16 * c + EG2PG1S]; }
            }
          }
          __syncthreads();
        }
        t_save = t_save + a.o.dt_save;
      }
    }
    publish_m();
    __syncthreads();
    const int fin = Nt == 0 ? 1 : cur;
    bool ns = false;
    if (a.o.out_mode == GAB1_OUT_FINAL4) {
      for (int c = 0; c < NC; ++c) {
        if (!ob[c]) continue;
        for (int n = tid; n < P; n += T) {
          const double v0 = at_node(fin, c, iSFK, n), v1 = at_node(fin, c, aSFK, n), v2 = ptot_at(fin, c, n), v3 = stot_at(fin, c, n);
          ob[c][n] = v0; ob[c][P + n] = v1; ob[c][2 * P + n] = v2; ob[c][3 * P + n] = v3;
          if (c == 0) ns |= isnan(v0) || isnan(v1) || isnan(v2) || isnan(v3);
        }
      }
    } else if (a.o.out_mode == GAB1_OUT_FINAL_STATE) {
      for (int c = 0; c < NC; ++c) {
        if (!ob[c]) continue;
        for (int q = 0; q < NCY; ++q)
          for (int n = tid; n < P; n += T) { const double v = at_node(fin, c, q, n); ob[c][(long long)q * P + n] = v; if (c == 0) ns |= isnan(v); }
        if (tid == 0)
          for (int j = 0; j < NMB; ++j) { const double v = Nt == 0 ? 0.0 : hdr[16 * c + j]; ob[c][(long long)NCY * P + j] = v; if (c == 0) ns |= isnan(v); }
      }
    } else if (a.o.out_mode == GAB1_OUT_PCT_BOUND) {      // param_fitting+inference_finitediff.jl:211-216
      const double R = a.o.R, R3 = R * R * R;
      const double ave_v = pct_ave.v * 3.0 / R3, mem_v = pct_memb.v * a.o.pct_mul / a.o.pct_div;
      const double pct_v = (ave_v + mem_v) / CoG1.v * 100.0;
      ns |= isnan(pct_v);
      if (lead && tid == 0) oset[0] = pct_v;
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        if (!ob[1 + n]) continue;
        const double ave_p = pct_ave.p[n] * 3.0 / R3, mem_p = pct_memb.p[n] * a.o.pct_mul / a.o.pct_div;
        if (tid == 0) ob[1 + n][0] = ((ave_p + mem_p) / CoG1.v - ((ave_v + mem_v) / CoG1.v) * (CoG1.p[n] / CoG1.v)) * 100.0;
      }
    }
    if (ns) status |= GAB1_ST_NAN;
    if (track_t && nts < Cn) {
      status |= GAB1_ST_SHORT;
      if (a.o.out_mode == GAB1_OUT_FULL) {
        for (int c = 0; c < NC; ++c) {
          if (!ob[c]) continue;
          long long o2 = 0;
          for (int mi = 0; mi < 12; ++mi) {
            if (!((a.o.matrix_mask >> mi) & 1u)) continue;
            for (long long i = (long long)nts * P + tid; i < (long long)Cn * P; i += T) ob[c][o2 + i] = 0.0;
            o2 += (long long)P * Cn;
          }
          for (int v = 0; v < GAB1_N_VECTORS; ++v)
            for (int cc = nts + tid; cc < Cn; cc += T) ob[c][o2 + (long long)v * Cn + cc] = 0.0;
        }
      }
    }
    int st_all = 0;
#pragma unroll
    for (unsigned b = 1u; b <= GAB1_ST_THROW; b <<= 1)
      if (__syncthreads_or((int)(status & b))) st_all |= (int)b;
    if (lead && mwarp && lane == 0) {
      if (a.status) a.status[set] = st_all;
      if (a.n_saved) a.n_saved[set] = track_t ? nts : 0;
      if (a.n_steps) a.n_steps[set] = Nt;
      if (a.n_bc) a.n_bc[set] = bc_total;
    }
  }
}

}  // namespace gab1
