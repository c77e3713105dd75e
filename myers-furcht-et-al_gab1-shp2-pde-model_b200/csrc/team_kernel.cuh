// team_kernel.cuh — LATENCY path: one CTA ("team" of W warps) per parameter set, one thread per radial node.
//
// Why: a single pdesolver call (run_base_model.jl:83 — BASELINE configs[0]) or a handful of sets is one warp per set in
// the throughput kernels: 140 116 steps of ~1.7 k cycles each at dr = 0.1 while 147 SMs idle.  Here the grid of ONE set
// is spread over ceil(Nr/32) warps, the state lives in shared memory as a ping-pong pair u[2][species][node] (the
// reference's own column-1 / column-2 scheme, basepdesolver.jl:115-133), and a step is: every thread updates its node
// from the old buffer; the LAST warp, which owns nodes Nr-31..Nr, then runs the lane-parallel membrane fixed point on
// its own inner-neighbour value and writes the boundary node; ONE __syncthreads(); swap.  Snapshots are already
// contiguous rows in shared memory.  Arithmetic forms, fixed point, event countdown: the fast path of
// solver_kernel.cuh, so the family is held to the same parity tests (GAB1_KERNEL=team).
//
// Used by the dispatcher when the batch is too small to fill the GPU with one warp per set (gab1pde.cu); 32 < Nr <= 256
// (teams of 2 to 8 warps).
#pragma once
#include "solver_kernel.cuh"

namespace gab1 {

template <int MODE>
__global__ void __launch_bounds__(256)
team_kernel(const KernelArgs a) {
  constexpr bool WHILE = MODE == MODE_FAST_WHILE;
  extern __shared__ double smem[];
  __shared__ unsigned s_item;
  __shared__ int s_dead[2];      // indexed by the parity of the step: a warp that runs ahead raises the OTHER slot
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = blockDim.x, W = T >> 5;
  const bool mwarp = warp == W - 1;                    // the warp that owns nodes Nr-31..Nr and runs the membrane block
  const int Nr = a.o.Nr, P = Nr + 1, Nts = a.o.Nts, Cn = Nts + 1;
  const int off = Nr - (T - 1);                        // node of thread `tid` = tid + off (<= 0 for padding threads)
  const int node = tid + off;
  double* U = smem;                                    // [2][NCY][T]
  double* hdr = smem + 2 * NCY * T;                    // 32 doubles: the membrane warp's exchange header
  double* row = hdr + WS_HDR;                          // P_pad doubles: derived rows for the final reductions
  auto u_at = [&](int buf, int q, int slot) -> double& { return U[(buf * NCY + q) * T + slot]; };

  // grid coefficients of this thread's node (legacy fast form: lap = cp*u[j+1] + cm*u[j-1] + c0*u[j])
  const bool interior = node >= 1 && node <= Nr - 1;
  double cpc = 0.0, cmc = 0.0, c0c = 0.0;
  {
    const double dr = a.o.dr, inv_dr2 = 1.0 / (dr * dr);
    const double r = (node >= 1 && node <= Nr) ? a.r[node] : 1.0;
    const double aj = (a.o.geometry == GAB1_GEOM_SPHERICAL) ? 1.0 / (r * dr) : 0.0;
    double cp = inv_dr2 + aj, cm = inv_dr2 - aj, c0 = -2.0 * inv_dr2;
    if (node == 1) { c0 += cm; cm = 0.0; }             // u[0] = u[1] (basepdesolver.jl:183-192)
    if (interior) { cpc = cp; cmc = cm; c0c = c0; }
  }
  const int sl = tid > 0 ? tid - 1 : 0, sr = tid < T - 1 ? tid + 1 : T - 1;
  if (tid < WS_HDR) hdr[tid] = 0.0;

  for (;;) {
    __syncthreads();
    if (tid == 0) { s_item = atomicAdd(a.counter, 1u); s_dead[0] = 0; s_dead[1] = 0; }
    __syncthreads();
    const unsigned item = s_item;
    if ((long long)item >= a.S) break;
    const long long set = a.order ? (long long)a.order[item] : (long long)item;

    double* oset = a.out + set * a.out_stride;
    unsigned status = 0;
    const double* Co = a.Co + set * a.Co_stride;
    const double* Dv = a.D + set * GAB1_N_D;
    const double* kv = a.k + set * GAB1_N_K;
    const double dt = a.dt[set];
    const double CoSFK = Co[0], CoG2 = Co[1], CoG1 = Co[2], CoS2 = Co[3], CoEGFR = Co[4];
    double D_Si = Dv[0], D_Sa = Dv[0];
    if (a.o.sfk_mode == GAB1_SFK_MEMBRANE) D_Sa = 1e-32;
    if (a.o.sfk_mode == GAB1_SFK_BOTH_FROZEN) { D_Si = 1e-32; D_Sa = 1e-32; }
    const bool track_t = (a.o.out_mode == GAB1_OUT_FULL || a.o.out_mode == GAB1_OUT_PCT_BOUND);
    const long long nout = a.out_stride;

    const double nt_f = ceil(__ddiv_rn(a.o.tf, dt));
    if (!(nt_f >= 0.0 && nt_f < 9.0e18)) {
      for (long long i = tid; i < nout; i += T) oset[i] = 0.0;
      if (tid == 0) {
        if (a.status) a.status[set] = GAB1_ST_THROW;
        if (a.n_saved) a.n_saved[set] = 0;
        if (a.n_steps) a.n_steps[set] = 0;
        if (a.n_bc) a.n_bc[set] = 0;
      }
      continue;
    }
    const long long Nt = (long long)nt_f;

    // ---- state: both buffers (the reference zero-initialises column 2) ----
    {
      const bool on = node >= 1 && node <= Nr;
#pragma unroll
      for (int q = 0; q < NCY; ++q) { u_at(0, q, tid) = 0.0; u_at(1, q, tid) = 0.0; }
      u_at(0, iSFK, tid) = on ? CoSFK : 0.0;
      u_at(0, GAB1, tid) = on ? CoG1 : 0.0;
      u_at(0, GRB2, tid) = on ? CoG2 : 0.0;
      u_at(0, SHP2, tid) = on ? CoS2 : 0.0;
    }
    __syncthreads();                                     // the initial state is visible to the neighbours
    int cur = 0;                                         // buffer holding the current time level
    // value of node n (0..Nr) of species q in buffer b; node 0 mirrors node 1
    auto at_node = [&](int b, int q, int n) -> double { return u_at(b, q, (n < 1 ? 1 : n) - off); };
    auto stot_at = [&](int b, int n) { return __dadd_rn(at_node(b, PG1S, n), at_node(b, G2PG1S, n)); };
    auto ptot_at = [&](int b, int n) {
      const double g2pg1 = at_node(b, G2PG1, n), pg1 = at_node(b, pGAB1, n), pg1s = at_node(b, PG1S, n), g2pg1s = at_node(b, G2PG1S, n);
      if (a.o.pg1tot_form == GAB1_PG1TOT_VIA_STOT) return __dadd_rn(__dadd_rn(g2pg1, pg1), __dadd_rn(pg1s, g2pg1s));
      return __dadd_rn(__dadd_rn(__dadd_rn(g2pg1, pg1), pg1s), g2pg1s);
    };

    if (a.o.out_mode == GAB1_OUT_FULL) {                 // initial column (basepdesolver.jl:94-97,111)
      long long o2 = 0;
      for (int mi = 0; mi < 12; ++mi) {
        if (!((a.o.matrix_mask >> mi) & 1u)) continue;
        const double v0 = mi == GAB1_M_iSFK ? CoSFK : mi == GAB1_M_GRB2 ? CoG2 : mi == GAB1_M_SHP2 ? CoS2 : mi == GAB1_M_GAB1 ? CoG1 : 0.0;
        for (int n = tid; n < P; n += T) oset[o2 + n] = v0;
        o2 += (long long)P * Cn;
      }
      if (tid < GAB1_N_VECTORS) oset[o2 + (long long)tid * Cn] = tid == GAB1_V_mE ? CoEGFR : 0.0;
    }

    double t = 0.0, t_save = a.o.dt_save;
    int nts = 1;
    const double modulus_step = (a.o.save_rule == GAB1_SAVE_MODULUS) ? rint(__ddiv_rn((double)Nt, (double)Nts)) : 0.0;
    long long bc_total = 0;
    double kp_now = kv[10];
    double pct_ave = 0.0, pct_memb = 0.0;
    bool dead = false;
    long long step = 1;

    const double kS2f_t = kv[0] * dt, kS2r_t = kv[1] * dt, kG1f_t = kv[2] * dt, kG1r_t = kv[3] * dt,
                 kG1p_t = kv[6] * dt, kG1dp_t = kv[7] * dt, kSi_t = kv[9] * dt;
    const double Dt_Si = D_Si * dt, Dt_Sa = D_Sa * dt, Dt_G1 = Dv[4] * dt, Dt_G2 = Dv[1] * dt, Dt_G2G1 = Dv[2] * dt,
                 Dt_S2 = Dv[6] * dt, Dt_G1S2 = Dv[5] * dt, Dt_G2G1S2 = Dv[3] * dt;

    // ---- membrane block: lane roles of the last warp (solver_kernel.cuh) ----
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
    const double ca = kv[8] * (a.o.dr / D_Sa);
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
    double x = (lane == ML + mE) ? CoEGFR : 0.0;         // meaningful in the membrane warp only

    bool pulse_pending = pulse;
    if (pulse_pending && a.o.t_prechase + dt > t && t >= a.o.t_prechase) {
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
    // the membrane values, broadcast from the membrane warp through the header (after a __syncthreads())
    auto publish_m = [&]() { if (mwarp && lane >= ML && lane < LE) hdr[16 + lane - ML] = x; };
    // one snapshot column of GAB1_OUT_FULL (basepdesolver.jl:268-294); `b` holds the new time level, hdr[16..23] the membrane values
    auto write_column = [&](int c, int b) {
      long long o2 = 0;
      constexpr int kSpecies[10] = {iSFK, aSFK, GRB2, GAB1, SHP2, G2G1, G2PG1, G2PG1S, pGAB1, PG1S};
      bool ns = false;
      for (int mi = 0; mi < 12; ++mi) {
        if (!((a.o.matrix_mask >> mi) & 1u)) continue;
        // 16-byte stores, pairs aligned by shedding the first node of a column that starts on an odd double (solver_kernel.cuh: store_row_v2)
        auto val = [&](int n) { return mi < 10 ? at_node(b, kSpecies[mi], n) : (mi == GAB1_M_PG1tot ? ptot_at(b, n) : stot_at(b, n)); };
        double* col = oset + o2 + (long long)c * P;
        const int head = (int)((reinterpret_cast<unsigned long long>(col) >> 3) & 1ull), npairs = (P - head) >> 1;
        for (int j = tid; j < npairs; j += T) stg2(col + head + 2 * j, val(head + 2 * j), val(head + 2 * j + 1));
        if (tid == T - 1) {
          if (head) col[0] = val(0);
          if (head + 2 * npairs < P) col[P - 1] = val(P - 1);
        }
        o2 += (long long)P * Cn;
      }
      for (int n = tid; n < P; n += T) ns |= isnan(at_node(b, PG1S, n));
      if (ns) status |= GAB1_ST_NAN;
      if (tid == 0) {
        const double* m = hdr + 16;
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
    // the staged row of PG1S + G2PG1S for the reductions (every thread then evaluates them redundantly)
    auto stage_stot = [&](int b) {
      __syncthreads();
      for (int n = tid; n < P; n += T) row[n] = stot_at(b, n);
      __syncthreads();
    };

    int countdown = plan();
    if (Nt >= 1)
    for (;;) {
      const int nxt = cur ^ 1;
      // ---- interior: every thread its node, old buffer -> new buffer (basepdesolver.jl:150-180) ----
      {
        const double Si = u_at(cur, iSFK, tid), Sa = u_at(cur, aSFK, tid), G1 = u_at(cur, GAB1, tid), pG1 = u_at(cur, pGAB1, tid),
                     G2 = u_at(cur, GRB2, tid), g2g1 = u_at(cur, G2G1, tid), g2pg1 = u_at(cur, G2PG1, tid), S2 = u_at(cur, SHP2, tid),
                     pg1s = u_at(cur, PG1S, tid), g2pg1s = u_at(cur, G2PG1S, tid);
        auto lap = [&](int q, double uc) { return fma(cpc, u_at(cur, q, sr), fma(cmc, u_at(cur, q, sl), c0c * uc)); };
        const double gb = kG1f_t * G2, ph = kG1p_t * Sa, sb = kS2f_t * S2;
        const double v1 = fma(gb, G1, -(kG1r_t * g2g1));
        const double v3 = fma(gb, pG1, -(kG1r_t * g2pg1));
        const double v5 = fma(gb, pg1s, -(kG1r_t * g2pg1s));
        const double v2 = fma(ph, G1, -(kG1dp_t * pG1));
        const double v6 = fma(ph, g2g1, -(kG1dp_t * g2pg1));
        const double v4 = fma(sb, pG1, -(kS2r_t * pg1s));
        const double v7 = fma(sb, g2pg1, -(kS2r_t * g2pg1s));
        u_at(nxt, iSFK, tid) = fma(Dt_Si, lap(iSFK, Si), fma(kSi_t, Sa, Si));
        u_at(nxt, aSFK, tid) = fma(Dt_Sa, lap(aSFK, Sa), fma(-kSi_t, Sa, Sa));
        u_at(nxt, GAB1, tid) = fma(Dt_G1, lap(GAB1, G1), G1 - v1 - v2);
        u_at(nxt, pGAB1, tid) = fma(Dt_G1, lap(pGAB1, pG1), pG1 - v3 + v2 - v4);
        u_at(nxt, GRB2, tid) = fma(Dt_G2, lap(GRB2, G2), G2 - v1 - v3 - v5);
        u_at(nxt, G2G1, tid) = fma(Dt_G2G1, lap(G2G1, g2g1), g2g1 + v1 - v6);
        u_at(nxt, G2PG1, tid) = fma(Dt_G2G1, lap(G2PG1, g2pg1), g2pg1 + v3 + v6 - v7);
        u_at(nxt, SHP2, tid) = fma(Dt_S2, lap(SHP2, S2), S2 - v4 - v7);
        u_at(nxt, PG1S, tid) = fma(Dt_G1S2, lap(PG1S, pg1s), pg1s + v4 - v5);
        u_at(nxt, G2PG1S, tid) = fma(Dt_G2G1S2, lap(G2PG1S, g2pg1s), g2pg1s + v5 + v7);
      }
      bool flag_check = false;
      if (mwarp) {
        // ---- membrane fixed point on the last warp: needs only u+[Nr-1], which lane 30 of this warp has just written ----
        const double m_old = x;
        const double m_next = shfl_down1(m_old);
        const double f = fma(m_old, fma(alpha2, m_old, alpha), -(beta * m_next));
        const double base = fma(dt, fma(s_own, f, s_src * shfl(f, f_src)), m_old);
        const double Md1 = shfl(m_old, src_den), Mn1 = shfl(m_old, src_num);
        const double A_t = kf_t * Md1;
        const double B_t = kr_t * Mn1;
        const double rden1 = fast_recip(fma(cf, Md1, 1.0));
        __syncwarp();
        const double Iq = lane < NCY ? u_at(nxt, lane, T - 2) : 0.0;
        const double Ii = u_at(nxt, iSFK, T - 2);
        const double cr = lane == aSFK ? fma(cf, Iq, ca * Ii) : cr_fixed;
        (void)iq_idx;
        int it = 1;
        bool unconverged = false, nan_exit = false;
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
        bc_total += it;
        if (lane < NCY) u_at(nxt, lane, T - 1) = x;            // the boundary node Nr
        flag_check = unconverged || nan_exit;
        if (flag_check && lane == 0) s_dead[step & 1] = 1;        // ask the team to test for an all-NaN state
      }
      __syncthreads();                                            // the new time level is complete
      cur = nxt;
      // every warp has passed this step's barrier, so nobody reads the previous step's slot any more: the membrane warp
      // clears it here, before it can raise it again at the next step (the slots alternate, so a membrane warp that is
      // a whole step ahead of a slow warp never touches the flag that warp is about to read)
      if (mwarp && lane == 0) s_dead[(step & 1) ^ 1] = 0;
      if (s_dead[step & 1]) {                                     // uniform: written before the barrier
        const bool mine = !(node >= 1 && node <= Nr) ||
                          (isnan(u_at(cur, iSFK, tid)) && isnan(u_at(cur, aSFK, tid)) && isnan(u_at(cur, GAB1, tid)) &&
                           isnan(u_at(cur, pGAB1, tid)) && isnan(u_at(cur, GRB2, tid)) && isnan(u_at(cur, G2G1, tid)) &&
                           isnan(u_at(cur, G2PG1, tid)) && isnan(u_at(cur, SHP2, tid)) && isnan(u_at(cur, PG1S, tid)) &&
                           isnan(u_at(cur, G2PG1S, tid)));
        const bool memb = !mwarp || (lane < ML || lane >= LE) || isnan(x);
        dead = __syncthreads_and(mine && memb);
        if (dead) countdown = 1;
      }
      t = t + dt;
      if (--countdown > 0) { ++step; continue; }
      // ---- rare path: exact event tests for the step just taken ----
      if (track_t) {
        const bool save = a.o.save_rule == GAB1_SAVE_T_GE_TSAVE ? (t >= t_save) : (fmod((double)step, modulus_step) == 0.0);
        if (save) {
          if (nts >= Cn) status |= GAB1_ST_OVERFLOW;
          else {
            const int c = nts++;
            publish_m();
            __syncthreads();
            if (a.o.out_mode == GAB1_OUT_FULL) write_column(c, cur);
            else if (c == Cn - 1) {
              stage_stot(cur);
              pct_ave = trapz_r2(a.r, row, P);
              pct_memb = hdr[16 + EG2PG1S];
            }
            __syncthreads();
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
    publish_m();
    __syncthreads();
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
          if (a.o.out_mode == GAB1_OUT_FULL) write_column(c, cur);
          else if (c == Cn - 1) { pct_ave = CUDART_NAN; pct_memb = CUDART_NAN; }
        }
        if (a.o.save_rule == GAB1_SAVE_T_GE_TSAVE) t_save = t_save + a.o.dt_save;
      }
    }
    // final-time view: column 2 of the reference's work arrays; after zero steps it is still all zeros (sapdesolver.jl:245)
    const int fin = Nt == 0 ? 1 : cur;
    double m[NMB];
#pragma unroll
    for (int j = 0; j < NMB; ++j) m[j] = Nt == 0 ? 0.0 : hdr[16 + j];
    bool ns = false;
    if (a.o.out_mode == GAB1_OUT_FINAL4) {
      for (int n = tid; n < P; n += T) {
        const double v0 = at_node(fin, iSFK, n), v1 = at_node(fin, aSFK, n), v2 = ptot_at(fin, n), v3 = stot_at(fin, n);
        oset[n] = v0; oset[P + n] = v1; oset[2 * P + n] = v2; oset[3 * P + n] = v3;
        ns |= isnan(v0) || isnan(v1) || isnan(v2) || isnan(v3);
      }
    } else if (a.o.out_mode == GAB1_OUT_FINAL_STATE) {
      for (int q = 0; q < NCY; ++q)
        for (int n = tid; n < P; n += T) { const double v = at_node(fin, q, n); oset[(long long)q * P + n] = v; ns |= isnan(v); }
      if (tid == 0) {
#pragma unroll
        for (int j = 0; j < NMB; ++j) { oset[(long long)NCY * P + j] = m[j]; ns |= isnan(m[j]); }
      }
    } else if (a.o.out_mode == GAB1_OUT_SIX) {
      double* rowB = row;                                         // PG1S + G2PG1S; aSFK is read in place
      stage_stot(fin);
      // aSFK row: contiguous in shared memory except node 0 (mirror of node 1): stage it behind rowB
      double* rowA = row + a.P_pad;
      for (int n = tid; n < P; n += T) rowA[n] = at_node(fin, aSFK, n);
      __syncthreads();
      bool threw = false;
      double six[6];
      const double R = a.o.R;
      six[0] = length_scale(a.r, rowA, P, 0.5, R, threw);
      six[1] = length_scale(a.r, rowA, P, 0.1, R, threw);
      six[2] = length_scale(a.r, rowB, P, 0.5, R, threw);
      six[3] = length_scale(a.r, rowB, P, 0.1, R, threw);
      six[4] = __ddiv_rn(rowB[0], rowB[P - 1]);
      six[5] = __ddiv_rn(__dmul_rn(trapz_r2(a.r, rowB, P), 3.0), a.R_pow3);
      if (threw) status |= GAB1_ST_THROW;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const double v = threw ? 0.0 : six[i];
        ns |= isnan(v);
        if (tid == 0) oset[i] = v;
      }
    }
    if (a.o.out_mode == GAB1_OUT_PCT_BOUND) {
      const double R = a.o.R;
      const double ave = __ddiv_rn(__dmul_rn(pct_ave, 3.0), __dmul_rn(__dmul_rn(R, R), R));
      const double mem = __ddiv_rn(__dmul_rn(pct_memb, a.o.pct_mul), a.o.pct_div);
      const double pct = __dmul_rn(__ddiv_rn(__dadd_rn(ave, mem), CoG1), 100.0);
      ns |= isnan(pct);
      if (tid == 0) oset[0] = pct;
    }
    if (ns) status |= GAB1_ST_NAN;
    if (track_t && nts < Cn) {
      status |= GAB1_ST_SHORT;
      if (a.o.out_mode == GAB1_OUT_FULL) {
        long long o2 = 0;
        for (int mi = 0; mi < 12; ++mi) {
          if (!((a.o.matrix_mask >> mi) & 1u)) continue;
          for (long long i = (long long)nts * P + tid; i < (long long)Cn * P; i += T) oset[o2 + i] = 0.0;
          o2 += (long long)P * Cn;
        }
        for (int v = 0; v < GAB1_N_VECTORS; ++v)
          for (int c = nts + tid; c < Cn; c += T) oset[o2 + (long long)v * Cn + c] = 0.0;
      }
    }
    // status bits of every thread (NaN seen in its share of a row, the membrane warp's cap flag) and the pass count
    int st_all = 0;                                               // __syncthreads_or is a logical OR: one bit at a time
#pragma unroll
    for (unsigned b = 1u; b <= GAB1_ST_THROW; b <<= 1)
      if (__syncthreads_or((int)(status & b))) st_all |= (int)b;
    if (mwarp && lane == 0) {
      if (a.status) a.status[set] = st_all;
      if (a.n_saved) a.n_saved[set] = track_t ? nts : 0;
      if (a.n_steps) a.n_steps[set] = Nt;
      if (a.n_bc) a.n_bc[set] = bc_total;
    }
  }
}

}  // namespace gab1
