// tangent_kernel.cuh — forward-mode (tangent) twin of the one-set-per-warp solver: the state of a parameter set AND
// its directional derivatives along NT input directions are stepped together, in registers, by one warp.
//
// What it stands in for: the reference differentiates its solver by running pdesolver_fitting on ForwardDiff dual
// numbers (basepdesolver.jl:674-932 generic in T; callers param_fitting+inference_finitediff.jl:128-151 testf /
// ForwardDiff.gradient, :188-240 loss under AutoForwardDiff, :308-370 turing_model under NUTS).  A dual number there
// is (value, partials): every arithmetic operation applies the sum/product/quotient rule to the partials, while every
// DECISION (membrane loop exit `error <= tol`, snapshot test `t >= t_save`, `ceil(tf/dt)`) looks at the value alone.
// So the control flow is the primal solve's, and the partials ride along.  dt is computed inside the solver from p
// (:696) and therefore carries partials too; the caller passes dt and its partials (gab1_default_dt_tangent).
//
// Mapping (same as solver_kernel.cuh): nodes 1..Nr right-aligned over the 32*K slots of a warp, all ten species of a
// node — value and NT partials — in that lane's registers; neighbours by shuffle; the membrane fixed point
// lane-parallel (lanes 0-9 one Robin closure each, lanes 10-17 one membrane species, lane 18 Etot), every tracked
// quantity a dual.  A request for n_dir directions is cut into ceil(n_dir/NT) work items per set; each item
// recomputes the primal (identically) and group 0 alone writes the value block and the diagnostics.
//
// Arithmetic: the `fast` forms of solver_kernel.cuh (hoisted reciprocals, rate constants pre-scaled by dt, net
// fluxes, FMA), differentiated term by term.  Bound asserted in tests: values 1e-9 relative (as the primal), partials
// 1e-9 of the largest magnitude of the same output array.
#pragma once
#include "solver_kernel.cuh"

namespace gab1 {

constexpr int TWS_HDR = 32;   // doubles per component in the per-warp smem header: [0,16) inner neighbour, [16,32) boundary

template <int NT>
struct Dn {
  double v;
  double p[NT];
};
template <int NT> __device__ __forceinline__ Dn<NT> dconst(double x) { Dn<NT> r; r.v = x;
#pragma unroll
  for (int n = 0; n < NT; ++n) r.p[n] = 0.0; return r; }
template <int NT> __device__ __forceinline__ Dn<NT> operator+(const Dn<NT>& a, const Dn<NT>& b) { Dn<NT> r; r.v = a.v + b.v;
#pragma unroll
  for (int n = 0; n < NT; ++n) r.p[n] = a.p[n] + b.p[n]; return r; }
template <int NT> __device__ __forceinline__ Dn<NT> operator-(const Dn<NT>& a, const Dn<NT>& b) { Dn<NT> r; r.v = a.v - b.v;
#pragma unroll
  for (int n = 0; n < NT; ++n) r.p[n] = a.p[n] - b.p[n]; return r; }
template <int NT> __device__ __forceinline__ Dn<NT> operator*(const Dn<NT>& a, const Dn<NT>& b) { Dn<NT> r; r.v = a.v * b.v;
#pragma unroll
  for (int n = 0; n < NT; ++n) r.p[n] = fma(a.p[n], b.v, a.v * b.p[n]); return r; }
// a*b + c
template <int NT> __device__ __forceinline__ Dn<NT> dfma(const Dn<NT>& a, const Dn<NT>& b, const Dn<NT>& c) { Dn<NT> r; r.v = fma(a.v, b.v, c.v);
#pragma unroll
  for (int n = 0; n < NT; ++n) r.p[n] = fma(a.p[n], b.v, fma(a.v, b.p[n], c.p[n])); return r; }
// a*b - c
template <int NT> __device__ __forceinline__ Dn<NT> dfms(const Dn<NT>& a, const Dn<NT>& b, const Dn<NT>& c) { Dn<NT> r; r.v = fma(a.v, b.v, -c.v);
#pragma unroll
  for (int n = 0; n < NT; ++n) r.p[n] = fma(a.p[n], b.v, fma(a.v, b.p[n], -c.p[n])); return r; }
// s*a + c, s without partials
template <int NT> __device__ __forceinline__ Dn<NT> daxpy(double s, const Dn<NT>& a, const Dn<NT>& c) { Dn<NT> r; r.v = fma(s, a.v, c.v);
#pragma unroll
  for (int n = 0; n < NT; ++n) r.p[n] = fma(s, a.p[n], c.p[n]); return r; }
template <int NT> __device__ __forceinline__ Dn<NT> dscale(double s, const Dn<NT>& a) { Dn<NT> r; r.v = s * a.v;
#pragma unroll
  for (int n = 0; n < NT; ++n) r.p[n] = s * a.p[n]; return r; }
// s / a, s without partials, true division (a may be 1e-32)
template <int NT> __device__ __forceinline__ Dn<NT> drdiv(double s, const Dn<NT>& a) { Dn<NT> r; r.v = s / a.v; const double w = -(r.v / a.v);
#pragma unroll
  for (int n = 0; n < NT; ++n) r.p[n] = w * a.p[n]; return r; }
template <int NT> __device__ __forceinline__ Dn<NT> dshfl(const Dn<NT>& a, int src) { Dn<NT> r; r.v = shfl(a.v, src);
#pragma unroll
  for (int n = 0; n < NT; ++n) r.p[n] = shfl(a.p[n], src); return r; }
template <int NT> __device__ __forceinline__ Dn<NT> dshfl_down1(const Dn<NT>& a) { Dn<NT> r; r.v = shfl_down1(a.v);
#pragma unroll
  for (int n = 0; n < NT; ++n) r.p[n] = shfl_down1(a.p[n]); return r; }
template <int NT> __device__ __forceinline__ Dn<NT> dshfl_up1(const Dn<NT>& a) { Dn<NT> r; r.v = shfl_up1(a.v);
#pragma unroll
  for (int n = 0; n < NT; ++n) r.p[n] = shfl_up1(a.p[n]); return r; }

template <int K, int NT>
__device__ void solve_set_tangent(const TangentArgs& ta, long long set, int group, int lane, double* ws, const Grid<K>& g) {
  typedef Dn<NT> T;
  const KernelArgs& a = ta.a;
  const int Nr = a.o.Nr, P = Nr + 1, Nts = a.o.Nts, Cn = Nts + 1;
  const long long nout = a.out_stride;
  double* rowA = ws + TWS_HDR * (1 + NT);
  double* rowB = rowA + a.P_pad;
  double* oset = a.out + set * nout * (1 + ta.n_dir);
  const bool lead = group == 0;                 // this work item owns the value block and the diagnostics
  unsigned status = 0, scratch_status = 0;

  // component c of this item: c = 0 value (stored by the lead item only), c = 1 + n partial along direction group*NT + n
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
  const T CoSFK = Cod(0), CoG2 = Cod(1), CoG1 = Cod(2), CoS2 = Cod(3), CoEGFR = Cod(4);
  T D_Si = Dd(0), D_Sa = Dd(0);
  if (a.o.sfk_mode == GAB1_SFK_MEMBRANE) D_Sa = dconst<NT>(1e-32);                                 // basepdesolver.jl:366
  if (a.o.sfk_mode == GAB1_SFK_BOTH_FROZEN) { D_Si = dconst<NT>(1e-32); D_Sa = dconst<NT>(1e-32); }  // basepdesolver_rect.jl:305-306

  const bool track_t = (a.o.out_mode == GAB1_OUT_FULL || a.o.out_mode == GAB1_OUT_PCT_BOUND);

  // Nt = Int64(ceil(tf/dt)): the value decides (basepdesolver.jl:729-735)
  const double nt_f = ceil(__ddiv_rn(a.o.tf, dt.v));
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

  // ---- state: SoA over components so that the output writers of solver_kernel.cuh serve every component ----
  double uv[NCY][K];
  double up[NT][NCY][K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const bool on = g.node[i] >= 1 && g.node[i] <= Nr;
#pragma unroll
    for (int q = 0; q < NCY; ++q) {
      uv[q][i] = 0.0;
#pragma unroll
      for (int n = 0; n < NT; ++n) up[n][q][i] = 0.0;
    }
    uv[iSFK][i] = on ? CoSFK.v : 0.0;      // basepdesolver.jl:776-779
    uv[GAB1][i] = on ? CoG1.v : 0.0;
    uv[GRB2][i] = on ? CoG2.v : 0.0;
    uv[SHP2][i] = on ? CoS2.v : 0.0;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      up[n][iSFK][i] = on ? CoSFK.p[n] : 0.0;
      up[n][GAB1][i] = on ? CoG1.p[n] : 0.0;
      up[n][GRB2][i] = on ? CoG2.p[n] : 0.0;
      up[n][SHP2][i] = on ? CoS2.p[n] : 0.0;
    }
  }
  auto ld = [&](int q, int i) { T r; r.v = uv[q][i];
#pragma unroll
    for (int n = 0; n < NT; ++n) r.p[n] = up[n][q][i]; return r; };
  auto st = [&](int q, int i, const T& x) { uv[q][i] = x.v;
#pragma unroll
    for (int n = 0; n < NT; ++n) up[n][q][i] = x.p[n]; };

  const int lane_b = g.G - 1;
  constexpr int idx_b = K - 1;
  const int lane_i = K >= 2 ? g.G - 1 : g.G - 2;
  constexpr int idx_i = K >= 2 ? K - 2 : 0;

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

  // ---- interior constants: rate constants and diffusivities pre-scaled by dt (duals) ----
  const T kS2f_t = kd(0) * dt, kS2r_t = kd(1) * dt, kG1f_t = kd(2) * dt, kG1r_t = kd(3) * dt, kG1p_t = kd(6) * dt,
          kG1dp_t = kd(7) * dt, kSi_t = kd(9) * dt;
  const T Dt_Si = D_Si * dt, Dt_Sa = D_Sa * dt, Dt_G1 = Dd(4) * dt, Dt_G2 = Dd(1) * dt, Dt_G2G1 = Dd(2) * dt,
          Dt_S2 = Dd(6) * dt, Dt_G1S2 = Dd(5) * dt, Dt_G2G1S2 = Dd(3) * dt;

  // ---- membrane block: lane roles (solver_kernel.cuh, fast path) ----
  constexpr int LZ = 31, LE = ML + NMB;
  T kf = dconst<NT>(0.0), kr = dconst<NT>(0.0), Dq = dconst<NT>(1.0);
  int src_num = LZ, src_den = LZ;
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
  const T cf = kf * drD;
  const T cr_fixed = kr * drD;
  const T ca = kd(8) * drdiv<NT>(a.o.dr, D_Sa);          // aSFK closure coefficient; a true division (D_Sa may be 1e-32)
  const bool is_flux = lane >= GAB1 && lane <= G2PG1S;
  const T kf_t = is_flux ? kf * dt : dconst<NT>(0.0), kr_t = is_flux ? kr * dt : dconst<NT>(0.0);
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
  T alpha = dconst<NT>(0.0), alpha2 = dconst<NT>(0.0), beta = dconst<NT>(0.0);
  double s_own = 0.0, s_src = 0.0;
  int f_src = LZ;
  switch (lane - ML) {
    case mE:     alpha = kd(12) * kd(14); beta = kd(13); s_own = -1.0; break;
    case mES:    alpha2 = kd(15);         beta = kd(16); s_own = -2.0; s_src = 1.0; f_src = ML + mE; break;
    case mESmES: alpha = kd(10);          beta = kd(11); s_own = -1.0; s_src = 1.0; f_src = ML + mES; break;
    case E:      s_src = 1.0; f_src = ML + mESmES; break;
    case NMB:    s_src = 2.0; f_src = ML + mESmES; break;
    default: break;
  }
  const double tol = a.o.tol;
  const bool untracked = lane >= LE;
  const int iq_idx = lane < NCY ? lane : 10;             // header slots 10..15 stay zero
  const int maxiters = a.o.maxiters;
  const unsigned ws_s = (unsigned)__cvta_generic_to_shared(ws);
  auto hdr_ld = [&](int slot) { T r; r.v = lds(ws_s + 8 * slot);
#pragma unroll
    for (int n = 0; n < NT; ++n) r.p[n] = lds(ws_s + 8 * (TWS_HDR * (1 + n) + slot)); return r; };
  auto hdr_st = [&](int slot, const T& x) { sts(ws_s + 8 * slot, x.v);
#pragma unroll
    for (int n = 0; n < NT; ++n) sts(ws_s + 8 * (TWS_HDR * (1 + n) + slot), x.p[n]); };

  // x: the dual this lane tracks across iterations and steps (boundary value / membrane species / Etot / zero)
  T x = (lane == ML + mE) ? CoEGFR : dconst<NT>(0.0);

  T t = dconst<NT>(0.0);
  double t_save = a.o.dt_save;
  int nts = 1;
  long long bc_total = 0;
  T pct_ave = dconst<NT>(0.0), pct_memb = dconst<NT>(0.0);

  // component view of the membrane column, gathered from the membrane lanes
  auto gather_m = [&](int c, double (&m)[NMB]) {
#pragma unroll
    for (int j = 0; j < NMB; ++j) {
      double val = x.v;
#pragma unroll
      for (int n = 0; n < NT; ++n) if (c == n + 1) val = x.p[n];
      m[j] = shfl(val, ML + j);
    }
  };
  // one snapshot column / the PCT capture, for every component this item owns
  auto snapshot = [&](int col) {
    double mv[NMB];
    gather_m(0, mv);
    if (a.o.out_mode == GAB1_OUT_FULL) {
      if (lead) write_full_column<K>(a, oset, col, uv, mv, t.v, CoEGFR.v, lane, g, rowA, status);
      const double Etot_v = 2.0 * (mv[E] + mv[EG2] + mv[EG2G1] + mv[EG2PG1] + mv[EG2PG1S]);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        double* ob = block_of(1 + n);
        if (!ob) continue;
        double mp[NMB];
        gather_m(1 + n, mp);
        write_full_column<K>(a, ob, col, up[n], mp, t.p[n], CoEGFR.v, lane, g, rowA, scratch_status);
        if (lane == 0) {      // the two outputs that divide by CoEGFR (basepdesolver.jl:287; basepdesolver_rect.jl:264): quotient rule
          double* v = ob + (long long)__popc(a.o.matrix_mask & GAB1_MASK_ALL_MATRICES) * P * Cn;
          const double Etot_p = 2.0 * (mp[E] + mp[EG2] + mp[EG2G1] + mp[EG2PG1] + mp[EG2PG1S]);
          const double w = CoEGFR.p[n] / CoEGFR.v;
          v[GAB1_V_pE * (long long)Cn + col] = (Etot_p * 100.0) / CoEGFR.v - (Etot_v * 100.0 / CoEGFR.v) * w;
          v[GAB1_V_EGFR_SHP2 * (long long)Cn + col] = (mp[EG2PG1S] * 100.0) / CoEGFR.v - (mv[EG2PG1S] * 100.0 / CoEGFR.v) * w;
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
        stage_row<K>(rowA, lane, g, Nr, [&](int i) { return derived_stot<K>(up[n], i); });
        pct_ave.p[n] = trapz_r2(a.r, rowA, P);
        pct_memb.p[n] = mp[EG2PG1S];
        __syncwarp();
      }
    }
  };

  for (long long step = 1; step <= Nt; ++step) {
    // ---- membrane prologue: everything that depends only on old-time values (independent of the interior) ----
    const T m_old = x;
    const T m_next = dshfl_down1<NT>(m_old);
    const T f = dfms<NT>(m_old, dfma<NT>(alpha2, m_old, alpha), beta * m_next);
    const T fsrc = dshfl<NT>(f, f_src);
    T dm;
    dm.v = fma(s_own, f.v, s_src * fsrc.v);
#pragma unroll
    for (int n = 0; n < NT; ++n) dm.p[n] = fma(s_own, f.p[n], s_src * fsrc.p[n]);
    const T base = dfma<NT>(dt, dm, m_old);
    const T Md1 = dshfl<NT>(m_old, src_den), Mn1 = dshfl<NT>(m_old, src_num);
    const T A_t = kf_t * Md1;
    const T B_t = kr_t * Mn1;

    // ---- interior (basepdesolver.jl:797-827), in place: fluxes of every node first, then species by species ----
    {
      T v1[K], v2[K], v3[K], v4[K], v5[K], v6[K], v7[K], sk[K];
#pragma unroll
      for (int i = 0; i < K; ++i) {
        const T Sa = ld(aSFK, i), G1 = ld(GAB1, i), pG1 = ld(pGAB1, i), G2 = ld(GRB2, i), g2g1 = ld(G2G1, i),
                g2pg1 = ld(G2PG1, i), S2 = ld(SHP2, i), pg1s = ld(PG1S, i), g2pg1s = ld(G2PG1S, i);
        const T gb = kG1f_t * G2, ph = kG1p_t * Sa, sb = kS2f_t * S2;
        v1[i] = dfms<NT>(gb, G1, kG1r_t * g2g1);         // GRB2 + GAB1   <-> G2G1
        v3[i] = dfms<NT>(gb, pG1, kG1r_t * g2pg1);       // GRB2 + pGAB1  <-> G2PG1
        v5[i] = dfms<NT>(gb, pg1s, kG1r_t * g2pg1s);     // GRB2 + PG1S   <-> G2PG1S
        v2[i] = dfms<NT>(ph, G1, kG1dp_t * pG1);         // GAB1  <-> pGAB1
        v6[i] = dfms<NT>(ph, g2g1, kG1dp_t * g2pg1);     // G2G1  <-> G2PG1
        v4[i] = dfms<NT>(sb, pG1, kS2r_t * pg1s);        // SHP2 + pGAB1  <-> PG1S
        v7[i] = dfms<NT>(sb, g2pg1, kS2r_t * g2pg1s);    // SHP2 + G2PG1  <-> G2PG1S
        sk[i] = kSi_t * Sa;                              // aSFK -> iSFK
      }
      // species q: u+ = Dt_q * lap(u_q) + (u_q + net kinetics); both Laplacians of a lane are taken before either node
      // is overwritten, so the update is in place
      auto advance = [&](int q, const T& Dt, auto kin) {
        const T hl = dshfl_up1<NT>(ld(q, K - 1)), hr = dshfl_down1<NT>(ld(q, 0));
        T L[K];
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const T um = i > 0 ? ld(q, i - 1) : hl, upn = i + 1 < K ? ld(q, i + 1) : hr, uc = ld(q, i);
          L[i].v = fma(g.cp[i], upn.v, fma(g.cm[i], um.v, g.c0[i] * uc.v));
#pragma unroll
          for (int n = 0; n < NT; ++n) L[i].p[n] = fma(g.cp[i], upn.p[n], fma(g.cm[i], um.p[n], g.c0[i] * uc.p[n]));
        }
#pragma unroll
        for (int i = 0; i < K; ++i) st(q, i, dfma<NT>(Dt, L[i], kin(i)));
      };
      advance(iSFK, Dt_Si, [&](int i) { return ld(iSFK, i) + sk[i]; });
      advance(aSFK, Dt_Sa, [&](int i) { return ld(aSFK, i) - sk[i]; });
      advance(GAB1, Dt_G1, [&](int i) { return ld(GAB1, i) - v1[i] - v2[i]; });
      advance(pGAB1, Dt_G1, [&](int i) { return ld(pGAB1, i) - v3[i] + v2[i] - v4[i]; });
      advance(GRB2, Dt_G2, [&](int i) { return ld(GRB2, i) - v1[i] - v3[i] - v5[i]; });
      advance(G2G1, Dt_G2G1, [&](int i) { return ld(G2G1, i) + v1[i] - v6[i]; });
      advance(G2PG1, Dt_G2G1, [&](int i) { return ld(G2PG1, i) + v3[i] + v6[i] - v7[i]; });
      advance(SHP2, Dt_S2, [&](int i) { return ld(SHP2, i) - v4[i] - v7[i]; });
      advance(PG1S, Dt_G1S2, [&](int i) { return ld(PG1S, i) + v4[i] - v5[i]; });
      advance(G2PG1S, Dt_G2G1S2, [&](int i) { return ld(G2PG1S, i) + v5[i] + v7[i]; });
    }
    // ---- inner-neighbour values u+[Nr-1] to the closure lanes ----
    if (lane == lane_i) {
#pragma unroll
      for (int q = 0; q < NCY; ++q) hdr_st(q, ld(q, idx_i));
    }
    __syncwarp();
    const T Iq = hdr_ld(iq_idx);
    // aSFK: I_a + ca*Etot*I_i/(1 + cf*Etot) = (I_a + (cf*I_a + ca*I_i)*Etot)/(1 + cf*Etot)   (basepdesolver.jl:853-854)
    const T cr = lane == aSFK ? dfma<NT>(cf, Iq, ca * hdr_ld(iSFK)) : cr_fixed;

    // ---- fixed-point iterations (basepdesolver.jl:844-885); exit decided by the values alone ----
    int it = 0;
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
      // |1 - new/old| <= tol  <=>  |old - new| <= tol*|old| on the VALUES (solver_kernel.cuh finish_pass)
      const bool ok = (fabs(x.v - xnew.v) < tol * fabs(x.v)) || untracked;
      x = xnew;
      if (__all_sync(FULL, ok)) break;
      if (it >= maxiters) break;
      Mn = dshfl<NT>(x, src_num);
      Md = dshfl<NT>(x, src_den);
    }
    bc_total += it;
    // ---- boundary values back to the lane that owns node Nr ----
    if (lane < NCY) hdr_st(16 + lane, x);
    __syncwarp();
    if (lane == lane_b) {
#pragma unroll
      for (int q = 0; q < NCY; ++q) st(q, idx_b, hdr_ld(16 + q));
    }
    t = t + dt;                                                     // basepdesolver.jl:908
    if (track_t && t.v >= t_save) {                                 // :912
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
          for (int n = 0; n < NT; ++n) up[n][q][i] = 0.0; }
      x = dconst<NT>(0.0);
    }
    double m[NMB];
    gather_m(0, m);
    if (lead) write_final<K>(a, oset, uv, m, lane, g, rowA, rowB, status);
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      double* ob = block_of(1 + n);
      if (!ob) continue;
      gather_m(1 + n, m);
      write_final<K>(a, ob, up[n], m, lane, g, rowA, rowB, scratch_status);
    }
  }
  if (a.o.out_mode == GAB1_OUT_PCT_BOUND) {      // param_fitting+inference_finitediff.jl:211-216
    const double R = a.o.R, R3 = R * R * R;
    const double ave_v = pct_ave.v * 3.0 / R3, mem_v = pct_memb.v * a.o.pct_mul / a.o.pct_div;
    const double pct_v = (ave_v + mem_v) / CoG1.v * 100.0;
    if (isnan(pct_v)) status |= GAB1_ST_NAN;
    if (lead && lane == 0) oset[0] = pct_v;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      double* ob = block_of(1 + n);
      if (!ob) continue;
      const double ave_p = pct_ave.p[n] * 3.0 / R3, mem_p = pct_memb.p[n] * a.o.pct_mul / a.o.pct_div;
      if (lane == 0) ob[0] = ((ave_p + mem_p) / CoG1.v - ((ave_v + mem_v) / CoG1.v) * (CoG1.p[n] / CoG1.v)) * 100.0;
    }
  }
  if (track_t && nts < Cn) {
    status |= GAB1_ST_SHORT;
    if (a.o.out_mode == GAB1_OUT_FULL) {           // columns never due stay zero (zeros(T, …), basepdesolver.jl:753-758)
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

// Persistent kernel: warps pull (set, direction group) items from a queue ordered by descending step count.
template <int K, int NT>
__global__ void __launch_bounds__(32 * GAB1_WARPS, K <= 2 ? 2 : 1)
tangent_kernel(const TangentArgs ta) {
  extern __shared__ double smem[];
  const KernelArgs& a = ta.a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* ws = smem + (size_t)warp * (TWS_HDR * (1 + NT) + 2 * a.P_pad);
  const int Nr = a.o.Nr;
  for (int i = lane; i < TWS_HDR * (1 + NT); i += 32) ws[i] = 0.0;
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
    solve_set_tangent<K, NT>(ta, set, group, lane, ws, g);
    __syncwarp();
  }
}

}  // namespace gab1
