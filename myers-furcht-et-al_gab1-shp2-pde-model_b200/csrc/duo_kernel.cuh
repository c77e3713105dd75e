// duo_kernel.cuh — LATENCY path of the register-resident grids (32 < Nr <= 128): TWO warps per parameter set.
//
// Why.  A parameter set is one serial chain of Nt explicit steps.  In the throughput kernel (solver_kernel.cuh) one warp
// runs the interior update (~150 FP64 warp instructions, issue-bound) and then the membrane fixed point (~230 cycles per
// pass, latency-bound: shuffle -> reciprocal -> shuffle -> compare -> vote) back to back: ~770 cycles per step alone on
// a scheduler, ~960 when it shares one.  With heavy-tailed priors (dt shrinks with the sum of the rate constants) ONE set
// of 6.4e5 steps then takes 310 ms while an eighth of a 10^5-set ensemble is 245 ms of work per GPU: the longest member
// bounds the 8-GPU run.  Small batches (fewer sets than warps) are in the same position.
//
// What.  The two halves of a step only meet at the membrane: the fixed point of step n needs u_n[Nr-1] and yields the
// boundary value u_n[Nr], which only node Nr-1 reads at step n+1.  So
//   warp A  owns nodes 1..Nr-2 in the throughput kernel's own layout (K nodes per lane, right-aligned, shuffled halo) and
//           runs its interior update, unchanged;
//   warp B  owns node Nr-1 (all lanes compute it redundantly from broadcast loads: the same expressions, so the same
//           bits), the boundary node and the eight membrane species: the throughput kernel's lane-parallel fixed point.
// Per step they exchange ten doubles each way through shared memory — u_n[Nr-2] (A -> B) and u_n[Nr-1] (B -> A), both
// produced during step n and consumed at step n+1 — guarded by two named barriers used as arrive / sync pairs, so the
// warps run concurrently on two schedulers with a step of slack and the period is max(interior, node + fixed point)
// instead of their sum.  Every arithmetic expression is the throughput kernel's, in the same order: results, iteration
// counts and status words are BIT-IDENTICAL to solve_kernel<K, MODE> (tests/test_gpu_parity.py, GAB1_KERNEL=duo), which
// is what lets the dispatcher route any subset of a batch here (gab1pde.cu: duo_plan_kernel; duo_solve_kernel below).
//
// Rare events (snapshot due, pulse-chase switch, last step, dead state) are planned by both warps from the same scalars
// (the countdown of solver_kernel.cuh); at an event B publishes the boundary and membrane values, the CTA meets at a
// full barrier, and warp A — which then holds the throughput kernel's complete register state — runs the throughput
// kernel's own output writers.
//
// Reference: basepdesolver.jl:149-296 (time loop), :150-180 (interior), :197-242 (membrane loop).
#pragma once
#include <stdio.h>

#include <type_traits>

#include "solver_kernel.cuh"

namespace gab1 {

// shared-memory exchange block of one duo (doubles): [0,16) [16,32) u[Nr-1] of even / odd steps (slot 10 stays zero: the
// closure lanes' "no inner neighbour" source; slot 15 = "the previous step's fixed point did not converge");
// [32,48) [48,64) u[Nr-2] of even / odd steps; [64,80) boundary values u[Nr]; [80,88) membrane species; [88,96) bookkeeping
constexpr int DUO_DX = 96;
// named barriers of a pair (64 threads each), ids bar0 + {0, 1, 2}: "u[Nr-2] of this step is posted" / "u[Nr-1] of this
// step is posted" / both warps meet.  Nothing in the duo path is a CTA-wide barrier: the pairs of a CTA are independent.
enum { DUO_BAR_A = 0, DUO_BAR_B = 1, DUO_BAR_ALL = 2 };

__device__ __forceinline__ void duo_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void duo_bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

template <int K, int MODE>
__device__ void duo_solve_set(const KernelArgs& a, long long set, int lane, bool roleA, double* ws, double* dx,
                              const Grid<K>& g, int* s_flags, long long* s_bc, int bar0) {
  constexpr bool WHILE = MODE == MODE_FAST_WHILE;
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

  const bool track_t = (a.o.out_mode == GAB1_OUT_FULL || a.o.out_mode == GAB1_OUT_PCT_BOUND);
  const long long nout = a.out_stride;

  // Nt = Int64(ceil(tf/dt)) (basepdesolver.jl:72)
  const double nt_f = ceil(__ddiv_rn(a.o.tf, dt));
  if (!(nt_f >= 0.0 && nt_f < 9.0e18)) {
    if (roleA) {
      for (long long i = lane; i < nout; i += 32) oset[i] = 0.0;
      if (lane == 0) {
        if (a.status) a.status[set] = GAB1_ST_THROW;
        if (a.n_saved) a.n_saved[set] = 0;
        if (a.n_steps) a.n_steps[set] = 0;
        if (a.n_bc) a.n_bc[set] = 0;
      }
    }
    return;
  }
  const long long Nt = (long long)nt_f;

  // ---- where nodes Nr, Nr-1, Nr-2 live in warp A's layout (flat slot of node n: n - 1 - off, off = Nr - G*K) ----
  const int lane_b = g.G - 1;
  constexpr int idx_b = K - 1;
  const int lane_i = (g.G * K - 2) / K;
  constexpr int idx_i = ((K - 2) % K + K) % K;
  const int lane_i2 = (g.G * K - 3) / K;
  constexpr int idx_i2 = ((K - 3) % K + K) % K;

  // ---- state ----
  double u[NCY][K];         // warp A: the throughput kernel's register state (nodes Nr-1 and Nr are copies of B's values)
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
  double c[NCY];            // warp B: node Nr-1, the same values in every lane
#pragma unroll
  for (int q = 0; q < NCY; ++q) c[q] = 0.0;
  c[iSFK] = CoSFK; c[GAB1] = CoG1; c[GRB2] = CoG2; c[SHP2] = CoS2;

  // initial column of the FULL output (basepdesolver.jl:94-97,111)
  if (roleA && a.o.out_mode == GAB1_OUT_FULL) {
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
  double pct_ave = 0.0, pct_memb = 0.0;
  bool dead = false;
  long long step = 1;

  // ---- interior constants: rate constants and diffusivities pre-scaled by dt ----
  const double kS2f_t = k.kS2f * dt, kS2r_t = k.kS2r * dt, kG1f_t = k.kG1f * dt, kG1r_t = k.kG1r * dt,
               kG1p_t = k.kG1p * dt, kG1dp_t = k.kG1dp * dt, kSi_t = k.kSi * dt;
  const double Dt_Si = D_Si * dt, Dt_Sa = D_Sa * dt, Dt_G1 = Dv[4] * dt, Dt_G2 = Dv[1] * dt, Dt_G2G1 = Dv[2] * dt,
               Dt_S2 = Dv[6] * dt, Dt_G1S2 = Dv[5] * dt, Dt_G2G1S2 = Dv[3] * dt;
  // stencil of node Nr-1 (warp B): the expressions of the Grid set-up in solve_kernel
  double cpN, cmN, c0N;
  {
    const double dr = a.o.dr;
    const double inv_dr2 = 1.0 / (dr * dr);
    const double rN = a.r[Nr - 1];
    const double aj = (a.o.geometry == GAB1_GEOM_SPHERICAL) ? 1.0 / (rN * dr) : 0.0;
    cpN = inv_dr2 + aj; cmN = inv_dr2 - aj; c0N = -2.0 * inv_dr2;
  }

  // ---- membrane block (warp B): lane roles of solver_kernel.cuh ----
  constexpr int LZ = 31, LE = ML + NMB;
  double kf = 0.0, kr = 0.0, Dq = 1.0;
  int src_num = LZ, src_den = LZ;
  switch (lane) {
    case iSFK:   kf = k.kSa;  Dq = D_Si; src_den = LE; break;
    case aSFK:   kf = k.kSa;  Dq = D_Si; src_num = LE; src_den = LE; break;
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
  const double ca = k.kSa * (a.o.dr / D_Sa);
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
    case mE:     alpha = k.kEGFf * k.EGF; beta = k.kEGFr; s_own = -1.0; break;
    case mES:    alpha2 = k.kdf;          beta = k.kdr;   s_own = -2.0; s_src = 1.0; f_src = ML + mE; break;
    case mESmES: alpha = kp_now;          beta = k.kdp;   s_own = -1.0; s_src = 1.0; f_src = ML + mES; break;
    case E:      s_src = 1.0; f_src = ML + mESmES; break;
    case NMB:    s_src = 2.0; f_src = ML + mESmES; break;
    default: break;
  }
  const double tol = a.o.tol;
  const bool untracked = lane >= LE;
  const int iq_idx = lane < NCY ? lane : 10;
  const bool pulse = a.o.t_prechase >= 0.0;
  const int maxiters = a.o.maxiters;
  const unsigned ws_s = (unsigned)__cvta_generic_to_shared(ws);
  const unsigned dx_s = (unsigned)__cvta_generic_to_shared(dx);
  const unsigned bufB_s = dx_s, bufA_s = dx_s + 8u * 32u, stage_s = dx_s + 8u * 64u, mem_s = dx_s + 8u * 80u;
  double x = (lane == ML + mE) ? CoEGFR : 0.0;
  bool prev_unconv = false;           // warp B: the fixed point of the previous step did not converge / met a NaN

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
  int countdown = plan();
  double m[NMB];
#pragma unroll
  for (int j = 0; j < NMB; ++j) m[j] = 0.0;
  long long bc_B = 0;
  unsigned status_B = 0;

  if (Nt >= 1) {
    // ---- step 0: post the initial values and meet once ----
    if (roleA) {
      if (lane == lane_i2) {
#pragma unroll
        for (int q = 0; q < NCY; ++q) sts(bufA_s + 8u * q, u[q][idx_i2]);
      }
      __syncwarp();
      duo_bar_sync(bar0 + DUO_BAR_B);
      duo_bar_arrive(bar0 + DUO_BAR_A);
    } else {
      sts(bufB_s + 8u * lane, 0.0);                     // both u[Nr-1] buffers: slots 10..15 zero
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NCY; ++q) { sts(bufB_s + 8u * q, c[q]); sts(stage_s + 8u * q, c[q]); }
      }
      __syncwarp();
      duo_bar_arrive(bar0 + DUO_BAR_B);
    }

    // the time loop, instantiated once per role so that each warp's live registers are its own role's only
    auto time_loop = [&](auto role_tag) {
    constexpr bool RA = decltype(role_tag)::value;
    for (;;) {
      bool flag;
      const unsigned par = ((unsigned)step & 1u) * 128u;       // byte offset of this step's exchange buffers
      if constexpr (RA) {
        // ================================================================================ warp A: interior
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
        double L[2][NCY];
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
          const double v1 = fma(gb, G1, -(kG1r_t * g2g1));
          const double v3 = fma(gb, pG1, -(kG1r_t * g2pg1));
          const double v5 = fma(gb, pg1s, -(kG1r_t * g2pg1s));
          const double v2 = fma(ph, G1, -(kG1dp_t * pG1));
          const double v6 = fma(ph, g2g1, -(kG1dp_t * g2pg1));
          const double v4 = fma(sb, pG1, -(kS2r_t * pg1s));
          const double v7 = fma(sb, g2pg1, -(kS2r_t * g2pg1s));
          u[iSFK][i] = fma(Dt_Si, lap[iSFK], fma(kSi_t, Sa, Si));
          u[aSFK][i] = fma(Dt_Sa, lap[aSFK], fma(-kSi_t, Sa, Sa));
          u[GAB1][i] = fma(Dt_G1, lap[GAB1], G1 - v1 - v2);
          u[pGAB1][i] = fma(Dt_G1, lap[pGAB1], pG1 - v3 + v2 - v4);
          u[GRB2][i] = fma(Dt_G2, lap[GRB2], G2 - v1 - v3 - v5);
          u[G2G1][i] = fma(Dt_G2G1, lap[G2G1], g2g1 + v1 - v6);
          u[G2PG1][i] = fma(Dt_G2G1, lap[G2PG1], g2pg1 + v3 + v6 - v7);
          u[SHP2][i] = fma(Dt_S2, lap[SHP2], S2 - v4 - v7);
          u[PG1S][i] = fma(Dt_G1S2, lap[PG1S], pg1s + v4 - v5);
          u[G2PG1S][i] = fma(Dt_G2G1S2, lap[G2PG1S], g2pg1s + v5 + v7);
          if (i == idx_i2) {
            // u+[Nr-2] for warp B as soon as it exists
            if (lane == lane_i2) {
#pragma unroll
              for (int q = 0; q < NCY; ++q) sts(bufA_s + par + 8u * q, u[q][i]);
            }
          }
        }
        __syncwarp();
        duo_bar_sync(bar0 + DUO_BAR_B);               // B's u+[Nr-1] of this step is posted (early in B's step: normally no wait)
        duo_bar_arrive(bar0 + DUO_BAR_A);             // my u+[Nr-2] is posted
        if (lane == lane_i) {
#pragma unroll
          for (int q = 0; q < NCY; ++q) u[q][idx_i] = lds(bufB_s + par + 8u * q);
        }
        flag = lds(bufB_s + par + 8u * 15u) != 0.0;
      } else {
        // ================================================================================ warp B: node Nr-1 + membrane
        // ---- membrane block prologue: everything that depends only on old-time values ----
        const double m_old = x;
        const double m_next = shfl_down1(m_old);
        const double f = fma(m_old, fma(alpha2, m_old, alpha), -(beta * m_next));
        const double base = fma(dt, fma(s_own, f, s_src * shfl(f, f_src)), m_old);
        const double Md1 = shfl(m_old, src_den), Mn1 = shfl(m_old, src_num);
        const double A_t = kf_t * Md1;
        const double B_t = kr_t * Mn1;
        const double rden1 = fast_recip(fma(cf, Md1, 1.0));

        // ---- node Nr-1: the interior update of solver_kernel.cuh on broadcast operands ----
        duo_bar_sync(bar0 + DUO_BAR_A);               // A's u[Nr-2] of the previous step is posted
        {
          const unsigned pa = bufA_s + (par ^ 128u);
          double lap[NCY];
#pragma unroll
          for (int q = 0; q < NCY; ++q) lap[q] = fma(cpN, lds(stage_s + 8u * q), fma(cmN, lds(pa + 8u * q), c0N * c[q]));
          const double Si = c[iSFK], Sa = c[aSFK], G1 = c[GAB1], pG1 = c[pGAB1], G2 = c[GRB2],
                       g2g1 = c[G2G1], g2pg1 = c[G2PG1], S2 = c[SHP2], pg1s = c[PG1S], g2pg1s = c[G2PG1S];
          const double gb = kG1f_t * G2, ph = kG1p_t * Sa, sb = kS2f_t * S2;
          const double v1 = fma(gb, G1, -(kG1r_t * g2g1));
          const double v3 = fma(gb, pG1, -(kG1r_t * g2pg1));
          const double v5 = fma(gb, pg1s, -(kG1r_t * g2pg1s));
          const double v2 = fma(ph, G1, -(kG1dp_t * pG1));
          const double v6 = fma(ph, g2g1, -(kG1dp_t * g2pg1));
          const double v4 = fma(sb, pG1, -(kS2r_t * pg1s));
          const double v7 = fma(sb, g2pg1, -(kS2r_t * g2pg1s));
          c[iSFK] = fma(Dt_Si, lap[iSFK], fma(kSi_t, Sa, Si));
          c[aSFK] = fma(Dt_Sa, lap[aSFK], fma(-kSi_t, Sa, Sa));
          c[GAB1] = fma(Dt_G1, lap[GAB1], G1 - v1 - v2);
          c[pGAB1] = fma(Dt_G1, lap[pGAB1], pG1 - v3 + v2 - v4);
          c[GRB2] = fma(Dt_G2, lap[GRB2], G2 - v1 - v3 - v5);
          c[G2G1] = fma(Dt_G2G1, lap[G2G1], g2g1 + v1 - v6);
          c[G2PG1] = fma(Dt_G2G1, lap[G2PG1], g2pg1 + v3 + v6 - v7);
          c[SHP2] = fma(Dt_S2, lap[SHP2], S2 - v4 - v7);
          c[PG1S] = fma(Dt_G1S2, lap[PG1S], pg1s + v4 - v5);
          c[G2PG1S] = fma(Dt_G2G1S2, lap[G2PG1S], g2pg1s + v5 + v7);
        }
        const unsigned pb = bufB_s + par;
        if (lane == 0) {
#pragma unroll
          for (int q = 0; q < NCY; ++q) sts(pb + 8u * q, c[q]);
          sts(pb + 8u * 15u, prev_unconv ? 1.0 : 0.0);
        }
        __syncwarp();
        duo_bar_arrive(bar0 + DUO_BAR_B);             // u+[Nr-1] is posted
        const double Iq = lds(pb + 8u * iq_idx);
        const double cr = lane == aSFK ? fma(cf, Iq, ca * lds(pb + 8u * iSFK)) : cr_fixed;

        // ---- fixed-point iterations (basepdesolver.jl:197-242); the first pass is peeled ----
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
        if (lane < NCY) sts(stage_s + 8u * lane, x);    // boundary values: the stencil of node Nr-1 at the next step
        __syncwarp();
        flag = prev_unconv;
        prev_unconv = unconverged || nan_exit;
      }
      t = t + dt;                                                   // basepdesolver.jl:265
      // a fixed point that failed at step n-1 sends BOTH warps through the event path after step n, where the
      // all-NaN test of the throughput kernel is made on the whole state (one step later than there: a state that is
      // all NaN steps exactly like the fast-forward loop counts, so the results are the same)
      if (flag) countdown = 1;
      if (--countdown > 0) { ++step; continue; }

      // ---- rare path: the CTA meets; warp A holds the complete state and does the throughput kernel's event work ----
      if constexpr (!RA) {
        if (lane >= ML && lane < LE) sts(mem_s + 8u * (lane - ML), x);
        if (lane == 0) { *s_bc = bc_total; s_flags[1] = (int)status; }
      }
      duo_bar_sync(bar0 + DUO_BAR_ALL);
      if constexpr (RA) {
        if (lane == lane_b) {
#pragma unroll
          for (int q = 0; q < NCY; ++q) u[q][idx_b] = lds(stage_s + 8u * q);
        }
#pragma unroll
        for (int j = 0; j < NMB; ++j) m[j] = lds(mem_s + 8u * j);
        bc_B = *s_bc;
        status_B = (unsigned)s_flags[1];
        bool all_nan = true;
#pragma unroll
        for (int j = 0; j < NMB; ++j) all_nan &= isnan(m[j]);
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
          for (int q = 0; q < NCY; ++q) all_nan &= !(g.node[i] >= 1 && g.node[i] <= Nr) || isnan(u[q][i]);
        const bool d = __all_sync(FULL, all_nan);
        if (lane == 0) s_flags[0] = d ? 1 : 0;
      }
      if (track_t) {
        const bool save = a.o.save_rule == GAB1_SAVE_T_GE_TSAVE ? (t >= t_save) : (fmod((double)step, modulus_step) == 0.0);
        if (save) {
          if (nts >= Cn) status |= GAB1_ST_OVERFLOW;
          else {
            const int cidx = nts++;
            if constexpr (RA) {
              if (a.o.out_mode == GAB1_OUT_FULL) write_full_column<K>(a, oset, cidx, u, m, t, CoEGFR, lane, g, rowA, status);
              else if (cidx == Cn - 1) {
                stage_row<K>(rowA, lane, g, Nr, [&](int i) { return derived_stot<K>(u, i); });
                pct_ave = trapz_r2(a.r, rowA, P);
                pct_memb = m[EG2PG1S];
                __syncwarp();
              }
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
      duo_bar_sync(bar0 + DUO_BAR_ALL);
      dead = s_flags[0] != 0;
      if (dead || step > Nt) break;
      countdown = plan();
    }
    };
    if (roleA) time_loop(std::true_type{});
    else {
      time_loop(std::false_type{});
      duo_bar_sync(bar0 + DUO_BAR_A);                 // consume warp A's last post: both barriers are idle again
    }
  }
  if (!roleA) return;

  // ================================================================================ warp A: the throughput kernel's epilogue
  status |= status_B & GAB1_ST_ITER_CAP;
  bc_total = bc_B;
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
        const int cidx = nts++;
        if (a.o.out_mode == GAB1_OUT_FULL) write_full_column<K>(a, oset, cidx, u, m, t, CoEGFR, lane, g, rowA, status);
        else if (cidx == Cn - 1) { pct_ave = CUDART_NAN; pct_memb = CUDART_NAN; }
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
      long long off = 0;
      for (int mi = 0; mi < 12; ++mi) {
        if (!((a.o.matrix_mask >> mi) & 1u)) continue;
        for (long long i = (long long)nts * P + lane; i < (long long)Cn * P; i += 32) oset[off + i] = 0.0;
        off += (long long)P * Cn;
      }
      for (int v = 0; v < GAB1_N_VECTORS; ++v)
        for (int cc = nts + lane; cc < Cn; cc += 32) oset[off + (long long)v * Cn + cc] = 0.0;
    }
  }
  if (lane == 0) {
    if (a.status) a.status[set] = (int)status;
    if (a.n_saved) a.n_saved[set] = track_t ? nts : 0;
    if (a.n_steps) a.n_steps[set] = Nt;
    if (a.n_bc) a.n_bc[set] = bc_total;
  }
}

// The product kernel of the register-resident grids with the latency lane in front: CTAs of four warps = two pairs.  Each
// pair first serves the head of the descending-work queue — items [0, *dyn_count), two warps per set — and when that is
// empty its warps join the one-set-per-warp queue (solve_set of solver_kernel.cuh), which starts at item *dyn_count.  One
// launch, so the long sets are resident from the first cycle whatever the block scheduler does (a separate latency
// kernel on a second stream was measured first: launched behind the persistent throughput grid it found no free
// registers until that grid had drained, and the two ran back to back).
template <int K, int MODE>
__global__ void __launch_bounds__(128, 2)
duo_solve_kernel(const KernelArgs a) {
  if (a.guard && *a.guard != a.guard_expect) return;      // the plan sent nothing to the latency lane: solve_kernel runs instead
  extern __shared__ __align__(16) double smem[];
  __shared__ unsigned s_item[2];
  __shared__ int s_flags[2][2];       // per pair: [0] the state is all NaN, [1] warp B's status bits
  __shared__ long long s_bc[2];       // per pair: warp B's membrane iteration count
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pair = warp >> 1;
  const bool roleA = (warp & 1) == 0;
  const int per_warp = WS_HDR + 2 * a.P_pad + WS_EX;
  double* ws = smem + (size_t)warp * per_warp;
  double* dx = smem + (size_t)4 * per_warp + (size_t)pair * DUO_DX;
  const int bar0 = 1 + 3 * pair;
  const int Nr = a.o.Nr;
  ws[lane] = 0.0;
  for (int i = threadIdx.x & 63; i < DUO_DX; i += 64) dx[i] = 0.0;
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
      g.a[i] = __ddiv_rn(1.0, __dmul_rn(r, dr));
      const double aj = (a.o.geometry == GAB1_GEOM_SPHERICAL) ? 1.0 / (r * dr) : 0.0;
      double cp = inv_dr2 + aj, cm = inv_dr2 - aj, c0 = -2.0 * inv_dr2;
      if (n == 1) { c0 += cm; cm = 0.0; }
      g.cp[i] = g.interior[i] ? cp : 0.0;
      g.cm[i] = g.interior[i] ? cm : 0.0;
      g.c0[i] = g.interior[i] ? c0 : 0.0;
    }
  }
  // ---- isolation (a.isolate, full batches only): a warp of the latency lane keeps the warps it would share issue slots
  //      with from taking new sets while its set runs.  Measured on B200: the longest set of the bench ensemble takes 248 ms
  //      alone on the GPU and 312 ms next to a second resident warp on its scheduler — and bounds the 8-GPU run.
  //      1: the whole SM is reserved; 2: the scheduler (%warpid mod 4) only.  Reservations are counters in global memory.
  int* resv = nullptr;
  if (a.isolate) {
    unsigned smid, wid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    resv = a.sm_resv + 4 * smid + (a.isolate == 2 ? (wid & 3u) : 0u);
  }
#ifdef GAB1_DUO_TIMING
  unsigned long long tl0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl0));
  atomicMin((unsigned long long*)(a.duo_counter + 3), tl0);      // counter words 8..9: earliest start (memset to 0xff by the host)
  int tl_sets = 0;
#endif
  // ---- latency lane: two warps per set ----
  const unsigned n_duo = *a.dyn_count;
  if (n_duo) {
    for (;;) {
      duo_bar_sync(bar0 + DUO_BAR_ALL);
      if (roleA && lane == 0) s_item[pair] = atomicAdd(a.duo_counter, 1u);
      duo_bar_sync(bar0 + DUO_BAR_ALL);
      const unsigned item = s_item[pair];
      if (item >= n_duo) break;
      const long long set = a.order ? (long long)a.order[item] : (long long)item;
      if (resv && lane == 0) {
        atomicAdd(resv, 1);
        __threadfence();
        if (roleA) atomicAdd(a.duo_counter + 1, 1u);      // this set's reservations are visible
      }
      duo_solve_set<K, MODE>(a, set, lane, roleA, ws, dx, g, s_flags[pair], &s_bc[pair], bar0);
      if (resv && lane == 0) atomicSub(resv, 1);
#ifdef GAB1_DUO_TIMING
      if (lane == 0 && roleA) {
        unsigned long long tl1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl1));
        printf("timeline duo item %u set %lld done at %.3f ms\n", item, set, (double)(tl1 - *(volatile unsigned long long*)(a.duo_counter + 3)) * 1e-6);
      }
#endif
    }
    if (roleA) ws[lane] = 0.0;        // the throughput kernel's header starts clean
    __syncwarp();
  }
  // ---- throughput lane: one warp per set; its queue starts behind the latency lane's sets ----
  if (resv && n_duo) {
    // the queue is in descending-work order, so the first set a warp takes is among the longest of the batch: no warp
    // starts one before every set of the latency lane has been claimed and its reservation is visible (microseconds)
    while (*(volatile unsigned*)(a.duo_counter + 1) < n_duo) __nanosleep(200);
  }
  for (;;) {
    if (resv) {                       // sleep while a set of the latency lane has this SM / scheduler reserved
      while (*(volatile int*)resv > 0) __nanosleep(4000);
    }
    unsigned item = 0;
    if (lane == 0) item = atomicAdd(a.counter, 1u);
    item = __shfl_sync(FULL, item, 0);
    if ((long long)item >= a.S) break;
    const long long set = a.order ? (long long)a.order[item] : (long long)item;
#ifdef GAB1_DUO_TIMING
    unsigned long long tls;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tls));
#endif
    solve_set<K, MODE>(a, set, lane, ws, g);
    __syncwarp();
#ifdef GAB1_DUO_TIMING
    ++tl_sets;
    if (lane == 0) {
      unsigned long long tle;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tle));
      const unsigned long long t00 = *(volatile unsigned long long*)(a.duo_counter + 3);
      if (tle - tls > 40000000ull || tle - t00 > 255000000ull)
        printf("timeline set %lld item %u start %.3f end %.3f ms\n", set, item, (double)(tls - t00) * 1e-6, (double)(tle - t00) * 1e-6);
    }
#endif
  }
#ifdef GAB1_DUO_TIMING
  if (lane == 0) {
    unsigned long long tl1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl1));
    printf("timeline warp exit %.3f ms sets %d\n", (double)(tl1 - *(volatile unsigned long long*)(a.duo_counter + 3)) * 1e-6, tl_sets);
  }
#endif
}

}  // namespace gab1
