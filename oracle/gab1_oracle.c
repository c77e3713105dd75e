/*
 * gab1_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * A plain-C, scalar, operation-for-operation restatement of the reference's explicit
 * finite-difference solvers, used ONLY as the parity checker (tests/, __graft_entry__.smoke())
 * and as the timed CPU baseline (bench.py cpu_baseline / --impl reference).  Nothing in the
 * shipped library links, imports or calls this file.
 *
 * PARITY PIN.  The reference ships no tests and no bit-level golden vectors, and no Julia runtime exists in the build
 * image or on the GPU box, so this restatement cannot be compared with the reference solver output value by value.
 * What the reference tree does hold is output COMPUTED BY ITS SOLVER: the eFAST sensitivity indices of
 * `GSA results/eFAST-GSA-res_concs_1000-spls-per-param_{S1,ST}.csv` (+ the `_memb-SFKs` pair), produced by
 * `gsa(fbatch_concs_mt, eFAST(), pbounds; samples=1000, batch=true)` (GSA_concs.jl:50-97) through sapdesolver /
 * sapdesolver_membSFK — functionals of 5000 full-length solves each.  This oracle reproduces them within the spread over
 * the design's random phases, including which outputs are exactly constant (tests/test_efast_pin.py; eFAST itself is
 * restated in oracle/efast.py from GlobalSensitivity 2.1.3).  That is a coarse pin (1e-2, not 1e-9): at the 1e-9 level
 * the oracle is additionally held by (i) a second independent NumPy restatement that must agree bit for bit
 * (oracle/numpy_oracle.py, tests/test_oracle_cross.py), (ii) exact discrete invariants of the scheme, (iii) the analytic
 * steady aSFK profile and the experimental % SHP2-bound GAB1 the reference was fitted to (tests/test_oracle_physics.py).
 *
 * What it follows (all under the reference's Julia/ directory):
 *   pdesolver                basepdesolver.jl:25-312
 *   pdesolver_membSFK        basepdesolver.jl:350-636      (D_Sa = 1e-32 at :366,477,530)
 *   pdesolver_rect           basepdesolver_rect.jl:23-294
 *   pdesolver_membSFK_rect   basepdesolver_rect.jl:298-569 (D_S = 1e-32 :305-306, modulus rule :336,526)
 *   sapdesolver              sapdesolver.jl:55-280
 *   sapdesolver_membSFK      sapdesolver_memb-SFK.jl:55-281 (`while error > tol` :175-177)
 *   pulsechase_solver        pulsechase_solver.jl:29-318    (kp switch :156-158)
 *   pmap_fun_dk reductions   sapdesolver.jl:343-356
 *   % SHP2-bound GAB1        run_base_model.jl:269-276
 *
 * Semantics mirrored from Julia: every `+ - * /` is one IEEE-754 binary64 operation in source order
 * (n-ary `a*b*c` and `a+b+c` fold left), no FMA contraction (build with -ffp-contract=off),
 * `x^2` is `x*x`, `maximum` propagates NaN, `NaN <= tol` is false, work arrays start at zero.
 * Third-party arithmetic on the path: NumericalIntegration v0.2.0 `integrate(x, y)` (Manifest.toml
 * :1811-1815; source not in the reference tree) restated as its published trapezoid rule
 * `0.5 * sum_i (x[i+1]-x[i])*(y[i]+y[i+1])` accumulated left to right.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "gab1pde.h"

enum { iSFK, aSFK, GAB1, pGAB1, GRB2, G2G1, G2PG1, SHP2, PG1S, G2PG1S }; /* basepdesolver.jl:199-202 */
enum { mE, mES, mESmES, E, EG2, EG2G1, EG2PG1, EG2PG1S };               /* basepdesolver.jl:203 */

/* cytosolic species behind each GAB1_OUT_FULL matrix 0..9 (basepdesolver.jl:271-280) */
static const int kMatrixSpecies[10] = { iSFK, aSFK, GRB2, GAB1, SHP2, G2G1, G2PG1, G2PG1S, pGAB1, PG1S };

static int64_t popcount12(uint32_t m) { int64_t c = 0; for (int i = 0; i < 12; ++i) c += (m >> i) & 1u; return c; }

int64_t gab1o_out_doubles_per_set(const gab1_opts* o) {
  const int64_t P = (int64_t)o->Nr + 1, C = (int64_t)o->Nts + 1;
  switch (o->out_mode) {
    case GAB1_OUT_FINAL4: return 4 * P;
    case GAB1_OUT_FULL: return popcount12(o->matrix_mask) * P * C + GAB1_N_VECTORS * C;
    case GAB1_OUT_SIX: return 6;
    case GAB1_OUT_PCT_BOUND: return 1;
    case GAB1_OUT_FINAL_STATE: return 10 * P + 8;
  }
  return 0;
}

static int64_t matrix_offset(const gab1_opts* o, int m) {
  if (!((o->matrix_mask >> m) & 1u)) return -1;
  const int64_t P = (int64_t)o->Nr + 1, C = (int64_t)o->Nts + 1;
  return popcount12(o->matrix_mask & ((1u << m) - 1u)) * P * C;
}
static int64_t vector_offset(const gab1_opts* o, int v) {
  const int64_t P = (int64_t)o->Nr + 1, C = (int64_t)o->Nts + 1;
  return popcount12(o->matrix_mask) * P * C + (int64_t)v * C;
}

/* NumericalIntegration.integrate(x, y), trapezoid, of y_j * x_j^2 scaled as the callers do. */
static double trapz_r2(const double* r, const double* y, int P) {
  double acc = 0.0;
  for (int i = 0; i + 1 < P; ++i) {
    const double yi = y[i] * (r[i] * r[i]);
    const double yn = y[i + 1] * (r[i + 1] * r[i + 1]);
    acc += (r[i + 1] - r[i]) * (yi + yn);
  }
  return 0.5 * acc;
}

/* `R - (r[y .>= f*maximum(y)] |> minimum)`; returns 0 and sets *threw when the selection is empty */
static double length_scale(const double* r, const double* y, int P, double f, double R, int* threw) {
  double mx = y[0]; int nan_seen = isnan(y[0]);
  for (int i = 1; i < P; ++i) { if (isnan(y[i])) nan_seen = 1; else if (!(mx >= y[i])) mx = y[i]; }
  if (nan_seen) mx = NAN;
  const double thr = f * mx;
  for (int i = 0; i < P; ++i) if (y[i] >= thr) return R - r[i];
  *threw = 1;
  return 0.0;
}

typedef struct { double* u[2][10]; double m[2][8]; } state_t;

/*
 * One parameter set.  Returns the status word.  `out` is the set's output block.
 */
static uint32_t solve_one(const gab1_opts* o, const double* Co, const double* D, const double* k,
                          double dt, const double* r, double* out,
                          int32_t* n_saved_out, int64_t* n_steps_out, int64_t* n_bc_out, double* work) {
  const int Nr = o->Nr, P = Nr + 1, Nts = o->Nts, C = Nts + 1;
  const double dr = o->dr, R = o->R, tf = o->tf, tol = o->tol;
  const int64_t nout = gab1o_out_doubles_per_set(o);
  uint32_t status = 0;
  memset(out, 0, (size_t)nout * sizeof(double));
  if (n_saved_out) *n_saved_out = 0;
  if (n_steps_out) *n_steps_out = 0;
  if (n_bc_out) *n_bc_out = 0;

  /* diffusivities (basepdesolver.jl:43-49; :366; basepdesolver_rect.jl:305-306) */
  double D_Si = D[0], D_Sa = D[0];
  if (o->sfk_mode == GAB1_SFK_MEMBRANE) D_Sa = 1e-32;
  if (o->sfk_mode == GAB1_SFK_BOTH_FROZEN) { D_Si = 1e-32; D_Sa = 1e-32; }
  const double D_G2 = D[1], D_G2G1 = D[2], D_G2G1S2 = D[3], D_G1 = D[4], D_G1S2 = D[5], D_S2 = D[6];
  const double Dsp[10] = { D_Si, D_Sa, D_G1, D_G1, D_G2, D_G2G1, D_G2G1, D_S2, D_G1S2, D_G2G1S2 };
  /* rate constants (basepdesolver.jl:52-68) */
  const double kS2f = k[0], kS2r = k[1], kG1f = k[2], kG1r = k[3], kG2f = k[4], kG2r = k[5],
               kG1p = k[6], kG1dp = k[7], kSa = k[8], kSi = k[9], kdp = k[11],
               kEGFf = k[12], kEGFr = k[13], EGF = k[14], kdf = k[15], kdr = k[16];
  double kp = k[10];
  const double CoSFK = Co[0], CoG2 = Co[1], CoG1 = Co[2], CoS2 = Co[3], CoEGFR = Co[4];

  /* Nt = Int64(ceil(tf/dt)) (basepdesolver.jl:72); Julia throws InexactError when not representable */
  const double nt_f = ceil(tf / dt);
  if (!(nt_f >= 0.0 && nt_f < 9.0e18)) return GAB1_ST_THROW;
  const int64_t Nt = (int64_t)nt_f;
  if (n_steps_out) *n_steps_out = Nt;

  state_t s;
  for (int c = 0; c < 2; ++c) for (int q = 0; q < 10; ++q) {
    s.u[c][q] = work + ((size_t)c * 10 + q) * P;
    for (int j = 0; j < P; ++j) s.u[c][q][j] = 0.0;
  }
  memset(s.m, 0, sizeof s.m);
  for (int j = 0; j < P; ++j) {           /* basepdesolver.jl:137-141 */
    s.u[0][iSFK][j] = CoSFK; s.u[0][GAB1][j] = CoG1; s.u[0][GRB2][j] = CoG2; s.u[0][SHP2][j] = CoS2;
  }
  s.m[0][mE] = CoEGFR;

  const int track_t = (o->out_mode == GAB1_OUT_FULL || o->out_mode == GAB1_OUT_PCT_BOUND);
  double* Mx[12]; double* Vx[GAB1_N_VECTORS];
  for (int m = 0; m < 12; ++m) Mx[m] = NULL;
  for (int v = 0; v < GAB1_N_VECTORS; ++v) Vx[v] = NULL;
  double* last_col = work + (size_t)20 * P;      /* PCT_BOUND: PG1S+G2PG1S of column Nts+1 (zeros if never saved) */
  double last_EG2PG1S = 0.0;
  for (int j = 0; j < P; ++j) last_col[j] = 0.0;
  if (o->out_mode == GAB1_OUT_FULL) {
    for (int m = 0; m < 12; ++m) { const int64_t off = matrix_offset(o, m); if (off >= 0) Mx[m] = out + off; }
    for (int v = 0; v < GAB1_N_VECTORS; ++v) Vx[v] = out + vector_offset(o, v);
    /* column 1 = initial state (basepdesolver.jl:94-97,111) */
    for (int j = 0; j < P; ++j) {
      if (Mx[GAB1_M_iSFK]) Mx[GAB1_M_iSFK][j] = CoSFK;
      if (Mx[GAB1_M_GRB2]) Mx[GAB1_M_GRB2][j] = CoG2;
      if (Mx[GAB1_M_SHP2]) Mx[GAB1_M_SHP2][j] = CoS2;
      if (Mx[GAB1_M_GAB1]) Mx[GAB1_M_GAB1][j] = CoG1;
    }
    Vx[GAB1_V_mE][0] = CoEGFR;
  }

  double t = 0.0, t_save = o->dt_save;
  int nts = 1;
  const double modulus_step = (o->save_rule == GAB1_SAVE_MODULUS) ? rint((double)Nt / (double)Nts) : 0.0;
  const double dr2 = dr * dr;
  int64_t bc_total = 0;
  const int cap = o->maxiters;

  double** u0 = s.u[0]; double** u1 = s.u[1];
  double* m0 = s.m[0]; double* m1 = s.m[1];

  for (int64_t step = 1; step <= Nt; ++step) {
    if (o->t_prechase >= 0.0) {            /* pulsechase_solver.jl:156-158 */
      if (o->t_prechase + dt > t && t >= o->t_prechase) kp = 0.0;
    }
    /* interior nodes j = 2..Nr (1-based) (basepdesolver.jl:150-180 / basepdesolver_rect.jl:131-161) */
    for (int j = 1; j < Nr; ++j) {
      double L[10];
      if (o->geometry == GAB1_GEOM_SPHERICAL) {
        const double a = 1 / (r[j] * dr);
        for (int q = 0; q < 10; ++q) {
          const double up = u0[q][j + 1], uc = u0[q][j], um = u0[q][j - 1];
          L[q] = Dsp[q] * (a * (up - um) + (up - 2.0 * uc + um) / dr2);
        }
      } else {
        for (int q = 0; q < 10; ++q) {
          const double up = u0[q][j + 1], uc = u0[q][j], um = u0[q][j - 1];
          L[q] = Dsp[q] * (up - 2.0 * uc + um) / dr2;
        }
      }
      const double Si = u0[iSFK][j], Sa = u0[aSFK][j], G1 = u0[GAB1][j], pG1 = u0[pGAB1][j], G2 = u0[GRB2][j],
                   g2g1 = u0[G2G1][j], g2pg1 = u0[G2PG1][j], S2 = u0[SHP2][j], pg1s = u0[PG1S][j], g2pg1s = u0[G2PG1S][j];
      u1[iSFK][j] = (L[iSFK] + kSi * Sa) * dt + Si;
      u1[aSFK][j] = (L[aSFK] - kSi * Sa) * dt + Sa;
      u1[GAB1][j] = (L[GAB1] - kG1f * G1 * G2 + kG1r * g2g1 - kG1p * Sa * G1 + kG1dp * pG1) * dt + G1;
      u1[pGAB1][j] = (L[pGAB1] - kG1f * pG1 * G2 + kG1r * g2pg1 + kG1p * Sa * G1 - kG1dp * pG1 - kS2f * S2 * pG1 + kS2r * pg1s) * dt + pG1;
      u1[GRB2][j] = (L[GRB2] - kG1f * G1 * G2 + kG1r * g2g1 - kG1f * pG1 * G2 + kG1r * g2pg1 - kG1f * G2 * pg1s + kG1r * g2pg1s) * dt + G2;
      u1[G2G1][j] = (L[G2G1] + kG1f * G1 * G2 - kG1r * g2g1 - kG1p * Sa * g2g1 + kG1dp * g2pg1) * dt + g2g1;
      u1[G2PG1][j] = (L[G2PG1] + kG1f * pG1 * G2 - kG1r * g2pg1 + kG1p * Sa * g2g1 - kG1dp * g2pg1 - kS2f * S2 * g2pg1 + kS2r * g2pg1s) * dt + g2pg1;
      u1[SHP2][j] = (L[SHP2] - kS2f * S2 * pG1 + kS2r * pg1s - kS2f * S2 * g2pg1 + kS2r * g2pg1s) * dt + S2;
      u1[PG1S][j] = (L[PG1S] + kS2f * S2 * pG1 - kS2r * pg1s - kG1f * G2 * pg1s + kG1r * g2pg1s) * dt + pg1s;
      u1[G2PG1S][j] = (L[G2PG1S] + kG1f * G2 * pg1s - kG1r * g2pg1s + kS2f * S2 * g2pg1 - kS2r * g2pg1s) * dt + g2pg1s;
    }
    /* zero flux at r = 0 (basepdesolver.jl:183-192) */
    for (int q = 0; q < 10; ++q) u1[q][0] = u1[q][1];

    /* reactive flux at r = R: fixed point over the Robin closures and membrane Euler step
       (basepdesolver.jl:197-242; while-form sapdesolver_memb-SFK.jl:175-222) */
    double err = tol * 2.;
    int it = 0;
    for (;;) {
      if (o->bc_loop == GAB1_BC_FOR_BREAK) { if (it >= o->maxiters) break; }
      else { if (!(err > tol)) break; if (it >= cap) { status |= GAB1_ST_ITER_CAP; break; } }
      ++it;
      double cold[10], mold[8];
      for (int q = 0; q < 10; ++q) cold[q] = u1[q][Nr];
      for (int q = 0; q < 8; ++q) mold[q] = m1[q];

      const double Etot = 2.0 * (m1[E] + m1[EG2] + m1[EG2G1] + m1[EG2PG1] + m1[EG2PG1S]);
      u1[iSFK][Nr] = u1[iSFK][Nr - 1] / (1 + kSa * Etot * dr / D_Si);
      u1[aSFK][Nr] = u1[aSFK][Nr - 1] + kSa * u1[iSFK][Nr] * Etot * dr / D_Sa;
      u1[GAB1][Nr] = (kG1r * m1[EG2G1] * dr / D_G1 + u1[GAB1][Nr - 1]) / (1 + kG1f * m1[EG2] * dr / D_G1);
      u1[pGAB1][Nr] = (kG1r * m1[EG2PG1] * dr / D_G1 + u1[pGAB1][Nr - 1]) / (1 + kG1f * m1[EG2] * dr / D_G1);
      u1[GRB2][Nr] = (kG2r * m1[EG2] * dr / D_G2 + u1[GRB2][Nr - 1]) / (1 + kG2f * m1[E] * dr / D_G2);
      u1[G2G1][Nr] = (kG2r * m1[EG2G1] * dr / D_G2G1 + u1[G2G1][Nr - 1]) / (1 + kG2f * m1[E] * dr / D_G2G1);
      u1[G2PG1][Nr] = (kG2r * m1[EG2PG1] * dr / D_G2G1 + u1[G2PG1][Nr - 1]) / (1 + kG2f * m1[E] * dr / D_G2G1);
      u1[SHP2][Nr] = (kS2r * m1[EG2PG1S] * dr / D_S2 + u1[SHP2][Nr - 1]) / (1 + kS2f * m1[EG2PG1] * dr / D_S2);
      u1[PG1S][Nr] = (kG1r * m1[EG2PG1S] * dr / D_G1S2 + u1[PG1S][Nr - 1]) / (1 + kG1f * m1[EG2] * dr / D_G1S2);
      u1[G2PG1S][Nr] = (kG2r * m1[EG2PG1S] * dr / D_G2G1S2 + u1[G2PG1S][Nr - 1]) / (1 + kG2f * m1[E] * dr / D_G2G1S2);

      const double bG1 = u1[GAB1][Nr], bpG1 = u1[pGAB1][Nr], bG2 = u1[GRB2][Nr], bg2g1 = u1[G2G1][Nr],
                   bg2pg1 = u1[G2PG1][Nr], bS2 = u1[SHP2][Nr], bpg1s = u1[PG1S][Nr], bg2pg1s = u1[G2PG1S][Nr];
      m1[mE] = (-kEGFf * EGF * m0[mE] + kEGFr * m0[mES]) * dt + m0[mE];
      m1[mES] = (kEGFf * EGF * m0[mE] - kEGFr * m0[mES] - 2 * kdf * m0[mES] * m0[mES] + 2 * kdr * m0[mESmES]) * dt + m0[mES];
      m1[mESmES] = (kdf * m0[mES] * m0[mES] - kdr * m0[mESmES] - kp * m0[mESmES] + kdp * m0[E]) * dt + m0[mESmES];
      m1[E] = (kp * m0[mESmES] - kdp * m0[E] - kG2f * m0[E] * bG2 + kG2r * m0[EG2] - kG2f * m0[E] * bg2g1 + kG2r * m0[EG2G1]
               - kG2f * m0[E] * bg2pg1 + kG2r * m0[EG2PG1] - kG2f * m0[E] * bg2pg1s + kG2r * m0[EG2PG1S]) * dt + m0[E];
      m1[EG2] = (kG2f * bG2 * m0[E] - kG2r * m0[EG2] - kG1f * bG1 * m0[EG2] + kG1r * m0[EG2G1] - kG1f * bpG1 * m0[EG2]
                 + kG1r * m0[EG2PG1] - kG1f * bpg1s * m0[EG2] + kG1r * m0[EG2PG1S]) * dt + m0[EG2];
      m1[EG2G1] = (kG2f * bg2g1 * m0[E] - kG2r * m0[EG2G1] + kG1f * bG1 * m0[EG2] - kG1r * m0[EG2G1]) * dt + m0[EG2G1];
      m1[EG2PG1] = (kG2f * bg2pg1 * m0[E] - kG2r * m0[EG2PG1] + kG1f * bpG1 * m0[EG2] - kG1r * m0[EG2PG1]
                    - kS2f * bS2 * m0[EG2PG1] + kS2r * m0[EG2PG1S]) * dt + m0[EG2PG1];
      m1[EG2PG1S] = (kS2f * bS2 * m0[EG2PG1] - kS2r * m0[EG2PG1S] + kG1f * bpg1s * m0[EG2] - kG1r * m0[EG2PG1S]
                     + kG2f * bg2pg1s * m0[E] - kG2r * m0[EG2PG1S]) * dt + m0[EG2PG1S];

      /* error = maximum(abs.(1 .- new./old)) over 18 values; Julia's maximum propagates NaN */
      double mx = -INFINITY; int nan_seen = 0;
      for (int q = 0; q < 10; ++q) { const double e = fabs(1.0 - u1[q][Nr] / cold[q]); if (isnan(e)) nan_seen = 1; else if (e > mx) mx = e; }
      for (int q = 0; q < 8; ++q) { const double e = fabs(1.0 - m1[q] / mold[q]); if (isnan(e)) nan_seen = 1; else if (e > mx) mx = e; }
      err = nan_seen ? NAN : mx;
      if (o->bc_loop == GAB1_BC_FOR_BREAK && err <= tol) break;
    }
    bc_total += it;

    /* advance: column 2 -> column 1 (basepdesolver.jl:245-262); column 2 keeps its values */
    for (int q = 0; q < 10; ++q) memcpy(u0[q], u1[q], (size_t)P * sizeof(double));
    for (int q = 0; q < 8; ++q) m0[q] = m1[q];

    if (track_t) {
      const double Etot = 2.0 * (m1[E] + m1[EG2] + m1[EG2G1] + m1[EG2PG1] + m1[EG2PG1S]);   /* basepdesolver.jl:263 */
      t += dt;
      int save;
      if (o->save_rule == GAB1_SAVE_T_GE_TSAVE) save = (t >= t_save);
      else save = (fmod((double)step, modulus_step) == 0.0);
      if (save) {
        if (nts >= C) {
          status |= GAB1_ST_OVERFLOW;
        } else {
          const int c = nts;      /* 0-based column index of the new snapshot */
          nts += 1;
          if (o->out_mode == GAB1_OUT_FULL) {
            for (int m = 0; m < 10; ++m) if (Mx[m]) memcpy(Mx[m] + (size_t)c * P, u1[kMatrixSpecies[m]], (size_t)P * sizeof(double));
            for (int j = 0; j < P; ++j) {
              const double stot = u1[PG1S][j] + u1[G2PG1S][j];
              double ptot;
              if (o->pg1tot_form == GAB1_PG1TOT_VIA_STOT) ptot = u1[G2PG1][j] + u1[pGAB1][j] + stot;
              else ptot = u1[G2PG1][j] + u1[pGAB1][j] + u1[PG1S][j] + u1[G2PG1S][j];
              if (Mx[GAB1_M_PG1Stot]) Mx[GAB1_M_PG1Stot][(size_t)c * P + j] = stot;
              if (Mx[GAB1_M_PG1tot]) Mx[GAB1_M_PG1tot][(size_t)c * P + j] = ptot;
              if (isnan(u1[PG1S][j])) status |= GAB1_ST_NAN;
            }
            Vx[GAB1_V_pE][c] = Etot * 100.0 / CoEGFR;
            Vx[GAB1_V_mE][c] = m1[mE]; Vx[GAB1_V_mES][c] = m1[mES]; Vx[GAB1_V_mESmES][c] = m1[mESmES];
            Vx[GAB1_V_E][c] = m1[E]; Vx[GAB1_V_EG2][c] = m1[EG2]; Vx[GAB1_V_EG2G1][c] = m1[EG2G1];
            Vx[GAB1_V_EG2PG1][c] = m1[EG2PG1]; Vx[GAB1_V_EG2PG1S][c] = m1[EG2PG1S];
            Vx[GAB1_V_EGFR_SHP2][c] = m1[EG2PG1S] * 100.0 / CoEGFR;
            Vx[GAB1_V_t_out][c] = t;
          } else if (c == C - 1) {
            for (int j = 0; j < P; ++j) last_col[j] = u1[PG1S][j] + u1[G2PG1S][j];
            last_EG2PG1S = m1[EG2PG1S];
          }
        }
        if (o->save_rule == GAB1_SAVE_T_GE_TSAVE) t_save += o->dt_save;
      }
    }
  }
  if (n_bc_out) *n_bc_out = bc_total;
  if (n_saved_out) *n_saved_out = track_t ? nts : 0;
  if (track_t && nts < C) status |= GAB1_ST_SHORT;

  /* final-time outputs: column `end` of the work arrays (sapdesolver.jl:245-256) */
  double** uf = u1;
  if (Nt == 0) uf = u1;   /* no step taken: column 2 is still all zeros, exactly as `X[:,end]` would be */
  if (o->out_mode == GAB1_OUT_FINAL4 || o->out_mode == GAB1_OUT_SIX) {
    double* y_stot = work + (size_t)21 * P;
    double* y_ptot = work + (size_t)22 * P;
    for (int j = 0; j < P; ++j) {
      y_stot[j] = uf[PG1S][j] + uf[G2PG1S][j];
      if (o->pg1tot_form == GAB1_PG1TOT_VIA_STOT) y_ptot[j] = uf[G2PG1][j] + uf[pGAB1][j] + y_stot[j];
      else y_ptot[j] = uf[G2PG1][j] + uf[pGAB1][j] + uf[PG1S][j] + uf[G2PG1S][j];
    }
    if (o->out_mode == GAB1_OUT_FINAL4) {
      memcpy(out, uf[iSFK], (size_t)P * sizeof(double));
      memcpy(out + P, uf[aSFK], (size_t)P * sizeof(double));
      memcpy(out + 2 * (size_t)P, y_ptot, (size_t)P * sizeof(double));
      memcpy(out + 3 * (size_t)P, y_stot, (size_t)P * sizeof(double));
      for (int64_t i = 0; i < 4 * (int64_t)P; ++i) if (isnan(out[i])) status |= GAB1_ST_NAN;
    } else {
      int threw = 0;
      double six[6];
      six[0] = length_scale(r, uf[aSFK], P, 0.5, R, &threw);
      six[1] = length_scale(r, uf[aSFK], P, 0.1, R, &threw);
      six[2] = length_scale(r, y_stot, P, 0.5, R, &threw);
      six[3] = length_scale(r, y_stot, P, 0.1, R, &threw);
      six[4] = y_stot[0] / y_stot[P - 1];
      six[5] = trapz_r2(r, y_stot, P) * 3.0 / pow(R, 3.0);
      if (threw) { status |= GAB1_ST_THROW; for (int i = 0; i < 6; ++i) out[i] = 0.0; }
      else for (int i = 0; i < 6; ++i) { out[i] = six[i]; if (isnan(six[i])) status |= GAB1_ST_NAN; }
    }
  } else if (o->out_mode == GAB1_OUT_FINAL_STATE) {
    for (int q = 0; q < 10; ++q) memcpy(out + (size_t)q * P, uf[q], (size_t)P * sizeof(double));
    for (int q = 0; q < 8; ++q) out[(size_t)10 * P + q] = m1[q];
    for (int64_t i = 0; i < nout; ++i) if (isnan(out[i])) status |= GAB1_ST_NAN;
  } else if (o->out_mode == GAB1_OUT_PCT_BOUND) {
    /* run_base_model.jl:272-276 on column `end` of the snapshot matrices */
    const double ave = trapz_r2(r, last_col, P) * 3.0 / (R * R * R);
    const double mem = last_EG2PG1S * o->pct_mul / o->pct_div;
    const double tot = ave + mem;
    out[0] = tot / CoG1 * 100.0;
    if (isnan(out[0])) status |= GAB1_ST_NAN;
  }
  return status;
}

/* doubles of scratch solve_one needs */
static size_t work_doubles(const gab1_opts* o) { return (size_t)23 * ((size_t)o->Nr + 1); }

static int check_opts(const gab1_opts* o) {
  if (!o || o->abi_version != GAB1_ABI_VERSION) return -1;
  if (o->Nr < 2 || o->Nts < 1 || o->maxiters < 0) return -2;
  if (o->geometry < 0 || o->geometry > 1 || o->sfk_mode < 0 || o->sfk_mode > 2 || o->bc_loop < 0 || o->bc_loop > 1 ||
      o->save_rule < 0 || o->save_rule > 1 || o->pg1tot_form < 0 || o->pg1tot_form > 1 || o->out_mode < 0 || o->out_mode > 4)
    return -3;
  return 0;
}

/* Same argument meaning as gab1_solve_batch (include/gab1pde.h); nthreads <= 0 => OpenMP default. */
int gab1o_solve_batch(const gab1_opts* o, int64_t S, const double* Co, int64_t Co_stride,
                      const double* D, const double* k, const double* dt, const double* r,
                      double* out, int32_t* status, int32_t* n_saved, int64_t* n_steps, int64_t* n_bc_iters,
                      int32_t nthreads) {
  const int rc = check_opts(o);
  if (rc) return rc;
  const int64_t nout = gab1o_out_doubles_per_set(o);
  const size_t wd = work_doubles(o);
  int failed = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel
  {
    double* work = (double*)malloc(wd * sizeof(double));
    if (!work) {
#pragma omp atomic write
      failed = 1;
    }
#pragma omp for schedule(dynamic, 1)
    for (int64_t i = 0; i < S; ++i) {
      if (!work) continue;
      const uint32_t st = solve_one(o, Co + i * Co_stride, D + i * GAB1_N_D, k + i * GAB1_N_K, dt[i], r,
                                    out + i * nout, n_saved ? n_saved + i : NULL, n_steps ? n_steps + i : NULL,
                                    n_bc_iters ? n_bc_iters + i : NULL, work);
      if (status) status[i] = (int32_t)st;
    }
    free(work);
  }
  return failed ? -4 : 0;
}

/* `1.0/(2.0*(maximum(D)/(dr.^2) + sum(k)/4))*0.99` (basepdesolver.jl:30), sum(k) left to right */
int gab1o_default_dt(int64_t S, const double* D, const double* k, double dr, double* dt) {
  for (int64_t i = 0; i < S; ++i) {
    double mx = D[i * GAB1_N_D];
    for (int q = 1; q < GAB1_N_D; ++q) if (D[i * GAB1_N_D + q] > mx) mx = D[i * GAB1_N_D + q];
    double sk = 0.0;
    for (int q = 0; q < GAB1_N_K; ++q) sk += k[i * GAB1_N_K + q];
    dt[i] = 1.0 / (2.0 * (mx / (dr * dr) + sk / 4)) * 0.99;
  }
  return 0;
}

int gab1o_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
