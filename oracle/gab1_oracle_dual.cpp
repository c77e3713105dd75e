/*
 * gab1_oracle_dual.cpp — CPU ORACLE for the forward-mode (tangent) path — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * The reference differentiates its solver by running it on ForwardDiff dual numbers:
 *   pdesolver_fitting(p::AbstractVector{T}) where T          basepdesolver.jl:674-932
 *   loss / OptimizationFunction(…, AutoForwardDiff())         param_fitting+inference_finitediff.jl:188-240
 *   ForwardDiff.gradient(testf, kvals[6:9])                   param_fitting+inference_finitediff.jl:128-151
 *   turing_model (NUTS gradients through the same solver)     param_fitting+inference_finitediff.jl:308-370
 * This file restates that: the scalar time loop of gab1_oracle.c re-typed over a dual number (value + N partials),
 * every `+ - * /` following the rules ForwardDiff uses (ForwardDiff v0.10.34, Manifest.toml; its source is NOT in
 * the reference tree): sum rule, product rule `a'*b + a*b'`, quotient rule `a'*(1/b) + b'*(-(a/(b*b)))`, `abs` by
 * sign, comparisons / `maximum` / `ceil` on the value alone.  Consequently the control flow (number of time steps,
 * membrane fixed-point iterations, snapshot schedule) is that of the primal solve, and the value component must be
 * BIT-IDENTICAL to gab1_oracle.c (tests/test_tangent_cpu.py asserts it).
 *
 * Note that in pdesolver_fitting the time step is computed INSIDE the solver from p (basepdesolver.jl:696), so dt
 * carries partials too (through sum(k) and maximum(D)); the caller passes dt and its partials, computed with the same
 * rules (gab1o_default_dt_dual below).
 *
 * PARITY UNPINNED: no Julia here, and ForwardDiff evaluates the partials with `muladd`, whose contraction into an FMA
 * is LLVM's choice, so the last bits of the partials are not defined by the reference source.  The pin used instead is
 * mathematical: the partials must equal the derivative of the (pinned-by-cross-restatement) primal oracle, checked by
 * central differences in tests/test_tangent_cpu.py.
 *
 * Output layout (same as the product's gab1_solve_tangent): per set, (1 + n_dir) consecutive blocks of
 * gab1_out_doubles_per_set doubles: block 0 = values, block 1 + d = partials along direction d.
 * Seeds: per set and direction 30 doubles in pdesolver_fitting's packed order p = [D(7); k(17); Co(5)], then dt.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "gab1pde.h"

namespace {

enum { iSFK, aSFK, GAB1, pGAB1, GRB2, G2G1, G2PG1, SHP2, PG1S, G2PG1S };
enum { mE, mES, mESmES, E, EG2, EG2G1, EG2PG1, EG2PG1S };
const int kMatrixSpecies[10] = {iSFK, aSFK, GRB2, GAB1, SHP2, G2G1, G2PG1, G2PG1S, pGAB1, PG1S};

template <int N>
struct Dual {
  double v;
  double p[N];
  Dual() : v(0.0) { for (int i = 0; i < N; ++i) p[i] = 0.0; }
  Dual(double x) : v(x) { for (int i = 0; i < N; ++i) p[i] = 0.0; }
};
template <int N> Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> c; c.v = a.v + b.v; for (int i = 0; i < N; ++i) c.p[i] = a.p[i] + b.p[i]; return c; }
template <int N> Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> c; c.v = a.v - b.v; for (int i = 0; i < N; ++i) c.p[i] = a.p[i] - b.p[i]; return c; }
template <int N> Dual<N> operator-(const Dual<N>& a) { Dual<N> c; c.v = -a.v; for (int i = 0; i < N; ++i) c.p[i] = -a.p[i]; return c; }
template <int N> Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> c; c.v = a.v * b.v; for (int i = 0; i < N; ++i) c.p[i] = a.p[i] * b.v + a.v * b.p[i]; return c; }
template <int N> Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> c; c.v = a.v / b.v;
  const double inv = 1.0 / b.v, w = -(a.v / (b.v * b.v));
  for (int i = 0; i < N; ++i) c.p[i] = a.p[i] * inv + b.p[i] * w;
  return c;
}
/* Real (op) Dual: the real operand has no partials (ForwardDiff dual.jl, binary ops with a Real) */
template <int N> Dual<N> operator+(double a, const Dual<N>& b) { Dual<N> c = b; c.v = a + b.v; return c; }
template <int N> Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> c = a; c.v = a.v + b; return c; }
template <int N> Dual<N> operator-(double a, const Dual<N>& b) { Dual<N> c; c.v = a - b.v; for (int i = 0; i < N; ++i) c.p[i] = -b.p[i]; return c; }
template <int N> Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> c = a; c.v = a.v - b; return c; }
template <int N> Dual<N> operator*(double a, const Dual<N>& b) { Dual<N> c; c.v = a * b.v; for (int i = 0; i < N; ++i) c.p[i] = a * b.p[i]; return c; }
template <int N> Dual<N> operator*(const Dual<N>& a, double b) { Dual<N> c; c.v = a.v * b; for (int i = 0; i < N; ++i) c.p[i] = a.p[i] * b; return c; }
template <int N> Dual<N> operator/(const Dual<N>& a, double b) { Dual<N> c; c.v = a.v / b; for (int i = 0; i < N; ++i) c.p[i] = a.p[i] / b; return c; }
template <int N> Dual<N> operator/(double a, const Dual<N>& b) {
  Dual<N> c; c.v = a / b.v;
  const double w = -(c.v / b.v);
  for (int i = 0; i < N; ++i) c.p[i] = w * b.p[i];
  return c;
}

int64_t popcount12(uint32_t m) { int64_t c = 0; for (int i = 0; i < 12; ++i) c += (m >> i) & 1u; return c; }
int64_t out_doubles(const gab1_opts* o) {
  const int64_t P = (int64_t)o->Nr + 1, C = (int64_t)o->Nts + 1;
  switch (o->out_mode) {
    case GAB1_OUT_FINAL4: return 4 * P;
    case GAB1_OUT_FULL: return popcount12(o->matrix_mask) * P * C + GAB1_N_VECTORS * C;
    case GAB1_OUT_PCT_BOUND: return 1;
    case GAB1_OUT_FINAL_STATE: return 10 * P + 8;
  }
  return 0;
}
int64_t matrix_offset(const gab1_opts* o, int m) {
  if (!((o->matrix_mask >> m) & 1u)) return -1;
  const int64_t P = (int64_t)o->Nr + 1, C = (int64_t)o->Nts + 1;
  return popcount12(o->matrix_mask & ((1u << m) - 1u)) * P * C;
}
int64_t vector_offset(const gab1_opts* o, int v) {
  const int64_t P = (int64_t)o->Nr + 1, C = (int64_t)o->Nts + 1;
  return popcount12(o->matrix_mask) * P * C + (int64_t)v * C;
}

/* one output slot of every component: block 0 value, block 1+d partial d (only the first n_dir are stored) */
template <int N>
struct Sink {
  double* out; int64_t nout; int n_dir; int d0;      /* this pass carries directions d0 .. d0 + n_dir - 1 */
  void put(int64_t at, const Dual<N>& x) const {
    out[at] = x.v;
    for (int d = 0; d < n_dir; ++d) out[(int64_t)(1 + d0 + d) * nout + at] = x.p[d];
  }
};

/* NumericalIntegration.integrate(r, y .* r.^2) (trapezoid); r carries no partials */
template <int N>
Dual<N> trapz_r2(const double* r, const Dual<N>* y, int P) {
  Dual<N> acc(0.0);
  for (int i = 0; i + 1 < P; ++i) {
    const Dual<N> yi = y[i] * (r[i] * r[i]);
    const Dual<N> yn = y[i + 1] * (r[i + 1] * r[i + 1]);
    acc = acc + (r[i + 1] - r[i]) * (yi + yn);
  }
  return 0.5 * acc;
}

template <int N>
uint32_t solve_one(const gab1_opts* o, int n_dir, int d0, const double* Co_, const double* D_, const double* k_, double dt_,
                   const double* seed /* n_dir x 30 */, const double* r, double* out, int32_t* n_saved_out,
                   int64_t* n_steps_out, int64_t* n_bc_out) {
  typedef Dual<N> T;
  const int Nr = o->Nr, P = Nr + 1, Nts = o->Nts, C = Nts + 1;
  const double dr = o->dr, R = o->R, tf = o->tf, tol = o->tol;
  const int64_t nout = out_doubles(o);
  uint32_t status = 0;
  if (n_saved_out) *n_saved_out = 0;
  if (n_steps_out) *n_steps_out = 0;
  if (n_bc_out) *n_bc_out = 0;
  const Sink<N> sink = {out, nout, n_dir, d0};

  /* p = [D; k; Co] with partials, and dt (basepdesolver.jl:690-696) */
  T D[7], k[17], Co[5], dt(dt_);
  for (int i = 0; i < 7; ++i) { D[i] = T(D_[i]); for (int d = 0; d < n_dir; ++d) D[i].p[d] = seed[d * 30 + i]; }
  for (int i = 0; i < 17; ++i) { k[i] = T(k_[i]); for (int d = 0; d < n_dir; ++d) k[i].p[d] = seed[d * 30 + 7 + i]; }
  for (int i = 0; i < 5; ++i) { Co[i] = T(Co_[i]); for (int d = 0; d < n_dir; ++d) Co[i].p[d] = seed[d * 30 + 24 + i]; }
  for (int d = 0; d < n_dir; ++d) dt.p[d] = seed[d * 30 + 29];

  T D_Si = D[0], D_Sa = D[0];
  if (o->sfk_mode == GAB1_SFK_MEMBRANE) D_Sa = T(1e-32);
  if (o->sfk_mode == GAB1_SFK_BOTH_FROZEN) { D_Si = T(1e-32); D_Sa = T(1e-32); }
  const T D_G2 = D[1], D_G2G1 = D[2], D_G2G1S2 = D[3], D_G1 = D[4], D_G1S2 = D[5], D_S2 = D[6];
  const T Dsp[10] = {D_Si, D_Sa, D_G1, D_G1, D_G2, D_G2G1, D_G2G1, D_S2, D_G1S2, D_G2G1S2};
  const T kS2f = k[0], kS2r = k[1], kG1f = k[2], kG1r = k[3], kG2f = k[4], kG2r = k[5], kG1p = k[6], kG1dp = k[7],
          kSa = k[8], kSi = k[9], kp = k[10], kdp = k[11], kEGFf = k[12], kEGFr = k[13], EGF = k[14], kdf = k[15],
          kdr = k[16];
  const T CoSFK = Co[0], CoG2 = Co[1], CoG1 = Co[2], CoS2 = Co[3], CoEGFR = Co[4];

  /* Nt = Int64(ceil(tf/dt)): ceil of a Dual is ceil of its value (basepdesolver.jl:729-735) */
  const double nt_f = ceil(tf / dt.v);
  if (!(nt_f >= 0.0 && nt_f < 9.0e18)) return GAB1_ST_THROW;
  const int64_t Nt = (int64_t)nt_f;
  if (n_steps_out) *n_steps_out = Nt;

  std::vector<T> store((size_t)21 * P);
  T* u0[10]; T* u1[10];
  for (int q = 0; q < 10; ++q) { u0[q] = store.data() + (size_t)q * P; u1[q] = store.data() + (size_t)(10 + q) * P; }
  T* last_col = store.data() + (size_t)20 * P;
  T m0[8], m1[8];
  for (int j = 0; j < P; ++j) { u0[iSFK][j] = CoSFK; u0[GAB1][j] = CoG1; u0[GRB2][j] = CoG2; u0[SHP2][j] = CoS2; }
  m0[mE] = CoEGFR;

  const int track_t = (o->out_mode == GAB1_OUT_FULL || o->out_mode == GAB1_OUT_PCT_BOUND);
  int64_t Mx[12], Vx[GAB1_N_VECTORS];
  for (int m = 0; m < 12; ++m) Mx[m] = -1;
  T last_EG2PG1S(0.0);
  if (o->out_mode == GAB1_OUT_FULL) {
    for (int m = 0; m < 12; ++m) Mx[m] = matrix_offset(o, m);
    for (int v = 0; v < GAB1_N_VECTORS; ++v) Vx[v] = vector_offset(o, v);
    for (int j = 0; j < P; ++j) {           /* column 1 = initial state (basepdesolver.jl:776-781) */
      if (Mx[GAB1_M_iSFK] >= 0) sink.put(Mx[GAB1_M_iSFK] + j, CoSFK);
      if (Mx[GAB1_M_GRB2] >= 0) sink.put(Mx[GAB1_M_GRB2] + j, CoG2);
      if (Mx[GAB1_M_SHP2] >= 0) sink.put(Mx[GAB1_M_SHP2] + j, CoS2);
      if (Mx[GAB1_M_GAB1] >= 0) sink.put(Mx[GAB1_M_GAB1] + j, CoG1);
    }
    sink.put(Vx[GAB1_V_mE], CoEGFR);
  }

  T t(0.0);
  double t_save = o->dt_save;
  int nts = 1;
  const double dr2 = dr * dr;
  int64_t bc_total = 0;

  for (int64_t step = 1; step <= Nt; ++step) {
    for (int j = 1; j < Nr; ++j) {          /* basepdesolver.jl:797-827 */
      T L[10];
      if (o->geometry == GAB1_GEOM_SPHERICAL) {
        const double a = 1 / (r[j] * dr);
        for (int q = 0; q < 10; ++q) {
          const T up = u0[q][j + 1], uc = u0[q][j], um = u0[q][j - 1];
          L[q] = Dsp[q] * (a * (up - um) + (up - 2.0 * uc + um) / dr2);
        }
      } else {
        for (int q = 0; q < 10; ++q) {
          const T up = u0[q][j + 1], uc = u0[q][j], um = u0[q][j - 1];
          L[q] = Dsp[q] * (up - 2.0 * uc + um) / dr2;
        }
      }
      const T Si = u0[iSFK][j], Sa = u0[aSFK][j], G1 = u0[GAB1][j], pG1 = u0[pGAB1][j], G2 = u0[GRB2][j],
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
    for (int q = 0; q < 10; ++q) u1[q][0] = u1[q][1];      /* basepdesolver.jl:830-839 */

    /* membrane fixed point (basepdesolver.jl:844-885); the exit test looks at values only */
    int it = 0;
    for (;;) {
      if (it >= o->maxiters) break;
      ++it;
      double cold[10], mold[8];
      for (int q = 0; q < 10; ++q) cold[q] = u1[q][Nr].v;
      for (int q = 0; q < 8; ++q) mold[q] = m1[q].v;
      const T Etot = 2.0 * (m1[E] + m1[EG2] + m1[EG2G1] + m1[EG2PG1] + m1[EG2PG1S]);
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
      const T bG1 = u1[GAB1][Nr], bpG1 = u1[pGAB1][Nr], bG2 = u1[GRB2][Nr], bg2g1 = u1[G2G1][Nr], bg2pg1 = u1[G2PG1][Nr],
              bS2 = u1[SHP2][Nr], bpg1s = u1[PG1S][Nr], bg2pg1s = u1[G2PG1S][Nr];
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
      double mx = -INFINITY; int nan_seen = 0;
      for (int q = 0; q < 10; ++q) { const double e = fabs(1.0 - u1[q][Nr].v / cold[q]); if (isnan(e)) nan_seen = 1; else if (e > mx) mx = e; }
      for (int q = 0; q < 8; ++q) { const double e = fabs(1.0 - m1[q].v / mold[q]); if (isnan(e)) nan_seen = 1; else if (e > mx) mx = e; }
      const double err = nan_seen ? NAN : mx;
      if (err <= tol) break;
    }
    bc_total += it;

    for (int q = 0; q < 10; ++q) for (int j = 0; j < P; ++j) u0[q][j] = u1[q][j];     /* basepdesolver.jl:887-905 */
    for (int q = 0; q < 8; ++q) m0[q] = m1[q];

    if (track_t) {
      const T Etot = 2.0 * (m1[E] + m1[EG2] + m1[EG2G1] + m1[EG2PG1] + m1[EG2PG1S]);
      t = t + dt;                                                                      /* :908 */
      if (t.v >= t_save) {                                                             /* :912 */
        if (nts >= C) status |= GAB1_ST_OVERFLOW;
        else {
          const int c = nts;
          nts += 1;
          if (o->out_mode == GAB1_OUT_FULL) {
            for (int m = 0; m < 10; ++m)
              if (Mx[m] >= 0) for (int j = 0; j < P; ++j) sink.put(Mx[m] + (int64_t)c * P + j, u1[kMatrixSpecies[m]][j]);
            for (int j = 0; j < P; ++j) {
              const T stot = u1[PG1S][j] + u1[G2PG1S][j];
              T ptot;
              if (o->pg1tot_form == GAB1_PG1TOT_VIA_STOT) ptot = u1[G2PG1][j] + u1[pGAB1][j] + stot;
              else ptot = u1[G2PG1][j] + u1[pGAB1][j] + u1[PG1S][j] + u1[G2PG1S][j];
              if (Mx[GAB1_M_PG1Stot] >= 0) sink.put(Mx[GAB1_M_PG1Stot] + (int64_t)c * P + j, stot);
              if (Mx[GAB1_M_PG1tot] >= 0) sink.put(Mx[GAB1_M_PG1tot] + (int64_t)c * P + j, ptot);
              if (isnan(u1[PG1S][j].v)) status |= GAB1_ST_NAN;
            }
            sink.put(Vx[GAB1_V_pE] + c, Etot * 100.0 / CoEGFR);
            sink.put(Vx[GAB1_V_mE] + c, m1[mE]); sink.put(Vx[GAB1_V_mES] + c, m1[mES]);
            sink.put(Vx[GAB1_V_mESmES] + c, m1[mESmES]); sink.put(Vx[GAB1_V_E] + c, m1[E]);
            sink.put(Vx[GAB1_V_EG2] + c, m1[EG2]); sink.put(Vx[GAB1_V_EG2G1] + c, m1[EG2G1]);
            sink.put(Vx[GAB1_V_EG2PG1] + c, m1[EG2PG1]); sink.put(Vx[GAB1_V_EG2PG1S] + c, m1[EG2PG1S]);
            sink.put(Vx[GAB1_V_EGFR_SHP2] + c, m1[EG2PG1S] * 100.0 / CoEGFR);
            sink.put(Vx[GAB1_V_t_out] + c, t);
          } else if (c == C - 1) {
            for (int j = 0; j < P; ++j) last_col[j] = u1[PG1S][j] + u1[G2PG1S][j];
            last_EG2PG1S = m1[EG2PG1S];
          }
        }
        t_save += o->dt_save;
      }
    }
  }
  if (n_bc_out) *n_bc_out = bc_total;
  if (n_saved_out) *n_saved_out = track_t ? nts : 0;
  if (track_t && nts < C) status |= GAB1_ST_SHORT;

  T** uf = u1;
  if (o->out_mode == GAB1_OUT_FINAL4) {
    for (int j = 0; j < P; ++j) {
      const T stot = uf[PG1S][j] + uf[G2PG1S][j];
      T ptot;
      if (o->pg1tot_form == GAB1_PG1TOT_VIA_STOT) ptot = uf[G2PG1][j] + uf[pGAB1][j] + stot;
      else ptot = uf[G2PG1][j] + uf[pGAB1][j] + uf[PG1S][j] + uf[G2PG1S][j];
      sink.put(j, uf[iSFK][j]); sink.put((int64_t)P + j, uf[aSFK][j]);
      sink.put(2 * (int64_t)P + j, ptot); sink.put(3 * (int64_t)P + j, stot);
    }
    for (int64_t i = 0; i < nout; ++i) if (isnan(out[i])) status |= GAB1_ST_NAN;
  } else if (o->out_mode == GAB1_OUT_FINAL_STATE) {
    for (int q = 0; q < 10; ++q) for (int j = 0; j < P; ++j) sink.put((int64_t)q * P + j, uf[q][j]);
    for (int q = 0; q < 8; ++q) sink.put((int64_t)10 * P + q, m1[q]);
    for (int64_t i = 0; i < nout; ++i) if (isnan(out[i])) status |= GAB1_ST_NAN;
  } else if (o->out_mode == GAB1_OUT_PCT_BOUND) {
    /* param_fitting+inference_finitediff.jl:211-216 (= run_base_model.jl:272-276) */
    const T ave = trapz_r2<N>(r, last_col, P) * 3.0 / (R * R * R);
    const T mem = last_EG2PG1S * o->pct_mul / o->pct_div;
    const T pct = (ave + mem) / CoG1 * 100.0;
    sink.put(0, pct);
    if (isnan(pct.v)) status |= GAB1_ST_NAN;
  }
  return status;
}

}  // namespace

extern "C" {

/* Same argument meaning as gab1_solve_tangent (include/gab1pde.h) plus a thread count. */
int gab1o_solve_tangent(const gab1_opts* o, int64_t S, int32_t n_dir, const double* Co, int64_t Co_stride, const double* D,
                        const double* k, const double* dt, const double* seeds, const double* r, double* out,
                        int32_t* status, int32_t* n_saved, int64_t* n_steps, int64_t* n_bc_iters, int32_t nthreads) {
  if (!o || o->abi_version != GAB1_ABI_VERSION) return -1;
  if (o->Nr < 2 || o->Nts < 1 || o->maxiters < 0 || n_dir < 1 || n_dir > 64) return -2;
  if (o->bc_loop != GAB1_BC_FOR_BREAK || o->save_rule != GAB1_SAVE_T_GE_TSAVE || o->t_prechase >= 0.0 ||
      o->out_mode == GAB1_OUT_SIX)
    return -3;
  const int64_t nout = out_doubles(o);
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t i = 0; i < S; ++i) {
    double* oi = out + i * nout * (1 + n_dir);
    const double* sd = seeds + i * (int64_t)n_dir * 30;
    int32_t* sv = n_saved ? n_saved + i : NULL;
    int64_t* ns = n_steps ? n_steps + i : NULL;
    int64_t* nb = n_bc_iters ? n_bc_iters + i : NULL;
    const double *Ci = Co + i * Co_stride, *Di = D + i * GAB1_N_D, *ki = k + i * GAB1_N_K;
    uint32_t st = 0;
    memset(oi, 0, (size_t)nout * (size_t)(1 + n_dir) * sizeof(double));
    for (int d0 = 0; d0 < n_dir; d0 += 8) {        /* at most 8 partials per pass; the value block is rewritten identically */
      const int n = n_dir - d0 < 8 ? n_dir - d0 : 8;
      const double* sdc = sd + (int64_t)d0 * 30;
      if (n == 1) st = solve_one<1>(o, n, d0, Ci, Di, ki, dt[i], sdc, r, oi, sv, ns, nb);
      else if (n == 2) st = solve_one<2>(o, n, d0, Ci, Di, ki, dt[i], sdc, r, oi, sv, ns, nb);
      else if (n <= 4) st = solve_one<4>(o, n, d0, Ci, Di, ki, dt[i], sdc, r, oi, sv, ns, nb);
      else st = solve_one<8>(o, n, d0, Ci, Di, ki, dt[i], sdc, r, oi, sv, ns, nb);
    }
    if (status) status[i] = (int32_t)st;
  }
  return 0;
}

/* dt = 1.0/(2.0*(maximum(D)/(dr_.^2) + sum(k)/4))*0.99 on duals (basepdesolver.jl:696): writes dt[S] and fills slot 29
 * of every seed row with dt's partial.  maximum picks the dual with the largest value; sum folds left to right. */
int gab1o_default_dt_dual(int64_t S, int32_t n_dir, const double* D, const double* k, double dr, double* dt, double* seeds) {
  for (int64_t i = 0; i < S; ++i) {
    int im = 0;
    for (int q = 1; q < GAB1_N_D; ++q) if (D[i * GAB1_N_D + q] > D[i * GAB1_N_D + im]) im = q;
    double sk = 0.0;
    for (int q = 0; q < GAB1_N_K; ++q) sk += k[i * GAB1_N_K + q];
    const double mx = D[i * GAB1_N_D + im];
    const double inner = mx / (dr * dr) + sk / 4;
    const double den = 2.0 * inner;
    const double inv = 1.0 / den;
    dt[i] = inv * 0.99;
    for (int d = 0; d < n_dir; ++d) {
      double* s = seeds + (i * n_dir + d) * 30;
      double dsk = 0.0;
      for (int q = 0; q < GAB1_N_K; ++q) dsk += s[7 + q];
      const double dinner = s[im] / (dr * dr) + dsk / 4;
      const double dden = 2.0 * dinner;
      s[29] = (-(inv / den) * dden) * 0.99;
    }
  }
  return 0;
}

}  /* extern "C" */
