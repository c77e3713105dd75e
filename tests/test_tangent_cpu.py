"""Pins for the forward-mode (dual-number) oracle, oracle/gab1_oracle_dual.cpp — CPU tier.

The reference differentiates its solver with ForwardDiff (pdesolver_fitting, basepdesolver.jl:674-932; callers
param_fitting+inference_finitediff.jl:128-151, 188-240, 308-370).  No Julia here, so the dual oracle is pinned by
  (1) its value component being BIT-IDENTICAL to the scalar C oracle (same operations in the same order),
  (2) its partials being the derivative of that pinned primal: central finite differences of the scalar oracle,
  (3) committed known-answer vectors (tests/golden/tangent_kat.npz).
"""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"

# directions used throughout: the four fitted rate constants kG1p, kG1dp, kSa, kSi = k[6:10] (param_fitting…:130-136)
FIT_K = (6, 7, 8, 9)


def unit_seeds(S, cols):
    s = np.zeros((S, len(cols), 30))
    for d, c in enumerate(cols):
        s[:, d, c] = 1.0
    return s


@pytest.mark.parametrize("mode", ["full", "final4", "pct", "state"])
def test_value_component_is_the_scalar_oracle(pkg, ofe, ensemble, mode):
    abi = pkg.abi
    Co = pkg.params.base_Co()
    D, k = ensemble[:3, :7], ensemble[:3, 7:]
    kw = dict(dr=0.5, tf=0.3, Nts=6, tol=1e-4, maxiters=20)
    om = dict(full=abi.OUT_FULL, final4=abi.OUT_FINAL4, pct=abi.OUT_PCT_BOUND, state=abi.OUT_FINAL_STATE)[mode]
    extra = dict(matrices=("aSFK", "PG1S", "G2PG1S")) if mode == "full" else {}
    seeds = unit_seeds(3, [7 + j for j in FIT_K])
    tan = ofe.pdesolver_tangent_batch(Co, D, k, seeds, out_mode=om, pct_mul=2.0, pct_div=3.0, **extra, **kw)
    ref = ofe.pdesolver_batch(Co, D, k, out_mode=om, pct_mul=2.0, pct_div=3.0, **extra, **kw)
    np.testing.assert_array_equal(tan.dt, ref.dt)
    np.testing.assert_array_equal(tan.out[:, 0], ref.out)
    np.testing.assert_array_equal(tan.n_steps, ref.n_steps)
    np.testing.assert_array_equal(tan.n_bc_iters, ref.n_bc_iters)
    np.testing.assert_array_equal(tan.n_saved, ref.n_saved)
    np.testing.assert_array_equal(tan.status, ref.status)
    assert np.abs(tan.out[:, 1:]).max() > 0


def _central_difference(ofe, Co, D, k, col, h, dt, **kw):
    def run(sign):
        D2, k2, Co2 = D.copy(), k.copy(), np.tile(Co, (D.shape[0], 1))
        if col < 7:
            D2[:, col] *= 1.0 + sign * h
        elif col < 24:
            k2[:, col - 7] *= 1.0 + sign * h
        else:
            Co2[:, col - 24] *= 1.0 + sign * h
        return ofe.pdesolver_batch(Co2, D2, k2, dt=dt, **kw)
    a, b = run(+1.0), run(-1.0)
    base = np.concatenate([D, k, np.tile(Co, (D.shape[0], 1))], axis=1)[:, col]
    assert np.array_equal(a.n_steps, b.n_steps)
    return (a.out - b.out) / (2.0 * h * base)[:, None]


@pytest.mark.parametrize("geometry,sfk", [(0, 0), (1, 0), (0, 1)])
def test_partials_are_the_derivative_of_the_primal_fixed_dt(pkg, ofe, ensemble, geometry, sfk):
    """dt held constant (seed slot 29 = 0).  The membrane fixed point is iterated to machine precision (tol 1e-13) so that
    the primal is a smooth function of the parameters and finite differences are meaningful."""
    abi = pkg.abi
    Co = pkg.params.base_Co()
    D, k = ensemble[:2, :7].copy(), ensemble[:2, 7:].copy()
    dt = pkg.params.default_dt(D, k, 0.5)
    kw = dict(dr=0.5, tf=0.25, Nts=5, tol=1e-13, maxiters=60, out_mode=abi.OUT_FINAL_STATE, geometry=geometry, sfk_mode=sfk)
    cols = [0, 4, 7 + 2, 7 + 6, 7 + 8, 7 + 9, 7 + 10, 7 + 15, 24 + 2, 24 + 4]     # D_S, D_G1, kG1f, kG1p, kSa, kSi, kp, kdf, CoG1, CoEGFR
    tan = ofe.pdesolver_tangent_batch(Co, D, k, unit_seeds(2, cols), dt=dt, **kw)
    for d, col in enumerate(cols):
        fd = _central_difference(ofe, Co, D, k, col, 1e-6, dt, **kw)
        ad = tan.out[:, 1 + d]
        scale = np.abs(fd).max(axis=1, keepdims=True)
        # membrane SFK: aSFK[Nr+1] ~ 1e28 (a division by 1e-32, basepdesolver.jl:530) dominates the scale and carries the
        # finite-difference noise of the fixed point's residual
        bound = 2e-6 if sfk == 0 else 5e-5
        assert np.abs(ad - fd).max() <= bound * scale.max(), (col, np.abs(ad - fd).max() / scale.max())


def test_partials_through_dt(pkg, ofe, ensemble):
    """dt = dt(D, k) as inside pdesolver_fitting (basepdesolver.jl:696): the chain through dt, the accumulated clock
    and the snapshot times."""
    abi = pkg.abi
    Co = pkg.params.base_Co()
    D, k = ensemble[1:2, :7].copy(), ensemble[1:2, 7:].copy()
    kw = dict(dr=0.5, tf=0.25, Nts=5, tol=1e-13, maxiters=60, out_mode=abi.OUT_FULL, matrices=("aSFK", "PG1S", "G2PG1S"))
    cols = [1, 7 + 1, 7 + 9]       # D_G2 (the largest D), kS2r = 480 (dominates sum(k)), kSi
    tan = ofe.pdesolver_tangent_batch(Co, D, k, unit_seeds(1, cols), **kw)
    assert np.all(tan.seeds[0, :, 29] < 0)          # every one of them shortens the step
    for d, col in enumerate(cols):
        h = 1e-6

        def run(sign):
            D2, k2 = D.copy(), k.copy()
            (D2 if col < 7 else k2)[:, col if col < 7 else col - 7] *= 1.0 + sign * h
            return ofe.pdesolver_batch(Co, D2, k2, **kw)       # dt recomputed from the perturbed parameters
        a, b = run(1.0), run(-1.0)
        assert np.array_equal(a.n_steps, b.n_steps) and np.array_equal(a.n_saved, b.n_saved)
        base = (D if col < 7 else k)[0, col if col < 7 else col - 7]
        fd = (a.out - b.out) / (2 * h * base)
        ad = tan.out[:, 1 + d]
        assert np.abs(ad - fd).max() <= 5e-6 * np.abs(fd).max(), (col, np.abs(ad - fd).max() / np.abs(fd).max())
        # the clock: t_out = n*dt, so d t_out = n * d dt
        t, dtd = tan.vector("t_out")[0, 0], tan.vector("t_out")[0, 1 + d]
        n = np.round(t / tan.dt[0])
        np.testing.assert_allclose(dtd, n * tan.seeds[0, d, 29], rtol=1e-10, atol=0)


def test_loss_gradient_matches_finite_differences(pkg, ofe):
    """loss(pvals_in) of param_fitting+inference_finitediff.jl:188-226 in log-parameters, gradient as AutoForwardDiff gives."""
    p0 = np.concatenate([pkg.params.DIFFS_BASE, pkg.params.KVALS_BASE])
    inds = [7 + j for j in FIT_K]
    x = np.log(p0[inds])[None, :]
    kw = dict(param_inds=inds, pvals0=p0, Co=pkg.params.base_Co(), dr=0.5, tf=0.5, Nts=10, tol=1e-13, maxiters=60)
    loss, grad, yhat = ofe.fitting_loss_and_gradient(x, 26.426, 9.363, **kw)
    assert 0 < yhat[0] < 100 and np.isfinite(loss[0])
    for i in range(4):
        h = 1e-6
        xp, xm = x.copy(), x.copy()
        xp[0, i] += h
        xm[0, i] -= h
        lp = ofe.fitting_loss_and_gradient(xp, 26.426, 9.363, **kw)[0][0]
        lm = ofe.fitting_loss_and_gradient(xm, 26.426, 9.363, **kw)[0][0]
        fd = (lp - lm) / (2 * h)
        assert abs(grad[0, i] - fd) <= 2e-5 * np.abs(grad).max(), (i, grad[0, i], fd)


def test_pdesolver_fitting_dual_surface(pkg, ofe):
    """Shapes and conventions of the Dual call of pdesolver_fitting (basepdesolver.jl:674-932, :929)."""
    p = np.concatenate([pkg.params.DIFFS_BASE, pkg.params.KVALS_BASE, pkg.params.base_Co()])
    dp = np.zeros((4, 29))
    for d, j in enumerate(FIT_K):
        dp[d, 7 + j] = 1.0
    (sol, dsol), r, (t, dtt), (dt, ddt) = ofe.pdesolver_fitting_dual(p, dp, dr=0.5, tf=0.2, Nts=4, tol=1e-3)
    sol0, r0, t0, dt0 = ofe.pdesolver_fitting(p, dr=0.5, tf=0.2, Nts=4, tol=1e-3)
    assert dt == dt0 and np.array_equal(t, t0) and np.array_equal(r, r0)
    for name in ("aSFK", "PG1S", "G2PG1S", "EG2PG1S"):
        np.testing.assert_array_equal(getattr(sol, name), getattr(sol0, name))
    assert dsol.aSFK.shape == (4, 21, 5) and dsol.EG2PG1S.shape == (4, 5) and dtt.shape == (4, 5) and ddt.shape == (4,)
    assert np.all(dsol.aSFK[:, :, 0] == 0)          # the initial column does not depend on k


def test_unsupported_options_are_rejected(pkg, ofe, ensemble):
    abi = pkg.abi
    with pytest.raises(RuntimeError):
        ofe.pdesolver_tangent_batch(pkg.params.base_Co(), ensemble[:1, :7], ensemble[:1, 7:], unit_seeds(1, [7]),
                                    dr=0.5, tf=0.1, Nts=2, out_mode=abi.OUT_SIX)


def test_tangent_kat(pkg, ofe, ensemble):
    """Committed known-answer vectors of the dual oracle (tests/golden/make_tangent_fixtures.py)."""
    kat = np.load(GOLD / "tangent_kat.npz")
    rows = kat["rows"]
    sub = ensemble[rows]
    seeds = unit_seeds(len(rows), [7 + j for j in FIT_K])
    res = ofe.pdesolver_tangent_batch(pkg.params.base_Co(), sub[:, :7], sub[:, 7:], seeds, dr=0.4, tf=0.5, Nts=5, tol=1e-3,
                                      maxiters=20, matrices=("aSFK", "PG1S", "G2PG1S"))
    np.testing.assert_array_equal(res.out, kat["full_dr04_tf05"])
    np.testing.assert_array_equal(res.n_bc_iters, kat["full_dr04_tf05_nbc"])
