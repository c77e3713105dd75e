"""Parity of the forward-mode CUDA kernels (gab1_solve_tangent, through the C ABI) against the dual-number oracle
(oracle/gab1_oracle_dual.cpp).  Needs a B200: run with -m gpu.

Bars:
  * value block: as the primal fast kernels — |gpu - ref| <= 1e-9 * max(|ref|, 1e-6 * max|field|), identical step counts,
    snapshot schedules, membrane-iteration counts and status words;
  * partials: |gpu - ref| <= TTOL * scale, TTOL = 1e-9, where scale is taken over the same output array (matrix / vector /
    profile) of the same set and direction: scale = max(max|ref partial|, 1e-4 * max|ref value| * max_i |seed_i / p_i|).
    A partial changes sign inside an array, so an element-wise relative bound is meaningless; and the second term says
    that a partial is resolved to 1e-13 * |value| per unit RELATIVE change of the parameter — the level to which the two
    implementations agree on the values themselves (observed 1e-14 .. 2e-13).  Partials far below that (d GAB1 / d D_S =
    1e-13 with membrane SFKs; d SHP2 / d D_G2, which acts only through dt) are residues of cancellations in either code.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-9
TTOL = 1e-9
FIT_K = (6, 7, 8, 9)


@pytest.fixture(scope="module")
def gfe(pkg):
    import __graft_entry__ as g
    g.build()
    lib = pkg.abi.load_library()
    assert lib.gab1_device_count() >= 1, "no CUDA device: the product path has no CPU fallback"
    return pkg.host.Frontend(pkg.abi.CudaBackend())


def unit_seeds(S, cols):
    s = np.zeros((S, len(cols), 30))
    for d, c in enumerate(cols):
        s[:, d, c] = 1.0
    return s


def blocks(o, abi):
    """(start, stop) of every output array inside one component block."""
    P, Cn = o.Nr + 1, o.Nts + 1
    if o.out_mode == abi.OUT_FULL:
        nm = bin(o.matrix_mask).count("1")
        return [(i * P * Cn, (i + 1) * P * Cn) for i in range(nm)] + [(nm * P * Cn + v * Cn, nm * P * Cn + (v + 1) * Cn) for v in range(11)]
    if o.out_mode == abi.OUT_FINAL4:
        return [(i * P, (i + 1) * P) for i in range(4)]
    if o.out_mode == abi.OUT_FINAL_STATE:
        return [(i * P, (i + 1) * P) for i in range(10)] + [(10 * P, 10 * P + 8)]      # ten profiles, the membrane column
    return [(0, 1)]


def value_err(a, b):
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    fin = np.isfinite(b)
    scale = np.where(fin, np.abs(b), 0).max(axis=-1, keepdims=True)
    den = np.maximum(np.abs(b), 1e-6 * scale)
    with np.errstate(invalid="ignore", divide="ignore"):
        return float(np.where(fin & (den > 0), np.abs(a - b) / den, 0.0).max())


def rel_seed(res, Co, D, k):
    """max_i |seed_i / p_i| per (set, direction): the relative size of the parameter perturbation a unit step along the
    direction makes."""
    S = D.shape[0]
    p = np.concatenate([D, k, np.broadcast_to(Co, (S, 5)), res.dt[:, None]], axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = np.where(res.seeds != 0, np.abs(res.seeds / p[:, None, :]), 0.0)
    return q.max(axis=-1)[:, :, None]


def tangent_err(res, ref, abi, rs=None):
    worst = 0.0
    for lo, hi in blocks(res.opts, abi):
        a, b = res.out[:, 1:, lo:hi], ref.out[:, 1:, lo:hi]
        assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern of the partials differs"
        fin = np.isfinite(b)
        scale = np.where(fin, np.abs(b), 0).max(axis=-1, keepdims=True)
        if rs is not None:
            v = ref.out[:, :1, lo:hi]
            vmax = np.where(np.isfinite(v), np.abs(v), 0).max(axis=-1, keepdims=True)
            scale = np.maximum(scale, 1e-4 * vmax * rs)
        with np.errstate(invalid="ignore", divide="ignore"):
            e = np.where(fin & (scale > 0), np.abs(a - b) / scale, np.where(fin, np.abs(a - b), 0.0))
        if float(e.max()) > worst:
            worst = float(e.max())
            tangent_err.where = (lo, hi) + tuple(int(i) for i in np.unravel_index(np.argmax(e), e.shape))
    return worst


def check(res, ref, abi, pars=None):
    np.testing.assert_array_equal(res.n_steps, ref.n_steps)
    np.testing.assert_array_equal(res.n_saved, ref.n_saved)
    # a diverging set's iteration count during the blow-up is not reproducible between contracted and strict arithmetic
    # (the primal fast kernels are held to the same rule, test_gpu_parity.py): compare the counts of the healthy sets
    good = (ref.status & abi.ST_NAN) == 0
    np.testing.assert_array_equal(res.n_bc_iters[good], ref.n_bc_iters[good])
    np.testing.assert_array_equal(res.status, ref.status)
    ev = value_err(res.out[:, 0], ref.out[:, 0])
    et = tangent_err(res, ref, abi, None if pars is None else rel_seed(res, *pars))
    assert ev < RTOL, f"value block: relative error {ev:.3e}"
    assert et < TTOL, f"partials: error {et:.3e} of the array scale at (block lo, hi, set, direction, element) = {getattr(tangent_err, 'where', None)}"
    return ev, et


GRIDS = {"K1": dict(dr=0.5, tf=0.4, Nts=8), "K2": dict(dr=0.2, tf=0.15, Nts=6), "K4": dict(dr=0.1, tf=0.04, Nts=4)}


@pytest.mark.parametrize("grid", list(GRIDS))
@pytest.mark.parametrize("mode", ["full", "final4", "pct", "state"])
def test_tangents_match_the_dual_oracle(pkg, gfe, ofe, ensemble, grid, mode):
    abi = pkg.abi
    Co = pkg.params.base_Co()
    rows = [0, 1, 2, 3, 4999]
    D, k = ensemble[rows, :7], ensemble[rows, 7:]
    om = dict(full=abi.OUT_FULL, final4=abi.OUT_FINAL4, pct=abi.OUT_PCT_BOUND, state=abi.OUT_FINAL_STATE)[mode]
    volCF, surfCF = pkg.params.conversion_factors()
    kw = dict(out_mode=om, tol=1e-4, maxiters=20, pct_mul=volCF, pct_div=surfCF, **GRIDS[grid])
    if mode == "full":
        kw["matrices"] = ("aSFK", "PG1S", "G2PG1S")
    seeds = unit_seeds(len(rows), [7 + j for j in FIT_K])
    res = gfe.pdesolver_tangent_batch(Co, D, k, seeds, **kw)
    ref = ofe.pdesolver_tangent_batch(Co, D, k, seeds, **kw)
    np.testing.assert_array_equal(res.seeds, ref.seeds)
    np.testing.assert_array_equal(res.dt, ref.dt)
    check(res, ref, abi, (Co, D, k))


@pytest.mark.parametrize("family", ["reg", "stream"])
@pytest.mark.parametrize("nt", ["1", "2", "4"])
@pytest.mark.parametrize("n_dir", [1, 3, 5])
def test_direction_grouping(pkg, gfe, ofe, ensemble, family, nt, n_dir, monkeypatch):
    """Both kernel families (partials in registers / streamed through shared memory) forced through the GAB1_TANGENT
    override: n_dir directions are cut into ceil(n_dir/NT) work items per set; every NT the kernels are built for, ragged
    last group, all 12 matrices, directions through D, k, Co and dt."""
    abi = pkg.abi
    if family == "stream" and nt == "1":
        pytest.skip("the streamed kernels carry 2 or 4 directions")
    monkeypatch.setenv("GAB1_TANGENT", family)
    monkeypatch.setenv("GAB1_TANGENT_NT", nt)
    Co = pkg.params.base_Co()
    D, k = ensemble[10:17, :7], ensemble[10:17, 7:]
    cols = [1, 7 + 8, 24 + 2, 7 + 1, 24 + 4][:n_dir]
    rng = np.random.default_rng(5)
    seeds = unit_seeds(7, cols) + (rng.standard_normal((7, n_dir, 30)) * 1e-3 if n_dir == 5 else 0.0)   # dense directions too
    for grid in ("K1", "K2", "K4") if family == "stream" else ("K1", "K2"):
        kw = dict(tol=1e-4, maxiters=20, **GRIDS[grid])
        res = gfe.pdesolver_tangent_batch(Co, D, k, seeds, **kw)
        ref = ofe.pdesolver_tangent_batch(Co, D, k, seeds, **kw)
        check(res, ref, abi, (Co, D, k))


@pytest.mark.parametrize("variant", [dict(geometry=1, pg1tot_form=1), dict(sfk_mode=1), dict(geometry=1, sfk_mode=2, pg1tot_form=1)])
def test_variants(pkg, gfe, ofe, ensemble, variant):
    """planar Laplacian, membrane aSFK (a true division by 1e-32 in the closure), both SFK frozen."""
    abi = pkg.abi
    Co = np.tile(pkg.params.hela_Co(), (4, 1)) * np.array([[1.0], [0.5], [2.0], [1.5]])      # per-set Co
    D, k = ensemble[20:24, :7], ensemble[20:24, 7:]
    seeds = unit_seeds(4, [0, 7 + 6, 7 + 9, 24 + 0])
    kw = dict(tol=1e-4, maxiters=20, dr=0.2, tf=0.1, Nts=4, **variant)
    res = gfe.pdesolver_tangent_batch(Co, D, k, seeds, **kw)
    ref = ofe.pdesolver_tangent_batch(Co, D, k, seeds, **kw)
    check(res, ref, abi, (Co, D, k))


def test_value_block_equals_the_primal_kernel_schedule(pkg, gfe, ensemble):
    """The value block of the tangent call and the primal product kernel: same control flow, values within 1e-9."""
    Co = pkg.params.base_Co()
    D, k = ensemble[:64, :7], ensemble[:64, 7:]
    kw = dict(dr=0.2, tf=0.3, Nts=6, tol=1e-4, maxiters=20, matrices=("aSFK", "PG1S", "G2PG1S"))
    tan = gfe.pdesolver_tangent_batch(Co, D, k, unit_seeds(64, [7 + 6, 7 + 7]), **kw)
    pri = gfe.pdesolver_batch(Co, D, k, **kw)
    np.testing.assert_array_equal(tan.n_bc_iters, pri.n_bc_iters)
    np.testing.assert_array_equal(tan.n_saved, pri.n_saved)
    np.testing.assert_array_equal(tan.status, pri.status)
    assert value_err(tan.out[:, 0], pri.out) < RTOL


def test_diverging_set_and_edge_cases(pkg, gfe, ofe, ensemble):
    """Row 75 diverges at dr = 0.2 (NaN reaches the outputs and the partials); unusable dt; snapshot overflow; empty batch."""
    abi = pkg.abi
    Co = pkg.params.base_Co()
    rows = [75, 0, 1]
    D, k = ensemble[rows, :7], ensemble[rows, 7:]
    seeds = unit_seeds(3, [7 + j for j in FIT_K])
    kw = dict(dr=0.2, tf=5.0, Nts=10, tol=1e-4, maxiters=20, out_mode=abi.OUT_FINAL4)
    dt = pkg.params.default_dt(D, k, 0.2)
    dt[2] = 0.0                                  # Int64(ceil(tf/0)) throws in the reference
    res = gfe.pdesolver_tangent_batch(Co, D, k, seeds, dt=dt, **kw)
    ref = ofe.pdesolver_tangent_batch(Co, D, k, seeds, dt=dt, **kw)
    assert res.status[0] & abi.ST_NAN and res.status[2] & abi.ST_THROW
    check(res, ref, abi, (Co, D, k))
    pri = gfe.pdesolver_batch(Co, D, k, dt=dt, **kw)               # the diverging set: same control flow as the primal kernel
    np.testing.assert_array_equal(res.n_bc_iters, pri.n_bc_iters)
    # more snapshots due than columns (dt_save tiny)
    kw2 = dict(dr=0.5, tf=0.2, Nts=3, dt_save=0.01, tol=1e-4, maxiters=20)
    res = gfe.pdesolver_tangent_batch(Co, D[1:2], k[1:2], seeds[1:2], **kw2)
    ref = ofe.pdesolver_tangent_batch(Co, D[1:2], k[1:2], seeds[1:2], **kw2)
    assert res.status[0] & abi.ST_OVERFLOW
    check(res, ref, abi, (Co, D[1:2], k[1:2]))
    # fewer snapshots than columns
    kw3 = dict(dr=0.5, tf=0.2, Nts=3, dt_save=0.15, tol=1e-4, maxiters=20)
    res = gfe.pdesolver_tangent_batch(Co, D[1:2], k[1:2], seeds[1:2], **kw3)
    ref = ofe.pdesolver_tangent_batch(Co, D[1:2], k[1:2], seeds[1:2], **kw3)
    assert res.status[0] & abi.ST_SHORT
    check(res, ref, abi, (Co, D[1:2], k[1:2]))
    empty = gfe.pdesolver_tangent_batch(Co, D[:0], k[:0], seeds[:0], dr=0.5, tf=0.1, Nts=2)
    assert empty.out.shape[0] == 0


def test_unsupported_options_fail_loudly(pkg, gfe, ensemble):
    abi = pkg.abi
    Co = pkg.params.base_Co()
    D, k = ensemble[:1, :7], ensemble[:1, 7:]
    with pytest.raises(abi.Gab1Error, match="SIX"):
        gfe.pdesolver_tangent_batch(Co, D, k, unit_seeds(1, [7]), dr=0.5, tf=0.1, Nts=2, out_mode=abi.OUT_SIX)
    with pytest.raises(abi.Gab1Error, match="128"):
        gfe.pdesolver_tangent_batch(Co, D, k, unit_seeds(1, [7]), dr=0.05, tf=0.001, Nts=2)


def test_golden_tangent_kat(pkg, gfe, ensemble):
    from pathlib import Path
    kat = np.load(Path(__file__).resolve().parent / "golden" / "tangent_kat.npz")
    sub = ensemble[kat["rows"]]
    seeds = unit_seeds(len(sub), [7 + j for j in FIT_K])
    res = gfe.pdesolver_tangent_batch(pkg.params.base_Co(), sub[:, :7], sub[:, 7:], seeds, dr=0.4, tf=0.5, Nts=5, tol=1e-3,
                                      maxiters=20, matrices=("aSFK", "PG1S", "G2PG1S"))

    class Ref:
        out = kat["full_dr04_tf05"]
    np.testing.assert_array_equal(res.n_bc_iters, kat["full_dr04_tf05_nbc"])
    assert value_err(res.out[:, 0], Ref.out[:, 0]) < RTOL
    assert tangent_err(res, Ref, pkg.abi) < TTOL


def test_fitting_gradient_full_length(pkg, gfe, ofe):
    """The quantity the reference's optimiser differentiates (param_fitting+inference_finitediff.jl:188-240): loss and
    gradient in log-parameters at the run_ensemble grid, full tf = 5, a multistart batch."""
    p0 = np.concatenate([pkg.params.DIFFS_BASE, pkg.params.KVALS_BASE])
    inds = [7 + j for j in FIT_K]
    rng = np.random.default_rng(123)
    x = np.log(p0[inds])[None, :] + rng.uniform(-1.0, 1.0, (6, 4))
    kw = dict(param_inds=inds, pvals0=p0, Co=pkg.params.base_Co(), dr=0.2, tf=5.0, Nts=100, tol=1e-3, maxiters=20)
    lg, gg, yg = gfe.fitting_loss_and_gradient(x, 26.426, 9.363, **kw)
    lo, go, yo = ofe.fitting_loss_and_gradient(x, 26.426, 9.363, **kw)
    assert np.abs(yg / yo - 1).max() < RTOL
    assert np.abs(gg - go).max() <= 1e-8 * np.abs(go).max()      # (mu - yhat) cancels: the loss gradient amplifies yhat's error


@pytest.mark.parametrize("family", ["reg", "stream"])
def test_tangent_runs_are_bitwise_repeatable(pkg, gfe, ensemble, family, monkeypatch):
    """Race surrogate (no compute-sanitizer on this pool): 1500 sets x 4 directions twice, bit for bit."""
    monkeypatch.setenv("GAB1_TANGENT", family)
    Co = pkg.params.base_Co()
    D, k = ensemble[:1500, :7], ensemble[:1500, 7:]
    seeds = unit_seeds(1500, [7 + j for j in FIT_K])
    for dr, tf in ((0.4, 0.05), (0.2, 0.03), (0.1, 0.01)):
        kw = dict(dr=dr, tf=tf, Nts=3, tol=1e-4, maxiters=20, out_mode=pkg.abi.OUT_FINAL4)
        a = gfe.pdesolver_tangent_batch(Co, D, k, seeds, **kw)
        b = gfe.pdesolver_tangent_batch(Co, D, k, seeds, **kw)
        same = (a.out.view(np.uint64) == b.out.view(np.uint64)) | (np.isnan(a.out) & np.isnan(b.out))
        assert same.all(), f"{family} dr={dr}: {np.count_nonzero(~same)} values differ between two runs"
        np.testing.assert_array_equal(a.n_bc_iters, b.n_bc_iters)


@pytest.mark.parametrize("mode", ["full", "final4", "pct", "state"])
def test_team_tangent_kernel(pkg, gfe, ofe, ensemble, mode, monkeypatch):
    """The latency path of forward mode (one CTA per set and direction, team_tangent_kernel.cuh) forced through
    GAB1_TANGENT=team: Nr = 100 (4 warps, spherical), Nr = 80 (3 warps, planar), per-set Co, directions through D, k, Co, dt.
    The same inputs through the register kernel (GAB1_TANGENT=reg) must pass the same bar."""
    abi = pkg.abi
    om = dict(full=abi.OUT_FULL, final4=abi.OUT_FINAL4, pct=abi.OUT_PCT_BOUND, state=abi.OUT_FINAL_STATE)[mode]
    volCF, surfCF = pkg.params.conversion_factors()
    rows = [0, 1, 2, 4999, 76]
    D, k = ensemble[rows, :7], ensemble[rows, 7:]
    Co = np.tile(pkg.params.base_Co(), (5, 1)) * np.array([[1.0], [0.7], [1.3], [1.0], [1.0]])
    seeds = unit_seeds(5, [1, 7 + 6, 7 + 9, 24 + 2, 24 + 4])
    for dr, tf, extra in ((0.1, 0.04, {}), (0.125, 0.05, dict(geometry=1, pg1tot_form=1))):
        kw = dict(out_mode=om, tol=1e-4, maxiters=20, pct_mul=volCF, pct_div=surfCF, dr=dr, tf=tf, Nts=4, **extra)
        if mode == "full":
            kw["matrices"] = ("aSFK", "PG1S", "G2PG1S", "PG1tot")
        ref = ofe.pdesolver_tangent_batch(Co, D, k, seeds, **kw)
        for family, nt in (("reg", ""), ("team", "1"), ("team", "2"), ("team", "4")):      # 1, 2 or 4 directions per CTA (5 directions: ragged)
            monkeypatch.setenv("GAB1_TANGENT", family)
            monkeypatch.setenv("GAB1_TANGENT_NT", nt)
            res = gfe.pdesolver_tangent_batch(Co, D, k, seeds, **kw)
            try:
                check(res, ref, abi, (Co, D, k))
            except AssertionError as e:
                raise AssertionError(f"{family} nt={nt} dr={dr}: {e}") from e


def test_fitting_gradient_full_length_final_stage_grid(pkg, gfe, ofe):
    """The final optimisation stage of the reference runs at dr = 0.1 (param_fitting+inference_finitediff.jl:240,263): two
    full-length (tf = 5, 1.4e5 steps) loss-and-gradient evaluations through the default dispatch (one CTA per set and
    direction) against the dual oracle."""
    p0 = np.concatenate([pkg.params.DIFFS_BASE, pkg.params.KVALS_BASE])
    inds = [7 + j for j in FIT_K]
    x = np.log(p0[inds])[None, :] + np.array([[0.0, 0.0, 0.0, 0.0], [0.4, -0.3, 0.2, -0.5]])
    kw = dict(param_inds=inds, pvals0=p0, Co=pkg.params.base_Co(), dr=0.1, tf=5.0, Nts=100, tol=1e-3, maxiters=20)
    lg, gg, yg = gfe.fitting_loss_and_gradient(x, 26.426, 9.363, **kw)
    lo, go, yo = ofe.fitting_loss_and_gradient(x, 26.426, 9.363, **kw)
    assert np.abs(yg / yo - 1).max() < RTOL
    assert np.abs(gg - go).max() <= 1e-8 * np.abs(go).max()
