"""Pins for the oracle that do not depend on another restatement (SURVEY.md §8c pins 1-3, 5)."""
import numpy as np


def test_discrete_invariants(pkg, ofe, ensemble):
    """EGFR conservation holds step by step (terms cancel pairwise in basepdesolver.jl:220-231) and
    iSFK + aSFK = CoSFK at every node when both share one diffusivity (MATLAB/finitediff_steady_state_BVP_comparison.m:81)."""
    Co = pkg.params.base_Co()
    res = ofe.pdesolver_batch(Co, ensemble[:4, :7], ensemble[:4, 7:], dr=0.4, tf=2.0, Nts=20, tol=1e-4, maxiters=20)
    v = {n: res.vector(n) for n in pkg.abi.VECTOR_NAMES}
    tot = v["mE"] + v["mES"] + 2 * (v["mESmES"] + v["E"] + v["EG2"] + v["EG2G1"] + v["EG2PG1"] + v["EG2PG1S"])
    assert np.abs(tot / Co[4] - 1).max() < 1e-12
    sfk = res.matrix("iSFK") + res.matrix("aSFK")
    assert np.abs(sfk / Co[0] - 1).max() < 1e-12
    # derived outputs
    np.testing.assert_array_equal(res.matrix("PG1Stot"), res.matrix("PG1S") + res.matrix("G2PG1S"))
    np.testing.assert_array_equal(v["pE"][:, 1:], (2.0 * (v["E"] + v["EG2"] + v["EG2G1"] + v["EG2PG1"] + v["EG2PG1S"]) * 100.0 / Co[4])[:, 1:])
    # column 1 is the initial state, t_out increases by ~dt_save
    assert np.all(res.matrix("iSFK")[:, :, 0] == Co[0]) and np.all(res.matrix("aSFK")[:, :, 0] == 0)
    assert np.allclose(np.diff(v["t_out"], axis=1), 0.1, atol=2e-3)


def test_steady_asfk_profile_matches_closed_form(pkg, ofe):
    """At t = 5 min the aSFK profile is close to the steady solution C(r) = C(R) * (R/r) * sinh(m r)/sinh(m R),
    m = sqrt(kSi/D_S)  (MATLAB/finitediff_steady_state_BVP_comparison.m:98-104)."""
    Co = pkg.params.base_Co()
    D, k = pkg.params.DIFFS_BASE, pkg.params.KVALS_BASE
    sol, r = ofe.sapdesolver(Co, D, k, dr=0.2, tf=5.0)
    m = np.sqrt(k[9] / D[0])
    rr = r[1:]
    shape = (10.0 / rr) * np.sinh(m * rr) / np.sinh(m * 10.0)
    assert np.abs(sol.aSFK[1:] / sol.aSFK[-1] / shape - 1).max() < 0.02


def test_pct_shp2_bound_gab1_near_fitting_target(pkg, ofe, ensemble):
    """Posterior-median parameters give % SHP2-bound GAB1 inside the experimental 26.4 +/- 9.4
    (exptl_pct_SHP2-bound-GAB1.csv; SURVEY.md §4 probe: rows 0,1 -> 23.46 %, 26.05 % at dr = 0.2)."""
    Co = pkg.params.base_Co()
    pct, _ = ofe.pct_shp2_bound_gab1(Co, ensemble[:2, :7], ensemble[:2, 7:])
    assert abs(pct[0] - 23.46) < 0.01 and abs(pct[1] - 26.05) < 0.01
    pct0, _ = ofe.pct_shp2_bound_gab1(Co, pkg.params.DIFFS_BASE[None], pkg.params.KVALS_BASE[None])
    assert 17.0 < pct0[0] < 36.0


def test_pct_equals_epilogue_of_full_solution(pkg, ofe, ensemble):
    Co = pkg.params.base_Co()
    volCF, surfCF = pkg.params.conversion_factors()
    D, k = ensemble[5:7, :7], ensemble[5:7, 7:]
    pct, _ = ofe.pct_shp2_bound_gab1(Co, D, k, dr=0.4, tf=1.0, Nts=10)
    full = ofe.pdesolver_batch(Co, D, k, dr=0.4, tf=1.0, Nts=10, tol=1e-4, maxiters=20)
    r = full.r
    for i in range(2):
        y = (full.matrix("PG1S")[i][:, -1] + full.matrix("G2PG1S")[i][:, -1]) * (r * r)
        acc = 0.0
        for j in range(len(r) - 1):
            acc += (r[j + 1] - r[j]) * (y[j] + y[j + 1])
        ave = 0.5 * acc * 3.0 / (10.0 * 10.0 * 10.0)
        tot = ave + full.vector("EG2PG1S")[i][-1] * volCF / surfCF
        assert pct[i] == tot / Co[2] * 100.0


def test_known_diverging_row_and_nan_count(pkg, ofe, ensemble):
    """Row 76 (kG1p = 71.3) diverges at dr = 0.2 but not at dr = 0.1; the NaN filter of
    run_ensemble (get_param_posteriors.jl:155) drops it."""
    Co = pkg.params.base_Co()
    rows = ofe.run_ensemble("pdesolver", ensemble[74:77], Co, Nts=10)       # 0-based rows 74, 75, 76
    assert [x.index for x in rows] == [1, 3]
    sol = ofe.pdesolver(Co, ensemble[75, :7], ensemble[75, 7:], dr=0.1, Nts=4, tol=1e-4, maxiters=20)[0]
    assert np.isfinite(sol.PG1S).all()


def test_run_ensemble_rows_are_the_batch_blocks_views(pkg, ofe, ensemble):
    """host.Frontend.run_ensemble / run_ensemble_pc (the code the GPU backend runs too): every row's 21 / 22 fields are the set's
    own blocks of the batched result, the rect solver's outputs are cut to 1 + #snapshots columns (basepdesolver_rect.jl:250-279),
    EGFR_SHP2 sits where basepdesolver_rect.jl:282-290 puts it, and `index` is 1-based (get_param_posteriors.jl:158)."""
    Co, abi = pkg.params.base_Co(), pkg.abi
    sub = ensemble[72:78]
    for name, variant in (("pdesolver", {}), ("pdesolver_rect", dict(geometry=abi.GEOM_RECT, pg1tot_form=abi.PG1TOT_CHAIN))):
        rows = ofe.run_ensemble(name, sub, Co, Nts=10, tf=0.5, show_prog=False)
        res = ofe.pdesolver_batch(Co, sub[:, :7], sub[:, 7:], R=10.0, dr=0.2, tf=0.5, Nts=10, tol=1e-4, maxiters=20, **variant)
        keep = np.flatnonzero((res.status & abi.ST_NAN) == 0)
        assert [x.index for x in rows] == (keep + 1).tolist()
        for x, j in zip(rows, keep):
            nc = int(res.n_saved[j]) if name == "pdesolver_rect" else 11
            assert type(x.sol).__name__ == ("Sol22" if name == "pdesolver_rect" else "Sol21")
            for n in abi.MATRIX_NAMES:
                np.testing.assert_array_equal(getattr(x.sol, n), res.matrix(n)[j][:, :nc])
            for n in abi.VECTOR_NAMES[:9] + (("EGFR_SHP2",) if name == "pdesolver_rect" else ()):
                np.testing.assert_array_equal(getattr(x.sol, n), res.vector(n)[j][:nc])
            np.testing.assert_array_equal(x.t_sol, res.vector("t_out")[j][:nc])
    rows = ofe.run_ensemble_pc("pulsechase_solver", sub[:3], Co, t_prechase=0.2, t_chase=0.1, Nts=6)
    assert [x.index for x in rows] == [1, 2, 3] and type(rows[0].sol).__name__ == "Sol22" and rows[0].sol.aSFK.shape == (51, 7)


def test_oracle_kat(pkg, ofe, ensemble):
    """Known-answer vectors written by tests/golden/make_fixtures.py: guards the oracle's arithmetic against drift."""
    from pathlib import Path
    kat = np.load(Path(__file__).parent / "golden" / "oracle_kat.npz")
    sub = ensemble[kat["rows"]]
    Co = pkg.params.base_Co()
    res = ofe.pdesolver_batch(Co, sub[:, :7], sub[:, 7:], dr=0.4, tf=1.0, Nts=10, tol=1e-4, maxiters=20)
    np.testing.assert_array_equal(res.out, kat["full_dr04_tf1"])
    np.testing.assert_array_equal(res.n_bc_iters, kat["full_dr04_tf1_nbc"])
    res = ofe.sapdesolver_batch(Co, sub[:, :7], sub[:, 7:], dr=0.2, tf=0.5)
    np.testing.assert_array_equal(res.out, kat["final4_dr02_tf05"])
    res = ofe.sapdesolver_batch(pkg.params.hela_Co(), sub[:, :7], sub[:, 7:], dr=0.2, tf=0.5, membSFK=True)
    np.testing.assert_array_equal(res.out, kat["final4_memb_dr02_tf05"])
    res = ofe.sapdesolver_batch(Co, sub[:, :7], sub[:, 7:], dr=0.2, tf=0.5, out_mode=pkg.abi.OUT_SIX)
    np.testing.assert_array_equal(res.out, kat["six_dr02_tf05"])
    res = ofe.pdesolver_batch(Co, sub[:, :7], sub[:, 7:], dr=0.25, tf=0.5, Nts=5, tol=1e-4, maxiters=20,
                              geometry=pkg.abi.GEOM_RECT, pg1tot_form=pkg.abi.PG1TOT_CHAIN)
    np.testing.assert_array_equal(res.out, kat["full_rect_dr025_tf05"])


def test_stored_mle_is_a_constrained_optimum_of_the_oracle_loss(pkg, ofe):
    """A pin on an artefact the reference itself computed: fitted_parameters.csv is the result of its bound-constrained
    LBFGS fit of `loss` (param_fitting+inference_finitediff.jl:188-226,254-270; final stage dr = 0.1, tol = 1e-3, box =
    prior mode x 10^+-2, :177-183).  Three of the four constants are stored AT a box bound (kG1p = 100 x 0.42, kG1dp =
    kSi = 9.5 / 100) and kSa in the interior.  Under the oracle's loss and forward-mode gradient the stored point must
    satisfy the optimality conditions of that problem: the gradient pushes every bounded constant outwards through its
    active bound, the free one is stationary (orders of magnitude below the others), and the loss is below the loss at
    the optimiser's starting point p0 and at the posterior medians."""
    P = pkg.params
    mu, sigma = 26.426, 9.363293460636593                      # exptl_pct_SHP2-bound-GAB1.csv
    names = ("kG1p", "kG1dp", "kSa", "kSi")
    mle = np.array([P.FITTED_MLE[n] for n in names])
    p0 = np.array([0.42, 9.5, 0.42, 9.5])                        # prior modes (exp(-0.86750), exp(2.25129))
    assert np.allclose(mle[[0, 1, 3]], [p0[0] * 100, p0[1] / 100, p0[3] / 100], rtol=1e-12)
    assert p0[2] / 100 < mle[2] < p0[2] * 100
    pv = np.concatenate([P.DIFFS_BASE, P.KVALS_BASE])
    x = np.log(np.stack([mle, p0, P.KVALS_BASE[6:10]]))
    loss, grad, yhat = ofe.fitting_loss_and_gradient(x, mu, sigma, param_inds=[13, 14, 15, 16], pvals0=pv, Co=P.base_Co(),
                                                    dr=0.1, tol=1e-3, maxiters=20)
    assert loss[0] < loss[1] and loss[0] < loss[2]
    assert loss[0] < 2.0e-3 and abs(yhat[0] - 26.03) < 0.05    # the plateau below the data mean: yhat cannot reach 26.43
    g = grad[0]
    assert g[0] < 0 and g[1] > 0 and g[3] > 0                    # upper bound active; lower bounds active
    assert abs(g[2]) < 1e-2 * min(abs(g[0]), abs(g[1]))          # free parameter: stationary
