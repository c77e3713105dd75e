"""The two independent CPU restatements (C, NumPy) must agree bit for bit (SURVEY.md §8c pin 4)."""
import numpy as np
import pytest

from oracle import numpy_oracle as npo


def _bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)).view(np.uint64)


def assert_bit_equal(a, b, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, what
    both_nan = np.isnan(a) & np.isnan(b)
    same = (_bits(a) == _bits(b)) | both_nan
    assert same.all(), f"{what}: {np.count_nonzero(~same)} of {same.size} differ; first at {np.argwhere(~same)[0]}"


CASES = [
    # name, row, kwargs for numpy solve, kwargs for frontend.pdesolver_batch
    ("sph_base", 0, dict(), dict()),
    ("sph_row1", 1, dict(), dict()),
    ("sph_membSFK", 2, dict(sfk_mode=1), dict(sfk_mode=1)),
    ("rect", 3, dict(rect=True, chain_pg1tot=True), dict(geometry=1, pg1tot_form=1)),
    ("rect_frozen_modulus", 4, dict(rect=True, chain_pg1tot=True, sfk_mode=2, modulus_rule=True),
     dict(geometry=1, pg1tot_form=1, sfk_mode=2, save_rule=1)),
    ("pulsechase", 5, dict(t_prechase=0.2), dict(t_prechase=0.2)),
    ("diverging_row76", 76, dict(), dict()),
]


@pytest.mark.parametrize("name,row,npkw,fekw", CASES, ids=[c[0] for c in CASES])
def test_full_solution_bitwise(pkg, ofe, ensemble, name, row, npkw, fekw):
    Co = pkg.params.base_Co()
    D, k = ensemble[row, :7], ensemble[row, 7:]
    dr, tf, Nts, tol, maxit = 0.4, 0.3, 6, 1e-4, 20
    r = pkg.params.julia_range(dr, 10.0)
    dt = float(pkg.params.default_dt(D, k, dr)[0])
    ref = npo.solve(Co, D, k, r, dr=dr, tf=tf, dt=dt, Nts=Nts, maxiters=maxit, tol=tol, **npkw)
    res = ofe.pdesolver_batch(Co, D[None], k[None], dr=dr, tf=tf, Nts=Nts, tol=tol, maxiters=maxit, **fekw)
    assert int(res.n_steps[0]) == ref["Nt"]
    assert int(res.n_saved[0]) == ref["n_saved"]
    assert int(res.n_bc_iters[0]) == ref["n_bc"]
    for n in pkg.abi.MATRIX_NAMES:
        assert_bit_equal(res.matrix(n)[0], ref["mats"][n], f"{name}:{n}")
    for n in pkg.abi.VECTOR_NAMES:
        assert_bit_equal(res.vector(n)[0], ref["vecs"][n], f"{name}:{n}")


@pytest.mark.parametrize("membSFK", [False, True])
def test_final_time_and_six_bitwise(pkg, ofe, ensemble, membSFK):
    Co = pkg.params.hela_Co() if membSFK else pkg.params.base_Co()
    rows = [7, 8]
    dr, tf = 0.4, 0.25
    r = pkg.params.julia_range(dr, 10.0)
    res4 = ofe.sapdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], dr=dr, tf=tf, membSFK=membSFK)
    res6 = ofe.sapdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], dr=dr, tf=tf, membSFK=membSFK,
                                 out_mode=pkg.abi.OUT_SIX)
    resS = ofe.sapdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], dr=dr, tf=tf, membSFK=membSFK,
                                 out_mode=pkg.abi.OUT_FINAL_STATE)
    P = len(r)
    for i, row in enumerate(rows):
        D, k = ensemble[row, :7], ensemble[row, 7:]
        dt = float(pkg.params.default_dt(D, k, dr)[0])
        ref = npo.solve(Co, D, k, r, dr=dr, tf=tf, dt=dt, maxiters=20, tol=1e-3, sfk_mode=1 if membSFK else 0,
                        while_loop=membSFK, chain_pg1tot=membSFK, snapshots=False)
        o = res4.out[i]
        assert_bit_equal(o[:P], ref["final"]["iSFK"])
        assert_bit_equal(o[P:2 * P], ref["final"]["aSFK"])
        assert_bit_equal(o[2 * P:3 * P], ref["PG1tot"])
        assert_bit_equal(o[3 * P:], ref["PG1Stot"])
        assert_bit_equal(res6.out[i], npo.six_scalars(r, ref["final"]["aSFK"], ref["PG1Stot"], 10.0))
        for q, n in enumerate(npo.CYTO):
            assert_bit_equal(resS.out[i][q * P:(q + 1) * P], ref["final"][n], n)
        assert_bit_equal(resS.out[i][10 * P:], [ref["memb"][n] for n in npo.MEMB])
        assert int(res4.n_bc_iters[i]) == ref["n_bc"]


def test_first_steps_hit_maxiters(pkg, ofe, ensemble):
    """Work arrays start at zero, so the first steps see 0/0 = NaN errors and run all `maxiters` iterations
    (basepdesolver.jl:238-241)."""
    Co = pkg.params.base_Co()
    D, k = ensemble[0, :7], ensemble[0, 7:]
    dt = float(pkg.params.default_dt(D, k, 0.4)[0])
    res = ofe.pdesolver_batch(Co, D[None], k[None], dr=0.4, tf=5 * dt * 0.999, Nts=2, tol=1e-4, maxiters=20)
    assert int(res.n_steps[0]) == 5 and int(res.n_bc_iters[0]) == 100
