"""Ensemble order statistics (reference: run_base_model.jl:103-174).

CPU tier: the oracle's restatement of Julia's `median` / `quantile` against NumPy's type-7 quantile (same definition up to
the order of the floating-point operations).  GPU tier: gab1_solve_ensemble_quantiles against the oracle — bit-identical
with strict arithmetic (the sorted values are the same doubles), within 1e-9 with the fast kernels.
"""
import numpy as np
import pytest

PROBS = ("median", 0.5 - 0.341, 0.5 + 0.341, 0.0, 1.0)
ROWS_ODD = [0, 1, 2, 3, 4, 2500, 4999]
ROWS_EVEN_WITH_NAN = [0, 1, 2, 3, 75, 4, 2500, 4999, 333]      # rows 75 and 333 diverge at dr = 0.2 and are dropped: 7 remain


def test_oracle_statistics_follow_the_type7_definition(pkg, ofe, ensemble):
    Co = pkg.params.base_Co()
    ens = ensemble[ROWS_ODD + [5]]                                   # 8 sets: even count, median = middle of two
    res, n_valid, r, status = ofe.ensemble_quantiles(ens, Co, probs=PROBS, dr=0.4, tf=0.4, Nts=5)
    assert n_valid == 8 and not (status & pkg.abi.ST_NAN).any()
    full = ofe.pdesolver_batch(Co, ens[:, :7], ens[:, 7:], dr=0.4, tf=0.4, Nts=5, tol=1e-4, maxiters=20)
    for name in ("aSFK", "PG1tot", "PG1Stot"):
        stack = full.matrix(name)                                    # (sets, node, column)
        assert res[name].shape == (len(PROBS), 6, 26)
        np.testing.assert_allclose(res[name][0], np.median(stack, axis=0).T, rtol=2e-16, atol=0)
        for j, p in enumerate(PROBS[1:], 1):
            np.testing.assert_allclose(res[name][j], np.quantile(stack, p, axis=0).T, rtol=1e-14, atol=0)
        np.testing.assert_array_equal(res[name][3], stack.min(axis=0).T)
        # quantile(v, 1) = v[n-1] + 1*(v[n] - v[n-1]) in Julia's formula: the maximum up to one rounding
        np.testing.assert_allclose(res[name][4], stack.max(axis=0).T, rtol=4e-16, atol=0)


def test_abi_rejects_bad_quantile_requests_without_a_gpu(pkg):
    import ctypes as C
    lib = pkg.abi.load_library()
    o = pkg.abi.make_opts(dr=0.4, tf=0.1, Nts=2, out_mode=pkg.abi.OUT_FINAL4)
    z = np.zeros(64)
    dp = C.POINTER(C.c_double)
    rc = lib.gab1_solve_ensemble_quantiles(C.byref(o), 1, z.ctypes.data_as(dp), 0, z.ctypes.data_as(dp), z.ctypes.data_as(dp),
                                           z.ctypes.data_as(dp), z.ctypes.data_as(dp), 1, 0, 1, 1, z.ctypes.data_as(dp),
                                           z.ctypes.data_as(dp), None, None, None, None, None)
    assert rc != 0 and b"GAB1_OUT_FULL" in lib.gab1_last_error()
    assert lib.gab1_quantiles_workspace_bytes(5000) >= 4 * 5000


@pytest.mark.gpu
@pytest.mark.parametrize("rows", [ROWS_ODD, ROWS_EVEN_WITH_NAN], ids=["odd", "nan_sets_dropped"])
def test_gpu_quantiles_match_the_oracle(pkg, ofe, ensemble, rows):
    assert pkg.abi.load_library().gab1_device_count() >= 1
    sfe = pkg.host.Frontend(pkg.abi.CudaBackend(arith=pkg.abi.ARITH_STRICT))
    gfe = pkg.host.Frontend(pkg.abi.CudaBackend(arith=pkg.abi.ARITH_FAST))
    Co = pkg.params.base_Co()
    ens = ensemble[rows]
    kw = dict(probs=PROBS, dr=0.2, Nts=6) if rows is ROWS_EVEN_WITH_NAN else dict(probs=PROBS, dr=0.2, tf=0.5, Nts=6)
    ref, nv_r, r_r, st_r = ofe.ensemble_quantiles(ens, Co, **kw)
    strict, nv_s, r_s, st_s = sfe.ensemble_quantiles(ens, Co, **kw)
    fast, nv_f, r_f, st_f = gfe.ensemble_quantiles(ens, Co, **kw)
    assert nv_r == nv_s == nv_f == (7 if rows is ROWS_EVEN_WITH_NAN else len(rows))
    np.testing.assert_array_equal(st_r & pkg.abi.ST_NAN, st_s & pkg.abi.ST_NAN)
    np.testing.assert_array_equal(st_r & pkg.abi.ST_NAN, st_f & pkg.abi.ST_NAN)
    for name in ref:
        a, b = strict[name], ref[name]
        assert (a.view(np.uint64) == b.view(np.uint64)).all(), f"{name}: strict statistics are not bit-identical"
        scale = np.abs(b).max()
        err = np.abs(fast[name] - b) / np.maximum(np.abs(b), 1e-6 * scale)
        assert err.max() < 1e-9, f"{name}: {err.max():.3e}"


@pytest.mark.gpu
def test_gpu_quantiles_of_the_whole_ensemble_are_ordered(pkg, ensemble):
    """All 5000 rows at run_ensemble's defaults: 4967 sets enter, and lower <= median <= upper everywhere."""
    gfe = pkg.host.Frontend(pkg.abi.CudaBackend(arith=pkg.abi.ARITH_FAST))
    res, n_valid, r, status = gfe.ensemble_quantiles(ensemble, pkg.params.base_Co(), probs=(0.159, "median", 0.841),
                                                     matrices=("aSFK", "PG1Stot"), columns=(90, 101))
    assert n_valid == 5000 - 33 and int((status & pkg.abi.ST_NAN != 0).sum()) == 33
    for name, q in res.items():
        assert q.shape == (3, 11, 51) and np.isfinite(q).all()
        assert (q[0] <= q[1]).all() and (q[1] <= q[2]).all()


def test_regrid_is_piecewise_linear_interpolation(pkg):
    """run_base_model.jl:108-119: summary surfaces interpolated from the dr = 0.2 grid onto 0:0.1:R."""
    r = pkg.params.julia_range(0.2, 10.0)
    y = np.stack([np.sin(r), r ** 2])                       # (2, 51)
    v, x = pkg.host.Frontend.regrid(y, r, dr_new=0.1, R=10.0)
    assert v.shape == (2, 101) and np.array_equal(x, pkg.params.julia_range(0.1, 10.0))
    np.testing.assert_array_equal(v[:, ::2], y)             # knots are reproduced exactly
    np.testing.assert_allclose(v[:, 1::2], 0.5 * (y[:, :-1] + y[:, 1:]), rtol=1e-13, atol=1e-15)    # midpoints (w = 0.5 up to the rounding of x - r[i])
    np.testing.assert_allclose(v, np.stack([np.interp(x, r, y[0]), np.interp(x, r, y[1])]), rtol=1e-13, atol=1e-15)
