"""The reference-named wrappers around the hot path that no other GPU test calls: each is
run through the CUDA library and through the oracle-bound twin of the same front end, and must agree to 1e-9 (grid-
quantised length scales exactly)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def gfe(pkg):
    import __graft_entry__ as g
    g.build()
    assert pkg.abi.load_library().gab1_device_count() >= 1, "no CUDA device: the product path has no CPU fallback"
    return pkg.host.Frontend(pkg.abi.CudaBackend())


def close(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = np.isfinite(b) & (b != 0)
    assert np.all(a[~m & np.isfinite(b)] == b[~m & np.isfinite(b)])
    return float(np.max(np.abs(a[m] - b[m]) / np.abs(b[m]))) if m.any() else 0.0


def field_err(a, b):
    """The suite's tolerance for profiles and time series: |a-b| / max(|b|, 1e-6 * max|b|) (values far below the field's
    scale — e.g. aSFK in the interior when its diffusivity is 1e-32 — are compared against that floor)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    scale = np.nanmax(np.abs(b)) if np.isfinite(b).any() else 0.0
    den = np.maximum(np.abs(b), 1e-6 * scale)
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.where(den > 0, np.abs(a - b) / den, 0.0)
    return float(np.nanmax(e)) if e.size else 0.0


def six_close(a, b):
    np.testing.assert_array_equal(a[:4], b[:4])          # r_1/2, r_1/10 of aSFK and PG1Stot: multiples of dr
    assert close(a[4:], b[4:]) < RTOL


@pytest.mark.parametrize("membSFK", [False, True])
def test_pmap_fun_variants(pkg, gfe, ofe, ensemble, membSFK):
    """pmap_fun_allpars / pmap_fun_dk / pmap_fun_dk_combD / pmap_fun_concs (sapdesolver.jl:288-451 and the membSFK twins,
    sapdesolver_memb-SFK.jl:288-449): packed parameter vector in, six scalars out."""
    Co = pkg.params.hela_Co() if membSFK else pkg.params.base_Co()
    D, k = ensemble[3, :7], ensemble[3, 7:]
    kw = dict(tf=1.0, membSFK=membSFK)
    six_close(gfe.pmap_fun_allpars(np.concatenate([Co, D, k]), **kw), ofe.pmap_fun_allpars(np.concatenate([Co, D, k]), **kw))
    six_close(gfe.pmap_fun_dk(np.concatenate([D, k]), Co=Co, **kw), ofe.pmap_fun_dk(np.concatenate([D, k]), Co=Co, **kw))
    p = np.concatenate([[1.3 * pkg.params.DIFFS_BASE[0]], k])
    six_close(gfe.pmap_fun_dk_combD(p, Co=Co, **kw), ofe.pmap_fun_dk_combD(p, Co=Co, **kw))
    six_close(gfe.pmap_fun_concs(Co * 0.7, **kw), ofe.pmap_fun_concs(Co * 0.7, **kw))


@pytest.mark.parametrize("membSFK", [False, True])
def test_fbatch_concs_mt(pkg, gfe, ofe, membSFK):
    """fbatch_concs_mt (sapdesolver.jl:460-476): 5 x S log-space concentrations over the eFAST range of GSA_concs.jl:62-71."""
    Co = pkg.params.base_Co()
    g = np.random.Generator(np.random.PCG64(11))
    p = np.log(Co)[:, None] + g.uniform(np.log(2e-4), np.log(2.0), size=(5, 24))
    Y, Yr = gfe.fbatch_concs_mt(p, tf=1.0, membSFK=membSFK), ofe.fbatch_concs_mt(p, tf=1.0, membSFK=membSFK)
    assert Y.shape == (6, 24)
    six_close(Y, Yr)


def test_pdesolver_membSFK_rect_takes_an_eight_element_D(pkg, gfe, ofe, ensemble):
    """pdesolver_membSFK_rect (basepdesolver_rect.jl:298-569): D[1] and D[3:8] of an 8-vector, both SFK diffusivities
    1e-32, modulus snapshot rule, one output column per snapshot taken, default dt from maximum(D) over all eight."""
    Co = pkg.params.base_Co()
    D7, k = ensemble[5, :7], ensemble[5, 7:]
    D8 = np.concatenate([D7[:1], [999.0], D7[1:]])          # D[2] is never read by the solver, but it enters the default dt
    a = gfe.pdesolver_membSFK_rect(Co, D8, k, dr=0.2, tf=0.4, Nts=9, tol=1e-4)
    b = ofe.pdesolver_membSFK_rect(Co, D8, k, dr=0.2, tf=0.4, Nts=9, tol=1e-4)
    assert a[3] == b[3] and np.array_equal(a[2], b[2]) and np.array_equal(a[1], b[1])
    assert a[0]._fields == b[0]._fields and "EGFR_SHP2" in a[0]._fields
    for name in a[0]._fields:
        x, y = getattr(a[0], name), getattr(b[0], name)
        assert x.shape == y.shape
        assert field_err(x, y) < RTOL, name
    with pytest.raises(IndexError):
        gfe.pdesolver_membSFK_rect(Co, D7, k, dr=0.2, tf=0.1)


def test_run_ensemble_pc(pkg, gfe, ofe, ensemble):
    """run_ensemble_pc (get_param_posteriors.jl:204-236) over pulsechase_solver: kp := 0 from t_prechase on, NaN sets dropped."""
    Co = pkg.params.base_Co()
    rows = [0, 1, 75, 2, 3]
    kw = dict(dr=0.2, t_prechase=0.3, t_chase=0.2, Nts=10)
    a = gfe.run_ensemble_pc("pulsechase_solver", ensemble[rows], Co, **kw)
    b = ofe.run_ensemble_pc("pulsechase_solver", ensemble[rows], Co, **kw)
    assert [r.index for r in a] == [r.index for r in b]
    for ra, rb in zip(a, b):
        assert np.array_equal(ra.t_sol, rb.t_sol)
        for name in ra.sol._fields:
            assert field_err(getattr(ra.sol, name), getattr(rb.sol, name)) < RTOL, name


def test_run_ensemble_raises_like_the_reference(pkg, gfe, ensemble):
    """A set whose Int64(ceil(tf/dt)) throws, or that has more snapshots due than columns, ends run_ensemble with an exception
    (the reference's threaded loop lets it out) instead of a zero or truncated row."""
    Co = pkg.params.base_Co()
    bad = ensemble[:3].copy()
    bad[1, 7:] = 0.0
    bad[1, :7] = 0.0                                         # dt = 0.99/(2*0) = Inf -> tf/dt = 0 -> fine; make dt NaN instead:
    bad[1, 7] = np.nan
    with pytest.raises(ArithmeticError):
        gfe.run_ensemble("pdesolver", bad, Co, dr=0.2, tf=0.1, Nts=4, show_prog=False)
