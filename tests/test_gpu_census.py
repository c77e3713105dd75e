"""Whole-configuration parity census of the product (fast) kernels against the oracle, at full length.

One test per BASELINE.json configuration: every parameter set of the configuration (or a stated sample
of it where the oracle would need minutes) is solved by the CUDA path through the C ABI and by the CPU oracle over the
FULL integration time, and the census (tests/census.py) counts, set by set, identical control flow and values within
1e-9.  Reports go to gpurun_out/census_*.json when that directory exists (copied to profiles/ by hand).

The bar (north_star: 1e-9 relative in FP64): every non-diverging, well-conditioned set has the oracle's step count,
snapshot schedule, status word AND membrane-iteration count, and its values are within 1e-9; diverging sets carry the
oracle's status.  A set is ill-conditioned when the ORACLE's own answer moves by 1e-10 or more (or changes its control
flow) under a one-ulp change of the initial concentrations (census.ill_conditioned): no arithmetic short of the
bit-identical strict kernels can track the reference there, so those sets are counted, listed, and held to the strict
kernel's bit-for-bit bar instead.  None exists in the posterior ensemble configurations; about 0.2 % of wide prior draws.
"""
import numpy as np
import pytest

from census import census, dump, ill_conditioned

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def gfe(pkg):
    import __graft_entry__ as g
    g.build()
    lib = pkg.abi.load_library()
    assert lib.gab1_device_count() >= 1, "no CUDA device: the product path has no CPU fallback"
    return pkg.host.Frontend(pkg.abi.CudaBackend(arith=pkg.abi.ARITH_FAST))


def perturbed(Co):
    return np.nextafter(np.asarray(Co, dtype=np.float64), np.inf)


def check(rep, max_ill=0):
    dump(rep)
    print(rep)
    assert rep["step_count_mismatches"] == 0 and rep["snapshot_count_mismatches"] == 0, rep
    assert rep["status_mismatches_live"] == 0, rep
    assert rep["ill_conditioned_sets"] <= max_ill, rep
    assert rep["flipped_sets"] == 0, f"membrane-iteration counts differ from the oracle on {rep['flipped']}"
    assert rep["same_flow_sets_over_rtol"] == 0, rep
    assert rep["live_sets_within_rtol"] == rep["live_sets"], rep


def test_census_config1_full_ensemble(pkg, gfe, ofe, ensemble):
    """configs[1]: pdesolver over ALL 5000 rows of parameter_ensemble.csv at run_ensemble's defaults
    (get_param_posteriors.jl:135-139), tf = 5: the complete final state of every set plus its counters."""
    Co = pkg.params.base_Co()
    kw = dict(dr=0.2, tf=5.0, Nts=100, tol=1e-4, maxiters=20, out_mode=pkg.abi.OUT_FINAL_STATE)
    res = gfe.pdesolver_batch(Co, ensemble[:, :7], ensemble[:, 7:], **kw)
    ref = ofe.pdesolver_batch(Co, ensemble[:, :7], ensemble[:, 7:], **kw)
    ref2 = ofe.pdesolver_batch(perturbed(Co), ensemble[:, :7], ensemble[:, 7:], **kw)
    rep = census(res, ref, name="config1_5000_rows_final_state", ill=ill_conditioned(ref, ref2))
    assert rep["diverging_sets"] == 33
    check(rep)                                    # no ill-conditioned set in the posterior ensemble


def test_census_config1_full_snapshots(pkg, gfe, ofe, ensemble):
    """configs[1] with the output the configuration names — all 12 matrices x 101 snapshots — on every fifth row."""
    Co = pkg.params.base_Co()
    rows = np.arange(0, 5000, 5)
    kw = dict(dr=0.2, tf=5.0, Nts=100, tol=1e-4, maxiters=20)
    res = gfe.pdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], **kw)
    ref = ofe.pdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], **kw)
    check(census(res, ref, name="config1_1000_rows_full_snapshots"))


def test_census_config2_synthetic_priors(pkg, gfe, ofe):
    """configs[2]: synthetic prior draws (get_param_priors.jl distributions) through fbatch_dk_mt's settings
    (sapdesolver.jl:330-387: dr = 0.2, tol = 1e-3, maxiters = 20), diverging sets included: final state, then the six
    GSA scalars (length scales identical, ratio and average within 1e-9 of their own magnitude)."""
    ens = pkg.params.synthetic_prior_ensemble(2048, seed=123)
    Co = pkg.params.base_Co()
    kw = dict(dr=0.2, tf=5.0, tol=1e-3, maxiters=20)
    res = gfe.sapdesolver_batch(Co, ens[:, :7], ens[:, 7:], out_mode=pkg.abi.OUT_FINAL_STATE, **kw)
    ref = ofe.sapdesolver_batch(Co, ens[:, :7], ens[:, 7:], out_mode=pkg.abi.OUT_FINAL_STATE, **kw)
    ref2 = ofe.sapdesolver_batch(perturbed(Co), ens[:, :7], ens[:, 7:], out_mode=pkg.abi.OUT_FINAL_STATE, **kw)
    ill = ill_conditioned(ref, ref2)
    rep = census(res, ref, name="config2_2048_prior_draws_final_state", ill=ill)
    assert rep["diverging_sets"] > 0              # wide priors: a few per cent blow up
    check(rep, max_ill=20)                        # < 1 % of the draws sit at the edge of the scheme's stability
    # ... and on exactly those sets the strict kernels still reproduce the oracle bit for bit
    bad = np.flatnonzero(ill)
    if len(bad):
        sfe = pkg.host.Frontend(pkg.abi.CudaBackend(arith=pkg.abi.ARITH_STRICT))
        s = sfe.sapdesolver_batch(Co, ens[bad, :7], ens[bad, 7:], out_mode=pkg.abi.OUT_FINAL_STATE, **kw)
        same = (s.out.view(np.uint64) == ref.out[bad].view(np.uint64)) | (np.isnan(s.out) & np.isnan(ref.out[bad]))
        assert same.all() and np.array_equal(s.n_bc_iters, ref.n_bc_iters[bad]) and np.array_equal(s.status, ref.status[bad])
    # certify=True (host.Frontend._certify): the sets that respond to a one-ulp input change are found by the library
    # itself and re-solved strictly, after which EVERY live set — no exemption — is within 1e-9 with the oracle's control flow
    cert = gfe.sapdesolver_batch(Co, ens[:, :7], ens[:, 7:], out_mode=pkg.abi.OUT_FINAL_STATE, certify=True, **kw)
    rep_c = census(cert, ref, name="config2_2048_prior_draws_certified")
    rep_c["resolved_strict"] = [int(i) for i in cert.resolved_strict]
    check(rep_c)
    live_ill = set(np.flatnonzero(ill & ((ref.status & 1) == 0)).tolist())
    assert live_ill <= set(rep_c["resolved_strict"]), (live_ill, rep_c["resolved_strict"])
    assert len(cert.resolved_strict) <= 60, "the certification should re-solve a few per cent of the draws at most"
    # certify=True runs the C ABI's gab1_solve_batch_certified; the Python restatement of the same procedure agrees with it
    py = gfe.sapdesolver_batch(Co, ens[:, :7], ens[:, 7:], out_mode=pkg.abi.OUT_FINAL_STATE, **kw)
    gfe._certify(py, Co, ens[:, :7], ens[:, 7:])
    np.testing.assert_array_equal(py.resolved_strict, cert.resolved_strict)
    assert ((py.out.view(np.uint64) == cert.out.view(np.uint64)) | (np.isnan(py.out) & np.isnan(cert.out))).all()
    np.testing.assert_array_equal(py.n_bc_iters, cert.n_bc_iters)
    # ... and with another output mode (the final states are then solved on the side): the re-solved rows are the oracle's bits
    six_c = gfe.sapdesolver_batch(Co, ens[:, :7], ens[:, 7:], out_mode=pkg.abi.OUT_SIX, certify=True, **kw)
    np.testing.assert_array_equal(six_c.resolved_strict, cert.resolved_strict)
    six = gfe.sapdesolver_batch(Co, ens[:, :7], ens[:, 7:], out_mode=pkg.abi.OUT_SIX, **kw)
    six_ref = ofe.sapdesolver_batch(Co, ens[:, :7], ens[:, 7:], out_mode=pkg.abi.OUT_SIX, **kw)
    rs = cert.resolved_strict
    assert ((six_c.out[rs].view(np.uint64) == six_ref.out[rs].view(np.uint64)) | (np.isnan(six_c.out[rs]) & np.isnan(six_ref.out[rs]))).all()
    np.testing.assert_array_equal(six_c.status[rs], six_ref.status[rs])
    np.testing.assert_array_equal(six.status[~ill], six_ref.status[~ill])
    sel = ~ill & ((ref.status & 1) == 0)
    np.testing.assert_array_equal(six.n_bc_iters[sel], six_ref.n_bc_iters[sel])
    live = ((six_ref.status & (pkg.abi.ST_NAN | pkg.abi.ST_THROW)) == 0) & ~ill
    np.testing.assert_array_equal(six.out[live, :4], six_ref.out[live, :4])
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.abs(six.out[live, 4:] - six_ref.out[live, 4:]) / np.abs(six_ref.out[live, 4:])
    e = np.where(six.out[live, 4:] == six_ref.out[live, 4:], 0.0, e)
    assert np.nanmax(e) < RTOL, f"ratio / average: {np.nanmax(e):.3e}"
    # the fbatch_dk_mt surface itself (log-space columns in, 6 x S out, zeros(6) where the reference's catch fires)
    Y = gfe.fbatch_dk_mt(np.log(ens[:256].T))
    Yr = ofe.fbatch_dk_mt(np.log(ens[:256].T))
    assert Y.shape == (6, 256)
    ok = ~ill[:256]
    np.testing.assert_array_equal(np.isnan(Y[:, ok]), np.isnan(Yr[:, ok]))
    np.testing.assert_array_equal(Y[:4, ok], Yr[:4, ok])


def test_census_config3_hela_membSFK(pkg, gfe, ofe, ensemble):
    """configs[3]: sapdesolver_membSFK (`while error > tol`, sapdesolver_memb-SFK.jl:175-222) with the HeLa
    concentrations (run_base_model_HeLa.jl:71-83) on 256 rows of the shipped ensemble, tf = 5."""
    Co = pkg.params.hela_Co()
    rows = np.arange(0, 5000, 5000 // 256)[:256]
    kw = dict(dr=0.2, tf=5.0, tol=1e-3, membSFK=True, out_mode=pkg.abi.OUT_FINAL_STATE, iter_cap=2000)
    res = gfe.sapdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], **kw)
    ref = ofe.sapdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], **kw)
    check(census(res, ref, name="config3_256_hela_membSFK_final_state"))


def test_census_config4_rect_fine_grid(pkg, gfe, ofe, ensemble):
    """configs[4]: pdesolver_rect at dr = 0.05 (Nr = 200), tf = 5, full snapshot output, 16 rows (5.5e5 steps each;
    the oracle needs ~7 s per row and core)."""
    Co = pkg.params.base_Co()
    rows = np.arange(0, 5000, 5000 // 16)[:16]
    kw = dict(dr=0.05, tf=5.0, Nts=100, tol=1e-4, maxiters=20, geometry=pkg.abi.GEOM_RECT, pg1tot_form=pkg.abi.PG1TOT_CHAIN)
    res = gfe.pdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], **kw)
    ref = ofe.pdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], **kw)
    check(census(res, ref, name="config4_16_rows_rect_dr005_full_snapshots"))
