"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Needs a B200: run with -m gpu.

Two bars (north_star: relative tolerance 1e-9 in FP64):
  * arith = strict  -> BIT-IDENTICAL to the oracle, including iteration counts and snapshot schedule;
  * arith = fast    -> |gpu - ref| <= RTOL * max(|ref|, 1e-6 * max|field|) with RTOL = 1e-9, and identical
                       membrane-iteration counts, step counts and snapshot schedules (the data-dependent control flow).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(scope="module")
def gfe(pkg):
    import __graft_entry__ as g
    g.build()
    lib = pkg.abi.load_library()
    assert lib.gab1_device_count() >= 1, "no CUDA device: the product path has no CPU fallback"
    return pkg.host.Frontend(pkg.abi.CudaBackend(arith=pkg.abi.ARITH_FAST))


@pytest.fixture(scope="module")
def sfe(pkg, gfe):
    return pkg.host.Frontend(pkg.abi.CudaBackend(arith=pkg.abi.ARITH_STRICT))


def assert_bits(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    same = (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))
    assert same.all(), f"{what}: {np.count_nonzero(~same)}/{same.size} values differ, first at {np.argwhere(~same)[0]}"


def rel_err(a, b):
    """max over sets of |a-b| / max(|b|, 1e-6*max|b| per set); NaN positions must coincide."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    fin = np.isfinite(b)
    scale = np.where(fin, np.abs(b), 0).max(axis=-1, keepdims=True)
    den = np.maximum(np.abs(b), 1e-6 * scale)
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.where(fin & (den > 0), np.abs(a - b) / den, 0.0)
    return float(e.max())


def check_control_flow(res, ref):
    np.testing.assert_array_equal(res.n_steps, ref.n_steps)
    np.testing.assert_array_equal(res.n_saved, ref.n_saved)
    np.testing.assert_array_equal(res.n_bc_iters, ref.n_bc_iters)
    np.testing.assert_array_equal(res.status, ref.status)


VARIANTS = {
    "pdesolver": dict(),
    "membSFK": dict(sfk_mode=1),
    "rect": dict(geometry=1, pg1tot_form=1),
    "rect_frozen_modulus": dict(geometry=1, pg1tot_form=1, sfk_mode=2, save_rule=1),
    "pulsechase": dict(t_prechase=0.2),
    "fitting_mask": dict(matrices=("aSFK", "PG1S", "G2PG1S")),
}


@pytest.mark.parametrize("dr", [0.4, 0.2, 0.1, 0.05])
@pytest.mark.parametrize("variant", list(VARIANTS))
def test_strict_is_bit_identical_full(pkg, sfe, ofe, ensemble, variant, dr):
    """All four lane layouts (K = 1, 2, 4, 8 nodes per lane) and every solver variant, full snapshot output."""
    Co = pkg.params.base_Co()
    rows = [0, 1, 2, 3, 75]
    tf = {0.4: 0.3, 0.2: 0.3, 0.1: 0.1, 0.05: 0.03}[dr]
    kw = dict(dr=dr, tf=tf, Nts=6, tol=1e-4, maxiters=20, **VARIANTS[variant])
    res = sfe.pdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], **kw)
    ref = ofe.pdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], **kw)
    check_control_flow(res, ref)
    assert_bits(res.out, ref.out, f"{variant} dr={dr}")


@pytest.mark.parametrize("dr", [0.4, 0.2, 0.1, 0.05])
@pytest.mark.parametrize("variant", list(VARIANTS))
def test_fast_within_tolerance_full(pkg, gfe, ofe, ensemble, variant, dr):
    Co = pkg.params.base_Co()
    rows = [0, 1, 2, 3, 4]
    tf = {0.4: 0.6, 0.2: 0.5, 0.1: 0.2, 0.05: 0.05}[dr]
    kw = dict(dr=dr, tf=tf, Nts=8, tol=1e-4, maxiters=20, **VARIANTS[variant])
    res = gfe.pdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], **kw)
    ref = ofe.pdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], **kw)
    check_control_flow(res, ref)
    e = rel_err(res.out, ref.out)
    assert e < RTOL, f"{variant} dr={dr}: rel err {e:.3e}"


@pytest.mark.parametrize("arith", ["strict", "fast"])
@pytest.mark.parametrize("membSFK", [False, True])
def test_final_time_outputs(pkg, gfe, sfe, ofe, ensemble, membSFK, arith):
    """sapdesolver / sapdesolver_membSFK final profiles, the six GSA scalars and the full final state (config 3/4 path)."""
    fe = sfe if arith == "strict" else gfe
    Co = pkg.params.hela_Co() if membSFK else pkg.params.base_Co()
    rows = list(range(10, 26))
    for mode in (pkg.abi.OUT_FINAL4, pkg.abi.OUT_SIX, pkg.abi.OUT_FINAL_STATE):
        kw = dict(dr=0.2, tf=1.0, membSFK=membSFK, out_mode=mode)
        res = fe.sapdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], **kw)
        ref = ofe.sapdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], **kw)
        check_control_flow(res, ref)
        if arith == "strict":
            assert_bits(res.out, ref.out, f"mode {mode}")
        elif mode == pkg.abi.OUT_SIX:
            # length scales are grid-quantised: must be identical; ratio and average within tolerance
            np.testing.assert_array_equal(res.out[:, :4], ref.out[:, :4])
            assert rel_err(res.out[:, 4:], ref.out[:, 4:]) < RTOL
        else:
            assert rel_err(res.out, ref.out) < RTOL


def test_full_length_solves_config2_sample(pkg, gfe, ofe, ensemble):
    """BASELINE config 2 settings (run_ensemble defaults: dr=0.2, tf=5, Nts=100, tol=1e-4, maxit=20) on a sample that
    includes a diverging set (row 75): ~3.7e4 steps each, every snapshot compared."""
    Co = pkg.params.base_Co()
    rows = [0, 1, 2, 75, 333, 4999, 2500, 1234]
    kw = dict(dr=0.2, tol=1e-4, maxiters=20)
    res = gfe.pdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], **kw)
    ref = ofe.pdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], **kw)
    np.testing.assert_array_equal(res.n_steps, ref.n_steps)
    np.testing.assert_array_equal(res.n_saved, ref.n_saved)
    good = (ref.status & pkg.abi.ST_NAN) == 0
    assert list(good) == [True, True, True, False, False, True, True, True]
    np.testing.assert_array_equal(res.status & pkg.abi.ST_NAN, ref.status & pkg.abi.ST_NAN)
    np.testing.assert_array_equal(res.n_bc_iters[good], ref.n_bc_iters[good])
    e = rel_err(res.out[good], ref.out[good])
    assert e < RTOL, f"rel err {e:.3e}"


def test_single_solve_config1(pkg, gfe, sfe, ofe):
    """BASELINE config 1: run_base_model.jl:83 — pdesolver(Co, Diffs, kvals; R, dr=0.1, tf=5, Nts=100, tol=1e-2)."""
    Co, D, k = pkg.params.base_Co(), pkg.params.DIFFS_BASE, pkg.params.KVALS_BASE
    sol_r, r_r, t_r, dt_r = ofe.pdesolver(Co, D, k, dr=0.1, tol=1e-2)
    sol_g, r_g, t_g, dt_g = gfe.pdesolver(Co, D, k, dr=0.1, tol=1e-2)
    assert dt_g == dt_r and np.array_equal(t_g, t_r) and np.array_equal(r_g, r_r)
    for name in sol_r._fields:
        e = rel_err(getattr(sol_g, name).T, getattr(sol_r, name).T)
        assert e < RTOL, f"{name}: {e:.3e}"
    sol_s = sfe.pdesolver(Co, D, k, dr=0.1, tol=1e-2)[0]
    for name in sol_r._fields:
        assert_bits(np.ascontiguousarray(getattr(sol_s, name)), np.ascontiguousarray(getattr(sol_r, name)), name)


def test_golden_kat(pkg, gfe, sfe, ensemble):
    """Committed known-answer vectors (tests/golden/oracle_kat.npz, written by the oracle in the build container)."""
    from pathlib import Path
    kat = np.load(Path(__file__).parent / "golden" / "oracle_kat.npz")
    sub = ensemble[kat["rows"]]
    Co = pkg.params.base_Co()
    calls = [
        ("full_dr04_tf1", lambda fe: fe.pdesolver_batch(Co, sub[:, :7], sub[:, 7:], dr=0.4, tf=1.0, Nts=10, tol=1e-4, maxiters=20)),
        ("final4_dr02_tf05", lambda fe: fe.sapdesolver_batch(Co, sub[:, :7], sub[:, 7:], dr=0.2, tf=0.5)),
        ("final4_memb_dr02_tf05", lambda fe: fe.sapdesolver_batch(pkg.params.hela_Co(), sub[:, :7], sub[:, 7:], dr=0.2, tf=0.5, membSFK=True)),
        ("six_dr02_tf05", lambda fe: fe.sapdesolver_batch(Co, sub[:, :7], sub[:, 7:], dr=0.2, tf=0.5, out_mode=pkg.abi.OUT_SIX)),
        ("full_rect_dr025_tf05", lambda fe: fe.pdesolver_batch(Co, sub[:, :7], sub[:, 7:], dr=0.25, tf=0.5, Nts=5, tol=1e-4,
                                                              maxiters=20, geometry=pkg.abi.GEOM_RECT, pg1tot_form=pkg.abi.PG1TOT_CHAIN)),
    ]
    for name, call in calls:
        assert_bits(call(sfe).out, kat[name], name)
        if name.startswith("six"):
            np.testing.assert_array_equal(call(gfe).out[:, :4], kat[name][:, :4])
        else:
            assert rel_err(call(gfe).out, kat[name]) < RTOL, name
    np.testing.assert_array_equal(calls[0][1](gfe).n_bc_iters, kat["full_dr04_tf1_nbc"])


def test_pct_bound_and_julia_surface(pkg, gfe, ofe, ensemble):
    Co = pkg.params.base_Co()
    pct_g, _ = gfe.pct_shp2_bound_gab1(Co, ensemble[:32, :7], ensemble[:32, 7:])
    pct_r, _ = ofe.pct_shp2_bound_gab1(Co, ensemble[:32, :7], ensemble[:32, 7:])
    assert rel_err(pct_g[None], pct_r[None]) < RTOL
    assert abs(pct_g[0] - 23.46) < 0.01 and abs(pct_g[1] - 26.05) < 0.01          # SURVEY.md §4 probe values
    # fbatch_dk_mt semantics: log-space 24 x S in, 6 x S out, zeros(6) for a column that throws (sapdesolver.jl:371-387)
    pb = np.log(ensemble[70:80].T)
    out_g = gfe.fbatch_dk_mt(pb, tf=2.0)
    out_r = ofe.fbatch_dk_mt(pb, tf=2.0)
    assert out_g.shape == (6, 10)
    np.testing.assert_array_equal(out_g[:4], out_r[:4])
    assert rel_err(out_g[4:].T, out_r[4:].T) < RTOL
    # run_ensemble drops NaN sets and keeps 1-based indices (get_param_posteriors.jl:155-161)
    rows_g = gfe.run_ensemble("pdesolver", ensemble[74:77], Co, Nts=10)
    assert [x.index for x in rows_g] == [1, 3]
    assert rows_g[0].sol.PG1S.shape == (51, 11) and rows_g[0].t_sol.shape == (11,)


def test_edge_cases(pkg, gfe, sfe, ofe, ensemble):
    Co = pkg.params.base_Co()
    D, k = ensemble[:3, :7], ensemble[:3, 7:]
    # empty batch
    res = gfe.pdesolver_batch(Co, np.zeros((0, 7)), np.zeros((0, 17)), dr=0.4, tf=0.1, Nts=2)
    assert res.out.shape[0] == 0
    # per-set Co (S x 5), ragged work (very different dt per set), a dt so large that no snapshot is due (SHORT)
    Cos = np.stack([Co, pkg.params.hela_Co(), Co * 0.5])
    dt = np.array([1e-4, 3e-4, 2.5e-3])
    kw = dict(dr=0.4, tf=0.2, Nts=4, dt=dt, tol=1e-4, maxiters=20)
    res, ref = sfe.pdesolver_batch(Cos, D, k, **kw), ofe.pdesolver_batch(Cos, D, k, **kw)
    check_control_flow(res, ref)
    assert_bits(res.out, ref.out)
    # dt_save smaller than dt: more snapshots are due than columns exist (OVERFLOW), identical handling
    kw = dict(dr=0.4, tf=0.05, Nts=40, dt_save=1e-5, tol=1e-4, maxiters=20)
    res, ref = sfe.pdesolver_batch(Co, D, k, **kw), ofe.pdesolver_batch(Co, D, k, **kw)
    assert (ref.status & pkg.abi.ST_OVERFLOW).all()
    check_control_flow(res, ref)
    assert_bits(res.out, ref.out)
    # unusable dt -> the reference throws InexactError; reported per set, zeros written
    kw = dict(dr=0.4, tf=0.05, Nts=4, dt=np.array([0.0, np.nan, 1e-4]))
    res, ref = gfe.pdesolver_batch(Co, D, k, **kw), ofe.pdesolver_batch(Co, D, k, **kw)
    assert list(res.status[:2] & pkg.abi.ST_THROW) == [pkg.abi.ST_THROW] * 2 and not res.out[:2].any()
    np.testing.assert_array_equal(res.status, ref.status)
    # maxiters = 0: the membrane loop body never runs
    kw = dict(dr=0.4, tf=0.05, Nts=4, maxiters=0)
    res, ref = sfe.pdesolver_batch(Co, D, k, **kw), ofe.pdesolver_batch(Co, D, k, **kw)
    assert_bits(res.out, ref.out)
    assert rel_err(gfe.pdesolver_batch(Co, D, k, **kw).out, ref.out) < RTOL
    # largest grid one warp holds: Nr = 250 (R = 100, dr = 0.4; length_scale_estimates.jl:55-56)
    kw = dict(R=100.0, dr=0.4, tf=0.05, Nts=2, tol=1e-4, maxiters=20)
    res, ref = sfe.pdesolver_batch(pkg.params.base_Co(100.0), D, k, **kw), ofe.pdesolver_batch(pkg.params.base_Co(100.0), D, k, **kw)
    check_control_flow(res, ref)
    assert_bits(res.out, ref.out)


def test_whole_ensemble_properties_config2(pkg, gfe, ensemble):
    """All 5000 rows of parameter_ensemble.csv at config-2 settings; size-independent properties only:
    exactly the 33 diverging rows the oracle finds, EGFR conservation and SFK conservation on the rest."""
    Co = pkg.params.base_Co()
    res = gfe.pdesolver_batch(Co, ensemble[:, :7], ensemble[:, 7:], dr=0.2, tol=1e-4, maxiters=20,
                              out_mode=pkg.abi.OUT_FINAL_STATE)
    bad = np.nonzero(res.status & pkg.abi.ST_NAN)[0]
    assert len(bad) == 33 and bad[0] == 75 and bad[-1] == 4656          # SURVEY.md §6 (33/5000), oracle run in DESIGN.md
    good = res.out[(res.status & pkg.abi.ST_NAN) == 0]
    P = 51
    m = good[:, 10 * P:]
    tot = m[:, 0] + m[:, 1] + 2 * m[:, 2:].sum(axis=1)
    assert np.abs(tot / Co[4] - 1).max() < 1e-11
    assert np.abs((good[:, :P] + good[:, P:2 * P]) / Co[0] - 1).max() < 1e-11
    ratio = res.n_bc_iters / res.n_steps
    assert 1.4 < np.median(ratio) < 1.6                                 # 1.51 measured with the oracle


def test_reciprocal_accuracy(pkg, gfe):
    """The Robin closures divide through a hardware reciprocal seed plus one cubic correction; its error must stay at
    the rounding level over the whole range the denominators 1 + kf*M*dr/D can take."""
    import ctypes as C
    lib = pkg.abi.load_library()
    seed, rec = C.c_double(), C.c_double()
    assert lib.gab1_debug_recip_error(0, 1e-3, 1e12, C.byref(seed), C.byref(rec)) == 0
    print(f"seed rel err {seed.value:.3e}, reciprocal rel err {rec.value:.3e}")
    assert seed.value < 2.0 ** -18
    assert rec.value < 4 * 2.0 ** -53


def test_pinned_output_is_written_in_place(pkg, gfe, ensemble):
    """`out` in pinned, device-mapped host memory: the kernels store the snapshots straight into it (no device copy,
    no D2H phase).  Must equal the staged (pageable) path bit for bit, including never-due columns (zero-filled) and
    a set with an unusable dt (whole block zero); the buffer is pre-filled with garbage to prove every element is written."""
    import ctypes as C
    lib = pkg.abi.load_library()
    Co = pkg.params.base_Co()
    D, k = np.ascontiguousarray(ensemble[:6, :7]), np.ascontiguousarray(ensemble[:6, 7:])
    dt = pkg.params.default_dt(D, k, 0.4)
    dt[2] = np.nan              # THROW
    dt[4] = 0.03                # so coarse that most snapshot columns are never due (SHORT)
    o = pkg.abi.make_opts(dr=0.4, tf=0.2, Nts=8, tol=1e-4, maxiters=20)
    r = pkg.params.julia_range(0.4, 10.0)
    rc, staged, st_s, *_ = pkg.abi.call_solve(lib.gab1_solve_batch, o, Co, D, k, dt, r)
    assert rc == 0 and (st_s[4] & pkg.abi.ST_SHORT) and (st_s[2] & pkg.abi.ST_THROW)
    n = pkg.abi.out_doubles_per_set(o)
    p, pinned = pkg.abi.pinned_empty(6 * n)
    try:
        pinned[:] = -12345.678
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        status = np.zeros(6, np.int32)
        rc = lib.gab1_solve_batch(C.byref(o), 6, Co.ctypes.data_as(dp), 0, D.ctypes.data_as(dp), k.ctypes.data_as(dp),
                                  dt.ctypes.data_as(dp), r.ctypes.data_as(dp), pinned.ctypes.data_as(dp),
                                  status.ctypes.data_as(ip), None, None, None)
        assert rc == 0, lib.gab1_last_error()
        assert_bits(pinned.reshape(6, n), staged, "pinned vs staged")
        np.testing.assert_array_equal(status, st_s)
    finally:
        pkg.abi.pinned_free(p)


FAMILIES = {
    # GAB1_KERNEL value -> grids (dr) it can hold; every family the library ships is checked against the oracle,
    # not only the one the dispatcher picks for a grid (gab1pde.cu: pick_fast)
    "legacy": [0.4, 0.2, 0.1],
    "group16": [0.4, 0.25, 0.2],
    "group32": [0.2, 0.1, 0.05],
    # state in shared memory (stream_kernel.cuh): K = 4, 8, 16 nodes per lane; dr = 0.025 (Nr = 400) exists only here
    "stream": [0.2, 0.1, 0.05, 0.025],
    # latency path: one CTA of 2 / 4 / 7 warps per set (team_kernel.cuh)
    "team": [0.25, 0.2, 0.1, 0.05],
    # latency lane of the register-resident grids: two warps per set (duo_kernel.cuh), K = 2 / 2 / 4
    "duo": [0.25, 0.2, 0.1],
    # several sets per warp, G lanes per set, the state in shared memory (gang_kernel.cuh): the default for Nr > 128;
    # G = 2 / 4 / 4 / 8 / 16 / 32 on these grids
    "gang": [0.4, 0.25, 0.2, 0.1, 0.05, 0.025],
}


@pytest.mark.parametrize("family", list(FAMILIES))
def test_every_kernel_family_matches_the_oracle(pkg, gfe, ofe, ensemble, family, monkeypatch):
    """Same bar as test_fast_within_tolerance_full for each kernel family forced through GAB1_KERNEL: values within
    1e-9, identical step counts, snapshot schedules, membrane-iteration counts and status words.  An odd number of sets
    (7) leaves one half-warp of the two-sets-per-warp kernels without a partner.  (Diverging rows are left to
    test_full_length_solves_config2_sample: on an unstable trajectory rounding differences grow exponentially, so
    iteration counts are only comparable while a set is stable.)"""
    monkeypatch.setenv("GAB1_KERNEL", family)
    Co = pkg.params.base_Co()
    rows = [0, 1, 2, 3, 4, 2500, 4999]
    for dr in FAMILIES[family]:
        tf = {0.4: 0.6, 0.25: 0.5, 0.2: 0.5, 0.1: 0.15, 0.05: 0.04, 0.025: 0.01}[dr]
        for variant in ("pdesolver", "rect", "pulsechase", "membSFK"):
            kw = dict(dr=dr, tf=tf, Nts=7, tol=1e-4, maxiters=20, **VARIANTS[variant])
            res = gfe.pdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], **kw)
            ref = ofe.pdesolver_batch(Co, ensemble[rows, :7], ensemble[rows, 7:], **kw)
            np.testing.assert_array_equal(res.n_steps, ref.n_steps)
            np.testing.assert_array_equal(res.n_saved, ref.n_saved)
            np.testing.assert_array_equal(res.status, ref.status)
            good = (ref.status & pkg.abi.ST_NAN) == 0
            np.testing.assert_array_equal(res.n_bc_iters[good], ref.n_bc_iters[good])
            e = rel_err(res.out[good], ref.out[good])
            assert e < RTOL, f"{family} {variant} dr={dr}: rel err {e:.3e}"
        # the while-loop form and the final-time reductions (sapdesolver_membSFK / fbatch semantics)
        for mode in (pkg.abi.OUT_FINAL4, pkg.abi.OUT_SIX):
            kw = dict(dr=dr, tf=tf, membSFK=True, out_mode=mode)
            res = gfe.sapdesolver_batch(pkg.params.hela_Co(), ensemble[rows, :7], ensemble[rows, 7:], **kw)
            ref = ofe.sapdesolver_batch(pkg.params.hela_Co(), ensemble[rows, :7], ensemble[rows, 7:], **kw)
            check_control_flow(res, ref)
            if mode == pkg.abi.OUT_SIX:
                np.testing.assert_array_equal(res.out[:, :4], ref.out[:, :4])
                assert rel_err(res.out[:, 4:], ref.out[:, 4:]) < RTOL
            else:
                assert rel_err(res.out, ref.out) < RTOL


def test_duo_kernel_is_bit_identical_to_the_one_warp_kernel(pkg, gfe, ensemble, monkeypatch):
    """The dispatcher sends any subset of a batch to the two-warps-per-set kernel (gab1pde.cu: duo_plan_kernel), which is only
    sound because that kernel evaluates the one-set-per-warp kernel's expressions in the same order: values, NaN patterns,
    iteration counts and status words must agree BIT FOR BIT — on posterior rows, on wide prior draws (diverging sets, sets that
    run into the iteration limit step after step), in both loop forms, on every output mode and on a K = 4 grid."""
    abi = pkg.abi
    Co = pkg.params.base_Co()
    prior = pkg.params.synthetic_prior_ensemble(192, seed=77)
    rows = np.r_[0:24, 2500:2508, 4990:5000]
    cases = []
    for variant in ("pdesolver", "rect", "pulsechase", "membSFK", "rect_frozen_modulus", "fitting_mask"):
        cases.append(("pdesolver_batch", Co, ensemble[rows], dict(dr=0.2, tf=0.7, Nts=9, tol=1e-4, maxiters=20, **VARIANTS[variant])))
    cases.append(("pdesolver_batch", Co, ensemble[rows[:12]], dict(dr=0.1, tf=0.2, Nts=5, tol=1e-4, maxiters=20)))
    cases.append(("pdesolver_batch", Co, ensemble[rows], dict(dr=0.2, tf=1.0, Nts=10, tol=1e-4, maxiters=20, out_mode=abi.OUT_PCT_BOUND)))
    cases.append(("pdesolver_batch", Co, prior, dict(dr=0.2, tf=5.0, Nts=4, tol=1e-4, maxiters=20)))
    for mode in (abi.OUT_SIX, abi.OUT_FINAL4, abi.OUT_FINAL_STATE):
        cases.append(("sapdesolver_batch", Co, prior, dict(dr=0.2, tf=5.0, out_mode=mode)))
        cases.append(("sapdesolver_batch", pkg.params.hela_Co(), prior, dict(dr=0.2, tf=2.0, membSFK=True, out_mode=mode)))
    cases.append(("sapdesolver_batch", Co, prior[:48], dict(dr=0.1, tf=0.5, out_mode=abi.OUT_FINAL_STATE)))
    for fn, co, ens, kw in cases:
        got = {}
        for family in ("legacy", "duo"):
            monkeypatch.setenv("GAB1_KERNEL", family)
            got[family] = getattr(gfe, fn)(co, ens[:, :7], ens[:, 7:], **kw)
        a, b = got["legacy"], got["duo"]
        assert_bits(a.out, b.out, f"{fn} {kw}")
        check_control_flow(b, a)
    # the automatic split (the head of the descending-work queue on two warps, the rest on one) changes nothing either
    monkeypatch.delenv("GAB1_KERNEL")
    big = pkg.params.synthetic_prior_ensemble(1500, seed=78)
    auto = gfe.sapdesolver_batch(Co, big[:, :7], big[:, 7:], dr=0.2, tf=5.0, out_mode=abi.OUT_SIX)
    monkeypatch.setenv("GAB1_DUO", "0")
    plain = gfe.sapdesolver_batch(Co, big[:, :7], big[:, 7:], dr=0.2, tf=5.0, out_mode=abi.OUT_SIX)
    assert_bits(auto.out, plain.out, "automatic split")
    check_control_flow(auto, plain)


def test_finest_grid_all_outputs_and_edge_cases(pkg, gfe, ofe, ensemble):
    """dr = 0.025 (Nr = 400, the reference's finest grid: SURVEY §8a4 stretch) runs on the shared-memory-resident kernel with
    16 nodes per lane; Nr = 250 (length_scale_estimates.jl: R = 100, dr = 0.4) with 8.  FINAL_STATE, PCT_BOUND, per-set Co,
    ragged dt, snapshot overflow, unusable dt, a grid too large for the build."""
    abi = pkg.abi
    rows = [0, 1, 2, 4999, 75]
    D, k = ensemble[rows, :7], ensemble[rows, 7:]
    Co = np.tile(pkg.params.base_Co(), (5, 1)) * np.array([[1.0], [0.5], [2.0], [1.0], [1.0]])
    volCF, surfCF = pkg.params.conversion_factors()
    for dr, R, tf in ((0.025, 10.0, 0.004), (0.4, 100.0, 0.6)):
        dt = pkg.params.default_dt(D, k, dr) * np.array([1.0, 0.9, 0.8, 1.0, 0.0])       # the last set: Int64(ceil(tf/0)) throws
        for om in (abi.OUT_FINAL_STATE, abi.OUT_PCT_BOUND, abi.OUT_FULL):
            kw = dict(R=R, dr=dr, tf=tf, Nts=4, dt=dt, tol=1e-4, maxiters=20, out_mode=om, pct_mul=volCF, pct_div=surfCF)
            res = gfe.pdesolver_batch(Co, D, k, **kw)
            ref = ofe.pdesolver_batch(Co, D, k, **kw)
            check_control_flow(res, ref)
            assert res.status[4] & abi.ST_THROW
            assert rel_err(res.out, ref.out) < RTOL
    kw = dict(dr=0.025, tf=0.004, Nts=3, dt_save=0.0002, tol=1e-4, maxiters=20)
    res = gfe.pdesolver_batch(Co[:2], D[:2], k[:2], **kw)
    ref = ofe.pdesolver_batch(Co[:2], D[:2], k[:2], **kw)
    assert np.all(res.status & abi.ST_OVERFLOW)
    check_control_flow(res, ref)
    assert rel_err(res.out, ref.out) < RTOL
    with pytest.raises(abi.Gab1Error, match="512"):
        gfe.pdesolver_batch(Co[:1], D[:1], k[:1], dr=0.01, tf=1e-5, Nts=2)
    strict = pkg.host.Frontend(abi.CudaBackend(arith=abi.ARITH_STRICT))
    with pytest.raises(abi.Gab1Error, match="strict"):
        strict.pdesolver_batch(Co[:1], D[:1], k[:1], dr=0.025, tf=1e-5, Nts=2)


def test_output_larger_than_the_device_budget_is_solved_in_pieces(pkg, gfe, ensemble, monkeypatch):
    """A staged output that does not fit the device (3e5 full solutions fill a B200) is solved piecewise by the host entry
    point; GAB1_MAX_DEVICE_BYTES shrinks the budget so that 96 sets x 72 KB need 16 pieces.  Results must not change."""
    Co = pkg.params.base_Co()
    kw = dict(dr=0.2, tf=0.2, Nts=12, tol=1e-4, maxiters=20)
    D, k = ensemble[:96, :7], ensemble[:96, 7:]
    whole = gfe.pdesolver_batch(Co, D, k, **kw)
    monkeypatch.setenv("GAB1_MAX_DEVICE_BYTES", str(6 * 72 * 1024))
    pieces = gfe.pdesolver_batch(Co, D, k, **kw)
    assert_bits(pieces.out, whole.out, "piecewise")
    check_control_flow(pieces, whole)


@pytest.mark.parametrize("family", ["", "legacy", "group16", "group32", "stream", "team", "gang"])
def test_runs_are_bitwise_repeatable(pkg, gfe, ensemble, family, monkeypatch):
    """compute-sanitizer is not available on this pool; a data race between lanes or warps (exchange headers, staged rows, the
    work queue) would show up as run-to-run differences.  2500 sets (more than two waves, every warp of a CTA busy, the dynamic
    queue in a different order each time) solved twice per kernel family must agree bit for bit."""
    monkeypatch.setenv("GAB1_KERNEL", family)
    Co = pkg.params.base_Co()
    rows = np.arange(2500)
    for dr, tf in ((0.4, 0.05), (0.2, 0.05), (0.05, 0.002)):
        if (family == "group16" and dr < 0.2) or (family == "legacy" and dr < 0.1) or (family in ("stream", "team") and dr > 0.2):
            continue
        sub = rows if dr >= 0.2 else rows[:1200]
        kw = dict(dr=dr, tf=tf, Nts=3, tol=1e-4, maxiters=20, matrices=("aSFK", "PG1Stot"))
        a = gfe.pdesolver_batch(Co, ensemble[sub, :7], ensemble[sub, 7:], **kw)
        b = gfe.pdesolver_batch(Co, ensemble[sub, :7], ensemble[sub, 7:], **kw)
        assert_bits(a.out, b.out, f"{family or 'default'} dr={dr}")
        check_control_flow(a, b)


@pytest.mark.parametrize("family", ["legacy", "group16", "group32", "stream", "team", "gang"])
def test_blow_up_takes_the_dead_state_path_in_every_family(pkg, gfe, ofe, ensemble, family, monkeypatch):
    """A time step 2.5x the stability limit blows every set up within a few hundred steps: values overflow, turn NaN, and the
    kernels' all-NaN fast-forward (clock and snapshot schedule only) takes over.  Status words, step counts, snapshot counts
    and the final column must be what the oracle produces by brute force."""
    monkeypatch.setenv("GAB1_KERNEL", family)
    Co = pkg.params.base_Co()
    rows = [0, 1, 2, 3, 4999]
    D, k = ensemble[rows, :7], ensemble[rows, 7:]
    for dr in {"legacy": (0.2, 0.1), "group16": (0.4, 0.2), "group32": (0.2, 0.05), "stream": (0.2, 0.05, 0.025), "team": (0.1, 0.05), "gang": (0.2, 0.05, 0.025)}[family]:
        dt = 2.5 * pkg.params.default_dt(D, k, dr)
        kw = dict(dr=dr, tf=3000 * float(dt.max()), Nts=6, dt=dt, tol=1e-4, maxiters=20)
        res = gfe.pdesolver_batch(Co, D, k, **kw)
        ref = ofe.pdesolver_batch(Co, D, k, **kw)
        assert np.all(ref.status & pkg.abi.ST_NAN), "the oracle did not blow up: the test is not testing anything"
        np.testing.assert_array_equal(res.status, ref.status)
        np.testing.assert_array_equal(res.n_steps, ref.n_steps)
        np.testing.assert_array_equal(res.n_saved, ref.n_saved)
        last_g, last_r = res.matrix("PG1S")[:, :, -1], ref.matrix("PG1S")[:, :, -1]
        np.testing.assert_array_equal(np.isnan(last_g), np.isnan(last_r))
        np.testing.assert_array_equal(res.vector("t_out"), ref.vector("t_out"))
