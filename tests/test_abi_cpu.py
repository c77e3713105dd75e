"""CPU-side checks of the boundary: the library loads and exports every symbol include/gab1pde.h declares,
the option struct matches the header, and the no-GPU failure mode is loud (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib(pkg):
    import __graft_entry__ as g
    g.build()
    return pkg.abi.load_library()


def test_exports_every_declared_symbol(lib):
    hdr = (ROOT / "include" / "gab1pde.h").read_text()
    names = set(re.findall(r"\b(gab1_[a-z0-9_]+)\s*\(", hdr))
    assert {"gab1_solve_batch", "gab1_solve_batch_device", "gab1_workspace_bytes", "gab1_default_dt"} <= names
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in gab1pde.h but not exported"


def test_struct_layout_matches_header(pkg, lib):
    o = pkg.abi.Opts()
    lib.gab1_opts_init(C.byref(o), 10.0, 0.2, 5.0, 100)
    assert (o.abi_version, o.Nr, o.Nts, o.maxiters) == (1, 50, 100, 100)
    assert (o.tol, o.R, o.dr, o.tf, o.dt_save, o.t_prechase) == (1e-6, 10.0, 0.2, 5.0, 0.05, -1.0)
    assert o.matrix_mask == 0xFFF and o.out_mode == pkg.abi.OUT_FULL
    for mode in range(5):
        o.out_mode = mode
        assert lib.gab1_out_doubles_per_set(C.byref(o)) == pkg.abi.out_doubles_per_set(o)
    o.out_mode = pkg.abi.OUT_FULL
    o.matrix_mask = pkg.abi.MASK_FITTING
    for m in range(12):
        assert lib.gab1_full_matrix_offset(C.byref(o), m) == pkg.abi.full_matrix_offset(o, m)
    for v in range(11):
        assert lib.gab1_full_vector_offset(C.byref(o), v) == pkg.abi.full_vector_offset(o, v)
    assert lib.gab1_version() == 1


def test_default_dt_matches_host_formula(pkg, lib, ensemble):
    D = np.ascontiguousarray(ensemble[:100, :7])
    k = np.ascontiguousarray(ensemble[:100, 7:])
    dt = np.zeros(100)
    dp = C.POINTER(C.c_double)
    assert lib.gab1_default_dt(100, D.ctypes.data_as(dp), k.ctypes.data_as(dp), 0.2, dt.ctypes.data_as(dp)) == 0
    np.testing.assert_array_equal(dt, pkg.params.default_dt(D, k, 0.2))
    assert np.ceil(5.0 / pkg.params.default_dt(ensemble[:, :7], ensemble[:, 7:], 0.2)).min() == 30688


def test_bad_options_are_rejected_with_a_message(pkg, lib):
    o = pkg.abi.make_opts(dr=0.2)
    o.abi_version = 99
    rc, *_ = pkg.abi.call_solve(lib.gab1_solve_batch, o, pkg.params.base_Co(), pkg.params.DIFFS_BASE, pkg.params.KVALS_BASE,
                                1e-4, pkg.params.julia_range(0.2, 10.0))
    assert rc < 0 and b"abi_version" in lib.gab1_last_error()


def test_no_gpu_means_loud_failure_not_cpu_fallback(pkg, lib):
    if lib.gab1_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.abi.Gab1Error, match="no CUDA device"):
        pkg.host.sapdesolver(pkg.params.base_Co(), pkg.params.DIFFS_BASE, pkg.params.KVALS_BASE, tf=0.01)


def test_certified_entry_point_validates_like_the_plain_one(pkg, lib):
    """gab1_solve_batch_certified (gab1pde.h): same option checks, an empty batch is a no-op that clears its counters, and without a
    GPU it fails as loudly as gab1_solve_batch — the certification is GPU work too (strict kernels), never the oracle."""
    Co, D, k, r = pkg.params.base_Co(), pkg.params.DIFFS_BASE, pkg.params.KVALS_BASE, pkg.params.julia_range(0.2, 10.0)
    n_res = C.c_int64(7)
    extra = (C.c_double(0.0), None, C.byref(n_res))
    o = pkg.abi.make_opts(dr=0.2)
    o.abi_version = 99
    rc, *_ = pkg.abi.call_solve(lib.gab1_solve_batch_certified, o, Co, D, k, 1e-4, r, *extra)
    assert rc < 0 and b"abi_version" in lib.gab1_last_error() and n_res.value == 0
    o = pkg.abi.make_opts(dr=0.2)
    n_res.value = 7
    rc, out, *_ = pkg.abi.call_solve(lib.gab1_solve_batch_certified, o, Co, np.zeros((0, 7)), np.zeros((0, 17)), np.zeros(0), r, *extra)
    assert rc == 0 and out.shape[0] == 0 and n_res.value == 0
    if lib.gab1_device_count() == 0:
        rc, *_ = pkg.abi.call_solve(lib.gab1_solve_batch_certified, o, Co, D, k, 1e-4, r, *extra)
        assert rc < 0 and b"no CUDA device" in lib.gab1_last_error()
        with pytest.raises(pkg.abi.Gab1Error, match="no CUDA device"):
            pkg.host.Frontend(pkg.abi.CudaBackend()).sapdesolver_batch(Co, D[None, :], k[None, :], tf=0.01, certify=True)


def test_julia_range_is_correctly_rounded(pkg):
    r = pkg.params.julia_range(0.2, 10.0)
    assert len(r) == 51 and r[3] == 0.6 and r[3] != 3 * 0.2 and r[-1] == 10.0     # SURVEY.md §7 "hard parts"
    assert len(pkg.params.julia_range(0.1, 10.0)) == 101 and len(pkg.params.julia_range(0.4, 100.0)) == 251
    with pytest.raises(IndexError):
        pkg.abi.call_solve(None, pkg.abi.make_opts(dr=0.3), pkg.params.base_Co(), pkg.params.DIFFS_BASE,
                           pkg.params.KVALS_BASE, 1e-4, pkg.params.julia_range(0.2, 10.0))
