"""The library's own multi-GPU path (gab1_solve_batch / gab1_solve_tangent with n_devices > 1: shards balanced by step count —
dealt for small per-set outputs, contiguous for full snapshots — one host thread and stream per device, no collective).  Needs at least two visible GPUs; skipped otherwise
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    """Bit-identical, NaN positions included; the sign and payload of a NaN are not compared: FP64 instructions propagate
    them from their operands, and a diverged set is recognised as all-NaN one step later on the two-warps-per-set lane than
    on the one-warp kernel (csrc/duo_kernel.cuh), which is the lane a small shard runs on."""
    a, b = np.asarray(a), np.asarray(b)
    return bool(((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))).all())


@pytest.fixture(scope="module")
def ndev(pkg):
    import __graft_entry__ as g
    g.build()
    n = pkg.abi.load_library().gab1_device_count()
    if n < 2:
        pytest.skip("needs two visible GPUs")
    return n


def test_sharded_solve_is_bitwise_the_single_device_solve(pkg, ndev, ensemble):
    Co = pkg.params.base_Co()
    D, k = ensemble[:3001, :7], ensemble[:3001, 7:]
    one = pkg.host.Frontend(pkg.abi.CudaBackend(n_devices=1))
    kw = dict(dr=0.2, tf=0.3, Nts=5, tol=1e-4, maxiters=20, matrices=("aSFK", "PG1Stot"))
    ref = one.pdesolver_batch(Co, D, k, **kw)
    for n in sorted({2, ndev}):
        res = pkg.host.Frontend(pkg.abi.CudaBackend(n_devices=n)).pdesolver_batch(Co, D, k, **kw)
        assert same_bits(res.out, ref.out), f"{n} devices"
        for f in ("status", "n_saved", "n_steps", "n_bc_iters"):
            np.testing.assert_array_equal(getattr(res, f), getattr(ref, f))
    six1 = one.sapdesolver_batch(Co, D, k, tf=0.3, out_mode=pkg.abi.OUT_SIX)
    six2 = pkg.host.Frontend(pkg.abi.CudaBackend(n_devices=2)).sapdesolver_batch(Co, D, k, tf=0.3, out_mode=pkg.abi.OUT_SIX)
    assert same_bits(six1.out, six2.out)


@pytest.mark.parametrize("plan", ["dealt", "contiguous"])
def test_both_shard_plans_scatter_every_set_to_its_own_row(pkg, ndev, ensemble, plan, monkeypatch):
    """Small per-set outputs are sharded by dealing the descending step-count order (gather, local batch, scatter); the
    contiguous plan is forced onto the same call for comparison.  Ragged work (a per-set dt that makes some sets 4x longer),
    per-set Co, every diagnostic array: each must come back in the caller's row order, bit-identical to one device."""
    monkeypatch.setenv("GAB1_SHARD_PLAN", plan)
    g = np.random.Generator(np.random.PCG64(3))
    D, k = ensemble[:1501, :7], ensemble[:1501, 7:]
    Co = pkg.params.base_Co()[None, :] * g.uniform(0.5, 2.0, size=(1501, 1))
    dt = pkg.params.default_dt(D, k, 0.2) * np.where(g.random(1501) < 0.2, 0.25, 1.0)
    kw = dict(dr=0.2, tf=0.3, dt=dt, out_mode=pkg.abi.OUT_FINAL_STATE)
    ref = pkg.host.Frontend(pkg.abi.CudaBackend(n_devices=1)).sapdesolver_batch(Co, D, k, **kw)
    for n in sorted({2, ndev}):
        res = pkg.host.Frontend(pkg.abi.CudaBackend(device_ids=list(range(n)))).sapdesolver_batch(Co, D, k, **kw)
        assert same_bits(res.out, ref.out), f"{plan}, {n} devices"
        for f in ("status", "n_saved", "n_steps", "n_bc_iters"):
            np.testing.assert_array_equal(getattr(res, f), getattr(ref, f))


def test_sharded_tangent_is_bitwise_the_single_device_tangent(pkg, ndev, ensemble):
    Co = pkg.params.base_Co()
    D, k = ensemble[:801, :7], ensemble[:801, 7:]
    seeds = np.zeros((801, 4, 30))
    for d in range(4):
        seeds[:, d, 7 + 6 + d] = 1.0
    kw = dict(dr=0.2, tf=0.1, Nts=3, tol=1e-4, maxiters=20, out_mode=pkg.abi.OUT_FINAL4)
    ref = pkg.host.Frontend(pkg.abi.CudaBackend(n_devices=1)).pdesolver_tangent_batch(Co, D, k, seeds, **kw)
    res = pkg.host.Frontend(pkg.abi.CudaBackend(n_devices=2)).pdesolver_tangent_batch(Co, D, k, seeds, **kw)
    assert np.array_equal(res.out.view(np.uint64), ref.out.view(np.uint64))
    np.testing.assert_array_equal(res.n_bc_iters, ref.n_bc_iters)
