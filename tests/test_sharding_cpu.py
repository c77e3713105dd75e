"""Host-side logic of the N > 1 path, on CPU: the shard plan of gab1_solve_batch, and a world_size-2 gloo run in which
each rank solves its shard (with the oracle standing in for the device) and the gathered result equals the single-rank one."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_plan_is_contiguous_balanced_and_complete(pkg, ensemble):
    import __graft_entry__ as g
    g.build()
    dt = pkg.params.default_dt(ensemble[:, :7], ensemble[:, 7:], 0.2)
    nt = np.ceil(5.0 / dt)
    for n in (1, 2, 3, 4, 8):
        b = pkg.abi.plan_shards(dt, 5.0, n)
        assert b[0] == 0 and b[-1] == len(dt) and np.all(np.diff(b) > 0)
        loads = np.array([nt[b[i]:b[i + 1]].sum() for i in range(n)])
        assert loads.max() / loads.mean() < 1.002            # within one set's work of perfect balance
    # ragged: one huge set dominates; more shards than sets; empty input
    dt2 = np.array([1e-6, 1e-3, 1e-3, 1e-3])
    b = pkg.abi.plan_shards(dt2, 1.0, 2)
    assert list(b) == [0, 1, 4]
    assert list(pkg.abi.plan_shards(np.array([1e-3, 1e-3]), 1.0, 4))[-1] == 2
    assert list(pkg.abi.plan_shards(np.zeros(0), 1.0, 2)) == [0, 0, 0]
    # unusable dt counts as one unit of work instead of poisoning the plan
    b = pkg.abi.plan_shards(np.array([np.nan, 0.0, 1e-3, 1e-3]), 1.0, 2)
    assert b[0] == 0 and b[-1] == 4 and np.all(np.diff(b) >= 0)


def test_dealt_plan_is_a_balanced_partition_with_the_same_mix_everywhere(pkg):
    """gab1_deal_shards (the plan for small per-set outputs, and the one bench.py's ranks use): a partition of the set
    indices, ascending inside a shard, equal total step counts, and the same distribution of solve lengths per shard."""
    import __graft_entry__ as g
    g.build()
    ens = pkg.params.synthetic_prior_ensemble(20000, seed=7)
    dt = pkg.params.default_dt(ens[:, :7], ens[:, 7:], 0.2)
    nt = np.ceil(5.0 / dt)
    for n in (1, 2, 3, 8):
        perm, b = pkg.abi.deal_shards(dt, 5.0, n)
        assert sorted(perm.tolist()) == list(range(len(dt)))
        assert b[0] == 0 and b[-1] == len(dt) and np.diff(b).max() - np.diff(b).min() <= max(2, len(dt) // (50 * n))
        loads, q90 = [], []
        for i in range(n):
            idx = perm[b[i]:b[i + 1]]
            assert np.all(np.diff(idx) > 0)
            loads.append(nt[idx].sum())
            q90.append(np.quantile(nt[idx], 0.9))
        assert max(loads) / np.mean(loads) < 1.0005
        assert max(q90) / min(q90) < 1.01
    # fewer sets than shards, empty input, unusable dt
    perm, b = pkg.abi.deal_shards(np.array([1e-3, 2e-3]), 1.0, 4)
    assert sorted(perm.tolist()) == [0, 1] and b[-1] == 2 and np.diff(b).max() == 1
    perm, b = pkg.abi.deal_shards(np.zeros(0), 1.0, 2)
    assert list(b) == [0, 0, 0]
    perm, b = pkg.abi.deal_shards(np.array([np.nan, 0.0, 1e-3, 1e-3]), 1.0, 2)
    assert sorted(perm.tolist()) == [0, 1, 2, 3]


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    import importlib
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
    from oracle import oracle
    ens = pkg.params.load_parameter_ensemble()[:12]
    Co = pkg.params.base_Co()
    dt = pkg.params.default_dt(ens[:, :7], ens[:, 7:], 0.4)
    dt[3] *= 0.25                                          # ragged work: the plan must move the boundary
    b = pkg.abi.plan_shards(dt, 0.2, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    fe = oracle.frontend(1)
    res = fe.sapdesolver_batch(Co, ens[lo:hi, :7], ens[lo:hi, 7:], dr=0.4, tf=0.2, dt=dt[lo:hi], out_mode=pkg.abi.OUT_SIX)
    mine = torch.zeros(12, 6, dtype=torch.float64)
    mine[lo:hi] = torch.from_numpy(res.out)
    dist.all_reduce(mine)                                  # disjoint shards: the sum is the gather
    t = torch.tensor([0.1 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)               # the max-over-ranks timing bench.py reports
    if rank == 0:
        q.put((mine.numpy(), float(t.item()), [int(x) for x in b]))
    dist.destroy_process_group()


def test_two_ranks_gloo_shards_gather_to_single_rank_result(pkg):
    import torch.multiprocessing as mp
    from oracle import oracle
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, tmax, bounds = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ens = pkg.params.load_parameter_ensemble()[:12]
    dt = pkg.params.default_dt(ens[:, :7], ens[:, 7:], 0.4)
    dt[3] *= 0.25
    ref = oracle.frontend(1).sapdesolver_batch(pkg.params.base_Co(), ens[:, :7], ens[:, 7:], dr=0.4, tf=0.2, dt=dt,
                                                out_mode=pkg.abi.OUT_SIX)
    np.testing.assert_array_equal(gathered, ref.out)
    assert tmax == pytest.approx(0.2)
    assert bounds[0] == 0 and bounds[2] == 12 and bounds[1] < 6     # the heavy set pulls the boundary left


def _ref_arm(env_extra, sets=None):
    import json
    import subprocess
    env = dict(os.environ, **env_extra)
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         env=env, capture_output=True, text=True, timeout=300, check=True).stdout
    return out, json


def test_reference_arm_under_torchrun_env():
    """bench.py --impl reference as the driver launches it for N > 1: rank 0 prints exactly one JSON line and uses every core of
    its affinity mask although torchrun exports OMP_NUM_THREADS=1; the other ranks print nothing and exit 0."""
    import __graft_entry__ as g
    g.build()
    out, json = _ref_arm({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0", "OMP_NUM_THREADS": "1"})
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["metric"] == "ensemble PDE solves/sec"
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0
    out, _ = _ref_arm({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1", "OMP_NUM_THREADS": "1"})
    assert out.strip() == ""
