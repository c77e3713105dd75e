"""Latency lane (duo_kernel.cuh) on one GPU: (1) the longest members of the bench ensemble alone, one warp against two;
(2) shard g of the 8-way deal of the 10^5-set bench ensemble (what one rank of the 8-GPU run solves), with and without
the automatic split; (3) small batches.  Prints one JSON object per line."""
import importlib
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
abi = pkg.abi


def env(**kv):
    for k in ("GAB1_KERNEL", "GAB1_DUO", "GAB1_ISOLATE"):
        os.environ.pop(k, None)
    for k, v in kv.items():
        os.environ[k] = v


def timed(fe, Co, ens, reps=3, **kw):
    best, res = 1e9, None
    for _ in range(reps):
        t0 = time.perf_counter()
        res = fe.sapdesolver_batch(Co, ens[:, :7], ens[:, 7:], **kw)
        best = min(best, time.perf_counter() - t0)
    return best * 1e3, res


def same(a, b):
    eq = (a.out.view(np.uint64) == b.out.view(np.uint64)) | (np.isnan(a.out) & np.isnan(b.out))
    return bool(eq.all() and np.array_equal(a.n_bc_iters, b.n_bc_iters) and np.array_equal(a.status, b.status))


def main():
    fe = pkg.host.Frontend(abi.CudaBackend(arith=abi.ARITH_FAST))
    Co = pkg.params.base_Co()
    ens = pkg.params.synthetic_prior_ensemble(100_000, seed=123)
    dt = pkg.params.default_dt(ens[:, :7], ens[:, 7:], 0.2)
    nt = np.ceil(5.0 / dt)
    kw = dict(dr=0.2, tf=5.0, tol=1e-3, maxiters=20, out_mode=abi.OUT_SIX)
    order = np.argsort(-nt)
    # warm-up (context, arena)
    fe.sapdesolver_batch(Co, ens[:64, :7], ens[:64, 7:], **kw)

    which = sys.argv[1:] or ["alone", "shards", "small"]
    for n in (1, 16, 148) if "alone" in which else ():
        top = ens[order[:n]]
        env(GAB1_KERNEL="legacy"); ms1, r1 = timed(fe, Co, top, **kw)
        env(GAB1_KERNEL="duo"); ms2, r2 = timed(fe, Co, top, **kw)
        print(json.dumps({"probe": "longest_sets_alone", "sets": n, "steps_longest": int(nt[order[0]]),
                          "one_warp_ms": ms1, "two_warps_ms": ms2, "gain": ms1 / ms2, "bit_identical": same(r1, r2),
                          "cycles_per_step_one_warp": ms1 * 1e-3 * 1.965e9 / nt[order[0]],
                          "cycles_per_step_two_warps": ms2 * 1e-3 * 1.965e9 / nt[order[0]]}), flush=True)

    perm, bounds = abi.deal_shards(dt, 5.0, 8)
    perm4, bounds4 = abi.deal_shards(dt, 5.0, 4)
    shards = [("shard_0_of_8", perm[bounds[0]:bounds[1]], 1962.6 / 8), ("shard_3_of_8", perm[bounds[3]:bounds[4]], 1962.6 / 8),
              ("shard_0_of_4", perm4[bounds4[0]:bounds4[1]], 1962.6 / 4)]
    for name, idx, ideal in shards if "shards" in which else ():
        sh = ens[idx]
        rec = {"probe": name, "sets": int(len(idx)), "longest_steps": int(nt[idx].max()), "ideal_ms_from_1gpu": ideal}
        env(GAB1_DUO="0"); ms0, r0 = timed(fe, Co, sh, **kw)
        rec["one_warp_only_ms"] = ms0
        for label, e in (("lane_no_isolation_ms", dict(GAB1_ISOLATE="0")), ("lane_sm_reserved_ms", dict(GAB1_ISOLATE="sm")),
                         ("lane_scheduler_reserved_ms", dict(GAB1_ISOLATE="sched"))):
            env(**e); ms, r = timed(fe, Co, sh, **kw)
            rec[label] = ms
            rec["bit_identical"] = rec.get("bit_identical", True) and same(r0, r)
        print(json.dumps(rec), flush=True)

    # small batches of posterior rows (run_ensemble-sized and below), full length
    post = pkg.params.load_parameter_ensemble()
    kw1 = dict(dr=0.2, tf=5.0, tol=1e-4, maxiters=20, out_mode=abi.OUT_FINAL4)
    for n in (1, 64, 296, 592, 900, 1184, 2000) if "small" in which else ():
        rows = post[:n]
        env(GAB1_DUO="0"); ms1, r1 = timed(fe, Co, rows, **kw1)
        env(); ms2, r2 = timed(fe, Co, rows, **kw1)
        env(GAB1_KERNEL="duo"); ms3, r3 = timed(fe, Co, rows, **kw1)
        print(json.dumps({"probe": "small_batch_posterior", "sets": n, "one_warp_only_ms": ms1, "automatic_ms": ms2,
                          "all_two_warps_ms": ms3, "bit_identical": same(r1, r2) and same(r1, r3)}), flush=True)
    env()


if __name__ == "__main__":
    main()
