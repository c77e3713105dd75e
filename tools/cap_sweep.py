#!/usr/bin/env python
"""configs[3] (sapdesolver_membSFK, HeLa concentrations): how many passes of `while error > tol`
(sapdesolver_memb-SFK.jl:175-222) does a step need when it converges at all?

The reference's loop has no iteration cap: a fixed point that never meets `tol` hangs the Julia solver.  The library
needs a finite cap (gab1_opts.maxiters with bc_loop = GAB1_BC_WHILE; status bit GAB1_ST_ITER_CAP).  This sweep solves the
same resampled ensemble under a range of caps and reports, per cap, the time of a pass and the set of flagged sets: the
default cap is chosen where that set stops shrinking (every set that converges does so below it), with a 10x margin.

  python tools/cap_sweep.py [--sets 20000] > gpurun_out/cap_sweep.jsonl
"""
import argparse
import importlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
PKG = "myers-furcht-et-al_gab1-shp2-pde-model_b200"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets", type=int, default=20000)
    ap.add_argument("--caps", default="20,40,80,160,320,640,1280,5000,20000,100000")
    ap.add_argument("--prior", action="store_true", help="wide synthetic prior draws instead of the resampled posterior ensemble")
    args = ap.parse_args()
    import __graft_entry__ as g
    g.build()
    pkg = importlib.import_module(PKG)
    abi, params = pkg.abi, pkg.params
    ens = params.synthetic_prior_ensemble(args.sets, seed=123) if args.prior else params.resampled_ensemble(args.sets, seed=123)
    Co = params.hela_Co()
    fe = pkg.host.Frontend(abi.CudaBackend())
    caps = sorted(int(c) for c in args.caps.split(","))
    fe.sapdesolver_batch(Co, ens[:256, :7], ens[:256, 7:], membSFK=True, iter_cap=20)
    runs = {}
    for cap in reversed(caps):
        t0 = time.perf_counter()
        res = fe.sapdesolver_batch(Co, ens[:, :7], ens[:, 7:], membSFK=True, iter_cap=cap)
        sec = time.perf_counter() - t0
        runs[cap] = (res, sec)
    ref, _ = runs[caps[-1]]
    ref_flag = (ref.status & abi.ST_ITER_CAP) != 0
    for cap in caps:
        res, sec = runs[cap]
        flag = (res.status & abi.ST_ITER_CAP) != 0
        clean = ~flag
        same = np.array_equal(res.out[clean].view(np.uint64), ref.out[clean].view(np.uint64)) and \
            np.array_equal(res.n_bc_iters[clean], ref.n_bc_iters[clean])
        live = clean & ((res.status & abi.ST_NAN) == 0)
        print(json.dumps({"cap": cap, "sets": args.sets, "ensemble": "prior" if args.prior else "resampled posterior",
                          "s_per_pass": sec, "solves_per_s": args.sets / sec,
                          "flagged_sets": int(flag.sum()), "flagged_beyond_largest_cap": int((flag & ~ref_flag).sum()),
                          "unflagged_sets_bitwise_equal_to_largest_cap": bool(same),
                          "nan_sets": int(((res.status & abi.ST_NAN) != 0).sum()),
                          "mean_passes_per_step_unflagged_live": float(res.n_bc_iters[live].sum() / max(res.n_steps[live].sum(), 1)),
                          "max_mean_passes_per_step_unflagged_live": float((res.n_bc_iters[live] / np.maximum(res.n_steps[live], 1)).max()) if live.any() else None}),
              flush=True)


if __name__ == "__main__":
    main()
