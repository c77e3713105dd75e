"""Throughput of the forward-mode path on the fitting workload: S parameter sets x 4 partials
(kG1p, kG1dp, kSa, kSi), run_ensemble grid dr = 0.2, tf = 5, % SHP2-bound GAB1 output — one loss-and-gradient evaluation
per set, as LBFGS multistart / NUTS would request them (param_fitting+inference_finitediff.jl:188-240,308-370).
Prints one JSON line per configuration; GAB1_TANGENT_NT selects directions per work item for A/B runs."""
import argparse
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets", type=int, default=4736)
    ap.add_argument("--dr", type=float, default=0.2)
    ap.add_argument("--tf", type=float, default=5.0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--ndir", type=int, default=4)
    a = ap.parse_args()
    ens = pkg.params.load_parameter_ensemble()
    ens = ens[np.arange(a.sets) % ens.shape[0]]
    Co = pkg.params.base_Co()
    volCF, surfCF = pkg.params.conversion_factors()
    seeds = np.zeros((a.sets, a.ndir, 30))
    for d in range(a.ndir):
        seeds[:, d, 7 + 6 + d % 4] = 1.0
    fe = pkg.host.Frontend(pkg.abi.CudaBackend())
    kw = dict(dr=a.dr, tf=a.tf, Nts=100, tol=1e-3, maxiters=20, out_mode=pkg.abi.OUT_PCT_BOUND, pct_mul=volCF, pct_div=surfCF)
    fe.pdesolver_tangent_batch(Co, ens[:256, :7], ens[:256, 7:], seeds[:256], **kw)      # warm-up
    best = 1e30
    for _ in range(a.reps):
        t0 = time.perf_counter()
        res = fe.pdesolver_tangent_batch(Co, ens[:, :7], ens[:, 7:], seeds, **kw)
        best = min(best, time.perf_counter() - t0)
    t0 = time.perf_counter()
    pri = fe.pdesolver_batch(Co, ens[:, :7], ens[:, 7:], **kw)
    t0 = time.perf_counter()
    pri = fe.pdesolver_batch(Co, ens[:, :7], ens[:, 7:], **kw)
    tp = time.perf_counter() - t0
    good = (res.status & 1) == 0
    flips = int((res.n_bc_iters[good] != pri.n_bc_iters[good]).sum())     # sets whose iteration count differs from the primal kernel's
    steps = float(res.n_steps.sum())
    print(json.dumps({"workload": f"{a.sets} sets x {a.ndir} partials, dr={a.dr}, tf={a.tf}, pct-bound + gradient",
                      "family": os.environ.get("GAB1_TANGENT", "default"), "nt": os.environ.get("GAB1_TANGENT_NT", "default"), "s_per_call": best,
                      "gradients_per_s": a.sets / best, "primal_s_per_call": tp, "cost_vs_primal": best / tp,
                      "node_steps_per_s": steps * (round(10 / a.dr) - 1) / best, "nan_sets": int((res.status & 1).sum()), "sets_with_other_iteration_count_than_primal_kernel": flips}))


if __name__ == "__main__":
    main()
