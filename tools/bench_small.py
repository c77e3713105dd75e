"""Latency of small batches (1 .. 1184 sets): the one-warp-per-set kernels against the one-CTA-per-set team kernel."""
import importlib, os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
ens = pkg.params.load_parameter_ensemble()
Co = pkg.params.base_Co()
fe = pkg.host.Frontend(pkg.abi.CudaBackend())
for dr, tf in ((0.2, 5.0), (0.1, 5.0), (0.05, 1.0)):
    for S in (1, 37, 148, 296, 592, 1184):
        if dr == 0.05 and S > 592:
            continue
        out = {}
        for fam in ("", "team"):
            os.environ["GAB1_KERNEL"] = fam
            os.environ["GAB1_TEAM_MAX_SETS"] = "0" if fam == "" else ""      # "": the throughput kernels, never the team kernel
            kw = dict(dr=dr, tf=tf, out_mode=pkg.abi.OUT_FINAL4)
            fe.sapdesolver_batch(Co, ens[:2, :7], ens[:2, 7:], dr=dr, tf=0.01, out_mode=pkg.abi.OUT_FINAL4)
            t0 = time.perf_counter()
            r = fe.sapdesolver_batch(Co, ens[:S, :7], ens[:S, 7:], **kw)
            out[fam] = (time.perf_counter() - t0, r)
        same = np.array_equal(out[""][1].n_bc_iters, out["team"][1].n_bc_iters)
        print(f"dr={dr} tf={tf} S={S:5d}: one warp per set {1e3*out[''][0]:9.2f} ms   team {1e3*out['team'][0]:9.2f} ms   ratio {out[''][0]/out['team'][0]:.2f}  iteration counts equal {same}", flush=True)
