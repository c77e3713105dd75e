#!/usr/bin/env python
"""One solve call for profiling under ncu: python tools/prof_one.py [sets] [tf] [dr] (kernel family via GAB1_KERNEL / GAB1_GANG)."""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
tf = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
dr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.2
ens = pkg.params.load_parameter_ensemble()
ens = ens[np.setdiff1d(np.arange(5000), [75])]
ens = np.tile(ens, (S // len(ens) + 1, 1))[:S]
fe = pkg.host.Frontend(pkg.abi.CudaBackend())
res = fe.pdesolver_batch(pkg.params.base_Co(), ens[:, :7], ens[:, 7:], dr=dr, tf=tf, tol=1e-4, maxiters=20, out_mode=pkg.abi.OUT_FINAL_STATE)
print("done", S, tf, dr, int(res.n_bc_iters.sum()))
