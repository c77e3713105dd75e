#!/usr/bin/env python
"""One solve of S synthetic prior draws with the bench settings (six GSA scalars, tol 1e-3) for profiling under ncu:
python tools/prof_prior.py [sets]"""
import importlib, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ens = pkg.params.synthetic_prior_ensemble(S, seed=123)
fe = pkg.host.Frontend(pkg.abi.CudaBackend())
res = fe.sapdesolver_batch(pkg.params.base_Co(), ens[:, :7], ens[:, 7:], dr=0.2, tol=1e-3, maxiters=20, out_mode=pkg.abi.OUT_SIX)
print("done", S, int(res.n_bc_iters.sum()), int((res.status & 1).sum()))
