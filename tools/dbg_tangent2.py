import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
from oracle import oracle
abi = pkg.abi
gfe = pkg.host.Frontend(abi.CudaBackend()); ofe = oracle.frontend()
ens = pkg.params.load_parameter_ensemble()
Co = pkg.params.base_Co()
rows = [75, 0, 1]
D, k = ens[rows, :7], ens[rows, 7:]
seeds = np.zeros((3, 4, 30))
for d, j in enumerate((6, 7, 8, 9)): seeds[:, d, 7 + j] = 1.0
kw = dict(dr=0.2, tf=5.0, Nts=10, tol=1e-4, maxiters=20, out_mode=abi.OUT_FINAL4)
dt = pkg.params.default_dt(D, k, 0.2); dt[2] = 0.0
ref = ofe.pdesolver_tangent_batch(Co, D, k, seeds, dt=dt, **kw)
res = gfe.pdesolver_tangent_batch(Co, D, k, seeds, dt=dt, **kw)
pri = gfe.pdesolver_batch(Co, D, k, dt=dt, **kw)
print("oracle ", ref.status, ref.n_steps, ref.n_bc_iters, ref.n_saved)
print("tangent", res.status, res.n_steps, res.n_bc_iters, res.n_saved)
print("primal ", pri.status, pri.n_steps, pri.n_bc_iters, pri.n_saved)
print("nan counts", np.isnan(res.out[0]).sum(axis=-1), np.isnan(ref.out[0]).sum(axis=-1))
