"""A tiny all-latency-lane batch for compute-sanitizer (synccheck / racecheck / initcheck / memcheck)."""
import importlib, os, sys, numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
abi = pkg.abi
ens = pkg.params.load_parameter_ensemble()[:9]
Co = pkg.params.base_Co()
os.environ["GAB1_KERNEL"] = "duo"
fe = pkg.host.Frontend(abi.CudaBackend(n_devices=1))
r1 = fe.sapdesolver_batch(Co, ens[:, :7], ens[:, 7:], dr=0.2, tf=0.02, out_mode=abi.OUT_FINAL_STATE)
r2 = fe.pdesolver_batch(Co, ens[:, :7], ens[:, 7:], dr=0.2, tf=0.02, Nts=4, tol=1e-4, maxiters=20)
print("ok", r1.n_steps[:3], r2.n_saved[:3])
