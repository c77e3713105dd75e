# round-2 session P (1 GPU): phase timing of the latency lane on the longest set (library built with -DGAB1_DUO_TIMING)
set -x
GAB1PDE_LIB=tools/_build/libgab1pde_timing.so timeout 300 python - <<'PY' 2>&1 | grep -E "duo timing|ms" | sort | uniq -c | head -20
import importlib, os, sys, time, numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
abi = pkg.abi
fe = pkg.host.Frontend(abi.CudaBackend(arith=abi.ARITH_FAST))
Co = pkg.params.base_Co()
ens = pkg.params.synthetic_prior_ensemble(100_000, seed=123)
dt = pkg.params.default_dt(ens[:, :7], ens[:, 7:], 0.2)
top = ens[np.argsort(dt)[:1]]
os.environ["GAB1_KERNEL"] = "duo"
for _ in range(2):
    t0 = time.perf_counter()
    fe.sapdesolver_batch(Co, top[:, :7], top[:, 7:], dr=0.2, tf=5.0, tol=1e-3, maxiters=20, out_mode=abi.OUT_SIX)
    print("ms", (time.perf_counter() - t0) * 1e3)
PY
