"""Second A/B of session Y: (1) K = 4 (dr = 0.1) with the pipelined halo against the library named by GAB1PDE_LIB;
(2) the 4736-row whole-wave batch through the plain kernel, the default path and the default path with the lane off."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
abi = pkg.abi
ens = pkg.params.load_parameter_ensemble()
Co = pkg.params.base_Co()
gfe = pkg.host.Frontend(abi.CudaBackend())
good = ens[np.setdiff1d(np.arange(5000), [75])][:4736]
tag0 = os.environ.get("AB_TAG", "lib")


def timeit(tag, f, n=3):
    f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    print(tag0, tag, "%.1f ms (min of %d; all: %s)" % (1e3 * min(ts), n, " ".join("%.1f" % (1e3 * t) for t in ts)), flush=True)


def env(**kv):
    for k in ("GAB1_KERNEL", "GAB1_DUO"):
        os.environ.pop(k, None)
    os.environ.update(kv)


k4 = lambda: gfe.pdesolver_batch(Co, good[:2368, :7], good[:2368, 7:], dr=0.1, tol=1e-4, maxiters=20, out_mode=abi.OUT_FINAL_STATE)
c1 = lambda: gfe.pdesolver_batch(Co, good[:, :7], good[:, 7:], dr=0.2, tol=1e-4, maxiters=20, out_mode=abi.OUT_FINAL_STATE)
for name, e in (("legacy", dict(GAB1_KERNEL="legacy")), ("default", {}), ("lane off", dict(GAB1_DUO="0"))):
    env(**e)
    timeit(f"{name}: dr=0.1 2368 sets", k4, n=3)
    timeit(f"{name}: dr=0.2 4736 sets (whole waves)", c1, n=6)
env()
