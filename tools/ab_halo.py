"""A/B timing of the one-set-per-warp product kernel (the library named by GAB1PDE_LIB, default the in-tree one) on the
bench workloads, resident outputs small (final state / six scalars) so that only the time loop is measured."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
abi = pkg.abi
ens = pkg.params.load_parameter_ensemble()
Co = pkg.params.base_Co()
gfe = pkg.host.Frontend(abi.CudaBackend())
pri = pkg.params.synthetic_prior_ensemble(20000, seed=123)
good = ens[np.setdiff1d(np.arange(5000), [75])][:4736]
tag0 = os.environ.get("AB_TAG", "lib")


def timeit(tag, f, n=3):
    f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    print(tag0, tag, "%.1f ms (min of %d; all: %s)" % (1e3 * min(ts), n, " ".join("%.1f" % (1e3 * t) for t in ts)), flush=True)


cfg1 = lambda e, **kw: (lambda: gfe.pdesolver_batch(Co, e[:, :7], e[:, 7:], dr=0.2, tol=1e-4, maxiters=20, out_mode=abi.OUT_FINAL_STATE, **kw))
cfg2 = lambda e: (lambda: gfe.sapdesolver_batch(Co, e[:, :7], e[:, 7:], dr=0.2, tol=1e-3, maxiters=20, out_mode=abi.OUT_SIX))
for kern in ("legacy", ""):
    if kern: os.environ["GAB1_KERNEL"] = kern
    else: os.environ.pop("GAB1_KERNEL", None)
    k = kern or "default"
    timeit(f"{k} config1 4736 whole waves", cfg1(good))
    timeit(f"{k} config1 5000", cfg1(ens))
    timeit(f"{k} config2 20000 prior", cfg2(pri))
    timeit(f"{k} dr=0.1 2368 sets", lambda: gfe.pdesolver_batch(Co, good[:2368, :7], good[:2368, 7:], dr=0.1, tol=1e-4, maxiters=20, out_mode=abi.OUT_FINAL_STATE), n=2)
    timeit(f"{k} dr=0.25 5000 sets", lambda: gfe.pdesolver_batch(Co, ens[:, :7], ens[:, 7:], dr=0.25, tol=1e-4, maxiters=20, out_mode=abi.OUT_FINAL_STATE), n=2)
