# round-2 session J (1 GPU): the express lane on the heaviest eighth of the 10^5-set prior ensemble
set -x
python - <<'PY' 2>&1 | tail -20
import importlib, os, sys, time, numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
abi = pkg.abi
ens = pkg.params.synthetic_prior_ensemble(100000, seed=123)
dt = pkg.params.default_dt(ens[:, :7], ens[:, 7:], 0.2)
nt = np.ceil(5.0 / dt)
perm, b = abi.deal_shards(dt, 5.0, 8)
g = int(np.argmax([nt[perm[b[i]:b[i + 1]]].max() for i in range(8)]))
mine = perm[b[g]:b[g + 1]]
print("shard", g, "sets", len(mine), "max Nt", nt[mine].max(), "sum Nt / 1184", nt[mine].sum() / 1184, "top 5 Nt", np.sort(nt[mine])[-5:], flush=True)
Co = pkg.params.base_Co()
fe = pkg.host.Frontend(abi.CudaBackend())
def run():
    return fe.sapdesolver_batch(Co, ens[mine, :7], ens[mine, 7:], dr=0.2, tol=1e-3, maxiters=20, out_mode=abi.OUT_SIX)
res = {}
for tag, env in (("express", {}), ("no express", {"GAB1_NO_EXPRESS": "1"})):
    os.environ.pop("GAB1_NO_EXPRESS", None); os.environ.update(env)
    run(); ts = []
    for _ in range(3):
        t0 = time.perf_counter(); r = run(); ts.append(time.perf_counter() - t0)
    res[tag] = r
    print(tag, "%.1f ms" % (1e3 * min(ts)), "ideal (1/8 of 1926 ms): 240.8 ms", flush=True)
a, c = res["express"], res["no express"]
print("bitwise equal:", np.array_equal(a.out.view(np.uint64), c.out.view(np.uint64)), np.array_equal(a.n_bc_iters, c.n_bc_iters), np.array_equal(a.status, c.status))
os.environ.pop("GAB1_NO_EXPRESS", None)
# whole ensemble on one GPU: the plan must find nothing to expedite
t0 = time.perf_counter(); fe.sapdesolver_batch(Co, ens[:, :7], ens[:, 7:], dr=0.2, tol=1e-3, maxiters=20, out_mode=abi.OUT_SIX); print("100000 sets: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
PY
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_census.py -m gpu -q -x 2>&1 | tail -3
