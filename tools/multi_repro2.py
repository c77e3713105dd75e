"""Diagnostic: the sequence of tests/test_gpu_multi.py::test_both_shard_plans_scatter_every_set_to_its_own_row, with details."""
import importlib, os, sys, numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
abi = pkg.abi
scenario = sys.argv[1]
ndev = abi.load_library().gab1_device_count()
ensemble = pkg.params.load_parameter_ensemble()
g = np.random.Generator(np.random.PCG64(3))
S = 1501
D, k = ensemble[:S, :7], ensemble[:S, 7:]
Co = pkg.params.base_Co()[None, :] * g.uniform(0.5, 2.0, size=(S, 1))
dt = pkg.params.default_dt(D, k, 0.2) * np.where(g.random(S) < 0.2, 0.25, 1.0)
kw = dict(dr=0.2, tf=0.3, dt=dt, out_mode=abi.OUT_FINAL_STATE)
os.environ["GAB1_SHARD_PLAN"] = "dealt"
perm, bounds = abi.deal_shards(dt, 0.3, ndev)
dev_of = np.zeros(S, int)
for d in range(ndev): dev_of[perm[bounds[d]:bounds[d + 1]]] = d
def solve(n, fam=None):
    if fam: os.environ["GAB1_KERNEL"] = fam
    else: os.environ.pop("GAB1_KERNEL", None)
    be = abi.CudaBackend(n_devices=1) if n == 1 else abi.CudaBackend(device_ids=list(range(n)))
    return pkg.host.Frontend(be).sapdesolver_batch(Co, D, k, **kw)
def report(name, a, b):
    eq = (a.out.view(np.uint64) == b.out.view(np.uint64)) | (np.isnan(a.out) & np.isnan(b.out))
    bad = np.flatnonzero(~eq.all(axis=1))
    msg = f"[{scenario}] {name}: rows differing {len(bad)}"
    if len(bad):
        r = bad[0]; cols = np.flatnonzero(~eq[r])
        rel = np.abs(a.out[bad] - b.out[bad]).max() / np.abs(a.out[bad]).max()
        msg += (f" rows {bad[:10]} devices {dev_of[bad[:10]]} ncols(first) {len(cols)} cols {cols[:6]} max rel {rel:.2e} "
                f"bc {a.n_bc_iters[r]}/{b.n_bc_iters[r]} steps {a.n_steps[r]} nt-rank {np.argsort(np.argsort(-a.n_steps))[bad[:10]]}")
    print(msg, flush=True)
leg = solve(1, "legacy")
if scenario == "as_test":
    ref = solve(1); report("auto one device vs legacy", leg, ref)
    r2 = solve(2); report("2 devices vs legacy", leg, r2)
    r8 = solve(ndev); report(f"{ndev} devices vs legacy", leg, r8); report(f"{ndev} devices vs auto ref", ref, r8)
    r8b = solve(ndev); report(f"{ndev} devices again vs legacy", leg, r8b)
elif scenario == "skip2":
    ref = solve(1); report("auto one device vs legacy", leg, ref)
    r8 = solve(ndev); report(f"{ndev} devices vs legacy", leg, r8)
elif scenario == "only2then8":
    r2 = solve(2); report("2 devices vs legacy", leg, r2)
    r8 = solve(ndev); report(f"{ndev} devices vs legacy", leg, r8)
elif scenario == "four":
    r4 = solve(4); report("4 devices vs legacy", leg, r4)
    r2 = solve(2); report("2 devices vs legacy", leg, r2)
    r4 = solve(4); report("4 devices after 2 vs legacy", leg, r4)
