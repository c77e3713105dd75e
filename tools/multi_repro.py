"""Diagnostic: small shards (every set on the latency lane) solved by several devices at once against the one-warp kernel."""
import importlib, os, sys, numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
abi = pkg.abi
ndev = abi.load_library().gab1_device_count()
ensemble = pkg.params.load_parameter_ensemble()
g = np.random.Generator(np.random.PCG64(3))
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1001
D, k = ensemble[:S, :7], ensemble[:S, 7:]
Co = pkg.params.base_Co()[None, :] * g.uniform(0.5, 2.0, size=(S, 1))
dt = pkg.params.default_dt(D, k, 0.2) * np.where(g.random(S) < 0.2, 0.25, 1.0)
kw = dict(dr=0.2, tf=0.3, dt=dt, out_mode=abi.OUT_FINAL_STATE)
os.environ["GAB1_KERNEL"] = "legacy"
ref = pkg.host.Frontend(abi.CudaBackend(n_devices=1)).sapdesolver_batch(Co, D, k, **kw)
os.environ.pop("GAB1_KERNEL")
perm, bounds = abi.deal_shards(dt, 0.3, ndev)
dev_of = np.zeros(S, int)
for d in range(ndev): dev_of[perm[bounds[d]:bounds[d + 1]]] = d
def report(name, b, rows=None):
    a_out = ref.out if rows is None else ref.out[rows]
    eq = (a_out.view(np.uint64) == b.out.view(np.uint64)) | (np.isnan(a_out) & np.isnan(b.out))
    bad = np.flatnonzero(~eq.all(axis=1))
    msg = f"{name}: rows differing {len(bad)}"
    if len(bad):
        r = bad[0]; cols = np.flatnonzero(~eq[r])
        rel = np.abs(a_out[bad] - b.out[bad]).max() / np.abs(a_out[bad]).max()
        gl = bad if rows is None else rows[bad]
        msg += f" first {gl[:8]} on devices {dev_of[gl[:8]]} ncols {len(cols)} cols {cols[:6]} max rel {rel:.2e} bc ref/got {ref.n_bc_iters[gl[0]]}/{b.n_bc_iters[r]} steps {ref.n_steps[gl[0]]}"
    print(msg, flush=True)
for rep in range(3):
    res = pkg.host.Frontend(abi.CudaBackend(device_ids=list(range(ndev)))).sapdesolver_batch(Co, D, k, **kw)
    report(f"{ndev} devices, rep {rep}", res)
# the same shards one at a time on device 0 and on device 1
for dev in range(min(ndev, 2)):
    for d in range(ndev):
        rows = np.sort(perm[bounds[d]:bounds[d + 1]])
        fe = pkg.host.Frontend(abi.CudaBackend(device_ids=[dev]))
        b = fe.sapdesolver_batch(Co[rows], D[rows], k[rows], **dict(kw, dt=dt[rows]))
        report(f"shard {d} ({len(rows)} sets) alone on device {dev}", b, rows)
os.environ["GAB1_KERNEL"] = "duo"
b = pkg.host.Frontend(abi.CudaBackend(n_devices=1)).sapdesolver_batch(Co, D, k, **kw)
report("all sets, two warps each, one device", b)
