"""Debug helper (GPU box): which golden KAT case differs from the fast path, and where."""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
fe = pkg.host.Frontend(pkg.abi.CudaBackend(arith=pkg.abi.ARITH_FAST))
kat = np.load(ROOT / "tests/golden/oracle_kat.npz")
ens = pkg.params.load_parameter_ensemble()
sub = ens[kat["rows"]]
Co = pkg.params.base_Co()
calls = [
    ("full_dr04_tf1", lambda: fe.pdesolver_batch(Co, sub[:, :7], sub[:, 7:], dr=0.4, tf=1.0, Nts=10, tol=1e-4, maxiters=20)),
    ("final4_dr02_tf05", lambda: fe.sapdesolver_batch(Co, sub[:, :7], sub[:, 7:], dr=0.2, tf=0.5)),
    ("final4_memb_dr02_tf05", lambda: fe.sapdesolver_batch(pkg.params.hela_Co(), sub[:, :7], sub[:, 7:], dr=0.2, tf=0.5, membSFK=True)),
    ("six_dr02_tf05", lambda: fe.sapdesolver_batch(Co, sub[:, :7], sub[:, 7:], dr=0.2, tf=0.5, out_mode=pkg.abi.OUT_SIX)),
    ("full_rect_dr025_tf05", lambda: fe.pdesolver_batch(Co, sub[:, :7], sub[:, 7:], dr=0.25, tf=0.5, Nts=5, tol=1e-4,
                                                        maxiters=20, geometry=pkg.abi.GEOM_RECT, pg1tot_form=pkg.abi.PG1TOT_CHAIN)),
]
print("rows", kat["rows"])
for name, call in calls:
    res = call()
    a, b = res.out, kat[name]
    na, nb = np.isnan(a), np.isnan(b)
    print(name, "shape", a.shape, "nan gpu/ref", na.sum(), nb.sum(), "status", res.status, "steps", res.n_steps, "bc", res.n_bc_iters)
    if not np.array_equal(na, nb):
        idx = np.argwhere(na != nb)
        print("  first mismatches", idx[:5], "gpu", a[tuple(idx[0])], "ref", b[tuple(idx[0])])
        sets = sorted(set(idx[:, 0]))
        print("  sets with mismatch", sets)
    else:
        fin = np.isfinite(b)
        scale = np.where(fin, np.abs(b), 0).max(axis=-1, keepdims=True)
        den = np.maximum(np.abs(b), 1e-6 * scale)
        e = np.where(fin & (den > 0), np.abs(a - b) / den, 0)
        print("  max rel err", e.max(), "at", np.unravel_index(e.argmax(), e.shape))
