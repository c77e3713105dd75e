"""Diagnostic: per-output-array error of the tangent kernels against the dual oracle for one variant (GPU box)."""
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
Co = np.tile(pkg.params.hela_Co(), (4, 1)) * np.array([[1.0], [0.5], [2.0], [1.5]])
D, k = ens[20:24, :7], ens[20:24, 7:]
cols = [0, 7 + 6, 7 + 9, 24 + 0]
seeds = np.zeros((4, 4, 30))
for d, c in enumerate(cols): seeds[:, d, c] = 1.0
kw = dict(tol=1e-4, maxiters=20, dr=0.2, tf=0.1, Nts=4, sfk_mode=1)
res = gfe.pdesolver_tangent_batch(Co, D, k, seeds, **kw); ref = ofe.pdesolver_tangent_batch(Co, D, k, seeds, **kw)
o = res.opts; P, Cn = o.Nr + 1, o.Nts + 1
names = list(abi.MATRIX_NAMES) + list(abi.VECTOR_NAMES)
for i, nme in enumerate(names):
    lo, hi = (i * P * Cn, (i + 1) * P * Cn) if i < 12 else (12 * P * Cn + (i - 12) * Cn, 12 * P * Cn + (i - 11) * Cn)
    a, b = res.out[:, 1:, lo:hi], ref.out[:, 1:, lo:hi]
    sc = np.abs(b).max(axis=-1, keepdims=True)
    e = np.abs(a - b) / np.where(sc > 0, sc, 1)
    idx = np.unravel_index(np.argmax(e), e.shape)
    print(f"{nme:10s} worst {e.max():.3e} at set {idx[0]} dir {idx[1]} elem {idx[2]} (node {idx[2] % P if i < 12 else '-'}, col {idx[2] // P if i < 12 else idx[2]}) gpu {a[idx]:.6e} ref {b[idx]:.6e} scale {sc[idx[0], idx[1], 0]:.3e}")
