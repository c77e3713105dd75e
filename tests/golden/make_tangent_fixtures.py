"""Regenerates tests/golden/tangent_kat.npz: known-answer vectors of the forward-mode oracle
(oracle/gab1_oracle_dual.cpp) — values and partials w.r.t. kG1p, kG1dp, kSa, kSi (with dt = dt(D, k) carrying
partials, basepdesolver.jl:696) on a few rows of parameter_ensemble.csv.  NOT reference outputs (no Julia here): the
dual oracle is pinned by tests/test_tangent_cpu.py (value component == scalar oracle, partials == finite differences)."""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

if __name__ == "__main__":
    from oracle import oracle
    pkg = oracle.pkg
    fe = oracle.frontend()
    ens = pkg.params.load_parameter_ensemble()
    rows = [0, 1, 76, 4999]
    sub = ens[rows]
    seeds = np.zeros((len(rows), 4, 30))
    for d, j in enumerate((6, 7, 8, 9)):
        seeds[:, d, 7 + j] = 1.0
    res = fe.pdesolver_tangent_batch(pkg.params.base_Co(), sub[:, :7], sub[:, 7:], seeds, dr=0.4, tf=0.5, Nts=5, tol=1e-3,
                                     maxiters=20, matrices=("aSFK", "PG1S", "G2PG1S"))
    np.savez_compressed(HERE / "tangent_kat.npz", rows=np.array(rows), full_dr04_tf05=res.out,
                        full_dr04_tf05_nbc=res.n_bc_iters)
    print("tangent fixtures written", res.out.shape)
