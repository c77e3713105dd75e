"""Regenerates the committed fixtures under tests/golden/.  Run in the build container (needs /root/reference).

  parameter_ensemble.npy   the reference's Julia/parameter_ensemble.csv (5000 x 24 float64, columns = pnames of
                           get_param_posteriors.jl:24-26) parsed with Python's correctly-rounded float().  It is the
                           benchmark/parity INPUT the reference's run scripts use; data, not source.
  oracle_kat.npz           known-answer vectors produced by the C oracle (oracle/gab1_oracle.c) on a few rows, so that
                           a change in the oracle's arithmetic is caught even when both restatements change together.
                           These are NOT reference outputs (no Julia here): parity stays "unpinned".
"""
import csv
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))


def ensemble():
    src = Path("/root/reference/Julia/parameter_ensemble.csv")
    with src.open() as f:
        rd = csv.reader(f)
        hdr = next(rd)
        rows = [[float(x) for x in row] for row in rd]
    a = np.array(rows, dtype=np.float64)
    assert a.shape == (5000, 24) and hdr[0] == "Dsfk" and hdr[-1] == "kdr"
    np.save(HERE / "parameter_ensemble.npy", a)
    return a


def kat(ens):
    from oracle import oracle
    pkg = oracle.pkg
    fe = oracle.frontend()
    Co = pkg.params.base_Co()
    rows = [0, 1, 2, 76, 4999]
    sub = ens[rows]
    out = {}
    res = fe.pdesolver_batch(Co, sub[:, :7], sub[:, 7:], dr=0.4, tf=1.0, Nts=10, tol=1e-4, maxiters=20)
    out["full_dr04_tf1"] = res.out
    out["full_dr04_tf1_nbc"] = res.n_bc_iters
    res = fe.sapdesolver_batch(Co, sub[:, :7], sub[:, 7:], dr=0.2, tf=0.5)
    out["final4_dr02_tf05"] = res.out
    res = fe.sapdesolver_batch(pkg.params.hela_Co(), sub[:, :7], sub[:, 7:], dr=0.2, tf=0.5, membSFK=True)
    out["final4_memb_dr02_tf05"] = res.out
    res = fe.sapdesolver_batch(Co, sub[:, :7], sub[:, 7:], dr=0.2, tf=0.5, out_mode=pkg.abi.OUT_SIX)
    out["six_dr02_tf05"] = res.out
    res = fe.pdesolver_batch(Co, sub[:, :7], sub[:, 7:], dr=0.25, tf=0.5, Nts=5, tol=1e-4, maxiters=20,
                             geometry=pkg.abi.GEOM_RECT, pg1tot_form=pkg.abi.PG1TOT_CHAIN)
    out["full_rect_dr025_tf05"] = res.out
    np.savez_compressed(HERE / "oracle_kat.npz", rows=np.array(rows), **out)


if __name__ == "__main__":
    e = ensemble()
    kat(e)
    print("fixtures written")
