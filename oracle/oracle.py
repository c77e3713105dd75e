"""ctypes harness for the C oracle (oracle/gab1_oracle.c) — TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.  `OracleBackend` has the same
`.solve` shape as the product's CudaBackend so that the Julia-surface frontend can be bound to either and the
results compared.  Parity pin: the reference's stored eFAST indices (coarse) — see the header of gab1_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import importlib
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB = HERE / "libgab1_oracle.so"

pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
abi = pkg.abi

_lib = None
_dual = None
LIB_DUAL = HERE / "libgab1_oracle_dual.so"


def build(force: bool = False) -> Path:
    src = HERE / "gab1_oracle.c"
    hdr = HERE.parent / "include" / "gab1pde.h"
    src2 = HERE / "gab1_oracle_dual.cpp"
    if force or not LIB.exists() or LIB.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime) or \
            not LIB_DUAL.exists() or LIB_DUAL.stat().st_mtime < max(src2.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run(["make", "-C", str(HERE), "-B"], check=True, capture_output=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(str(LIB))
        lib.gab1o_solve_batch.argtypes = abi.SOLVE_ARGTYPES + [C.c_int32]
        lib.gab1o_solve_batch.restype = C.c_int
        lib.gab1o_max_threads.restype = C.c_int
        lib.gab1o_out_doubles_per_set.argtypes = [C.POINTER(abi.Opts)]
        lib.gab1o_out_doubles_per_set.restype = C.c_int64
        _lib = lib
    return _lib


def load_dual():
    """The forward-mode (dual-number) restatement, oracle/gab1_oracle_dual.cpp."""
    global _dual
    if _dual is None:
        build()
        lib = C.CDLL(str(LIB_DUAL))
        lib.gab1o_solve_tangent.argtypes = abi.TANGENT_ARGTYPES + [C.c_int32]
        lib.gab1o_solve_tangent.restype = C.c_int
        dp = C.POINTER(C.c_double)
        lib.gab1o_default_dt_dual.argtypes = [C.c_int64, C.c_int32, dp, dp, C.c_double, dp, dp]
        lib.gab1o_default_dt_dual.restype = C.c_int
        _dual = lib
    return _dual


def max_threads() -> int:
    return int(load().gab1o_max_threads())


class OracleBackend:
    name = "oracle"

    def __init__(self, nthreads: int = 0):
        self.nthreads = nthreads

    def solve(self, o, Co, D, k, dt, r):
        lib = load()
        rc, out, status, n_saved, n_steps, n_bc = abi.call_solve(lib.gab1o_solve_batch, o, Co, D, k, dt, r,
                                                                 C.c_int32(self.nthreads))
        if rc != 0:
            raise RuntimeError(f"oracle rejected the options ({rc})")
        return out, status, n_saved, n_steps, n_bc


    def solve_tangent(self, o, Co, D, k, dt, seeds, r):
        lib = load_dual()
        rc, out, status, n_saved, n_steps, n_bc = abi.call_solve_tangent(lib.gab1o_solve_tangent, o, Co, D, k, dt, seeds, r,
                                                                         C.c_int32(self.nthreads))
        if rc != 0:
            raise RuntimeError(f"dual oracle rejected the options ({rc})")
        return out, status, n_saved, n_steps, n_bc

    def default_dt_tangent(self, D, k, dr, seeds):
        return abi._default_dt_tangent(load_dual().gab1o_default_dt_dual, D, k, dr, seeds)

    def solve_quantiles(self, o, Co, D, k, dt, r, matrices, c0, c1, probs):
        """CPU restatement of gab1_solve_ensemble_quantiles: the oracle's FULL result, NaN sets dropped, then the order
        statistics with the published definitions of Julia's Statistics stdlib (not in the reference tree: unpinned)."""
        import numpy as np
        out, status, n_saved, n_steps, n_bc = self.solve(o, Co, D, k, dt, r)
        keep = (status & abi.ST_NAN) == 0
        P, Cn = o.Nr + 1, o.Nts + 1
        ms = [m for m in range(abi.N_MATRICES) if (matrices >> m) & 1]
        p = abi.encode_probs(probs)
        q = np.zeros((len(ms), len(p), c1 - c0, P))
        for i, m in enumerate(ms):
            off = abi.full_matrix_offset(o, m)
            stack = out[keep, off:off + P * Cn].reshape(-1, Cn, P)[:, c0:c1, :]      # (sets, column, node)
            v = np.sort(stack, axis=0)
            n = v.shape[0]
            for j, pj in enumerate(p):
                if n == 0:
                    q[i, j] = np.nan
                elif pj < 0:                                   # median!: middle(a, b) = a/2 + b/2
                    q[i, j] = v[(n - 1) // 2] if n % 2 else v[n // 2 - 1] / 2 + v[n // 2] / 2
                elif n == 1:
                    q[i, j] = v[0]
                else:                                          # Statistics._quantile, alpha = beta = 1
                    aleph = n * pj + (1.0 + pj * (1.0 - 1.0 - 1.0))
                    jj = min(max(int(np.trunc(aleph)), 1), n - 1)
                    g = min(max(aleph - jj, 0.0), 1.0)
                    q[i, j] = v[jj - 1] + g * (v[jj] - v[jj - 1])
        return q, int(keep.sum()), status, n_saved, n_steps, n_bc


def frontend(nthreads: int = 0):
    return pkg.host.Frontend(OracleBackend(nthreads))
