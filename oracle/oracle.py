"""ctypes harness for the C oracle (oracle/gab1_oracle.c) — TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.  `OracleBackend` has the same
`.solve` shape as the product's CudaBackend so that the Julia-surface frontend can be bound to either and the
results compared.  PARITY UNPINNED (no reference outputs exist; see the header of gab1_oracle.c).
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


def build(force: bool = False) -> Path:
    src = HERE / "gab1_oracle.c"
    hdr = HERE.parent / "include" / "gab1pde.h"
    if force or not LIB.exists() or LIB.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
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


def frontend(nthreads: int = 0):
    return pkg.host.Frontend(OracleBackend(nthreads))
