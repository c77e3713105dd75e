"""ctypes view of include/gab1pde.h and the loader for the CUDA library.

The product path has exactly one implementation: libgab1pde.so (hand-written sm_100a kernels).
If the library is missing or no CUDA device is usable the calls raise — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ABI_VERSION = 1
N_CO, N_D, N_K = 5, 7, 17
N_MATRICES, N_VECTORS = 12, 11

GEOM_SPHERICAL, GEOM_RECT = 0, 1
SFK_DIFFUSIBLE, SFK_MEMBRANE, SFK_BOTH_FROZEN = 0, 1, 2
BC_FOR_BREAK, BC_WHILE = 0, 1
SAVE_T_GE_TSAVE, SAVE_MODULUS = 0, 1
PG1TOT_VIA_STOT, PG1TOT_CHAIN = 0, 1
OUT_FINAL4, OUT_FULL, OUT_SIX, OUT_PCT_BOUND, OUT_FINAL_STATE = 0, 1, 2, 3, 4
ARITH_FAST, ARITH_STRICT = 0, 1

ST_NAN, ST_ITER_CAP, ST_SHORT, ST_OVERFLOW, ST_THROW = 1, 2, 4, 8, 16

MATRIX_NAMES = ("iSFK", "aSFK", "GRB2", "GAB1", "SHP2", "G2G1", "G2PG1", "G2PG1S", "PG1", "PG1S", "PG1tot", "PG1Stot")
VECTOR_NAMES = ("pE", "mE", "mES", "mESmES", "E", "EG2", "EG2G1", "EG2PG1", "EG2PG1S", "EGFR_SHP2", "t_out")
MASK_ALL = 0xFFF
MASK_FITTING = (1 << 1) | (1 << 9) | (1 << 7)
N_SEED = 30          # one seed row: partials of [D(7); k(17); Co(5); dt] along one direction (gab1pde.h)


class Opts(C.Structure):
    """struct gab1_opts (include/gab1pde.h)."""

    _fields_ = [
        ("abi_version", C.c_int32),
        ("geometry", C.c_int32),
        ("sfk_mode", C.c_int32),
        ("bc_loop", C.c_int32),
        ("save_rule", C.c_int32),
        ("pg1tot_form", C.c_int32),
        ("out_mode", C.c_int32),
        ("matrix_mask", C.c_uint32),
        ("maxiters", C.c_int32),
        ("Nr", C.c_int32),
        ("Nts", C.c_int32),
        ("arith", C.c_int32),
        ("tol", C.c_double),
        ("R", C.c_double),
        ("dr", C.c_double),
        ("tf", C.c_double),
        ("dt_save", C.c_double),
        ("t_prechase", C.c_double),
        ("pct_mul", C.c_double),
        ("pct_div", C.c_double),
        ("n_devices", C.c_int32),
        ("reserved", C.c_int32),
        ("device_ids", C.POINTER(C.c_int32)),
    ]


def make_opts(*, R=10.0, dr=0.1, tf=5.0, Nts=100, Nr=None, dt_save=None, maxiters=100, tol=1e-6,
              geometry=GEOM_SPHERICAL, sfk_mode=SFK_DIFFUSIBLE, bc_loop=BC_FOR_BREAK,
              save_rule=SAVE_T_GE_TSAVE, pg1tot_form=PG1TOT_VIA_STOT, out_mode=OUT_FULL,
              matrix_mask=MASK_ALL, arith=ARITH_FAST, t_prechase=-1.0, pct_mul=1.0, pct_div=1.0,
              n_devices=1) -> Opts:
    o = Opts()
    o.abi_version = ABI_VERSION
    o.geometry, o.sfk_mode, o.bc_loop = geometry, sfk_mode, bc_loop
    o.save_rule, o.pg1tot_form, o.out_mode = save_rule, pg1tot_form, out_mode
    o.matrix_mask = matrix_mask
    o.maxiters = int(maxiters)
    o.Nr = int(np.ceil(R / dr)) if Nr is None else int(Nr)   # Nr = Int64(ceil(R/dr)), basepdesolver.jl:71
    o.Nts = int(Nts)
    o.arith = arith
    o.tol = float(tol)
    o.R, o.dr, o.tf = float(R), float(dr), float(tf)
    o.dt_save = float(tf) / int(Nts) if dt_save is None else float(dt_save)  # dt_save = tf/Nts, basepdesolver.jl:31
    o.t_prechase = float(t_prechase)
    o.pct_mul, o.pct_div = float(pct_mul), float(pct_div)
    o.n_devices = int(n_devices)
    o.reserved = 0
    o.device_ids = None
    return o


def out_doubles_per_set(o: Opts) -> int:
    P, Cn = o.Nr + 1, o.Nts + 1
    nm = bin(o.matrix_mask & MASK_ALL).count("1")
    return {OUT_FINAL4: 4 * P, OUT_FULL: nm * P * Cn + N_VECTORS * Cn, OUT_SIX: 6, OUT_PCT_BOUND: 1,
            OUT_FINAL_STATE: 10 * P + 8}[o.out_mode]


def full_matrix_offset(o: Opts, m: int) -> int:
    if not (o.matrix_mask >> m) & 1:
        return -1
    return bin(o.matrix_mask & ((1 << m) - 1)).count("1") * (o.Nr + 1) * (o.Nts + 1)


def full_vector_offset(o: Opts, v: int) -> int:
    return bin(o.matrix_mask & MASK_ALL).count("1") * (o.Nr + 1) * (o.Nts + 1) + v * (o.Nts + 1)


_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)

SOLVE_ARGTYPES = [C.POINTER(Opts), C.c_int64, _dp, C.c_int64, _dp, _dp, _dp, _dp, _dp, _i32p, _i32p, _i64p, _i64p]


def _ptr(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def call_solve(fn, o: Opts, Co, D, k, dt, r, *extra, alloc_out=None):
    """Marshal numpy arrays into a gab1_solve_batch-shaped function and return its outputs.
    alloc_out(n_doubles) -> flat float64 array supplies the output block (default: zero-filled pageable memory)."""
    D = np.ascontiguousarray(D, dtype=np.float64).reshape(-1, N_D)
    k = np.ascontiguousarray(k, dtype=np.float64).reshape(-1, N_K)
    S = D.shape[0]
    if k.shape[0] != S:
        raise ValueError("D and k must hold the same number of parameter sets")
    Co = np.ascontiguousarray(Co, dtype=np.float64)
    if Co.ndim == 1:
        if Co.shape[0] != N_CO:
            raise ValueError("Co must have 5 entries")
        co_stride = 0
    else:
        if Co.shape != (S, N_CO):
            raise ValueError("Co must be (5,) or (S, 5)")
        co_stride = N_CO
    dt = np.ascontiguousarray(np.broadcast_to(np.asarray(dt, dtype=np.float64), (S,)))
    r = np.ascontiguousarray(r, dtype=np.float64)
    if r.shape != (o.Nr + 1,):
        # the reference indexes r[Nr+1] (basepdesolver.jl:151,206); a shorter grid is a BoundsError there
        raise IndexError(f"r has {r.shape[0]} nodes but Nr+1 = {o.Nr + 1}")
    n = out_doubles_per_set(o)
    out = np.zeros((S, n), dtype=np.float64) if alloc_out is None else alloc_out(S * n).reshape(S, n)
    status = np.zeros(S, dtype=np.int32)
    n_saved = np.zeros(S, dtype=np.int32)
    n_steps = np.zeros(S, dtype=np.int64)
    n_bc = np.zeros(S, dtype=np.int64)
    rc = fn(C.byref(o), S, _ptr(Co, _dp), co_stride, _ptr(D, _dp), _ptr(k, _dp), _ptr(dt, _dp), _ptr(r, _dp),
            _ptr(out, _dp), _ptr(status, _i32p), _ptr(n_saved, _i32p), _ptr(n_steps, _i64p), _ptr(n_bc, _i64p), *extra)
    return rc, out, status, n_saved, n_steps, n_bc


TANGENT_ARGTYPES = [C.POINTER(Opts), C.c_int64, C.c_int32, _dp, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _i32p, _i32p, _i64p,
                    _i64p]


def call_solve_tangent(fn, o: Opts, Co, D, k, dt, seeds, r, *extra):
    """Marshal numpy arrays into a gab1_solve_tangent-shaped function.  seeds: (S, n_dir, 30).
    Returns rc, out (S, 1 + n_dir, doubles_per_set), status, n_saved, n_steps, n_bc."""
    D = np.ascontiguousarray(D, dtype=np.float64).reshape(-1, N_D)
    k = np.ascontiguousarray(k, dtype=np.float64).reshape(-1, N_K)
    S = D.shape[0]
    seeds = np.ascontiguousarray(seeds, dtype=np.float64)
    if seeds.ndim != 3 or seeds.shape[0] != S or seeds.shape[2] != N_SEED:
        raise ValueError("seeds must be (S, n_dir, 30)")
    n_dir = seeds.shape[1]
    Co = np.ascontiguousarray(Co, dtype=np.float64)
    co_stride = 0 if Co.ndim == 1 else N_CO
    if Co.shape not in ((N_CO,), (S, N_CO)):
        raise ValueError("Co must be (5,) or (S, 5)")
    dt = np.ascontiguousarray(np.broadcast_to(np.asarray(dt, dtype=np.float64), (S,)))
    r = np.ascontiguousarray(r, dtype=np.float64)
    if r.shape != (o.Nr + 1,):
        raise IndexError(f"r has {r.shape[0]} nodes but Nr+1 = {o.Nr + 1}")
    out = np.zeros((S, 1 + n_dir, out_doubles_per_set(o)), dtype=np.float64)
    status = np.zeros(S, dtype=np.int32)
    n_saved = np.zeros(S, dtype=np.int32)
    n_steps = np.zeros(S, dtype=np.int64)
    n_bc = np.zeros(S, dtype=np.int64)
    rc = fn(C.byref(o), S, n_dir, _ptr(Co, _dp), co_stride, _ptr(D, _dp), _ptr(k, _dp), _ptr(dt, _dp), _ptr(seeds, _dp),
            _ptr(r, _dp), _ptr(out, _dp), _ptr(status, _i32p), _ptr(n_saved, _i32p), _ptr(n_steps, _i64p), _ptr(n_bc, _i64p),
            *extra)
    return rc, out, status, n_saved, n_steps, n_bc


class Gab1Error(RuntimeError):
    pass


_LIB = None
LIB_NAME = "libgab1pde.so"


def lib_path() -> Path:
    return Path(__file__).resolve().parent / LIB_NAME


def load_library():
    """dlopen libgab1pde.so (built in-tree by __graft_entry__.build()); raises if it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = Path(os.environ.get("GAB1PDE_LIB", lib_path()))
    if not p.exists():
        raise Gab1Error(f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                        "There is no CPU fallback.")
    lib = C.CDLL(str(p))
    lib.gab1_solve_batch.argtypes = SOLVE_ARGTYPES
    lib.gab1_solve_batch.restype = C.c_int
    lib.gab1_solve_batch_certified.argtypes = SOLVE_ARGTYPES + [C.c_double, _i32p, C.POINTER(C.c_int64)]
    lib.gab1_solve_batch_certified.restype = C.c_int
    lib.gab1_solve_batch_device.argtypes = [C.POINTER(Opts), C.c_int32, C.c_void_p, C.c_int64,
                                            C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gab1_solve_batch_device.restype = C.c_int
    lib.gab1_workspace_bytes.argtypes = [C.c_int64]
    lib.gab1_workspace_bytes.restype = C.c_size_t
    lib.gab1_out_doubles_per_set.argtypes = [C.POINTER(Opts)]
    lib.gab1_out_doubles_per_set.restype = C.c_int64
    lib.gab1_full_matrix_offset.argtypes = [C.POINTER(Opts), C.c_int32]
    lib.gab1_full_matrix_offset.restype = C.c_int64
    lib.gab1_full_vector_offset.argtypes = [C.POINTER(Opts), C.c_int32]
    lib.gab1_full_vector_offset.restype = C.c_int64
    lib.gab1_opts_init.argtypes = [C.POINTER(Opts), C.c_double, C.c_double, C.c_double, C.c_int32]
    lib.gab1_opts_init.restype = None
    lib.gab1_default_dt.argtypes = [C.c_int64, _dp, _dp, C.c_double, _dp]
    lib.gab1_default_dt.restype = C.c_int
    lib.gab1_plan_shards.argtypes = [C.c_int64, _dp, C.c_double, C.c_int32, _i64p]
    lib.gab1_plan_shards.restype = C.c_int
    lib.gab1_deal_shards.argtypes = [C.c_int64, _dp, C.c_double, C.c_int32, _i64p, _i64p]
    lib.gab1_deal_shards.restype = C.c_int
    lib.gab1_host_alloc.argtypes = [C.c_size_t]
    lib.gab1_host_alloc.restype = C.c_void_p
    lib.gab1_host_alloc_near.argtypes = [C.c_size_t, C.c_int32]
    lib.gab1_host_alloc_near.restype = C.c_void_p
    lib.gab1_device_numa_node.argtypes = [C.c_int32]
    lib.gab1_device_numa_node.restype = C.c_int
    lib.gab1_host_free.argtypes = [C.c_void_p]
    lib.gab1_host_free.restype = None
    lib.gab1_release_device_memory.argtypes = []
    lib.gab1_release_device_memory.restype = None
    lib.gab1_quantiles_workspace_bytes.argtypes = [C.c_int64]
    lib.gab1_quantiles_workspace_bytes.restype = C.c_size_t
    lib.gab1_ensemble_quantiles_device.argtypes = [C.POINTER(Opts), C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                                   C.c_uint32, C.c_int32, C.c_int32, C.c_int32, _dp, C.c_void_p, C.c_void_p,
                                                   C.c_void_p]
    lib.gab1_ensemble_quantiles_device.restype = C.c_int
    lib.gab1_solve_ensemble_quantiles.argtypes = [C.POINTER(Opts), C.c_int64, _dp, C.c_int64, _dp, _dp, _dp, _dp, C.c_uint32,
                                                  C.c_int32, C.c_int32, C.c_int32, _dp, _dp, _i32p, _i32p, _i64p, _i64p, _i64p]
    lib.gab1_solve_ensemble_quantiles.restype = C.c_int
    lib.gab1_solve_tangent.argtypes = TANGENT_ARGTYPES
    lib.gab1_solve_tangent.restype = C.c_int
    lib.gab1_solve_tangent_device.argtypes = [C.POINTER(Opts), C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64]
    lib.gab1_solve_tangent_device.argtypes += [C.c_void_p] * 11
    lib.gab1_solve_tangent_device.restype = C.c_int
    lib.gab1_default_dt_tangent.argtypes = [C.c_int64, C.c_int32, _dp, _dp, C.c_double, _dp, _dp]
    lib.gab1_default_dt_tangent.restype = C.c_int
    lib.gab1_sample_prior.argtypes = [C.c_int64, C.c_uint64, _dp, _dp, C.c_double, C.c_double, _dp, _dp]
    lib.gab1_sample_prior.restype = C.c_int
    lib.gab1_sample_prior_device.argtypes = [C.c_int32, C.c_void_p, C.c_int64, C.c_uint64, _dp, _dp, C.c_double, C.c_double,
                                             C.c_void_p, C.c_void_p]
    lib.gab1_sample_prior_device.restype = C.c_int
    lib.gab1_measure_fp64_tflops.argtypes = [C.c_int32, C.c_double]
    lib.gab1_measure_fp64_tflops.restype = C.c_double
    lib.gab1_debug_recip_error.argtypes = [C.c_int32, C.c_double, C.c_double, _dp, _dp]
    lib.gab1_debug_recip_error.restype = C.c_int
    lib.gab1_kernel_launches.argtypes = []
    lib.gab1_kernel_launches.restype = C.c_int64
    lib.gab1_device_count.argtypes = []
    lib.gab1_device_count.restype = C.c_int
    lib.gab1_version.argtypes = []
    lib.gab1_version.restype = C.c_int
    lib.gab1_last_error.argtypes = []
    lib.gab1_last_error.restype = C.c_char_p
    _LIB = lib
    return lib


class CudaBackend:
    """The product backend: every solve goes through gab1_solve_batch in libgab1pde.so."""

    name = "cuda"

    # output blocks of at least this many bytes come from the pinned pool: the kernels then write the caller's buffer
    # directly while the time loop runs (gab1pde.h, gab1_host_alloc) instead of a staged copy into pageable memory
    PINNED_FROM_BYTES = 32 << 20

    def __init__(self, n_devices: int = 1, arith: int = ARITH_FAST, device_ids=None):
        self.n_devices = n_devices if device_ids is None else len(device_ids)
        self.arith = arith
        self.device_ids = None if device_ids is None else (C.c_int32 * len(device_ids))(*device_ids)

    def strict_twin(self):
        """The same devices with arith = 1 (every operation of the reference in source order: bit-identical to the oracle)."""
        t = CudaBackend(self.n_devices, ARITH_STRICT)
        t.device_ids = self.device_ids
        return t

    def _bind(self, o: Opts):
        o.n_devices = self.n_devices
        o.device_ids = None if self.device_ids is None else C.cast(self.device_ids, C.POINTER(C.c_int32))

    def _alloc_out(self, n: int):
        if n * 8 >= self.PINNED_FROM_BYTES:
            return pinned_pool_array(n, np.float64, device=self.device_ids[0] if self.device_ids is not None else 0)
        return np.empty(n, dtype=np.float64)           # the kernels write every element of every set's block

    def solve(self, o: Opts, Co, D, k, dt, r):
        lib = load_library()
        self._bind(o)
        o.arith = self.arith
        rc, out, status, n_saved, n_steps, n_bc = call_solve(lib.gab1_solve_batch, o, Co, D, k, dt, r, alloc_out=self._alloc_out)
        if rc != 0:
            raise Gab1Error(f"gab1_solve_batch failed ({rc}): {lib.gab1_last_error().decode(errors='replace')}")
        return out, status, n_saved, n_steps, n_bc

    def solve_certified(self, o: Opts, Co, D, k, dt, r, response: float = 0.0):
        """gab1_solve_batch_certified: the fast solve, then a strict re-solve of every set whose final state responds to a
        one-ulp change of Co.  Returns (out, status, n_saved, n_steps, n_bc, indices of the re-solved sets)."""
        lib = load_library()
        self._bind(o)
        o.arith = self.arith
        S = np.asarray(D).reshape(-1, N_D).shape[0]
        resolved = np.zeros(S, dtype=np.int32)
        n_res = C.c_int64(0)
        rc, out, status, n_saved, n_steps, n_bc = call_solve(lib.gab1_solve_batch_certified, o, Co, D, k, dt, r,
                                                             C.c_double(response), _ptr(resolved, _i32p), C.byref(n_res),
                                                             alloc_out=self._alloc_out)
        if rc != 0:
            raise Gab1Error(f"gab1_solve_batch_certified failed ({rc}): {lib.gab1_last_error().decode(errors='replace')}")
        idx = np.flatnonzero(resolved)
        assert len(idx) == n_res.value
        return out, status, n_saved, n_steps, n_bc, idx

    def solve_tangent(self, o: Opts, Co, D, k, dt, seeds, r):
        """Values and forward-mode partials along the seed directions (gab1_solve_tangent)."""
        lib = load_library()
        self._bind(o)
        rc, out, status, n_saved, n_steps, n_bc = call_solve_tangent(lib.gab1_solve_tangent, o, Co, D, k, dt, seeds, r)
        if rc != 0:
            raise Gab1Error(f"gab1_solve_tangent failed ({rc}): {lib.gab1_last_error().decode(errors='replace')}")
        return out, status, n_saved, n_steps, n_bc

    def default_dt_tangent(self, D, k, dr, seeds):
        lib = load_library()
        return _default_dt_tangent(lib.gab1_default_dt_tangent, D, k, dr, seeds)

    def solve_quantiles(self, o: Opts, Co, D, k, dt, r, matrices: int, c0: int, c1: int, probs):
        """Solve on one GPU, keep the FULL result in HBM, return order statistics across the sets (gab1pde.h)."""
        lib = load_library()
        self._bind(o)
        o.n_devices = 1
        o.arith = self.arith
        D = np.ascontiguousarray(D, dtype=np.float64).reshape(-1, N_D)
        k = np.ascontiguousarray(k, dtype=np.float64).reshape(-1, N_K)
        S = D.shape[0]
        Co = np.ascontiguousarray(Co, dtype=np.float64)
        co_stride = 0 if Co.ndim == 1 else N_CO
        dt = np.ascontiguousarray(np.broadcast_to(np.asarray(dt, dtype=np.float64), (S,)))
        r = np.ascontiguousarray(r, dtype=np.float64)
        p = encode_probs(probs)
        nm = bin(matrices).count("1")
        q = np.zeros((nm, len(p), c1 - c0, o.Nr + 1), dtype=np.float64)
        status = np.zeros(S, dtype=np.int32)
        n_saved = np.zeros(S, dtype=np.int32)
        n_steps = np.zeros(S, dtype=np.int64)
        n_bc = np.zeros(S, dtype=np.int64)
        n_valid = np.zeros(1, dtype=np.int64)
        rc = lib.gab1_solve_ensemble_quantiles(C.byref(o), S, _ptr(Co, _dp), co_stride, _ptr(D, _dp), _ptr(k, _dp), _ptr(dt, _dp),
                                               _ptr(r, _dp), matrices, c0, c1, len(p), _ptr(p, _dp), _ptr(q, _dp),
                                               _ptr(status, _i32p), _ptr(n_saved, _i32p), _ptr(n_steps, _i64p),
                                               _ptr(n_bc, _i64p), _ptr(n_valid, _i64p))
        if rc != 0:
            raise Gab1Error(f"gab1_solve_ensemble_quantiles failed ({rc}): {lib.gab1_last_error().decode(errors='replace')}")
        return q, int(n_valid[0]), status, n_saved, n_steps, n_bc


def _default_dt_tangent(fn, D, k, dr, seeds):
    """dt and its partials by the rules of basepdesolver.jl:696 on duals; returns (dt (S,), seeds with slot 29 filled)."""
    D = np.ascontiguousarray(D, dtype=np.float64).reshape(-1, N_D)
    k = np.ascontiguousarray(k, dtype=np.float64).reshape(-1, N_K)
    seeds = np.array(seeds, dtype=np.float64, order="C", copy=True)
    S, n_dir = seeds.shape[0], seeds.shape[1]
    dt = np.zeros(S, dtype=np.float64)
    if fn(S, n_dir, _ptr(D, _dp), _ptr(k, _dp), float(dr), _ptr(dt, _dp), _ptr(seeds, _dp)) != 0:
        raise Gab1Error("default_dt_tangent failed")
    return dt, seeds


def sample_prior(S: int, seed: int, mu, sigma, EGF: float, Kdd: float):
    """gab1_sample_prior: S synthetic prior sets drawn on the device -> (D (S, 7), k (S, 17))."""
    lib = load_library()
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    if mu.shape != (22,) or sigma.shape != (22,):
        raise ValueError("mu and sigma must hold 22 entries")
    D = np.zeros((S, N_D))
    k = np.zeros((S, N_K))
    rc = lib.gab1_sample_prior(S, seed, _ptr(mu, _dp), _ptr(sigma, _dp), float(EGF), float(Kdd), _ptr(D, _dp), _ptr(k, _dp))
    if rc != 0:
        raise Gab1Error(f"gab1_sample_prior failed ({rc}): {lib.gab1_last_error().decode(errors='replace')}")
    return D, k


def encode_probs(probs) -> np.ndarray:
    """'median' -> -1.0 (Julia's median rule), numbers in [0, 1] -> quantile(v, p)."""
    return np.array([-1.0 if (isinstance(p, str) and p == "median") else float(p) for p in probs], dtype=np.float64)


def plan_shards(dt, tf: float, n_shards: int) -> np.ndarray:
    """bounds[g]..bounds[g+1]: the contiguous, step-count-balanced range of sets that device/rank g solves."""
    lib = load_library()
    dt = np.ascontiguousarray(dt, dtype=np.float64)
    bounds = np.zeros(n_shards + 1, dtype=np.int64)
    if lib.gab1_plan_shards(dt.shape[0], _ptr(dt, _dp), float(tf), n_shards, _ptr(bounds, _i64p)) != 0:
        raise Gab1Error(lib.gab1_last_error().decode())
    return bounds


def deal_shards(dt, tf: float, n_shards: int):
    """(perm, bounds): perm[bounds[g]:bounds[g+1]] are the set indices device/rank g solves when the sets are dealt from
    the descending step-count order (gab1_deal_shards: the plan for small per-set outputs)."""
    lib = load_library()
    dt = np.ascontiguousarray(dt, dtype=np.float64)
    perm = np.zeros(dt.shape[0], dtype=np.int64)
    bounds = np.zeros(n_shards + 1, dtype=np.int64)
    if lib.gab1_deal_shards(dt.shape[0], _ptr(dt, _dp), float(tf), n_shards, _ptr(perm, _i64p), _ptr(bounds, _i64p)) != 0:
        raise Gab1Error(lib.gab1_last_error().decode())
    return perm, bounds


class _PinnedPool:
    """Pinned, device-mapped host blocks kept between calls: pinning 2.5 GB costs ~0.5 s, re-using a block nothing.
    A block returns here when the array that wraps it (and every view of it) has been garbage-collected."""

    def __init__(self, keep_bytes: int = 8 << 30):
        self.keep_bytes = keep_bytes
        self.free = {}          # (nbytes, device) -> [pointer]
        self.held = 0

    def take(self, nbytes: int, device: int):
        lst = self.free.get((nbytes, device))
        if lst:
            self.held -= nbytes
            return lst.pop()
        lib = load_library()
        p = lib.gab1_host_alloc_near(nbytes, device)
        if not p:
            self.trim(0)
            p = lib.gab1_host_alloc_near(nbytes, device)
        if not p:
            raise MemoryError(lib.gab1_last_error().decode())
        return p

    def give(self, p, nbytes: int, device: int):
        if self.held + nbytes > self.keep_bytes:
            load_library().gab1_host_free(p)
            return
        self.free.setdefault((nbytes, device), []).append(p)
        self.held += nbytes

    def trim(self, keep_bytes: int):
        lib = load_library()
        for key, lst in list(self.free.items()):
            while lst and self.held > keep_bytes:
                lib.gab1_host_free(lst.pop())
                self.held -= key[0]


_POOL = _PinnedPool()


def pinned_pool_array(n: int, dtype=np.float64, device: int = 0) -> np.ndarray:
    """A flat array of pinned, device-mapped host memory near `device` that goes back to the pool when it is collected."""
    import weakref
    nbytes = max(int(n) * np.dtype(dtype).itemsize, 8)
    p = _POOL.take(nbytes, device)
    base = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))
    weakref.finalize(base, _POOL.give, p, nbytes, device)
    return base.view(dtype)[:n]


def release_pinned_pool() -> None:
    _POOL.trim(0)


def pinned_empty(n: int, dtype=np.float64):
    """numpy view of pinned, device-mapped host memory from gab1_host_alloc (caller frees with pinned_free)."""
    lib = load_library()
    nbytes = int(n) * np.dtype(dtype).itemsize
    p = lib.gab1_host_alloc(nbytes)
    if not p:
        raise MemoryError(lib.gab1_last_error().decode())
    a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,)).view(dtype)
    return p, a


def pinned_free(p) -> None:
    load_library().gab1_host_free(p)
