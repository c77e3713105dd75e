#!/usr/bin/env python
"""Benchmark of the hot path: ensemble PDE solves/sec (BASELINE.json metric).

Workload (config.workload): BASELINE configs[1] — the base spherical model (`pdesolver`) over all 5000 rows of the
reference's parameter_ensemble.csv at run_ensemble's defaults (dr=0.2 => Nr=50, tf=5, Nts=100, tol=1e-4, maxit=20,
get_param_posteriors.jl:135-139), full snapshot output (12 matrices 51x101 + 11 vectors per set).  One "step" is one
pass over the whole ensemble.  With --gpus N every rank solves the whole ensemble on its own GPU (weak scaling, no
data-path collective: parameter sets are independent).

  value      solves/s with inputs and outputs resident in HBM (gab1_solve_batch_device), CUDA-event timed
  e2e        solves/s through the reference-facing entry point gab1_solve_batch with HOST (pinned) buffers:
             H2D of the parameters and D2H of every snapshot inside the timed region
  roofline   FP64 pipe: algorithmic flops of the reference's expressions (229 per interior node-step, 241 per membrane
             iteration) / time, against a DFMA peak measured live on the same GPU
  cpu_baseline  the C restatement of the reference Julia solver (oracle/) on the host cores, bounded sample

`--impl reference` times that CPU restatement alone (Julia itself is not installable here).
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
PKG = "myers-furcht-et-al_gab1-shp2-pde-model_b200"

F_INT_SPH, F_BC = 229.0, 241.0          # flops as written in basepdesolver.jl:151-179 / :205-238
CFG = dict(dr=0.2, R=10.0, tf=5.0, Nts=100, tol=1e-4, maxiters=20)


def workload(pkg):
    ens = pkg.params.load_parameter_ensemble()
    if CFG.get("sets"):
        ens = ens[:CFG["sets"]]
    D = np.ascontiguousarray(ens[:, :7])
    k = np.ascontiguousarray(ens[:, 7:])
    Co = pkg.params.base_Co(CFG["R"])
    dt = pkg.params.default_dt(D, k, CFG["dr"])
    r = pkg.params.julia_range(CFG["dr"], CFG["R"])
    o = pkg.abi.make_opts(R=CFG["R"], dr=CFG["dr"], tf=CFG["tf"], Nts=CFG["Nts"], maxiters=CFG["maxiters"], tol=CFG["tol"],
                          out_mode=pkg.abi.OUT_FULL)
    return o, Co, D, k, dt, r


CONFIG = {"workload": "configs[1]: pdesolver over all 5000 rows of parameter_ensemble.csv, run_ensemble defaults "
                      "(spherical, dr=0.2/Nr=50, tf=5, Nts=100, tol=1e-4, maxit=20), full snapshot output",
          "sets_per_gpu": 5000, "grid_nodes": 51, "median_steps_per_solve": 37239,
          "out_bytes_per_set": (12 * 51 + 11) * 101 * 8,
          "l2": "2.5 GB of snapshot output per step exceeds L2 (126 MB) and a 256 MB buffer is rewritten between steps",
          "parallelism": "one rank per GPU, whole ensemble per rank, no collective"}


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v == "Active":
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(pkg, nsets_per_thread=32):
    """Times the oracle (C restatement of the reference CPU path) on a bounded sample of the same workload: 32 sets per host
    thread (about 20 CPU-seconds on 16 threads; enough sets per thread for the dynamic schedule to balance)."""
    from oracle import oracle
    o, Co, D, k, dt, r = workload(pkg)
    # every core this process may run on: torchrun exports OMP_NUM_THREADS=1 by default, which omp_get_max_threads() would
    # follow and turn the N>1 reference arm into a one-thread run
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else oracle.max_threads()
    n = min(D.shape[0], nsets_per_thread * threads)
    be = oracle.OracleBackend(threads)
    t0 = time.perf_counter()
    be.solve(o, Co, D[:n], k[:n], dt[:n], r)
    sec = time.perf_counter() - t0
    return n / sec, threads, n, sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = importlib.import_module(PKG)
    for _ in range(args.warmup):
        cpu_sample(pkg, 2)
    vals, secs, n, threads = [], 0.0, 0, 1
    for _ in range(args.steps):
        v, threads, n, s = cpu_sample(pkg)
        vals.append(v)
        secs += s
    value = float(np.mean(vals))
    sample = f"first {n} rows of the ensemble per step, OpenMP schedule(dynamic) over sets"
    line = {"impl": "reference", "metric": "ensemble PDE solves/sec", "value": value, "unit": "solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "reference parameter_ensemble.csv (tests/golden/parameter_ensemble.npy)", "config": CONFIG,
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "C restatement of the reference Julia CPU path (oracle/gab1_oracle.c); Julia is not installable offline"}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    pkg = importlib.import_module(PKG)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries the one JSON line: NCCL prints its version banner to stdout while the communicator is created
        # (NCCL_DEBUG=VERSION on the box; NCCL_DEBUG_FILE is not honoured at that level), so fd 1 points at stderr meanwhile
        libc = C.CDLL(None)
        sys.stdout.flush()
        libc.fflush(None)
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            libc.fflush(None)
            os.dup2(saved, 1)
            os.close(saved)
    lib = pkg.abi.load_library()
    o, Co, D, k, dt, r = workload(pkg)
    S = D.shape[0]
    nout = pkg.abi.out_doubles_per_set(o)
    dev = torch.device("cuda", local)

    def to_dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    dCo, dD, dk, ddt, dr_ = map(to_dev, (Co, D, k, dt, r))
    dout = torch.empty(S * nout, dtype=torch.float64, device=dev)
    dstatus = torch.zeros(S, dtype=torch.int32, device=dev)
    dsaved = torch.zeros(S, dtype=torch.int32, device=dev)
    dsteps = torch.zeros(S, dtype=torch.int64, device=dev)
    dbc = torch.zeros(S, dtype=torch.int64, device=dev)
    ws = torch.empty(int(lib.gab1_workspace_bytes(S)), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step_resident():
        rc = lib.gab1_solve_batch_device(C.byref(o), local, C.c_void_p(stream.cuda_stream), S, dCo.data_ptr(), 0,
                                         dD.data_ptr(), dk.data_ptr(), ddt.data_ptr(), dr_.data_ptr(), dout.data_ptr(),
                                         dstatus.data_ptr(), dsaved.data_ptr(), dsteps.data_ptr(), dbc.data_ptr(),
                                         ws.data_ptr())
        if rc != 0:
            raise RuntimeError(lib.gab1_last_error().decode())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # FP64 roofline denominator, measured live on this GPU (MEASURED_PEAKS.json has no FP64 entry)
    peak_tf = lib.gab1_measure_fp64_tflops(local, 1.0)

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = lib.gab1_kernel_launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0 = time.perf_counter()
    for a, b in ev:
        flush.fill_(1)                      # evict L2 between timed iterations
        a.record(stream)
        step_resident()
        b.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    launches = lib.gab1_kernel_launches() - launches0
    clocks = sampler.stop()
    ms = [a.elapsed_time(b) for a, b in ev]
    tsum = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tsum, op=dist.ReduceOp.MAX)
    total_ms = float(tsum.item())
    value = world * S * args.steps / (total_ms * 1e-3)

    # algorithmic flops of one pass, from the per-set step and membrane-iteration counts the kernel returns
    n_steps = dsteps.cpu().numpy().astype(np.float64)
    n_bc = dbc.cpu().numpy().astype(np.float64)
    flops = float((n_steps * (o.Nr - 1) * F_INT_SPH).sum() + n_bc.sum() * F_BC)
    kernel_ms = float(np.mean(ms))
    achieved = flops / (kernel_ms * 1e-3) / 1e12
    nan_sets = int((dstatus.cpu().numpy() & 1).sum())

    # ---- end to end through the host entry point: pinned host buffers, copies inside the timed region ----
    hb = {}
    if args.no_e2e:
        if rank == 0:
            print(json.dumps({"profiling_run": True, "value": value, "ms_per_step": total_ms / args.steps,
                              "roofline": {"achieved": achieved, "peak": peak_tf, "frac": achieved / peak_tf}}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    def pinned(name, arr_or_n, dtype):
        n = arr_or_n if isinstance(arr_or_n, int) else arr_or_n.size
        nbytes = n * np.dtype(dtype).itemsize
        p = lib.gab1_host_alloc(nbytes)
        if not p:
            raise MemoryError(name)
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,)).view(dtype)
        if not isinstance(arr_or_n, int):
            a[:] = np.asarray(arr_or_n, dtype=dtype).ravel()
        hb[name] = (p, a)
        return a

    hCo, hD, hk, hdt, hr = (pinned(n, a, np.float64) for n, a in (("Co", Co), ("D", D), ("k", k), ("dt", dt), ("r", r)))
    hout = pinned("out", S * nout, np.float64)
    hstatus, hsaved = pinned("status", S, np.int32), pinned("saved", S, np.int32)
    hsteps, hbc = pinned("steps", S, np.int64), pinned("bc", S, np.int64)
    dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    o.n_devices = 1
    ids = (C.c_int32 * 1)(local)
    o.device_ids = C.cast(ids, C.POINTER(C.c_int32))

    def step_e2e():
        rc = lib.gab1_solve_batch(C.byref(o), S, hCo.ctypes.data_as(dp), 0, hD.ctypes.data_as(dp), hk.ctypes.data_as(dp),
                                  hdt.ctypes.data_as(dp), hr.ctypes.data_as(dp), hout.ctypes.data_as(dp),
                                  hstatus.ctypes.data_as(ip), hsaved.ctypes.data_as(ip), hsteps.ctypes.data_as(lp),
                                  hbc.ctypes.data_as(lp))
        if rc != 0:
            raise RuntimeError(lib.gab1_last_error().decode())

    e2e_steps = max(2, min(args.steps, 3))
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_sec = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_sec, op=dist.ReduceOp.MAX)
    e2e_value = world * S * e2e_steps / float(e2e_sec.item())
    checksum = float(np.nansum(hout[::1009]))       # read of the result on the host
    h2d = int(hCo.nbytes + hD.nbytes + hk.nbytes + hdt.nbytes + hr.nbytes)
    d2h = int(hout.nbytes + hstatus.nbytes + hsaved.nbytes + hsteps.nbytes + hbc.nbytes)
    same = bool(np.array_equal(hbc, n_bc.astype(np.int64)))
    for p, _ in hb.values():
        lib.gab1_host_free(p)

    if rank == 0:
        traffic = None          # DRAM bytes of the dominant kernel per launch, from the committed ncu --set full capture
        try:
            tj = json.loads((ROOT / "profiles" / "r1_traffic.json").read_text())
            if not (CFG.get("sets") or args.dr):
                traffic = tj["traffic_bytes_per_launch"]
        except (OSError, ValueError, KeyError):
            pass
        cpu = None
        if world == 1 and not args.no_cpu:
            v, threads, n, sec = cpu_sample(pkg)
            cpu = {"value": v, "unit": "solves/s", "cores": threads, "kind": "port",
                   "sample": f"first {n} rows of the same ensemble, full output, {sec:.1f} s; C restatement of the "
                             "reference Julia solver (oracle/), OpenMP dynamic over sets"}
        line = {"metric": "ensemble PDE solves/sec", "value": value, "unit": "solves/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "reference parameter_ensemble.csv (tests/golden/parameter_ensemble.npy); 33 of 5000 rows diverge as in the reference",
                "config": CONFIG,
                "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                             "frac": achieved / peak_tf if peak_tf > 0 else None, "traffic": traffic,
                             "traffic_note": "dram__bytes_read+write of one launch (profiles/r1_traffic.json); the "
                                             "algorithmic output is 2.517 GB of snapshots, inputs 1 MB: HBM is not the bound",
                             "peak_source": "DFMA micro-benchmark run live by bench.py (gab1_measure_fp64_tflops); "
                                            "2 flop per FMA; MEASURED_PEAKS.json has no FP64 entry",
                             "flops_per_launch": flops, "kernel_ms": kernel_ms,
                             "flop_model": "229 per interior node-step + 241 per membrane iteration, as written in the reference"},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "host_checksum": checksum, "matches_resident_run": same},
                "gpu_launches": int(launches), "clocks": clocks, "nan_sets": nan_sets,
                "wall_s_timed_region": wall}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs only)")
    ap.add_argument("--dr", type=float, default=None, help="experiment: other radial step (not the reported workload)")
    ap.add_argument("--sets", type=int, default=None, help="experiment: use only the first N rows")
    args = ap.parse_args()
    import __graft_entry__ as g
    g.build()
    if args.dr is not None:
        CFG["dr"] = args.dr
        CONFIG["workload"] += f" [EXPERIMENT dr={args.dr}]"
    if args.sets is not None:
        CFG["sets"] = args.sets
        CONFIG["workload"] += f" [EXPERIMENT first {args.sets} rows]"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
