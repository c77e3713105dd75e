#!/usr/bin/env python
"""Throughput of the BASELINE.json configs other than the headline one (which bench.py measures), through the C ABI.

  python tools/bench_configs.py [--devices N] [--configs 1,3,4,5] [--scale F]

One JSON line per config.  `resident` = gab1_solve_batch_device on device 0 with inputs and outputs in HBM (CUDA events);
`e2e` = gab1_solve_batch with pinned host buffers sharded over N devices by the library itself (one host thread and one
stream per device, no collective) — the multi-GPU path a Julia caller gets.  Synthetic ensembles: params.py.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
PKG = "myers-furcht-et-al_gab1-shp2-pde-model_b200"


def configs(pkg, scale, cap=1_000_000):
    abi, params = pkg.abi, pkg.params
    ens = params.load_parameter_ensemble()
    out = {}
    # 1: run_base_model.jl:83 — one pdesolver call, dr = 0.1, tol = 1e-2 (latency)
    out[1] = dict(name="configs[0]: single pdesolver solve, dr=0.1 (Nr=100), Nts=100, tol=1e-2, maxiters=100",
                  o=abi.make_opts(dr=0.1, tol=1e-2, maxiters=100), Co=params.base_Co(),
                  D=params.DIFFS_BASE[None, :], k=params.KVALS_BASE[None, :], dr=0.1, f_int=229.0)
    # 3: fbatch_dk_mt semantics on synthetic prior draws (GSA sweep)
    S3 = int(100000 * scale)
    e3 = params.synthetic_prior_ensemble(S3, seed=123)
    out[3] = dict(name=f"configs[2]: sapdesolver + six GSA scalars on {S3} synthetic prior draws, dr=0.2, tol=1e-3, maxiters=20",
                  o=abi.make_opts(dr=0.2, Nts=1, tol=1e-3, maxiters=20, out_mode=abi.OUT_SIX), Co=params.base_Co(),
                  D=e3[:, :7], k=e3[:, 7:], dr=0.2, f_int=229.0)
    # 4: sapdesolver_membSFK with HeLa concentrations
    S4 = int(100000 * scale)
    e4 = params.resampled_ensemble(S4, seed=123)
    o4 = abi.make_opts(dr=0.2, Nts=1, tol=1e-3, maxiters=cap, out_mode=abi.OUT_FINAL4, sfk_mode=abi.SFK_MEMBRANE,
                       bc_loop=abi.BC_WHILE, pg1tot_form=abi.PG1TOT_CHAIN)
    out[4] = dict(name=f"configs[3]: sapdesolver_membSFK, HeLa concentrations, {S4} resampled sets, dr=0.2, tol=1e-3, while-loop cap {cap}",
                  o=o4, Co=params.hela_Co(), D=e4[:, :7], k=e4[:, 7:], dr=0.2, f_int=229.0)
    # 5: rectangular geometry at 4x refinement
    S5 = max(2, int(1184 * scale))
    out[5] = dict(name=f"configs[4]: pdesolver_rect, dr=0.05 (Nr=200), first {S5} rows of parameter_ensemble.csv, tol=1e-4, maxit=20, full output",
                  o=abi.make_opts(dr=0.05, tol=1e-4, maxiters=20, geometry=abi.GEOM_RECT, pg1tot_form=abi.PG1TOT_CHAIN),
                  Co=params.base_Co(), D=ens[:S5, :7], k=ens[:S5, 7:], dr=0.05, f_int=179.0)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", type=int, default=1)
    ap.add_argument("--configs", default="1,3,4,5")
    ap.add_argument("--scale", type=float, default=1.0, help="scale the ensemble sizes (quick runs)")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--cap", type=int, default=1_000_000, help="configs[3]: safety cap of the `while error > tol` loop (wrappers default: 100000)")
    ap.add_argument("--tf", type=float, default=None, help="profiling runs only: a shorter final time than the configs' tf = 5")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    args = ap.parse_args()
    import torch
    import __graft_entry__ as g
    g.build()
    pkg = importlib.import_module(PKG)
    abi, params = pkg.abi, pkg.params
    lib = abi.load_library()
    assert lib.gab1_device_count() >= args.devices, "not enough CUDA devices"
    peak = lib.gab1_measure_fp64_tflops(0, 0.5)
    dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    cfgs = configs(pkg, args.scale, args.cap)
    for ci in [int(x) for x in args.configs.split(",")]:
        if ci == 2:
            # configs[1] reduced on the device: median surface and the +-1 sigma band of three matrices over all snapshots
            ens = params.load_parameter_ensemble()
            fe = pkg.host.Frontend(abi.CudaBackend())
            fe.ensemble_quantiles(ens[:64], params.base_Co(), columns=(100, 101))
            t0 = time.perf_counter()
            res, n_valid, _, status = fe.ensemble_quantiles(ens, params.base_Co())
            sec = time.perf_counter() - t0
            nbytes = sum(v.nbytes for v in res.values())
            print(json.dumps({"config": "configs[1] + on-device statistics: median, 15.9 % and 84.1 % quantiles of aSFK, PG1tot, PG1Stot "
                                        "at all 101 snapshots x 51 nodes over the sets without NaN (gab1_solve_ensemble_quantiles)",
                              "sets": int(ens.shape[0]), "n_valid": int(n_valid), "e2e": {"s_per_call": sec, "solves_per_s": ens.shape[0] / sec,
                              "d2h_bytes": int(nbytes + 28 * ens.shape[0]), "full_result_bytes_kept_on_device": int(ens.shape[0]) * 3 * 51 * 101 * 8 + ens.shape[0] * 11 * 101 * 8}}), flush=True)
            continue
        c = cfgs[ci]
        o = c["o"]
        if args.tf is not None:
            o.tf = args.tf
            o.dt_save = args.tf / o.Nts
            c["name"] += f" [tf = {args.tf}: NOT the config, a profiling run]"
        D = np.ascontiguousarray(c["D"], dtype=np.float64)
        k = np.ascontiguousarray(c["k"], dtype=np.float64)
        Co = np.ascontiguousarray(c["Co"], dtype=np.float64)
        S = D.shape[0]
        dt = params.default_dt(D, k, c["dr"])
        r = params.julia_range(c["dr"], 10.0)
        nout = abi.out_doubles_per_set(o)
        line = {"config": c["name"], "sets": S, "out_bytes_per_set": nout * 8, "devices": args.devices}
        # ---- resident, device 0
        dev = torch.device("cuda", 0)
        td = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        dCo, dD, dk, ddt, dr_ = map(td, (Co, D, k, dt, r))
        dout = torch.empty(S * nout, dtype=torch.float64, device=dev)
        dst = torch.zeros(S, dtype=torch.int32, device=dev)
        dsv = torch.zeros(S, dtype=torch.int32, device=dev)
        dsteps = torch.zeros(S, dtype=torch.int64, device=dev)
        dbc = torch.zeros(S, dtype=torch.int64, device=dev)
        ws = torch.empty(int(lib.gab1_workspace_bytes(S)), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev)

        def step():
            rc = lib.gab1_solve_batch_device(C.byref(o), 0, C.c_void_p(stream.cuda_stream), S, dCo.data_ptr(), 0, dD.data_ptr(),
                                             dk.data_ptr(), ddt.data_ptr(), dr_.data_ptr(), dout.data_ptr(), dst.data_ptr(),
                                             dsv.data_ptr(), dsteps.data_ptr(), dbc.data_ptr(), ws.data_ptr())
            if rc:
                raise RuntimeError(lib.gab1_last_error().decode())

        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        n_steps = dsteps.cpu().numpy().astype(np.float64)
        n_bc = dbc.cpu().numpy().astype(np.float64)
        status = dst.cpu().numpy()
        flops = float((n_steps * (o.Nr - 1) * c["f_int"]).sum() + n_bc.sum() * 241.0)
        line["resident"] = {"ms_per_pass": ms, "solves_per_s": S / (ms * 1e-3), "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
                            "frac_of_measured_fp64_peak": flops / (ms * 1e-3) / 1e12 / peak, "peak_tflops": peak,
                            "median_steps": float(np.median(n_steps)), "membrane_passes_per_step": float(n_bc.sum() / max(n_steps.sum(), 1)),
                            "nan_sets": int((status & abi.ST_NAN != 0).sum()), "threw_sets": int((status & abi.ST_THROW != 0).sum()),
                            "iter_cap_sets": int((status & abi.ST_ITER_CAP != 0).sum())}
        del dout
        torch.cuda.empty_cache()
        if args.no_e2e:
            print(json.dumps(line), flush=True)
            continue
        # ---- end to end through the host entry point, sharded by the library over --devices GPUs
        bufs = []

        def pinned(a_or_n, dtype):
            n = a_or_n if isinstance(a_or_n, int) else a_or_n.size
            p, a = abi.pinned_empty(n, dtype)
            bufs.append(p)
            if not isinstance(a_or_n, int):
                a[:] = np.asarray(a_or_n, dtype=dtype).ravel()
            return a

        hCo, hD, hk, hdt, hr = (pinned(a, np.float64) for a in (Co, D, k, dt, r))
        hout = pinned(S * nout, np.float64)
        hst, hsv = pinned(S, np.int32), pinned(S, np.int32)
        hsteps, hbc = pinned(S, np.int64), pinned(S, np.int64)
        o.n_devices = args.devices

        def host_step():
            rc = lib.gab1_solve_batch(C.byref(o), S, hCo.ctypes.data_as(dp), 0, hD.ctypes.data_as(dp), hk.ctypes.data_as(dp),
                                      hdt.ctypes.data_as(dp), hr.ctypes.data_as(dp), hout.ctypes.data_as(dp),
                                      hst.ctypes.data_as(ip), hsv.ctypes.data_as(ip), hsteps.ctypes.data_as(lp), hbc.ctypes.data_as(lp))
            if rc:
                raise RuntimeError(lib.gab1_last_error().decode())

        host_step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            host_step()
        sec = (time.perf_counter() - t0) / args.steps
        line["e2e"] = {"s_per_pass": sec, "solves_per_s": S / sec, "matches_resident": bool(np.array_equal(hbc, n_bc.astype(np.int64))),
                       "h2d_bytes": int(hCo.nbytes + hD.nbytes + hk.nbytes + hdt.nbytes + hr.nbytes), "d2h_bytes": int(hout.nbytes + 24 * S)}
        for p in bufs:
            abi.pinned_free(p)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
