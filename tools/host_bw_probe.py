#!/usr/bin/env python
"""Device-to-host bandwidth ceiling of the box: N GPUs copying 1 GiB each into pinned host memory at the same time
(cudaMemcpyAsync on one stream per device), for N = 1, 2, 4, 8.  This is the ceiling for configs[1]'s end-to-end number
(2.52 GB of snapshots per rank and pass): whatever the kernels do, the snapshots have to cross PCIe into host DRAM.

  python tools/host_bw_probe.py > gpurun_out/host_bw.jsonl
"""
import json
import time

import torch

GiB = 1 << 30
n_dev = torch.cuda.device_count()
src = [torch.empty(GiB, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n_dev)]
dst = [torch.empty(GiB, dtype=torch.uint8, pin_memory=True) for _ in range(n_dev)]
streams = [torch.cuda.Stream(device=d) for d in range(n_dev)]


def run(n, reps=5):
    for d in range(n):
        with torch.cuda.stream(streams[d]):
            dst[d].copy_(src[d], non_blocking=True)
    for d in range(n):
        streams[d].synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for d in range(n):
            with torch.cuda.stream(streams[d]):
                dst[d].copy_(src[d], non_blocking=True)
    for d in range(n):
        streams[d].synchronize()
    sec = time.perf_counter() - t0
    return n * reps * GiB / sec / 1e9


for n in (1, 2, 4, 8):
    if n <= n_dev:
        gbs = run(n)
        print(json.dumps({"gpus_copying": n, "aggregate_d2h_GBps": round(gbs, 1), "per_gpu_GBps": round(gbs / n, 1),
                          "configs1_solves_per_s_ceiling": round(gbs * 1e9 / 503384.0)}), flush=True)
