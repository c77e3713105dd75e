# round-2 session G (2 GPUs): the sharded bench (strong scaling), the library's own multi-device path, debug timeline
set -x
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2g_multi_tests.log 2>&1; tail -3 gpurun_out/r2g_multi_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_r2g_2gpu.json 2> gpurun_out/bench_r2g_2gpu.err; tail -c 2500 gpurun_out/bench_r2g_2gpu.json; tail -5 gpurun_out/bench_r2g_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r2g_2gpu_reference.json 2> gpurun_out/bench_r2g_2gpu_reference.err; cut -c1-300 gpurun_out/bench_r2g_2gpu_reference.json
GAB1_DEBUG_TIMING=1 python - <<'PY' 2>&1 | tail -30
import importlib, sys, time, numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
ens = pkg.params.synthetic_prior_ensemble(100000, seed=123)
pb = np.ascontiguousarray(np.log(ens).T)
for nd in (1, 2):
    fe = pkg.host.Frontend(pkg.abi.CudaBackend(device_ids=list(range(nd))))
    fe.fbatch_dk_mt(pb[:, :2000])
    for rep in range(2):
        t0 = time.perf_counter(); Y = fe.fbatch_dk_mt(pb); print("devices", nd, "rep", rep, "%.1f ms" % (1e3 * (time.perf_counter() - t0)), flush=True)
PY
