# round-2 session I (8 GPUs): the sharded bench at N = 8 (strong scaling), reference arm, library-sharded call timeline, NUMA placement A/B
set -x
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc; lscpu | grep -i "numa\|socket\|model name" | head -8
python - <<'PY'
import importlib, sys
sys.path.insert(0, ".")
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
lib = pkg.abi.load_library()
print("numa node per device:", [lib.gab1_device_numa_node(d) for d in range(lib.gab1_device_count())])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 > gpurun_out/bench_r2i_8gpu.json 2> gpurun_out/bench_r2i_8gpu.err; tail -c 3000 gpurun_out/bench_r2i_8gpu.json; tail -3 gpurun_out/bench_r2i_8gpu.err
GAB1_NUMA=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 3 > gpurun_out/bench_r2i_8gpu_nonuma.json 2> gpurun_out/bench_r2i_8gpu_nonuma.err; tail -c 1200 gpurun_out/bench_r2i_8gpu_nonuma.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench_r2i_8gpu_reference.json 2> gpurun_out/bench_r2i_8gpu_reference.err; cut -c1-300 gpurun_out/bench_r2i_8gpu_reference.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --steps 3 > gpurun_out/bench_r2i_4gpu.json 2> gpurun_out/bench_r2i_4gpu.err; tail -c 1500 gpurun_out/bench_r2i_4gpu.json
GAB1_DEBUG_TIMING=1 python - <<'PY' 2>&1 | tail -45
import importlib, sys, time, numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
ens = pkg.params.synthetic_prior_ensemble(100000, seed=123)
pb = np.ascontiguousarray(np.log(ens).T)
for nd in (8,):
    fe = pkg.host.Frontend(pkg.abi.CudaBackend(device_ids=list(range(nd))))
    fe.fbatch_dk_mt(pb[:, :8000])
    for rep in range(2):
        t0 = time.perf_counter(); Y = fe.fbatch_dk_mt(pb); print("devices", nd, "rep", rep, "%.1f ms" % (1e3 * (time.perf_counter() - t0)), flush=True)
PY
