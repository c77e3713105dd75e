# round-2 session L (8 GPUs): host-path ceiling, final bench lines at N = 8 (both arms)
set -x
timeout 300 python tools/host_bw_probe.py > gpurun_out/r2_host_bw.jsonl 2> gpurun_out/r2_host_bw.err; cat gpurun_out/r2_host_bw.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 > gpurun_out/bench_r2l_8gpu.json 2> gpurun_out/bench_r2l_8gpu.err; tail -c 1500 gpurun_out/bench_r2l_8gpu.json; tail -2 gpurun_out/bench_r2l_8gpu.err
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2
