# round-2 session U (8 GPUs): bench line at N = 8 with the latency lane, multi-GPU tests
set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 > gpurun_out/bench_r2u_8gpu.json 2> gpurun_out/bench_r2u_8gpu.err; tail -c 1800 gpurun_out/bench_r2u_8gpu.json | cut -c1-1800; tail -2 gpurun_out/bench_r2u_8gpu.err
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2
