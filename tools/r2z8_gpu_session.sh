# round-2 session Z8 (8 GPUs): final tree — bench at N = 8
set -x
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 8 > gpurun_out/bench_r2z_8gpu.json 2> gpurun_out/bench_r2z_8gpu.err; cut -c1-300 gpurun_out/bench_r2z_8gpu.json; tail -2 gpurun_out/bench_r2z_8gpu.err
