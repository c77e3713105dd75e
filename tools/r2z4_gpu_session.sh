# round-2 session Z4 (4 GPUs): final tree — bench at N = 4
set -x
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 4 > gpurun_out/bench_r2z_4gpu.json 2> gpurun_out/bench_r2z_4gpu.err; cut -c1-300 gpurun_out/bench_r2z_4gpu.json; tail -2 gpurun_out/bench_r2z_4gpu.err
