# round-2 session Z2 (2 GPUs): final tree — multi-GPU tests, both bench arms at N = 2
set -x
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 > gpurun_out/bench_r2z_2gpu.json 2> gpurun_out/bench_r2z_2gpu.err; cut -c1-300 gpurun_out/bench_r2z_2gpu.json; tail -2 gpurun_out/bench_r2z_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --impl reference > gpurun_out/bench_r2z_2gpu_reference.json 2> gpurun_out/bench_r2z_2gpu_reference.err; cut -c1-200 gpurun_out/bench_r2z_2gpu_reference.json
