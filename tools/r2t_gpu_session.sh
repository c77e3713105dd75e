# round-2 session T (1 GPU): whole GPU test tier with the latency lane as default, bench at N = 1
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2t_gpu_tests.log 2>&1; tail -5 gpurun_out/r2t_gpu_tests.log
timeout 600 python bench.py > gpurun_out/bench_r2t_1gpu.json 2> gpurun_out/bench_r2t_1gpu.err; tail -c 900 gpurun_out/bench_r2t_1gpu.json; tail -3 gpurun_out/bench_r2t_1gpu.err
