# round-2 session Y (1 GPU): the tree after the pipelined K >= 4 halo, 16-byte snapshot stores, certified entry point and the
# latency-lane threshold: GPU tests, both bench arms, A/B of the K = 4 halo, ncu launch list + full captures
set -x
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_gpu_tests.log 2>&1; tail -3 gpurun_out/r2y_gpu_tests.log
timeout 500 python bench.py > gpurun_out/bench_r2y_1gpu.json 2> gpurun_out/bench_r2y_1gpu.err; cut -c1-400 gpurun_out/bench_r2y_1gpu.json; tail -2 gpurun_out/bench_r2y_1gpu.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_r2y_reference.json 2> gpurun_out/bench_r2y_reference.err; cut -c1-200 gpurun_out/bench_r2y_reference.json
AB_TAG=pipe timeout 200 python tools/ab_halo2.py 2>&1 | tee gpurun_out/r2y_ab_pipe.txt | tail -8
AB_TAG=shfl GAB1PDE_LIB=tools/_build/libgab1pde_nopipe.so timeout 200 python tools/ab_halo2.py 2>&1 | tee gpurun_out/r2y_ab_shfl.txt | tail -8
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2y_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/r2y_ncu_launches.log 2>&1; tail -2 gpurun_out/r2y_ncu_launches.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:duo_solve_kernel -s 1 -c 1 -f -o gpurun_out/r2y_duo_solve_kernel_config2 python bench.py --steps 1 --warmup 1 --no-e2e > gpurun_out/r2y_ncu_full.log 2>&1; tail -2 gpurun_out/r2y_ncu_full.log
GAB1_KERNEL=legacy timeout 300 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -f -o gpurun_out/r2y_solve_kernel_k4_pipe python tools/prof_one.py 2368 0.25 0.1 > gpurun_out/r2y_ncu_k4.log 2>&1; tail -1 gpurun_out/r2y_ncu_k4.log
ls -la gpurun_out/*.ncu-rep
