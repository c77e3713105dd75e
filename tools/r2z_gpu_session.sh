# round-2 session Z (1 GPU): final tree — GPU tests, both bench arms, small-batch policy of the latency lane, ncu launch list
set -x
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_gpu_tests.log 2>&1; tail -3 gpurun_out/r2z_gpu_tests.log
timeout 500 python bench.py > gpurun_out/bench_r2z_1gpu.json 2> gpurun_out/bench_r2z_1gpu.err; cut -c1-300 gpurun_out/bench_r2z_1gpu.json; tail -2 gpurun_out/bench_r2z_1gpu.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_r2z_reference.json 2> gpurun_out/bench_r2z_reference.err; cut -c1-200 gpurun_out/bench_r2z_reference.json
timeout 200 python tools/duo_probe.py small > gpurun_out/r2z_duo_small.jsonl 2> gpurun_out/r2z_duo_small.err; cut -c1-220 gpurun_out/r2z_duo_small.jsonl
timeout 200 python tools/ab_halo2.py 2>&1 | tee gpurun_out/r2z_ab.txt | tail -8
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/r2z_ncu_launches.log 2>&1; tail -2 gpurun_out/r2z_ncu_launches.log
python -c "import __graft_entry__ as g; g.smoke()"
