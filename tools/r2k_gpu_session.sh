# round-2 session K (1 GPU): GPU tests at the current tree, both bench arms, ncu evidence
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2k_gpu_tests.log 2>&1; tail -3 gpurun_out/r2k_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2k_reference.json 2> gpurun_out/bench_r2k_reference.err; cut -c1-200 gpurun_out/bench_r2k_reference.json
timeout 600 python bench.py > gpurun_out/bench_r2k_1gpu.json 2> gpurun_out/bench_r2k_1gpu.err; tail -c 600 gpurun_out/bench_r2k_1gpu.json; tail -3 gpurun_out/bench_r2k_1gpu.err
bash tools/r2h_gpu_session.sh
timeout 600 python tools/bench_configs.py --configs 1,4 --steps 2 --cap 2000 2>/dev/null | grep "^{" > gpurun_out/r2k_configs_0_3.jsonl; cut -c1-500 gpurun_out/r2k_configs_0_3.jsonl
