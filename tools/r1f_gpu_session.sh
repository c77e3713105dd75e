# round-1 closing session: every GPU test, smoke, both bench arms, latency lines
set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1f_gpu_tests.log 2>&1; tail -3 gpurun_out/r1f_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1f_reference.json 2> gpurun_out/bench_r1f_reference.err
timeout 900 python bench.py > gpurun_out/bench_r1f_1gpu.json 2> gpurun_out/bench_r1f_1gpu.err; tail -c 300 gpurun_out/bench_r1f_1gpu.json
timeout 600 python tools/bench_small.py > gpurun_out/small_batches_r1f.txt 2>&1; grep -c ratio gpurun_out/small_batches_r1f.txt
rm -f gpurun_out/tangent_latency_r1f.jsonl
for cfg in "1 0.2" "1 0.1" "101 0.2" "101 0.1"; do set -- $cfg; timeout 300 python tools/bench_tangent.py --reps 2 --dr $2 --sets $1 --tf 5.0 2>&1 | tail -1 >> gpurun_out/tangent_latency_r1f.jsonl; done; cut -c1-200 gpurun_out/tangent_latency_r1f.jsonl
timeout 300 python tools/bench_configs.py --configs 1 --steps 3 2>/dev/null | grep "^{" > gpurun_out/config0_r1f.jsonl; cut -c1-260 gpurun_out/config0_r1f.jsonl
