set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r1c_gpu_tests.log 2>&1; tail -3 gpurun_out/r1c_gpu_tests.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1c_reference.json 2> gpurun_out/bench_r1c_reference.err; tail -c 600 gpurun_out/bench_r1c_reference.json
python bench.py > gpurun_out/bench_r1c_1gpu.json 2> gpurun_out/bench_r1c_1gpu.err; tail -c 1500 gpurun_out/bench_r1c_1gpu.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for cfg in "0.2 4736 5.0" "0.1 1184 5.0" "0.4 4736 5.0"; do set -- $cfg; python tools/bench_tangent.py --reps 2 --dr $1 --sets $2 --tf $3 2>&1 | tail -1 >> gpurun_out/tangent_r1c.jsonl; done; cat gpurun_out/tangent_r1c.jsonl
