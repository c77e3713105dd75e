# round-1 final session: every GPU test, both bench arms, smoke, launch list of the bench command, other configs, forward mode
set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1e_gpu_tests.log 2>&1; tail -3 gpurun_out/r1e_gpu_tests.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1e_reference.json 2> gpurun_out/bench_r1e_reference.err
timeout 900 python bench.py > gpurun_out/bench_r1e_1gpu.json 2> gpurun_out/bench_r1e_1gpu.err; tail -c 400 gpurun_out/bench_r1e_1gpu.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches_r1e.log 2>&1
timeout 900 python tools/bench_configs.py --configs 1,3,5 --steps 2 2>/dev/null | grep "^{" > gpurun_out/configs_r1e_1gpu.jsonl; cut -c1-160 gpurun_out/configs_r1e_1gpu.jsonl
rm -f gpurun_out/tangent_r1e.jsonl; for cfg in "0.2 4736 5.0" "0.1 1184 5.0" "0.4 4736 5.0"; do set -- $cfg; timeout 300 python tools/bench_tangent.py --reps 2 --dr $1 --sets $2 --tf $3 2>&1 | tail -1 >> gpurun_out/tangent_r1e.jsonl; done; cut -c1-250 gpurun_out/tangent_r1e.jsonl
