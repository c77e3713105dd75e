# round-2 session F (1 GPU): whole GPU test-suite with the gang family as the default for large grids; configs[4] and the finest grid
set -x
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_gpu_tests.log 2>&1; tail -5 gpurun_out/r2f_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python tools/bench_configs.py --configs 5 --steps 2 2>/dev/null | grep "^{" > gpurun_out/r2f_config5.jsonl; cut -c1-600 gpurun_out/r2f_config5.jsonl
GAB1_KERNEL=stream timeout 900 python tools/bench_configs.py --configs 5 --steps 2 --no-e2e 2>/dev/null | grep "^{" > gpurun_out/r2f_config5_stream.jsonl; cut -c1-400 gpurun_out/r2f_config5_stream.jsonl
