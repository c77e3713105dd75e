# round-2 session N (1 GPU): the two-warps-per-set latency kernel — parity tests, then the probe
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "duo or family" > gpurun_out/r2n_duo_tests.log 2>&1; tail -15 gpurun_out/r2n_duo_tests.log
timeout 900 python tools/duo_probe.py > gpurun_out/r2n_duo_probe.jsonl 2> gpurun_out/r2n_duo_probe.err; cat gpurun_out/r2n_duo_probe.jsonl; tail -5 gpurun_out/r2n_duo_probe.err
