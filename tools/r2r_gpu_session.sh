# round-2 session R (1 GPU): latency lane with isolation — parity tests, probe
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "duo or family" > gpurun_out/r2n_duo_tests.log 2>&1; tail -4 gpurun_out/r2n_duo_tests.log
timeout 900 python tools/duo_probe.py shards > gpurun_out/r2r_duo_probe.jsonl 2> gpurun_out/r2r_duo_probe.err; cut -c1-500 gpurun_out/r2r_duo_probe.jsonl; tail -5 gpurun_out/r2r_duo_probe.err
