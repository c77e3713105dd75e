# round-2 session Q (1 GPU): latency lane — parity tests, phase timing, probe
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "duo or family" > gpurun_out/r2n_duo_tests.log 2>&1; tail -15 gpurun_out/r2n_duo_tests.log
bash tools/r2p_gpu_session.sh
timeout 900 python tools/duo_probe.py > gpurun_out/r2n_duo_probe.jsonl 2> gpurun_out/r2n_duo_probe.err; cut -c1-400 gpurun_out/r2n_duo_probe.jsonl; tail -5 gpurun_out/r2n_duo_probe.err
