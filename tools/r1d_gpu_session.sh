# round-1 session D: parity after the plain-shared-memory change, streamed-kernel timings, ncu capture on a short run
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1d_gpu_tests.log 2>&1; tail -3 gpurun_out/r1d_gpu_tests.log
timeout 300 python tools/bench_configs.py --configs 5 --scale 1.0 --steps 2 --no-e2e 2>/dev/null | tail -1 > gpurun_out/cfg5_stream_plain.json; cut -c1-520 gpurun_out/cfg5_stream_plain.json
for cfg in "stream 4 0.2 4736 5.0" "stream 2 0.2 4736 5.0" "stream 4 0.1 1184 1.0"; do set -- $cfg; GAB1_TANGENT=$1 GAB1_TANGENT_NT=$2 timeout 300 python tools/bench_tangent.py --reps 2 --dr $3 --sets $4 --tf $5 2>&1 | tail -1 | cut -c1-330; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_kernel --launch-skip 1 -c 1 -o gpurun_out/prof_r1_stream_k8 -f python tools/bench_configs.py --configs 5 --scale 1.0 --steps 1 --tf 0.02 --no-e2e > gpurun_out/ncu_stream.log 2>&1; tail -2 gpurun_out/ncu_stream.log | cut -c1-300
timeout 300 python tools/bench_configs.py --configs 5 --scale 1.0 --steps 2 --tf 0.02 --no-e2e 2>/dev/null | tail -1 > gpurun_out/cfg5_stream_short.json
