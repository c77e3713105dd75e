# round-2 session Z5 (1 GPU): the other BASELINE configs at the final tree: configs[0] (one solve), configs[3] (HeLa membSFK), configs[4] (planar Nr = 200)
set -x
timeout 200 python tools/bench_configs.py --configs 1,4,5 --steps 2 --cap 2000 2>/dev/null | grep "^{" > gpurun_out/r2z_configs_0_3_4.jsonl; cut -c1-420 gpurun_out/r2z_configs_0_3_4.jsonl
