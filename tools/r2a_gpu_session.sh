# round-2 session A (1 GPU): every GPU test incl. the parity census, the new bench line, cap sweep, strict-kernel timing
set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
nproc
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_census.py > gpurun_out/r2a_gpu_tests.log 2>&1; tail -3 gpurun_out/r2a_gpu_tests.log
timeout 900 python -m pytest tests/test_gpu_census.py -m gpu -q -s > gpurun_out/r2a_census.log 2>&1; tail -15 gpurun_out/r2a_census.log | cut -c1-600
timeout 600 python bench.py > gpurun_out/bench_r2a_1gpu.json 2> gpurun_out/bench_r2a_1gpu.err; tail -c 1500 gpurun_out/bench_r2a_1gpu.json; tail -5 gpurun_out/bench_r2a_1gpu.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2a_reference.json 2> gpurun_out/bench_r2a_reference.err; cut -c1-300 gpurun_out/bench_r2a_reference.json
timeout 600 python tools/cap_sweep.py --sets 20000 > gpurun_out/cap_sweep_posterior.jsonl 2> gpurun_out/cap_sweep.err; cut -c1-400 gpurun_out/cap_sweep_posterior.jsonl
timeout 600 python tools/cap_sweep.py --sets 20000 --prior --caps 20,40,80,160,320,640,1280,5000,20000 > gpurun_out/cap_sweep_prior.jsonl 2>> gpurun_out/cap_sweep.err; cut -c1-400 gpurun_out/cap_sweep_prior.jsonl
python - <<'PY' 2>&1 | tail -4
import importlib, time, numpy as np, sys
sys.path.insert(0, ".")
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
ens = pkg.params.load_parameter_ensemble()
Co = pkg.params.base_Co()
for arith, name in ((1, "strict"), (0, "fast")):
    fe = pkg.host.Frontend(pkg.abi.CudaBackend(arith=arith))
    for n in (1, 16, 1184):
        fe.pdesolver_batch(Co, ens[:n, :7], ens[:n, 7:], dr=0.2, tol=1e-4, maxiters=20, out_mode=pkg.abi.OUT_FINAL_STATE)
        t0 = time.perf_counter()
        fe.pdesolver_batch(Co, ens[:n, :7], ens[:n, 7:], dr=0.2, tol=1e-4, maxiters=20, out_mode=pkg.abi.OUT_FINAL_STATE)
        print(name, n, "sets:", round(1e3 * (time.perf_counter() - t0), 1), "ms", flush=True)
PY
