# round-2 session S (1 GPU): timeline of shard 0 of 8 (library built with -DGAB1_DUO_TIMING)
set -x
GAB1_DUO_TIMELINE=1 GAB1PDE_LIB=tools/_build/libgab1pde_timing.so timeout 300 python - > gpurun_out/r2s_timeline.txt 2>&1 <<'PY'
import importlib, os, sys, time, numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
abi = pkg.abi
fe = pkg.host.Frontend(abi.CudaBackend(arith=abi.ARITH_FAST))
Co = pkg.params.base_Co()
ens = pkg.params.synthetic_prior_ensemble(100_000, seed=123)
dt = pkg.params.default_dt(ens[:, :7], ens[:, 7:], 0.2)
perm, bounds = abi.deal_shards(dt, 5.0, 8)
idx = perm[bounds[0]:bounds[1]]
nt = np.ceil(5.0 / dt[idx])
print("top steps in shard 0:", np.sort(nt)[::-1][:12].astype(int), "total", nt.sum(), "per warp", nt.sum() / 1184)
sh = ens[idx]
t0 = time.perf_counter()
res = fe.sapdesolver_batch(Co, sh[:, :7], sh[:, 7:], dr=0.2, tf=5.0, tol=1e-3, maxiters=20, out_mode=abi.OUT_SIX)
print("ms", (time.perf_counter() - t0) * 1e3)
ok = (res.status & 1) == 0
print("nan sets", (~ok).sum(), "bc/step of live", res.n_bc_iters[ok].sum() / res.n_steps[ok].sum())
o = np.argsort(-res.n_steps)
print("passes per step of the 12 longest:", (res.n_bc_iters[o[:12]] / res.n_steps[o[:12]]).round(2), res.status[o[:12]])
np.savez("gpurun_out/r2s_shard0.npz", steps=res.n_steps, bc=res.n_bc_iters, status=res.status)
PY
grep -c "warp exit" gpurun_out/r2s_timeline.txt; grep "duo item\|^ms\|top steps\|nan sets\|passes per" gpurun_out/r2s_timeline.txt | head -20
grep "warp exit" gpurun_out/r2s_timeline.txt | awk '{print $4}' | sort -n | awk '{a[NR]=$1} END {print "exit times ms: min", a[1], "p10", a[int(NR*0.1)], "p50", a[int(NR*0.5)], "p90", a[int(NR*0.9)], "p99", a[int(NR*0.99)], "max", a[NR]}'
grep "timeline set" gpurun_out/r2s_timeline.txt | sort -k9 -n | tail -40
