set -x
python - <<'PY' 2>&1 | tail -30
import importlib, os, numpy as np, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from census import per_set_rel_err
from oracle import oracle
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
abi = pkg.abi
Co = pkg.params.base_Co()
ofe = oracle.frontend()
gfe = pkg.host.Frontend(abi.CudaBackend())
pri = pkg.params.synthetic_prior_ensemble(20000, seed=123)
np.set_printoptions(linewidth=200)
for name, Cox, k in (("prior512", Co, dict(dr=0.2, tf=5.0, tol=1e-3, maxiters=20, out_mode=abi.OUT_FINAL_STATE)),
                     ("hela256", pkg.params.hela_Co(), dict(dr=0.2, tf=5.0, tol=1e-3, membSFK=True, out_mode=abi.OUT_FINAL_STATE, iter_cap=300))):
    n = 512 if name == "prior512" else 256
    ref = ofe.sapdesolver_batch(Cox, pri[:n, :7], pri[:n, 7:], **k)
    out = {}
    for kern, extra in (("legacy", {}), ("gang", {}), ("gang", {"GAB1_GANG_NO_RETRY": "1"})):
        os.environ["GAB1_KERNEL"] = kern; os.environ["GAB1_GANG"] = "8,7"
        for kk in ("GAB1_GANG_NO_RETRY",): os.environ.pop(kk, None)
        os.environ.update(extra)
        res = gfe.sapdesolver_batch(Cox, pri[:n, :7], pri[:n, 7:], **k)
        e = per_set_rel_err(res.out, ref.out)
        live = (ref.status & 1) == 0
        bad = np.flatnonzero(live & ((e >= 1e-9) | (res.n_bc_iters != ref.n_bc_iters) | (res.status != ref.status)))
        print(name, kern, extra, "offenders:", [(int(i), float("%.2e" % e[i]), int(res.n_bc_iters[i]), int(ref.n_bc_iters[i]), int(res.status[i]), int(ref.status[i])) for i in bad][:12], flush=True)
        print("   diverging sets with different status:", int(((res.status != ref.status) & ~live).sum()), "of", int((~live).sum()), flush=True)
PY
export GAB1_KERNEL=gang GAB1_GANG=8,7 GAB1_GANG_WARPS=4
python tools/prof_one.py 2368 0.25 && ncu --set full --clock-control none --import-source on -k regex:gang_kernel -c 1 -f -o gpurun_out/r2_gang87_w4 python tools/prof_one.py 2368 0.25 > gpurun_out/ncu_gang87_w4.log 2>&1; tail -2 gpurun_out/ncu_gang87_w4.log
unset GAB1_GANG_WARPS
ncu --set full --clock-control none --import-source on -k regex:gang_kernel -c 1 -f -o gpurun_out/r2_gang87_w8 python tools/prof_one.py 4736 0.25 > gpurun_out/ncu_gang87_w8.log 2>&1; tail -2 gpurun_out/ncu_gang87_w8.log
ls -la gpurun_out/*.ncu-rep
