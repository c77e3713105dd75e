# round-2 session E (1 GPU): gang kernels v2 (membrane constants in shared memory, bank-skewed exchange areas)
set -x
python - <<'PY' 2>&1 | tail -40
import importlib, os, time, numpy as np, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from census import per_set_rel_err
from oracle import oracle
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
abi = pkg.abi
ens = pkg.params.load_parameter_ensemble()
Co = pkg.params.base_Co()
ofe = oracle.frontend()
gfe = pkg.host.Frontend(abi.CudaBackend())
def cmp(tag, res, ref):
    live = (ref.status & 1) == 0
    e = per_set_rel_err(res.out, ref.out)
    print(tag, "sets", len(e), "steps_eq", np.array_equal(res.n_steps, ref.n_steps), "saved_eq", np.array_equal(res.n_saved, ref.n_saved), "status_eq", np.array_equal(res.status, ref.status),
          "iters_eq_live", np.array_equal(res.n_bc_iters[live], ref.n_bc_iters[live]), "max_err_live %.3e" % (e[live].max() if live.any() else 0), flush=True)
os.environ["GAB1_KERNEL"] = "gang"
for shape, dr, tf in (("8,7", 0.2, 0.5), ("4,13", 0.2, 0.5), ("16,13", 0.05, 0.05), ("2,13", 0.4, 0.6), ("32,13", 0.025, 0.012)):
    os.environ["GAB1_GANG"] = shape
    rows = [0, 1, 2, 3, 75, 333, 4999, 10, 11, 12, 13]
    for name, kw in (("full", dict()), ("membSFK", dict(sfk_mode=1)), ("rect_mod", dict(geometry=1, pg1tot_form=1, sfk_mode=2, save_rule=1)), ("pulse", dict(t_prechase=tf * 0.4)), ("pct", dict(out_mode=abi.OUT_PCT_BOUND, pct_mul=2.0, pct_div=3.0))):
        k = dict(dr=dr, tf=tf, Nts=8, tol=1e-4, maxiters=20, **kw)
        cmp(f"gang {shape} {name}", gfe.pdesolver_batch(Co, ens[rows, :7], ens[rows, 7:], **k), ofe.pdesolver_batch(Co, ens[rows, :7], ens[rows, 7:], **k))
    for memb in (False, True):
        k = dict(dr=dr, tf=tf, membSFK=memb, out_mode=abi.OUT_SIX, iter_cap=500)
        cmp(f"gang {shape} six memb={memb}", gfe.sapdesolver_batch(pkg.params.hela_Co() if memb else Co, ens[rows, :7], ens[rows, 7:], **k), ofe.sapdesolver_batch(pkg.params.hela_Co() if memb else Co, ens[rows, :7], ens[rows, 7:], **k))
PY
python - <<'PY' 2>&1 | tail -40
import importlib, os, time, numpy as np, sys
sys.path.insert(0, ".")
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
abi = pkg.abi
ens = pkg.params.load_parameter_ensemble()
Co = pkg.params.base_Co()
gfe = pkg.host.Frontend(abi.CudaBackend())
pri = pkg.params.synthetic_prior_ensemble(20000, seed=123)
good = ens[np.setdiff1d(np.arange(5000), [75])][:4736]
def timeit(tag, f, n=2):
    f(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    print(tag, "%.1f ms" % (1e3 * min(ts)), flush=True)
cfg1 = lambda e: (lambda: gfe.pdesolver_batch(Co, e[:, :7], e[:, 7:], dr=0.2, tol=1e-4, maxiters=20, out_mode=abi.OUT_FINAL_STATE))
cfg2 = lambda e: (lambda: gfe.sapdesolver_batch(Co, e[:, :7], e[:, 7:], dr=0.2, tol=1e-3, maxiters=20, out_mode=abi.OUT_SIX))
os.environ["GAB1_KERNEL"] = "legacy"
timeit("legacy config1 5000", cfg1(ens)); timeit("legacy config2 20000 prior", cfg2(pri)); timeit("legacy config1-tol 18944 tiled posterior", cfg1(np.tile(good, (4, 1))), n=1)
os.environ["GAB1_KERNEL"] = "gang"; os.environ["GAB1_GANG"] = "8,7"
for w in ("", "4"):
    if w: os.environ["GAB1_GANG_WARPS"] = w
    else: os.environ.pop("GAB1_GANG_WARPS", None)
    timeit(f"gang 8,7 warps={w or 'max'} config1 5000", cfg1(ens))
    timeit(f"gang 8,7 warps={w or 'max'} config2 20000 prior", cfg2(pri))
    timeit(f"gang 8,7 warps={w or 'max'} config1-tol 18944 tiled posterior", cfg1(np.tile(good, (4, 1))), n=1)
os.environ.pop("GAB1_GANG_WARPS", None)
os.environ["GAB1_GANG"] = "4,13"
timeit("gang 4,13 config1-tol 18944", cfg1(np.tile(good, (4, 1))), n=1)
for kern, shape, dr, n in (("legacy", "", 0.1, 2368), ("gang", "8,13", 0.1, 2368), ("gang", "16,7", 0.1, 2368), ("stream", "", 0.05, 1184), ("gang", "16,13", 0.05, 1184), ("gang", "32,7", 0.05, 1184), ("group16", "", 0.4, 9472), ("gang", "4,7", 0.4, 9472), ("gang", "2,13", 0.4, 9472), ("stream", "", 0.025, 592), ("gang", "32,13", 0.025, 592)):
    os.environ["GAB1_KERNEL"] = kern; os.environ["GAB1_GANG"] = shape
    geo = dict(geometry=1, pg1tot_form=1) if dr == 0.05 else {}
    tf = 1.0 if dr <= 0.05 else 5.0
    e = np.tile(good, (2, 1))[:n]
    timeit(f"{kern} {shape} dr={dr} {n} sets tf={tf} final_state", lambda: gfe.pdesolver_batch(Co, e[:, :7], e[:, 7:], dr=dr, tf=tf, tol=1e-4, maxiters=20, out_mode=abi.OUT_FINAL_STATE, **geo), n=1)
PY
export GAB1_KERNEL=gang GAB1_GANG=8,7
ncu --set full --clock-control none --import-source on -k regex:gang_kernel -c 1 -f -o gpurun_out/r2_gang87v2_w8 python tools/prof_one.py 4736 0.25 > gpurun_out/ncu_gang87v2_w8.log 2>&1; tail -1 gpurun_out/ncu_gang87v2_w8.log
