# round-2 session B (1 GPU): the gang kernels — parity against the oracle, then timing against the other families
set -x
export PYTHONPATH=.
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
    print(tag, "sets", len(e), "steps_eq", np.array_equal(res.n_steps, ref.n_steps), "saved_eq", np.array_equal(res.n_saved, ref.n_saved),
          "status_eq", np.array_equal(res.status, ref.status), "iters_eq_live", np.array_equal(res.n_bc_iters[live], ref.n_bc_iters[live]),
          "max_err_live %.3e" % (e[live].max() if live.any() else 0), flush=True)
# parity: short runs, every output mode / variant, gang shapes
for shape, dr in (("8,7", 0.2), ("4,13", 0.2), ("8,13", 0.1), ("4,7", 0.4), ("2,13", 0.4), ("4,10", 0.25), ("16,13", 0.05), ("32,13", 0.025), ("16,7", 0.1)):
    os.environ["GAB1_KERNEL"] = "gang"; os.environ["GAB1_GANG"] = shape
    rows = [0, 1, 2, 3, 75, 333, 4999, 10, 11, 12, 13]
    tf = {0.4: 0.6, 0.25: 0.5, 0.2: 0.5, 0.1: 0.2, 0.05: 0.05, 0.025: 0.012}[dr]
    for name, kw in (("full", dict()), ("membSFK", dict(sfk_mode=1)), ("rect", dict(geometry=1, pg1tot_form=1)),
                     ("rect_mod", dict(geometry=1, pg1tot_form=1, sfk_mode=2, save_rule=1)), ("pulse", dict(t_prechase=tf * 0.4)),
                     ("mask", dict(matrices=("aSFK", "PG1S", "G2PG1S"))), ("final_state", dict(out_mode=abi.OUT_FINAL_STATE)),
                     ("pct", dict(out_mode=abi.OUT_PCT_BOUND, pct_mul=2.0, pct_div=3.0))):
        k = dict(dr=dr, tf=tf, Nts=8, tol=1e-4, maxiters=20, **kw)
        res = gfe.pdesolver_batch(Co, ens[rows, :7], ens[rows, 7:], **k)
        ref = ofe.pdesolver_batch(Co, ens[rows, :7], ens[rows, 7:], **k)
        cmp(f"gang {shape} dr={dr} {name}", res, ref)
    for memb in (False, True):
        for mode in (abi.OUT_FINAL4, abi.OUT_SIX):
            k = dict(dr=dr, tf=tf, membSFK=memb, out_mode=mode, iter_cap=500)
            res = gfe.sapdesolver_batch(pkg.params.hela_Co() if memb else Co, ens[rows, :7], ens[rows, 7:], **k)
            ref = ofe.sapdesolver_batch(pkg.params.hela_Co() if memb else Co, ens[rows, :7], ens[rows, 7:], **k)
            cmp(f"gang {shape} dr={dr} sa memb={memb} mode={mode}", res, ref)
PY
python - <<'PY' 2>&1 | tail -30
import importlib, os, time, numpy as np, sys
sys.path.insert(0, ".")
pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
abi = pkg.abi
ens = pkg.params.load_parameter_ensemble()
Co = pkg.params.base_Co()
gfe = pkg.host.Frontend(abi.CudaBackend())
pri = pkg.params.synthetic_prior_ensemble(20000, seed=123)
def timeit(tag, f, n=2):
    f(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    print(tag, "%.1f ms" % (1e3 * min(ts)), flush=True)
for kern, shape in (("legacy", ""), ("gang", "8,7"), ("gang", "4,13")):
    os.environ["GAB1_KERNEL"] = kern; os.environ["GAB1_GANG"] = shape
    for w in (("", ) if kern == "legacy" else ("", "4", "6")):
        if w: os.environ["GAB1_GANG_WARPS"] = w
        else: os.environ.pop("GAB1_GANG_WARPS", None)
        timeit(f"{kern} {shape} warps={w or 'max'} config1 5000 sets final_state", lambda: gfe.pdesolver_batch(Co, ens[:, :7], ens[:, 7:], dr=0.2, tol=1e-4, maxiters=20, out_mode=abi.OUT_FINAL_STATE))
        timeit(f"{kern} {shape} warps={w or 'max'} config2 20000 prior draws six", lambda: gfe.sapdesolver_batch(Co, pri[:, :7], pri[:, 7:], dr=0.2, tol=1e-3, maxiters=20, out_mode=abi.OUT_SIX))
os.environ.pop("GAB1_GANG_WARPS", None)
for kern, shape, dr, n in (("legacy", "", 0.1, 2368), ("gang", "8,13", 0.1, 2368), ("gang", "16,7", 0.1, 2368), ("stream", "", 0.05, 592), ("gang", "16,13", 0.05, 592), ("legacy", "", 0.4, 5000), ("gang", "4,7", 0.4, 5000), ("gang", "2,13", 0.4, 5000)):
    os.environ["GAB1_KERNEL"] = kern; os.environ["GAB1_GANG"] = shape
    geo = dict(geometry=1, pg1tot_form=1) if dr == 0.05 else {}
    timeit(f"{kern} {shape} dr={dr} {n} sets final_state", lambda: gfe.pdesolver_batch(Co, ens[:n, :7], ens[:n, 7:], dr=dr, tol=1e-4, maxiters=20, out_mode=abi.OUT_FINAL_STATE, **geo), n=1)
PY
