"""Whole-configuration parity census: the product (fast) kernels against the CPU oracle, set by set.

For every parameter set of a batch it records whether the data-dependent control flow (step count, snapshot count,
status word, membrane-iteration count) is the oracle's and how far the values are from the oracle's, so that the
"within 1e-9" claim is a count over ALL sets of a BASELINE configuration instead of a sample:

  flipped set      a non-diverging set whose total membrane-iteration count differs from the oracle's: at some step the
                   fixed point's exit test `|1 - new/old| <= tol` (basepdesolver.jl:238-241) fell on the other side
                   because the two arithmetic forms differ in the last bits;
  diverging set    the oracle's status carries GAB1_ST_NAN (the reference drops / zeroes such sets:
                   get_param_posteriors.jl:155, sapdesolver.jl:378-382); only its NaN pattern and status are compared.

Used by tests/test_gpu_census.py (asserts) and tools/parity_census.py (writes the JSON report kept under profiles/).
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

ST_NAN = 1


def per_set_rel_err(a, b):
    """Per set: max |a-b| / max(|b|, 1e-6 * max|b| over the set's block); positions where both are NaN count as equal,
    a NaN on one side only gives inf."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    fin = np.isfinite(b)
    scale = np.where(fin, np.abs(b), 0).max(axis=-1, keepdims=True)
    den = np.maximum(np.abs(b), 1e-6 * scale)
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.where(fin & (den > 0), np.abs(a - b) / den, 0.0)
    e = np.where(np.isnan(a) != np.isnan(b), np.inf, e)
    e = np.where(np.isnan(e), np.inf, e)
    return e.max(axis=-1)


def ill_conditioned(ref, ref_perturbed, rtol: float = 1e-9):
    """Sets whose REFERENCE answer is not defined to `rtol`: the oracle solved them twice, the second time with every
    initial concentration moved by one unit in the last place, and the two answers differ by a tenth of `rtol` or more
    (or in their control flow).  The explicit scheme's time step ignores the second-order rate x concentration terms
    (basepdesolver.jl:30), so some wide-prior sets sit at the edge of stability, where an alternating mode amplifies
    rounding differences by 1e5..1e13 over the 4e4 steps without blowing up; any arithmetic that is not bit-identical
    to the reference's (a different `sum(k)` order in Julia itself would do) lands somewhere else on such a set."""
    e = per_set_rel_err(ref_perturbed.out, ref.out)
    return (e >= 0.1 * rtol) | (ref.n_bc_iters != ref_perturbed.n_bc_iters) | (ref.status != ref_perturbed.status)


def census(res, ref, *, name: str, rtol: float = 1e-9, exact_columns=None, max_listed: int = 50, ill=None) -> dict:
    """res / ref: BatchResult of the CUDA path and of the oracle on the same inputs.  `ill`: mask from ill_conditioned();
    those sets are reported separately and excluded from the well-conditioned counts."""
    S = ref.out.shape[0]
    div = (ref.status & ST_NAN) != 0
    if ill is None:
        ill = np.zeros(S, dtype=bool)
    live = ~div & ~ill
    err = per_set_rel_err(res.out, ref.out)
    flipped = live & (res.n_bc_iters != ref.n_bc_iters)
    same_flow = live & ~flipped
    nan_pattern_same = np.array([np.array_equal(np.isnan(res.out[i]), np.isnan(ref.out[i])) for i in np.flatnonzero(div)], dtype=bool)
    ill_live = ill & ~div
    rep = {
        "config": name, "sets": int(S), "rtol": rtol,
        "ill_conditioned_sets": int(ill_live.sum()),
        "ill_conditioned_same_iteration_count": int((res.n_bc_iters[ill_live] == ref.n_bc_iters[ill_live]).sum()),
        "ill_conditioned_within_rtol": int((err[ill_live] < rtol).sum()),
        "ill_conditioned": [{"set": int(i), "iters_gpu": int(res.n_bc_iters[i]), "iters_oracle": int(ref.n_bc_iters[i]),
                             "rel_err": float(err[i])} for i in np.flatnonzero(ill_live)[:max_listed]],
        "diverging_sets": int(div.sum()),
        "diverging_sets_same_nan_pattern": int(nan_pattern_same.sum()),
        "diverging_sets_same_status": int((res.status[div] == ref.status[div]).sum()),
        "diverging_sets_same_iteration_count": int((res.n_bc_iters[div] == ref.n_bc_iters[div]).sum()),
        "step_count_mismatches": int((res.n_steps != ref.n_steps).sum()),
        "snapshot_count_mismatches": int((res.n_saved != ref.n_saved).sum()),
        "status_mismatches": int((res.status != ref.status).sum()),
        "status_mismatches_live": int((res.status[live] != ref.status[live]).sum()),
        "flipped_sets": int(flipped.sum()),
        "same_flow_sets": int(same_flow.sum()),
        "max_rel_err_same_flow": float(err[same_flow].max()) if same_flow.any() else 0.0,
        "same_flow_sets_over_rtol": int((err[same_flow] >= rtol).sum()),
        "max_rel_err_flipped": float(err[flipped].max()) if flipped.any() else 0.0,
        "flipped_sets_over_rtol": int((err[flipped] >= rtol).sum()),
        "live_sets_within_rtol": int((err[live] < rtol).sum()),
        "live_sets": int(live.sum()),
        "mean_iterations_per_step": float(ref.n_bc_iters[live].sum() / max(ref.n_steps[live].sum(), 1)),
        "flipped": [{"set": int(i), "iters_gpu": int(res.n_bc_iters[i]), "iters_oracle": int(ref.n_bc_iters[i]),
                     "rel_err": float(err[i])} for i in np.flatnonzero(flipped)[:max_listed]],
    }
    if exact_columns is not None:     # outputs that are grid-quantised (the four length scales): must be identical
        c = list(exact_columns)
        same = (res.out[:, c] == ref.out[:, c]) | (np.isnan(res.out[:, c]) & np.isnan(ref.out[:, c]))
        rep["exact_column_mismatches_live"] = int((~same[live]).any(axis=1).sum())
    return rep


def dump(rep: dict, directory="gpurun_out") -> None:
    d = Path(__file__).resolve().parent.parent / directory
    if d.is_dir():
        (d / f"census_{rep['config']}.json").write_text(json.dumps(rep, indent=1))
