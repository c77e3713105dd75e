#!/usr/bin/env python
"""Extracts the call surface of the reference's hot-path functions from its Julia sources into
tests/golden/reference_signatures.json: positional arguments with their type annotations, keyword names with annotations
and default expressions (white space removed).  Run in the build container, where /root/reference exists:

    python tests/golden/make_reference_signatures.py
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT / "tools"))
from julia_signatures import signatures  # noqa: E402

REF = Path("/root/reference/Julia")
FILES = {
    "basepdesolver.jl": ["pdesolver", "pdesolver_membSFK", "pdesolver_fitting"],
    "basepdesolver_rect.jl": ["pdesolver_rect", "pdesolver_membSFK_rect"],
    "pulsechase_solver.jl": ["pulsechase_solver"],
    "get_param_posteriors.jl": ["run_ensemble", "run_ensemble_pc"],
    "sapdesolver.jl": ["sapdesolver", "pmap_fun_allpars", "fbatch", "pmap_fun_dk", "fbatch_dk", "fbatch_dk_mt", "pmap_fun_dk_combD",
                       "fbatch_dk_combD", "pmap_fun_concs", "fbatch_concs", "fbatch_concs_mt"],
    "sapdesolver_memb-SFK.jl": ["sapdesolver_membSFK", "pmap_fun_allpars", "fbatch", "pmap_fun_dk", "fbatch_dk", "fbatch_dk_mt",
                                "pmap_fun_dk_combD", "fbatch_dk_combD", "pmap_fun_concs", "fbatch_concs", "fbatch_concs_mt"],
}


def extract():
    out = {}
    for fname, names in FILES.items():
        sigs = signatures((REF / fname).read_text())
        for n in names:
            assert len(sigs[n]) == 1, (fname, n)
            out[f"{fname}:{n}"] = sigs[n][0]
    return out


if __name__ == "__main__":
    (Path(__file__).resolve().parent / "reference_signatures.json").write_text(json.dumps(extract(), indent=1))
