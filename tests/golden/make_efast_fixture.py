#!/usr/bin/env python
"""Generates tests/golden/efast_concs.npz: the ORACLE's outputs on the eFAST design of the reference's concentration
sensitivity analysis (GSA_concs.jl:50-81: `gsa(fbatch_concs_mt, eFAST(), pbounds; samples=1000, batch=true)` with
pbounds = log.([2e-4*Co, 2*Co]) for the five initial concentrations), for `sapdesolver` and `sapdesolver_membSFK`.

  python tests/golden/make_efast_fixture.py        (about 4 minutes on 8 cores: 4 x 5000 full-length oracle solves)

Two replicates per solver: the design's five phase shifts are random in GlobalSensitivity (Julia's default RNG, not
reproducible here), so each replicate uses five phases from NumPy's PCG64(2024).  tests/test_efast_pin.py compares the
sensitivity indices computed from these outputs with the ones the reference stored (GSA results/*.csv) and re-computes
a random subset of the columns with the oracle to prove the file is the oracle's output.
"""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
SAMPLES, REPS, SEED = 1000, 2, 2024


def designs(pkg, efast):
    Co = pkg.params.base_Co()
    bounds = np.log(np.stack([Co * 0.0002, Co * 2.0], axis=1))          # GSA_concs.jl:62-71
    g = np.random.Generator(np.random.PCG64(SEED))
    phases = np.array([2 * np.pi * g.random(5) for _ in range(REPS)])
    return phases, [efast.design(bounds, SAMPLES, ph) for ph in phases]


def main():
    from oracle import efast, oracle
    pkg = importlib.import_module("myers-furcht-et-al_gab1-shp2-pde-model_b200")
    fe = oracle.frontend()
    phases, ps = designs(pkg, efast)
    out = {"phases": phases}
    for name, memb in (("base", False), ("membSFK", True)):
        out["Y_" + name] = np.array([fe.fbatch_concs_mt(p, membSFK=memb) for p in ps])
    np.savez_compressed(Path(__file__).resolve().parent / "efast_concs.npz", **out)


if __name__ == "__main__":
    main()
