"""Parameter supply for the batched solver: grids, time step, baseline vectors, synthetic ensembles.

Everything here is host-side set-up that runs once per batch; none of it is on the hot path.
"""
from __future__ import annotations

from fractions import Fraction
from pathlib import Path

import numpy as np

PNAMES = ("Dsfk", "Dg2", "Dg2g1", "Dg2g1s2", "Dg1", "Dg1s2", "Ds2",
          "kS2f", "kS2r", "kG1f", "kG1r", "kG2f", "kG2r", "kG1p", "kG1dp",
          "kSa", "kSi", "kp", "kdp", "kEGFf", "kEGFr", "EGF", "kdf", "kdr")   # get_param_posteriors.jl:24-26

# Baseline diffusivities / rate constants.  The reference derives them by running its prior pipeline
# (get_param_priors.jl:270-301) and taking posterior medians in log space for the four fitted constants
# (get_param_posteriors.jl:17-34).  That pipeline needs Julia packages; the values below are the rounded
# prior modes the reference itself records in MATLAB/run_base_model.m:8-37 together with the posterior
# medians recomputed from Turing results/Turing_res_5-chains_1000-spls_posteriors.csv.
DIFFS_BASE = np.array([84.0, 136.0, 62.0, 56.0, 67.0, 57.0, 80.0])
KVALS_BASE = np.array([1.594, 480.0, 8.842e-4, 0.1227, 1.594, 480.0,
                       1.266519301152554, 3.117916270722422, 0.7924254053962531, 4.665683702990119,
                       13.84, 41.21, 55.84, 0.1301, 0.00167, 1.2, 0.456])
# fitted_parameters.csv (kG1p, kG1dp, kSa, kSi)
FITTED_MLE = {"kG1p": 41.999999999999964, "kG1dp": 0.09499999999999997, "kSa": 16.175675458812922, "kSi": 0.09499999999999997}


def conversion_factors(R: float = 10.0):
    """volCF, surfCF of run_base_model.jl:67-68."""
    volCF = 1 / (4 / 3 * np.pi * R**3)
    surfCF = 1 / (4 * np.pi * R**2)
    return volCF, surfCF


def base_Co(R: float = 10.0) -> np.ndarray:
    """Co of run_base_model.jl:71-76: 6e5 copies of every protein."""
    volCF, surfCF = conversion_factors(R)
    return np.array([6.0e5 * volCF] * 4 + [6.0e5 * surfCF])


def hela_Co(R: float = 10.0) -> np.ndarray:
    """Co_hela of run_base_model_HeLa.jl:71-83."""
    volCF, surfCF = conversion_factors(R)
    return np.array([166000.0 * volCF, 628000.0 * volCF, 1530.0 * volCF, 3.0e5 * volCF, 93000.0 * surfCF])


def julia_range(dr: float, R: float) -> np.ndarray:
    """`collect(0.0:dr:R)` (basepdesolver.jl:73).

    Julia builds float ranges in twice precision from the rational form of the step, so element j is the
    correctly rounded value of j*step_num/step_den, not fl(j*dr).  For the decimal steps the reference uses
    (0.4, 0.25, 0.2, 0.1, 0.05, 0.025) that is j/den with one IEEE division.
    """
    f = Fraction(dr).limit_denominator(10_000)
    if float(f) != dr:
        raise ValueError(f"dr={dr!r} is not a short decimal; pass r explicitly")
    n = int(np.floor(Fraction(R).limit_denominator(10_000) / f))
    j = np.arange(n + 1, dtype=np.float64)
    return j * f.numerator / f.denominator if f.numerator != 1 else j / f.denominator


def default_dt(D, k, dr: float) -> np.ndarray:
    """`1.0/(2.0*(maximum(D)/(dr.^2) + sum(k)/4))*0.99` (basepdesolver.jl:30) per row, sum(k) left to right."""
    D = np.asarray(D, dtype=np.float64).reshape(-1, 7)
    k = np.asarray(k, dtype=np.float64).reshape(-1, 17)
    sk = np.zeros(k.shape[0])
    for q in range(17):
        sk = sk + k[:, q]
    return 1.0 / (2.0 * (D.max(axis=1) / (dr * dr) + sk / 4)) * 0.99


def load_parameter_ensemble() -> np.ndarray:
    """The reference's parameter_ensemble.csv (5000 x 24, columns PNAMES) as committed under tests/golden."""
    p = Path(__file__).resolve().parent.parent / "tests" / "golden" / "parameter_ensemble.npy"
    return np.load(p)


# (mu, sigma) of log-normal priors reproduced from get_param_priors.jl:19-198 via calcModeSpread.
_UV = {
    "Dsfk": (4.42968, 0.04759), "Dg2": (4.91133, 0.05144), "Dg2g1": (4.12597, 0.04836), "Dg2g1s2": (4.02392, 0.05353),
    "Dg1": (4.20291, 0.05965), "Dg1s2": (4.04167, 0.05260), "Ds2": (4.38078, 0.04997),
    "kp": (2.62771, 0.28676), "kdp": (3.71872, 0.93048),
    "kG1p": (-0.86750, 2.30257), "kSa": (-0.86750, 2.30257), "kG1dp": (2.25129, 2.30257), "kSi": (2.25129, 2.30257),
}
_EGF = 1.67e-3   # get_param_priors.jl:14


def synthetic_prior_ensemble(S: int, seed: int = 123) -> np.ndarray:
    """S x 24 synthetic parameter sets drawn from the reference's prior distributions.

    Independent log-normals exp(N(mu, sigma)); binding pairs drawn as (Kd, k_r) with k_f = k_r/Kd as
    generate_ensemble does (get_param_posteriors.jl:75-76).  RNG: NumPy PCG64(seed) (the reference seeds 123).
    """
    g = np.random.Generator(np.random.PCG64(seed))
    def ln(mu, sd):
        return np.exp(g.normal(mu, sd, size=S))
    p = np.zeros((S, 24))
    for i, n in enumerate(PNAMES[:7]):
        p[:, i] = ln(*_UV[n])
    Kd_S2, kS2r = ln(4.09800, 1.09860), ln(6.17379, 0.09531)
    Kd_G2, kG2r = ln(4.09800, 1.09860), ln(6.17379, 0.09531)
    kG1f, kG1r = ln(-7.20617, 2.88008), ln(-2.09794, 1.14776)
    kEGFf, kEGFr = ln(4.02250, 0.49119), ln(-1.96105, 0.50689)
    kdf = ln(0.18232, 0.09531)
    kdr = kdf * 0.38
    cols = {"kS2f": kS2r / Kd_S2, "kS2r": kS2r, "kG1f": kG1f, "kG1r": kG1r, "kG2f": kG2r / Kd_G2, "kG2r": kG2r,
            "kG1p": ln(*_UV["kG1p"]), "kG1dp": ln(*_UV["kG1dp"]), "kSa": ln(*_UV["kSa"]), "kSi": ln(*_UV["kSi"]),
            "kp": ln(*_UV["kp"]), "kdp": ln(*_UV["kdp"]), "kEGFf": kEGFf, "kEGFr": kEGFr,
            "EGF": np.full(S, _EGF), "kdf": kdf, "kdr": kdr}
    for i, n in enumerate(PNAMES[7:]):
        p[:, 7 + i] = cols[n]
    return p


# the 22 log-normal draws of one synthetic prior set, in the order gab1_sample_prior takes them (include/gab1pde.h):
# D(7); Kd_S2, kS2r, Kd_G2, kG2r, kG1f, kG1r, kEGFf, kEGFr, kdf; kG1p, kG1dp, kSa, kSi, kp, kdp
PRIOR_MU_SIGMA = ([_UV[n] for n in PNAMES[:7]] +
                  [(4.09800, 1.09860), (6.17379, 0.09531), (4.09800, 1.09860), (6.17379, 0.09531), (-7.20617, 2.88008),
                   (-2.09794, 1.14776), (4.02250, 0.49119), (-1.96105, 0.50689), (0.18232, 0.09531)] +
                  [_UV[n] for n in ("kG1p", "kG1dp", "kSa", "kSi", "kp", "kdp")])
PRIOR_KDD = 0.38     # kdr = kdf * Kdd


def synthetic_prior_ensemble_device(S: int, seed: int = 123) -> np.ndarray:
    """The distributions of synthetic_prior_ensemble drawn ON THE DEVICE (gab1_sample_prior: Philox4x32-10 + Box-Muller,
    the library's own reproducible stream).  Returns S x 24 like its host twin (different stream)."""
    from . import abi
    mu, sigma = (np.array(x) for x in zip(*PRIOR_MU_SIGMA))
    D, k = abi.sample_prior(S, seed, mu, sigma, _EGF, PRIOR_KDD)
    return np.concatenate([D, k], axis=1)


def resampled_ensemble(S: int, seed: int = 123) -> np.ndarray:
    """S x 24 synthetic sets with the marginals of the shipped ensemble: column-wise log-normal refit
    of parameter_ensemble.csv ("fully data-driven alternative")."""
    base = load_parameter_ensemble()
    g = np.random.Generator(np.random.PCG64(seed))
    lg = np.log(base)
    mu, sd = lg.mean(axis=0), lg.std(axis=0)
    return np.exp(mu + sd * g.standard_normal((S, 24)))
