"""eFAST design and estimators — TEST INFRASTRUCTURE, NOT PRODUCT (only tests/ and tests/golden/make_*.py import this).

The reference's global sensitivity analyses call `gsa(fbatch_concs_mt, eFAST(), pbounds; samples=1000, batch=true)`
(GSA_concs.jl:62-81; GSA_diffs+kinetic-params_MoL.jl:66-85).  `gsa`/`eFAST` live in GlobalSensitivity.jl, pinned at
v2.1.3 by the reference's Manifest.toml and ABSENT from /root/reference, so its algorithm (extended Fourier amplitude
sensitivity test, Saltelli et al. 1999, as that package version implements it) is restated here:

  * frequencies: omega_1 = floor((samples-1)/(2M)), M = num_harmonics = 4; m = floor(omega_1/(2M)); the other
    parameters get floor.(range(1, m, length = n-1)) when m >= n-1, else (0:n-2) .% m .+ 1;
  * design, for each parameter i: parameter i oscillates with omega_1 and the others with the complementary set,
    s_k = 2*pi*k/samples, one random phase phi_i = 2*pi*rand() shared by every parameter of block i,
    x = 0.5 + asin(sin(omega*s + phi))/pi, value = lb + x*(ub - lb)  (quantile of Uniform(lb, ub));
  * estimators from the block's outputs y: F = fft(y)[2 : samples/2], P_k = (|F_k|/samples)^2, V = 2*sum(P),
    S1 = 2*sum(P[omega_1*(1:M)])/V,  ST = 1 - 2*sum(P[1 : omega_1/2])/V;  a constant output gives 0/0 = NaN, which
    the reference's script replaces by 0 (GSA_concs.jl:82-83).

The random phases come from Julia's default RNG (`Random.seed!(123)`, GSA_concs.jl:26), which cannot be reproduced
outside Julia: the stored S1/ST (GSA results/*.csv) are therefore pinned within the spread over phases, not bit for bit.
"""
from __future__ import annotations

import numpy as np

NUM_HARMONICS = 4


def frequencies(num_params: int, samples: int, num_harmonics: int = NUM_HARMONICS):
    omega1 = float(np.floor((samples - 1) / (2 * num_harmonics)))
    m = float(np.floor(omega1 / (2 * num_harmonics)))
    if m >= num_params - 1:
        rest = np.floor(np.linspace(1.0, m, num_params - 1))
    else:
        rest = np.arange(num_params - 1) % m + 1
    return omega1, rest


def design(bounds, samples: int, phases, num_harmonics: int = NUM_HARMONICS) -> np.ndarray:
    """bounds: (n, 2) lower/upper per parameter; phases: n values in [0, 2*pi).  Returns ps (n, samples*n): block i
    (columns i*samples:(i+1)*samples) is the search curve on which parameter i carries the high frequency."""
    bounds = np.asarray(bounds, dtype=np.float64)
    n = bounds.shape[0]
    omega1, rest = frequencies(n, samples, num_harmonics)
    s = (2 * np.pi / samples) * np.arange(samples)
    ps = np.zeros((n, samples * n))
    for i in range(n):
        om = np.zeros(n)
        om[i] = omega1
        om[[k for k in range(n) if k != i]] = rest
        for j in range(n):
            x = 0.5 + (1 / np.pi) * np.arcsin(np.sin(om[j] * s + phases[i]))
            ps[j, i * samples:(i + 1) * samples] = bounds[j, 0] + x * (bounds[j, 1] - bounds[j, 0])
    return ps


def indices(all_y, num_params: int, samples: int, num_harmonics: int = NUM_HARMONICS, nan_to_zero: bool = True):
    """all_y: (n_out, samples*num_params) model outputs on `design`.  Returns S1, ST of shape (n_out, num_params) — the
    layout of the reference's efast.S1 / efast.ST, whose transposes are the rows of the stored CSV files."""
    all_y = np.atleast_2d(np.asarray(all_y, dtype=np.float64))
    omega1, _ = frequencies(num_params, samples, num_harmonics)
    w1 = int(omega1)
    S1 = np.zeros((all_y.shape[0], num_params))
    ST = np.zeros_like(S1)
    for i in range(num_params):
        y = all_y[:, i * samples:(i + 1) * samples]
        ft = np.fft.fft(y, axis=1)[:, 1:samples // 2]           # Julia's [2:Int(floor(samples/2))]
        P = (np.abs(ft) / samples) ** 2                          # P[:, k-1] is frequency k
        V = 2 * P.sum(axis=1)
        # an exactly constant block has an exactly zero spectrum (FFTW returns exact zeros for it; pocketfft leaves
        # ~1e-13 of round-off): variance 0, so both indices are 0/0 = NaN as in the reference
        V = np.where((y == y[:, :1]).all(axis=1), 0.0, V)
        P = np.where(V[:, None] == 0.0, 0.0, P)
        with np.errstate(invalid="ignore", divide="ignore"):
            S1[:, i] = 2 * P[:, [h * w1 - 1 for h in range(1, num_harmonics + 1)]].sum(axis=1) / V
            ST[:, i] = 1 - 2 * P[:, :w1 // 2].sum(axis=1) / V
    if nan_to_zero:
        S1[np.isnan(S1)] = 0.0
        ST[np.isnan(ST)] = 0.0
    return S1, ST


def read_reference_csv(path):
    """A stored `*_S1.csv` / `*_ST.csv` of the reference (rows = parameters, six output columns, then param, type) as an
    array (n_out, n_params) in efast.S1 orientation."""
    import csv
    with open(path, newline="") as fh:
        rows = list(csv.reader(fh))
    return np.array([[float(v) for v in row[:6]] for row in rows[1:]]).T
