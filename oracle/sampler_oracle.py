"""NumPy restatement of the on-device prior sampler — TEST INFRASTRUCTURE, NOT PRODUCT.

What it stands in for: the prior half of `generate_ensemble` (get_param_posteriors.jl:38-86): per parameter set, independent
log-normal draws for the 7 diffusivities and kG1p, kG1dp, kSa, kSi, kp, kdp (uvpars, :60-72) and (Kd, k_r)-style pairs for the five
binding reactions with k_f = k_r / Kd (:75-76); the distributions' (mu, sigma) come from get_param_priors.jl:19-198 through
calcModeSpread (values in params.py, reproduced in SURVEY §8d).  The reference draws with Julia's default RNG (unseeded in
generate_ensemble; `Random.seed!(123)` in the GSA scripts) — a stream that cannot be reproduced outside Julia — so the library
defines its own: Philox4x32-10 (Salmon et al., SC'11; counter = (set, draw pair, 0, 0), key = seed), two 53-bit uniforms per
call, Box–Muller.  PARITY UNPINNED against the reference's draws by construction; pinned here by Random123's published
known-answer vectors for Philox4x32-10 and by the moments of the draws.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
N_NORMALS = 22      # 7 D; Kd_S2, kS2r, Kd_G2, kG2r, kG1f, kG1r, kEGFf, kEGFr, kdf; kG1p, kG1dp, kSa, kSi, kp, kdp


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32."""
    c = [np.asarray(ctr[..., i], dtype=np.uint32).copy() for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32).copy()
    k1 = np.asarray(key[..., 1], dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = k0 + W0
            k1 = k1 + W1
    return np.stack(c, axis=-1)


def normals(S: int, seed: int) -> np.ndarray:
    """(S, 22) standard normals: call j of set s uses counter (s_lo, s_hi, j, 0) and yields normals 2j, 2j+1."""
    s = np.arange(S, dtype=np.uint64)
    out = np.zeros((S, N_NORMALS))
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    for j in range(N_NORMALS // 2):
        ctr = np.stack([(s & np.uint64(0xFFFFFFFF)).astype(np.uint32), (s >> np.uint64(32)).astype(np.uint32),
                        np.full(S, j, dtype=np.uint32), np.zeros(S, dtype=np.uint32)], axis=-1)
        r = philox4x32_10(ctr, np.broadcast_to(key, (S, 2))).astype(np.uint64)
        u1 = (((r[:, 0] << np.uint64(21)) | (r[:, 1] >> np.uint64(11))).astype(np.float64) + 0.5) * 2.0 ** -53
        u2 = (((r[:, 2] << np.uint64(21)) | (r[:, 3] >> np.uint64(11))).astype(np.float64) + 0.5) * 2.0 ** -53
        rad = np.sqrt(-2.0 * np.log(u1))
        th = 2.0 * np.pi * u2
        out[:, 2 * j] = rad * np.cos(th)
        out[:, 2 * j + 1] = rad * np.sin(th)
    return out


def sample_prior(S: int, seed: int, mu, sigma, EGF: float, Kdd: float):
    """-> D (S, 7), k (S, 17) in the reference's order (basepdesolver.jl:43-68)."""
    z = normals(S, seed)
    v = np.exp(np.asarray(mu)[None, :] + np.asarray(sigma)[None, :] * z)
    D = v[:, :7].copy()
    Kd_S2, kS2r, Kd_G2, kG2r, kG1f, kG1r, kEGFf, kEGFr, kdf, kG1p, kG1dp, kSa, kSi, kp, kdp = (v[:, 7 + i] for i in range(15))
    k = np.stack([kS2r / Kd_S2, kS2r, kG1f, kG1r, kG2r / Kd_G2, kG2r, kG1p, kG1dp, kSa, kSi, kp, kdp, kEGFf, kEGFr,
                  np.full(S, EGF), kdf, kdf * Kdd], axis=1)
    return D, k
