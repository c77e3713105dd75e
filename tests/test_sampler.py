"""Device prior sampler: the NumPy restatement is pinned by Random123's published known-answer vectors
for Philox4x32-10 and by the moments of its draws (CPU tier); the CUDA kernel is held to the restatement (GPU tier)."""
import numpy as np
import pytest


def test_philox_known_answer_vectors():
    """Random123 kat_vectors, philox4x32 with 10 rounds."""
    from oracle import sampler_oracle as so
    u = lambda *a: np.array([a], dtype=np.uint32)
    f = 0xFFFFFFFF
    cases = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
             ((f, f, f, f), (f, f), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
             ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for ctr, key, want in cases:
        assert tuple(int(x) for x in so.philox4x32_10(u(*ctr), u(*key))[0]) == want


def test_draws_have_the_prior_moments(pkg):
    from oracle import sampler_oracle as so
    mu, sigma = (np.array(x) for x in zip(*pkg.params.PRIOR_MU_SIGMA))
    S = 200_000
    D, k = so.sample_prior(S, 123, mu, sigma, 1.67e-3, pkg.params.PRIOR_KDD)
    z = so.normals(S, 123)
    assert np.abs(z.mean(axis=0)).max() < 5 / np.sqrt(S) and np.abs(z.std(axis=0) - 1).max() < 5 / np.sqrt(2 * S)
    assert np.abs(np.corrcoef(z.T) - np.eye(22)).max() < 0.012
    # parameters: log-means and log-sds of the independent ones, the derived ones
    lD = np.log(D)
    assert np.abs(lD.mean(axis=0) - mu[:7]).max() < 5 * sigma[:7].max() / np.sqrt(S)
    np.testing.assert_allclose(np.log(k[:, 6]).std(), sigma[16], rtol=0.01)          # kG1p
    np.testing.assert_allclose(np.log(k[:, 0]).mean(), mu[8] - mu[7], atol=0.02)       # kS2f = kS2r / Kd
    assert np.all(k[:, 14] == 1.67e-3) and np.allclose(k[:, 16], k[:, 15] * 0.38, rtol=1e-15)
    # a different seed is a different stream; the same seed the same
    assert not np.array_equal(so.normals(8, 124), z[:8]) and np.array_equal(so.normals(8, 123), z[:8])


@pytest.mark.gpu
def test_device_sampler_matches_the_restatement(pkg):
    import __graft_entry__ as g
    g.build()
    from oracle import sampler_oracle as so
    mu, sigma = (np.array(x) for x in zip(*pkg.params.PRIOR_MU_SIGMA))
    for S, seed in ((1, 0), (1000, 123), (70001, 2 ** 40 + 5)):
        D, k = pkg.abi.sample_prior(S, seed, mu, sigma, 1.67e-3, pkg.params.PRIOR_KDD)
        Dr, kr = so.sample_prior(S, seed, mu, sigma, 1.67e-3, pkg.params.PRIOR_KDD)
        # same integers, same formulas; libm differences of log / sincos / exp are a few ulp, amplified by sigma <= 2.9
        np.testing.assert_allclose(D, Dr, rtol=1e-12)
        np.testing.assert_allclose(k, kr, rtol=1e-12)
    ens = pkg.params.synthetic_prior_ensemble_device(4096, seed=7)
    assert ens.shape == (4096, 24) and np.all(ens > 0)
    # the draws go straight into a solve
    res = pkg.host.sapdesolver_batch(pkg.params.base_Co(), ens[:256, :7], ens[:256, 7:], tf=0.2, out_mode=pkg.abi.OUT_SIX)
    assert res.out.shape == (256, 6)
