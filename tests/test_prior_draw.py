"""GP.sample(t).max() on the device (fab_gp_prior_draw_batch; quirk Q5 of the reference: the incumbent of EI is the
maximum of ONE prior draw at the representer points, fabolas.py:198-199, entropy_search.py:137).

The draw is a pure function of the seed: the test replays the Philox / Box-Muller normals on the host, factorises the
ORACLE's prior covariance k(t, t) + 1.25e-12 I (george_oracle, the same matrix GP.sample uses) with numpy and compares
sample = mean + L z entry by entry; then checks the law of the draws (parity with numpy's own generator can only be
distributional)."""
import numpy as np
import pytest

from oracle import george_oracle as go
from tests.mcmc_replay import philox4x32_10, uniform53

PURPOSE = 7


def _normals(seed, b, Z):
    z = np.empty(Z + 1)
    for pair in range((Z + 1) // 2):
        r = philox4x32_10(seed, b, pair, PURPOSE, 0)
        u1, u2 = uniform53(r[0], r[1]), uniform53(r[2], r[3])
        rad = np.sqrt(-2.0 * np.log(u1))
        z[2 * pair], z[2 * pair + 1] = rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)
    return z[:Z]


def _check(eng, family, Z, B=5, N=40, seed=1234):
    rng = np.random.default_rng(Z + family)
    D = 3
    X = rng.uniform(-1, 1, (N, D)); y = np.sin(X.sum(1))
    d = D - 1 if family == 1 else D
    Pk = d + 3 if family == 1 else d + 1
    theta = rng.uniform(-1, 1, (B, Pk))
    eng.set_data(family, X, y, float(y.mean()), D)
    eng.fit(theta, np.full(B, 1e-2))
    reps = rng.uniform(-1, 1, (B, Z, D))
    ymax, sample, info = eng.prior_draw_max(reps, seed=seed, want_sample=True)
    ymax, sample, info = ymax.cpu().numpy(), sample.cpu().numpy(), info.cpu().numpy()
    assert np.all(info == 0)
    white = []
    for b in range(B):
        K = go.build_kernel(family, theta[b], D).get_value(reps[b])
        K[np.diag_indices_from(K)] += go.TINY
        L = np.linalg.cholesky(K)
        ref = y.mean() + L @ _normals(seed, b, Z)
        scale = np.sqrt(np.diag(K)).max()
        # the factor of a covariance with condition number up to ~1e10 amplifies rounding differences between the two
        # Cholesky orders; compare at 1e-7 of the prior standard deviation and, exactly, through the whitened draw
        assert np.max(np.abs(sample[b] - ref)) < 1e-6 * scale, (b, np.max(np.abs(sample[b] - ref)))
        assert ymax[b] == sample[b].max()
        white.append(np.linalg.solve(L, sample[b] - y.mean()))
    return np.concatenate(white)


def test_prior_draw_emulated(sim_engine):
    for family, Z in ((0, 1), (1, 7), (1, 50), (0, 64)):
        _check(sim_engine, family, Z, B=3)
    # a different seed gives a different draw, the same seed the same one
    a = _check(sim_engine, 1, 10, B=2, seed=5); b = _check(sim_engine, 1, 10, B=2, seed=5); c = _check(sim_engine, 1, 10, B=2, seed=6)
    assert np.array_equal(a, b) and not np.allclose(a, c)


@pytest.mark.gpu
def test_prior_draw_gpu(gpu_engine):
    for family, Z in ((0, 1), (1, 7), (1, 50), (0, 64)):
        _check(gpu_engine, family, Z)
    # law of the draws: the whitened values L^-1 (sample - mean) of many models are i.i.d. standard normal
    w = np.concatenate([_check(gpu_engine, 1, 50, B=128, seed=s) for s in range(6)])
    w = w[np.abs(w) < 50]                                  # (ill-conditioned tails of the whitening do not count)
    assert abs(w.mean()) < 0.03 and abs(w.var() - 1.0) < 0.05, (w.mean(), w.var())
    # duplicated points: positive semi-definite covariance -> pivots dropped, finite draw (numpy's SVD-based generator
    # accepts those as well)
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (30, 3)); y = np.sin(X.sum(1))
    gpu_engine.set_data(0, X, y, float(y.mean()), 3)
    gpu_engine.fit(np.array([[5.0, 0.0, 0.0, 0.0]]), np.array([1e-2]))
    reps = rng.uniform(-1, 1, (1, 12, 3)); reps[0, 5:] = reps[0, 4]
    ymax, sample, info = gpu_engine.prior_draw_max(reps, seed=3, want_sample=True)
    assert np.all(np.isfinite(sample.cpu().numpy())) and float(ymax[0]) == float(sample.max())
