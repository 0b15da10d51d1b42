"""GP prediction (mean / variance / full covariance), EI and argmax: CUDA path vs the oracle.
Tolerances (north star): predictive mean/variance within 1e-9 -- relative to the prior variance
k(x,x) for variances/covariances (k** - V.V cancels to ~0 next to observations, where a relative
error in the difference is meaningless for any algorithm)."""
import numpy as np
import pytest

from fabolas_b200.workloads import gp_workload
from oracle import acq_oracle as ao
from oracle import george_oracle as go
from tests.conftest import golden

TOL = 1e-9


def _setup(eng, N, D, family, B, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (N, D))
    if family == 1:
        X[:, -1] = rng.uniform(0, 1, N)
    y = np.sin(X.sum(1)) + 0.05 * rng.standard_normal(N)
    d = D - 1 if family == 1 else D
    Pk = d + 3 if family == 1 else d + 1
    theta = rng.uniform(-1, 1, (B, Pk))
    theta[:, 1:1 + d] = rng.uniform(-1, 2, (B, d))
    yerr2 = rng.uniform(1e-3, 0.1, B)
    eng.set_data(family, X, y, float(y.mean()), D)
    eng.fit(theta, yerr2)
    models = []
    for b in range(B):
        gp = go.GP(go.build_kernel(family, theta[b], D), mean=y.mean())
        gp.compute(X, np.sqrt(yerr2[b]))
        models.append(gp)
    return X, y, models, rng


def _check_predict(eng, N, D, family, B, M, seed, cov=False):
    X, y, models, rng = _setup(eng, N, D, family, B, seed)
    T = rng.uniform(-1, 1, (M, D))
    if family == 1:
        T[:, -1] = rng.uniform(0, 1, M)
    T[0] = X[0]                                   # a candidate on top of an observation (cancellation case)
    out = eng.predict(T, want_var=True, want_cov=cov)
    mu, var = out[0].cpu().numpy(), out[1].cpu().numpy()
    for b, gp in enumerate(models):
        rmu, rcov = gp.predict(y, T, return_cov=True)
        scale = np.sqrt(np.outer(np.diag(gp.kernel.get_value(T)), np.diag(gp.kernel.get_value(T))))
        assert np.max(np.abs(mu[b] - rmu) / np.maximum(1.0, np.abs(rmu))) < TOL
        assert np.max(np.abs(var[b] - np.diag(rcov)) / np.diag(scale)) < TOL
        if cov:
            c = out[2].cpu().numpy()[b]
            assert np.max(np.abs(c - rcov) / scale) < TOL
            assert np.max(np.abs(c - c.T) / scale) < 1e-12


@pytest.mark.parametrize("N,D,family,B,M,cov", [(20, 2, 0, 2, 5, True), (70, 3, 1, 2, 51, True), (130, 4, 1, 1, 70, False)])
def test_predict_emulated(sim_engine, N, D, family, B, M, cov):
    _check_predict(sim_engine, N, D, family, B, M, seed=N, cov=cov)


def test_ei_argmax_emulated(sim_engine):
    rng = np.random.default_rng(0)
    mu, var = rng.standard_normal((3, 300)), rng.uniform(-0.01, 1, (3, 300))
    var[1, 7] = 0.0
    ymax = np.array([0.3, -0.2, 1.0])
    ei, bv, bi = sim_engine.expected_improvement(mu, var, ymax)
    for b in range(3):
        ok = var[b] >= 0
        ref = ao.expected_improvement(mu[b][ok], np.diag(var[b][ok]), np.array([ymax[b]]))
        assert np.allclose(ei.numpy()[b][ok], ref, rtol=1e-9, atol=1e-14)   # u*Phi(u)+phi(u) cancels for u << 0
        assert np.all(np.isnan(ei.numpy()[b][~ok]))           # sqrt of a negative variance, like numpy
        # np.argmax semantics (expected_improvement.py:92-93): the first NaN is the maximum
        assert int(bi[b]) == int(np.argmax(ei.numpy()[b])) and np.isnan(float(bv[b]))
    ei2, bv2, bi2 = sim_engine.expected_improvement(mu, np.abs(var), ymax)
    for b in range(3):
        assert int(bi2[b]) == int(np.argmax(ei2.numpy()[b])) and float(bv2[b]) == ei2.numpy()[b].max()


def test_ei_argmax_chunked_emulated(sim_engine):
    """More candidates than one CTA scans: per-chunk partials + final reduction; the FIRST maximum wins across chunks."""
    rng = np.random.default_rng(1)
    M = 5000
    mu, var = rng.standard_normal((2, M)), rng.uniform(0.1, 1, (2, M))
    for c in (4100, 1300, 2999):                              # the same (largest) EI in three different chunks
        mu[0, c], var[0, c] = 9.0, 1.0
    mu[1, M - 1], var[1, M - 1] = 9.0, 1.0                    # maximum in the last, ragged chunk
    ei, bv, bi = sim_engine.expected_improvement(mu, var, np.array([0.0, 0.0]))
    assert int(bi[0]) == 1300 and int(bi[1]) == M - 1
    assert np.array_equal(bi.numpy(), np.argmax(ei.numpy(), axis=1)) and np.array_equal(bv.numpy(), ei.numpy().max(axis=1))
    _, bv2, bi2 = sim_engine.expected_improvement(mu, var, np.array([0.0, 0.0]), want_ei=False)
    assert np.array_equal(bi2.numpy(), bi.numpy()) and np.array_equal(bv2.numpy(), bv.numpy())


def test_predict_requires_fit_emulated(sim_engine):
    from fabolas_b200 import FabolasError
    sim_engine.set_data(0, np.zeros((4, 1)) + np.arange(4)[:, None], np.arange(4.0), 0.0, 1)
    with pytest.raises(FabolasError):
        sim_engine.predict(np.zeros((2, 1)))


@pytest.mark.parametrize("N,D,family,B,M,cov,jc", [(130, 4, 1, 2, 70, False, 1), (200, 3, 0, 1, 40, True, 2), (260, 3, 1, 1, 33, False, 3)])
def test_predict_chunked_panel_emulated(sim_engine, N, D, family, B, M, cov, jc):
    """The N > 512 path of the prediction kernel (cross-covariance panel generated `jc` tile rows at a time, partial
    rows of V parked in the per-CTA scratch, persistent grid) forced at small N."""
    try:
        sim_engine.debug_set_predict_chunk(jc)
        _check_predict(sim_engine, N, D, family, B, M, seed=N, cov=cov)
    finally:
        sim_engine.debug_set_predict_chunk(0)


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("N,M,cov,jc", [(600, 100, False, 0), (1000, 50, True, 0), (1000, 300, False, 0), (2048, 70, False, 0),
                                        (512, 257, False, 3), (300, 64, True, 2)])
def test_predict_large_history_gpu(gpu_engine, N, M, cov, jc):
    """Prediction beyond N = 512 (the history of a BO run grows every iteration; acquisitions.py:31 predicts at
    whatever size it has): chunked cross-covariance panel, against the oracle; jc > 0 forces the path at smaller N."""
    try:
        gpu_engine.debug_set_predict_chunk(jc)
        _check_predict(gpu_engine, N, 6, 1, 2, M, seed=N + M, cov=cov)
    finally:
        gpu_engine.debug_set_predict_chunk(0)


@pytest.mark.gpu
@pytest.mark.parametrize("N,D,family,B,M,cov", [(1, 1, 0, 1, 3, True), (20, 2, 0, 3, 1, True), (100, 3, 1, 4, 50, True),
                                                (100, 3, 1, 4, 51, True), (256, 6, 1, 4, 64, True),
                                                (256, 6, 1, 3, 1000, False), (300, 4, 0, 2, 333, False),
                                                (512, 6, 1, 2, 257, False)])
def test_predict_parity_gpu(gpu_engine, N, D, family, B, M, cov):
    _check_predict(gpu_engine, N, D, family, B, M, seed=N + M, cov=cov)


@pytest.mark.gpu
def test_golden_predict_svm_mnist_gpu(gpu_engine):
    g = golden("gp_oracle_svm_mnist.npz")
    gpu_engine.set_data(1, g["X"], g["y"], float(g["y"].mean()), 3)
    gpu_engine.fit(g["theta"], g["yerr2"])
    mu, var, cov = gpu_engine.predict(g["T"], want_cov=True)
    for b in range(g["theta"].shape[0]):
        scale = np.abs(np.diag(g["cov"][b])).max()
        assert np.max(np.abs(mu[b].cpu().numpy() - g["mu"][b])) < TOL * max(1.0, np.abs(g["mu"][b]).max())
        assert np.max(np.abs(cov[b].cpu().numpy() - g["cov"][b])) < TOL * max(scale, 1.0) * 10


@pytest.mark.gpu
def test_ei_sweep_argmax_gpu(gpu_engine):
    """expected_improvement.get_candidate shape (predict sweep + EI + argmax), 200k candidates."""
    import torch
    w = gp_workload(2, n_walkers=2)
    gpu_engine.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])
    gpu_engine.fit(w["theta"], w["yerr2"])
    rng = np.random.default_rng(1)
    M = 200_000
    T = np.column_stack([rng.uniform(4, 9, (M, 3)), rng.uniform(32, 512, M), rng.uniform(-6, -1, M), rng.uniform(0, 1, M)])
    mu, var = gpu_engine.predict(T)
    ymax = np.array([w["y"].max()] * 2)
    ei, bv, bi = gpu_engine.expected_improvement(mu, var, ymax)
    mu, var, ei = mu.cpu().numpy(), var.cpu().numpy(), ei.cpu().numpy()
    assert np.all(var > -1e-9)
    for b in range(2):
        sd = np.sqrt(np.maximum(var[b], 0))
        u = (mu[b] - ymax[b]) / sd
        ref = sd * (u * ao.norm_cdf(u) + ao.norm_pdf(u))
        ok = var[b] > 0
        assert np.allclose(ei[b][ok], ref[ok], rtol=1e-9, atol=1e-14)
        assert int(bi[b]) == int(np.argmax(ei[b]))            # np.argmax semantics incl. NaN (none expected here)
    # oracle spot check of the sweep on a sample of the candidates
    gp = go.GP(go.build_kernel(1, w["theta"][0], 6), mean=w["mean"]); gp.compute(w["X"], np.sqrt(w["yerr2"][0]))
    idx = rng.choice(M, 64, replace=False)
    rmu, rvar = gp.predict(w["y"], T[idx], return_cov=False, return_var=True)
    kss = np.array([gp.kernel.get_value(T[i:i + 1])[0, 0] for i in idx])
    assert np.max(np.abs(mu[0][idx] - rmu) / np.maximum(1, np.abs(rmu))) < TOL
    assert np.max(np.abs(var[0][idx] - rvar) / kss) < TOL
