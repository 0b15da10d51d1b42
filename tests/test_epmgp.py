"""EPMGP joint_min: CUDA path vs the reference's golden vectors and the C oracle.
Tolerance (north star): P_min within 1e-6 absolute; derivatives checked at 1e-7 relative."""
import numpy as np
import pytest

from oracle import epmgp_oracle as eo
from tests.conftest import golden, rel_err


def _check_case(eng, mu, S, ref, tol_p=1e-6, tol_d=1e-7):
    logP, dMu, dSg, dMM, info = eng.joint_min(mu, S)
    assert int(info.cpu().numpy().max()) == 0
    logP = logP.cpu().numpy()
    for b in range(logP.shape[0]):
        r = ref[b]
        assert np.max(np.abs(np.exp(logP[b]) - np.exp(r[0]))) < tol_p
        assert abs(np.exp(logP[b]).sum() - 1.0) < 1e-12
        assert rel_err(dMu.cpu().numpy()[b], r[1]) < tol_d
        assert rel_err(dSg.cpu().numpy()[b], r[2]) < tol_d
        assert rel_err(dMM.cpu().numpy()[b], r[3]) < tol_d
    lp2, info2 = eng.joint_min(mu, S, with_derivatives=False)
    assert np.array_equal(lp2.cpu().numpy(), logP)


@pytest.mark.parametrize("name", ["z5", "z12", "z8_tight", "gp15"])
def test_epmgp_golden_emulated(sim_engine, name):
    g = golden("epmgp_ref.npz")
    ref = [(g[name + "_logP"], g[name + "_dMu"], g[name + "_dSg"], g[name + "_dMM"])]
    _check_case(sim_engine, g[name + "_mu"][None], g[name + "_S"][None], ref)


def test_epmgp_degenerate_emulated(sim_engine):
    # Z = 1 (nothing to compare against: P = 1) and a far-apart pair that takes the z > 6 / z < -6 exits
    lp, info = sim_engine.joint_min(np.array([[0.3]]), np.array([[[2.0]]]), with_derivatives=False)
    assert float(lp[0, 0]) == 0.0 and int(info[0]) == 0
    mu = np.array([[0.0, 50.0, -40.0]]); S = np.eye(3)[None] * 0.5
    lp, dMu, dSg, dMM, info = sim_engine.joint_min(mu, S)
    ref = eo.joint_min(mu[0], S[0], with_derivatives=True)
    assert np.allclose(lp.numpy()[0], ref[0], atol=1e-12) and int(info[0]) == 0
    assert np.allclose(dMu.numpy()[0], ref[1], atol=1e-12) and np.allclose(dMM.numpy()[0], ref[3], atol=1e-12)


@pytest.mark.gpu
def test_epmgp_golden_gpu(gpu_engine):
    g = golden("epmgp_ref.npz")
    for name in [str(n) for n in g["names"]]:
        ref = [(g[name + "_logP"], g[name + "_dMu"], g[name + "_dSg"], g[name + "_dMM"])]
        _check_case(gpu_engine, g[name + "_mu"][None], g[name + "_S"][None], ref)


@pytest.mark.gpu
def test_epmgp_batch_vs_oracle_gpu(gpu_engine):
    """Config-3 shape: Z = 50 representers, a batch of hyper-samples, posterior-like covariances."""
    rng = np.random.default_rng(0)
    B, Z = 24, 50
    mus, Ss, refs = [], [], []
    for b in range(B):
        A = rng.standard_normal((Z, Z + 5))
        S = A @ A.T / Z * rng.uniform(0.2, 3.0) + 1e-3 * np.eye(Z)
        if b % 3 == 0:
            S = np.clip(S, np.finfo(float).eps, np.inf)               # fabolas.py:192
        mu = rng.standard_normal(Z) * rng.uniform(0.1, 3.0)
        mus.append(mu); Ss.append(S)
    mus, Ss = np.array(mus), np.array(Ss)
    logP, dMu, dSg, dMM, info, sweeps = gpu_engine.joint_min(mus, Ss, return_sweeps=True)
    assert int(info.cpu().numpy().max()) == 0
    for b in range(B):
        r, sw = eo.joint_min(mus[b], Ss[b], with_derivatives=True, return_sweeps=True)
        assert np.array_equal(sweeps.cpu().numpy()[b], sw)
        assert np.max(np.abs(np.exp(logP.cpu().numpy()[b]) - np.exp(r[0]))) < 1e-6
        assert rel_err(dMu.cpu().numpy()[b], r[1]) < 1e-7
        assert rel_err(dSg.cpu().numpy()[b], r[2]) < 1e-7
        assert rel_err(dMM.cpu().numpy()[b], r[3]) < 1e-7


@pytest.mark.gpu
def test_epmgp_nan_is_data_gpu(gpu_engine):
    mu = np.zeros((2, 4)); S = np.stack([np.eye(4), np.full((4, 4), np.nan)])
    logP, info = gpu_engine.joint_min(mu, S, with_derivatives=False)
    assert int(info[0]) == 0 and abs(np.exp(logP.cpu().numpy()[0]).sum() - 1) < 1e-12
    # NaN covariance: z is NaN -> -inf -> EP aborts -> every logZ = -inf -> -500 -> uniform (epmgp.py:164-175, 102)
    assert np.allclose(np.exp(logP.cpu().numpy()[1]), 0.25)
