"""The oracle against the reference's golden vectors (tests/golden/*_ref.npz were produced by the
reference's own epmgp.py / acquisitions.py / horseshoe.py / ard.py) and against independent checks
(scikit-learn GPR, central finite differences) for the george restatement."""
import numpy as np
import pytest

from oracle import acq_oracle as ao
from oracle import epmgp_oracle as eo
from oracle import george_oracle as go
from oracle.ard_oracle import ard_kernel
from tests.conftest import golden, rel_err


# ---------------------------------------------------------------- EPMGP (pinned)
@pytest.mark.parametrize("name", ["z5", "z12", "z20_spread", "z50", "z8_tight", "gp15"])
def test_epmgp_oracle_matches_reference(name):
    g = golden("epmgp_ref.npz")
    logP, dMu, dSg, dMM = eo.joint_min(g[name + "_mu"], g[name + "_S"], with_derivatives=True)
    assert np.max(np.abs(np.exp(logP) - np.exp(g[name + "_logP"]))) < 1e-10      # P_min
    assert np.max(np.abs(logP - g[name + "_logP"])) < 1e-8
    assert rel_err(dMu, g[name + "_dMu"]) < 1e-8
    assert rel_err(dSg, g[name + "_dSg"]) < 1e-8
    assert rel_err(dMM, g[name + "_dMM"]) < 1e-8
    assert abs(np.exp(logP).sum() - 1.0) < 1e-12
    assert np.array_equal(eo.joint_min(g[name + "_mu"], g[name + "_S"]), logP)


# ---------------------------------------------------------------- acquisitions (pinned)
def _acq_setup():
    g = golden("acq_ref.npz")
    B = g["theta"].shape[0]
    D = g["Xtrain"].shape[1]
    models, cost_models = [], []
    for b in range(B):
        m = go.GP(go.build_kernel(1, g["theta"][b], D), mean=g["y"].mean()); m.compute(g["Xtrain"], np.sqrt(g["yerr2"][b]))
        c = go.GP(go.build_kernel(1, g["ctheta"][b], D), mean=g["c"].mean()); c.compute(g["Xraw"], np.sqrt(g["yerr2"][b]))
        models.append(m); cost_models.append(c)
    pmin = [(g["pmin_logP"][b], g["pmin_dMu"][b], g["pmin_dSg"][b], g["pmin_dMM"][b]) for b in range(B)]
    Omega = [[g["Omega"][b, p][:, None] for p in range(g["Omega"].shape[1])] for b in range(B)]
    return g, models, cost_models, pmin, Omega


def test_information_gain_oracle_matches_reference():
    g, models, cost_models, pmin, Omega = _acq_setup()
    ds = {"y": g["y"], "c": g["c"]}
    I = g["Omega"].shape[1]
    for t, ig, igc in zip(g["tests"], g["ig"], g["igc"]):
        v = ao.information_gain(t, models, pmin, list(g["reps"]), list(g["U"]), Omega, ds, n_innovations=I)
        assert abs(v - ig) < 1e-9 * max(1.0, abs(ig))
        vc = ao.information_gain_cost(t, cost_models, models, ds, pmin, list(g["reps"]), list(g["U"]), Omega, n_innovations=I)
        assert abs(vc - igc) < 1e-9 * max(1.0, abs(igc))


def test_predict_innovations_ei_oracle_match_reference():
    g, models, cost_models, pmin, Omega = _acq_setup()
    for t, pm, pc in zip(g["tests"], g["pred_model"], g["pred_cost"]):
        assert np.allclose(ao.predict_testpoint_george(models, g["y"], t), pm, rtol=1e-11, atol=1e-13)
        assert np.allclose(ao.predict_testpoint_george(cost_models, g["c"], t), pc, rtol=1e-11, atol=1e-13)
    dmu, dsg = ao.compute_innovations(g["tests"][0].reshape(1, -1), g["y"], models[0], g["reps"][0],
                                      np.array([[g["pred_model"][0, 1]]]))
    assert np.allclose(dmu, g["innov_dmu"], rtol=1e-11, atol=1e-14)
    assert np.allclose(dsg, g["innov_dsigma"], rtol=1e-11, atol=1e-14)
    assert np.allclose(ao.expected_improvement(g["ei_mu"], g["ei_cov"], g["y"]), g["ei"], rtol=1e-10, atol=1e-300)


def test_priors_oracle_match_reference():
    g = golden("priors_ref.npz")
    for x, a, b in zip(g["xs"], g["horseshoe01"], g["horseshoe1"]):
        for scale, ref in ((0.1, a), (1.0, b)):
            v = ao.horseshoe_logpdf(x, scale)
            assert (np.isinf(ref) and v == ref) or abs(v - ref) <= 1e-12 * max(1.0, abs(ref))
    for c, ref in zip(g["covs"], g["lognorm"]):
        assert abs(ao.lognorm_logpdf(c, 1.0) - ref) < 1e-12 * max(1.0, abs(ref))


def test_ard_oracle_matches_reference_on_square_inputs():
    g = golden("ard_ref.npz")
    for n in (2, 4, 6):
        X, Y, ls = g[f"n{n}_X"], g[f"n{n}_Y"], g[f"n{n}_ls"]
        for nu in (0.5, 1.5, 2.5, np.inf):
            tag = f"n{n}_nu{nu}"
            assert np.allclose(ard_kernel(X, Y, ls, nu), g[tag + "_Kxy"], rtol=1e-12, atol=1e-15)
            if nu != 0.5:
                K, G = ard_kernel(X, None, ls, nu, eval_gradient=True)
                assert np.allclose(K, g[tag + "_K"], rtol=1e-12, atol=1e-15)
                assert np.allclose(G, g[tag + "_G"], rtol=1e-12, atol=1e-15)


# ---------------------------------------------------------------- george restatement (unpinned: independent checks)
def _toy(family, N=35, D=4, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (N, D))
    if family == 1:
        X[:, -1] = rng.uniform(0, 1, N)
    y = np.sin(X.sum(1)) + 0.05 * rng.standard_normal(N)
    d = D - 1 if family == 1 else D
    theta = np.concatenate([[0.3], rng.uniform(-0.5, 1.5, d), [-0.2, 0.4][: 2 * family]])
    return X, y, theta


def test_george_oracle_vs_sklearn():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel as C, Matern, WhiteKernel
    X, y, theta = _toy(0)
    D = X.shape[1]
    ll, g = go.loglik_batch(0, X, y, 0.0, theta[None], np.array([0.01]), grad=True)
    k = C(D * np.exp(theta[0])) * Matern(length_scale=np.sqrt(np.exp(theta[1:])), nu=2.5) + WhiteKernel(0.01 + 1.25e-12)
    gpr = GaussianProcessRegressor(kernel=k, alpha=0, optimizer=None).fit(X, y)
    l, gr = gpr.log_marginal_likelihood(gpr.kernel_.theta, eval_gradient=True)
    assert abs(ll[0] - l) < 1e-11 * abs(l)
    assert np.allclose(g[0, 0], gr[0], rtol=1e-10)
    assert np.allclose(2 * g[0, 1:1 + D], gr[1:1 + D], rtol=1e-9)                 # d/dlog l = 2 d/dlog M
    assert np.allclose(g[0, -1] * (0.01 + 1.25e-12), gr[-1], rtol=1e-9)           # d/dlog noise
    gp = go.GP(go.build_kernel(0, theta, D), mean=0.0); gp.compute(X, 0.1)
    T = np.random.default_rng(1).uniform(-1, 1, (7, D))
    mu, cov = gp.predict(y, T, return_cov=True)
    k2 = C(D * np.exp(theta[0])) * Matern(length_scale=np.sqrt(np.exp(theta[1:])), nu=2.5)
    gpr2 = GaussianProcessRegressor(kernel=k2, alpha=0.01 + 1.25e-12, optimizer=None).fit(X, y)
    m2, c2 = gpr2.predict(T, return_cov=True)
    assert np.allclose(mu, m2, rtol=1e-9, atol=1e-11) and np.allclose(cov, c2, rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("family", [0, 1])
def test_george_oracle_gradient_vs_finite_differences(family):
    X, y, theta = _toy(family, seed=3)
    f = lambda t, n2=0.02: go.loglik_batch(family, X, y, y.mean(), t[None], np.array([n2]))[0]
    _, g = go.loglik_batch(family, X, y, y.mean(), theta[None], np.array([0.02]), grad=True)
    for k in range(theta.size):
        e = np.zeros_like(theta); e[k] = 1e-5
        fd = (f(theta + e) - f(theta - e)) / 2e-5
        assert abs(fd - g[0, k]) < 2e-7 * max(1.0, abs(fd))
    fd = (f(theta, 0.02 + 1e-6) - f(theta, 0.02 - 1e-6)) / 2e-6
    assert abs(fd - g[0, -1]) < 1e-6 * max(1.0, abs(fd))


def test_george_oracle_semantics():
    # scalar * kernel has amplitude = scalar; writing v into the slot gives ndim * exp(v)   
    k = 2.5 * go.Matern52Kernel([0.5, 0.7], ndim=3, axes=[0, 1])
    assert np.isclose(k.get_value(np.zeros((1, 3)))[0, 0], 2.5)
    k.set_parameter_vector([0.3, 0.1, 0.2])
    assert np.isclose(k.get_value(np.zeros((1, 3)))[0, 0], 3 * np.exp(0.3))
    kk = 1 * go.Matern52Kernel([0.1, 0.1], ndim=3, axes=[0, 1]) * (go.LinearKernel(0.0, order=1, ndim=3, axes=[2])
                                                                    + go.ConstantKernel(0.0, ndim=3, axes=[2]))
    assert len(kk) == 5 and kk.ndim == 3
    x = np.array([[0.1, 0.2, 0.5], [0.3, -0.1, 0.25]])
    r2 = (0.2 ** 2 + 0.3 ** 2) / 0.1
    r = np.sqrt(5 * r2)
    assert np.isclose(kk.get_value(x)[0, 1], (1 + r + 5 * r2 / 3) * np.exp(-r) * (0.5 * 0.25 + 1.0))
    # golden (oracle-generated) fixture still reproduces
    g = golden("gp_oracle_svm_mnist.npz")
    ll, gr = go.loglik_batch(1, g["X"], g["y"], g["y"].mean(), g["theta"], g["yerr2"], grad=True)
    assert np.allclose(ll, g["ll"], rtol=1e-12) and np.allclose(gr, g["grad"], rtol=1e-9, atol=1e-9)
    # non positive definite -> -inf, never an exception
    bad = go.loglik_batch(0, np.zeros((4, 1)), np.zeros(4), 0.0, np.array([[0.0, 0.0]]), np.array([-1e-3]))
    assert bad[0] == -np.inf
