"""Expected-Improvement baseline on the CUDA engine (reference: expected_improvement.py:70-93).

The reference fits sklearn's GaussianProcessRegressor(1 * Matern(nu=2.5)) by L-BFGS on the marginal
likelihood WITH gradient, then predicts a candidate sweep, computes EI and takes the argmax.  Here the
likelihood + gradient come from fab_gp_loglik_batch, the sweep from fab_gp_predict_batch and EI +
argmax from fab_ei_batch; scipy's L-BFGS-B stays the optimiser, as inside sklearn."""
import numpy as np
from scipy.optimize import minimize

from . import runtime
from ._capi import FAB_FAMILY_M52

ALPHA = 1e-10          # GaussianProcessRegressor default jitter
TINY = 1.25e-12


def fit_gpr(engine, X, y, theta0=(0.0, 0.0), bounds=((np.log(1e-5), np.log(1e5)), (np.log(1e-10), np.log(1e10)))):
    """max over (log constant, log length_scale) of the LML of C * Matern52(isotropic), mean 0."""
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    D = X.shape[1]
    engine.set_data(FAB_FAMILY_M52, X, y, 0.0, float(D))

    def to_engine(t):
        return np.concatenate([[t[0] - np.log(D)], np.full(D, 2.0 * t[1])])       # amp = C ; M_k = l^2

    def nll(t):
        ll, g, info = engine.loglik(to_engine(t)[None], np.array([ALPHA - TINY]), grad=True)
        if int(info[0]) != 0:
            return 1e25, np.zeros(2)
        g = g[0].cpu().numpy()
        return -float(ll[0]), -np.array([g[0], 2.0 * g[1:1 + D].sum()])

    res = minimize(nll, np.asarray(theta0, dtype=np.float64), jac=True, method="L-BFGS-B", bounds=bounds)
    theta = to_engine(res.x)
    engine.fit(theta[None], np.array([ALPHA - TINY]))
    return res.x, -res.fun


def get_candidate(prior, bounds, n_gen_samples=100, engine=None):
    engine = engine or runtime.new_engine()
    X = np.asarray(prior["X"], dtype=np.float64)
    y = np.asarray(prior["y"], dtype=np.float64).reshape(-1)
    fit_gpr(engine, X, y)
    X_samples = np.random.uniform(low=[b[0] for b in bounds], high=[b[1] for b in bounds],
                                  size=(n_gen_samples, X.shape[1]))
    mean, var = engine.predict(X_samples, want_var=True)
    _, _, index = engine.expected_improvement(mean, var, np.array([y.max()]), want_ei=False)
    if int(index[0]) < 0:
        raise ArithmeticError("expected improvement has no maximiser over the candidate set")
    return X_samples[int(index[0])]
