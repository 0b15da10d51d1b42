"""Entropy Search candidate search on the CUDA engine (reference: entropy_search.py), Matern-only
kernel, no dataset-size dimension.  Quirk numbers refer to SURVEY.md section 3.4."""
import logging

import numpy as np
import torch
from scipy.optimize import minimize

from . import epmgp, runtime
from .acquisitions import information_gain, information_gain_batch, prepare_entropy_search
from ._capi import FAB_FAMILY_M52, FAB_PRIOR_ES
from .fabolas import lognorm_logpdf
from .horseshoe import Horseshoe
from .models import ModelBatch
from .sampler import EnsembleSampler


def log_prob_batch(theta, engine):
    """entropy_search.py:40-77 for a (B, P) block of walkers: box 0 < cov < 100, -9 < lamb < 2,
    lognormal + horseshoe priors, set_parameter_vector([cov, *lamb]) (Q3), yerr = sqrt(noise)."""
    theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
    cov, lamb, noise = theta[:, 0], theta[:, 1:-1], theta[:, -1]
    ok = (cov > 0) & (cov < 100) & np.all(lamb > -9, axis=1) & np.all(lamb < 2, axis=1) & (noise > -20) & (noise < 20)
    ok &= noise >= 0                                  # sqrt(noise) of a negative number: NaN -> exception -> -inf
    out = np.full(theta.shape[0], -np.inf)
    if not np.any(ok):
        return out
    p = theta[ok]
    prior = lognorm_logpdf(p[:, 0], 1.0) + Horseshoe(scale=0.1).logpdf(p[:, -1])
    ll, info = engine.loglik(p[:, :-1], p[:, -1])
    ll = ll.cpu().numpy()
    ll[info.cpu().numpy() != 0] = -np.inf
    out[ok] = prior + ll
    return out


def sample_hypers(X, y, K=20, burn_in=100, iterations=500, engine=None, on_device=True):
    """entropy_search.py:20-85.  on_device (default): the whole chain is one persistent kernel."""
    engine = engine or runtime.new_engine()
    if on_device:
        from .fabolas import _collect_chain, _enqueue_chain
        return _collect_chain(*_enqueue_chain(engine, X, y, K, burn_in + iterations, FAB_FAMILY_M52, FAB_PRIOR_ES))
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    engine.set_data(FAB_FAMILY_M52, X, y, float(np.mean(y)), float(X.shape[1]))
    ndim = X.shape[1] + 2
    sampler = EnsembleSampler(K, ndim, log_prob_batch, args=[engine], vectorize=True)
    state = sampler.run_mcmc(np.random.rand(K, ndim), burn_in, rstate0=np.random.get_state())
    sampler.reset()
    sampler.run_mcmc(state, iterations)
    return sampler.chain[:, -1]


def entropy_search(dataset, bounds, n_hyper_samples=20, n_gen_samples=50, n_innovations=20, burn_in=100,
                   iterations=500, maxiter=100):
    """entropy_search.py:88-169."""
    X = np.asarray(dataset["X"], dtype=np.float64)
    y = np.asarray(dataset["y"], dtype=np.float64).reshape(-1)
    D = X.shape[1]
    logging.info("Sampling hypeparameters..")
    h = sample_hypers(X, y, K=n_hyper_samples, burn_in=burn_in, iterations=iterations)
    with np.errstate(all="ignore"):
        theta = np.column_stack([np.log(h[:, 0] / D), h[:, 1:-1]])     # kernel_cov * Matern52(metric=e**lamb)
    models = ModelBatch(FAB_FAMILY_M52, X, y, theta, h[:, -1] ** 2, amp_mult=D)   # Q3: compute(X, yerr=noise)
    eng = models.engine
    K = len(models)
    reps = np.random.uniform(low=[b[0] for b in bounds], high=[b[1] for b in bounds], size=(K, n_gen_samples, D))
    mean, _, cov = eng.predict(reps, want_var=True, want_cov=True)
    y_max, _ = eng.prior_draw_max(reps, seed=int(np.random.randint(0, 2 ** 31 - 1)))     # Q5: regressor.sample(reps).max()
    U, _, _ = eng.expected_improvement(mean, torch.diagonal(cov, dim1=1, dim2=2).contiguous(), y_max)
    p_min = epmgp.joint_min_batch(mean, cov, with_derivatives=True, engine=eng)
    Omega = np.random.normal(size=(K, n_innovations, n_gen_samples))
    prepare_entropy_search(models, reps, cov, p_min, U, Omega)
    logging.info("Ready to optimize Information Gain")
    return minimize(fun=lambda x: -information_gain(x, models), x0=X[np.argmax(y)], method="L-BFGS-B",
                    bounds=bounds, options={"maxiter": maxiter})
