"""FABOLAS candidate search on the CUDA engine: same functions, arguments and quirks as the
reference's fabolas.py, with the MCMC likelihood, model fitting, EPMGP and information gain
evaluated in batches on the GPU.  Quirk numbers refer to SURVEY.md section 3.4."""
import logging

import numpy as np
import torch
from scipy.optimize import minimize

from . import epmgp, runtime
from .acquisitions import information_gain_cost, information_gain_cost_batch, prepare_entropy_search
from ._capi import FAB_FAMILY_M52_BASIS, FAB_PRIOR_FABOLAS
from .george import GP, ConstantKernel, LinearKernel, Matern52Kernel
from .horseshoe import Horseshoe
from .models import ModelBatch
from .sampler import EnsembleSampler


def lognorm_logpdf(x, s=1.0):
    """scipy.stats.lognorm.logpdf(x, s, 0) for x > 0 (fabolas.py:47)."""
    with np.errstate(all="ignore"):
        return -np.log(s * x * np.sqrt(2 * np.pi)) - np.log(x) ** 2 / (2 * s * s)


def log_likelihood(params, gp, X, y):
    """fabolas.py:16-66, one walker, george-shaped GP facade."""
    cov, lamb, noise = params[0], params[1:-1], params[-1]
    if not 0 < cov < 20:
        return -np.inf
    if not (np.all(np.array(lamb) > -9) and np.all(np.array(lamb) < 2)):
        return -np.inf
    if not -20 < noise < 20:
        return -np.inf
    prior = lognorm_logpdf(cov, 1.0) + Horseshoe(scale=0.1).logpdf(noise)
    gp.kernel.set_parameter_vector([cov, *(np.e ** lamb[:-2]), *lamb[-2:]])      # Q1
    try:
        if noise < 0:
            return -np.inf
        gp.compute(X, yerr=np.sqrt(noise))
    except Exception:
        return -np.inf
    return prior + gp.log_likelihood(y.reshape(-1), quiet=True)


def log_likelihood_batch(params, engine):
    """The same function for a (B, P) block of walkers in one GPU call (engine holds X, y)."""
    params = np.atleast_2d(np.asarray(params, dtype=np.float64))
    cov, lamb, noise = params[:, 0], params[:, 1:-1], params[:, -1]
    ok = (cov > 0) & (cov < 20) & np.all(lamb > -9, axis=1) & np.all(lamb < 2, axis=1) & (noise > -20) & (noise < 20)
    ok &= ~(noise < 0)
    out = np.full(params.shape[0], -np.inf)
    if not np.any(ok):
        return out
    p = params[ok]
    prior = lognorm_logpdf(p[:, 0], 1.0) + Horseshoe(scale=0.1).logpdf(p[:, -1])
    theta = np.column_stack([p[:, 0], np.e ** p[:, 1:-3], p[:, -3:-1]])          # Q1: raw cov, e**lamb, raw basis params
    ll, info = engine.loglik(theta, p[:, -1])                                    # yerr = sqrt(noise) -> yerr^2 = noise
    ll = ll.cpu().numpy()
    ll[info.cpu().numpy() != 0] = -np.inf
    out[ok] = prior + ll
    return out


def _enqueue_chain(engine, X, y, K, nsteps, family, prior):
    """Enqueue the whole stretch-move chain of one model (burn-in and main run are ONE launch: emcee's reset() between
    them only forgets the stored samples, fabolas.py:106-116) on the engine's stream; nothing is awaited.  numpy's
    global state picks the seed and the initial positions, in the order DeviceEnsembleSampler consumes it."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    # kernel = 1 * Matern52(axes 0..d-1) [* (Linear(axis d) + Constant(axis d))]; the scalar becomes a ConstantKernel
    # over all ndim axes, hence amp_mult = ndim
    engine.set_data(family, X, y, float(np.mean(y)), float(X.shape[1]))
    ndim = engine.Pk + 1                       # len(kernel) + 1 (the noise)
    seed = int(np.random.randint(0, 2 ** 31 - 1))
    p0 = np.random.rand(K, ndim)
    coords, lp, _, _, _ = engine.mcmc_run(prior, p0, nsteps, seed=seed, store_chain=False)
    return coords, lp


def _collect_chain(coords, lp):
    lp = lp.cpu().numpy()                      # synchronises the engine's stream
    if np.any(np.isnan(lp)):
        raise ValueError("Probability function returned NaN")
    return coords.cpu().numpy()


def sample_hypers(X, y, K=20, burn_in=100, iterations=500, engine=None, on_device=True):
    """fabolas.py:69-117: K walkers, 100 burn-in + 500 steps, last positions (K, P).  on_device (default): the
    whole chain is one persistent kernel; otherwise the emcee-shaped host loop calls the batched GPU likelihood
    once per half-step."""
    engine = engine or runtime.new_engine()
    if on_device:
        return _collect_chain(*_enqueue_chain(engine, X, y, K, burn_in + iterations, FAB_FAMILY_M52_BASIS, FAB_PRIOR_FABOLAS))
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    engine.set_data(FAB_FAMILY_M52_BASIS, X, y, float(np.mean(y)), float(X.shape[1]))
    ndim = X.shape[1] + 3                      # len(kernel) + 1
    sampler = EnsembleSampler(K, ndim, log_likelihood_batch, args=[engine], vectorize=True)
    state = sampler.run_mcmc(np.random.rand(K, ndim), burn_in, rstate0=np.random.get_state())
    sampler.reset()
    sampler.run_mcmc(state, iterations)
    return sampler.chain[:, -1]


def sample_hypers_pair(Xf, y, X, c, K=20, burn_in=100, iterations=500):
    """The function-model and the cost-model chains of get_candidate (fabolas.py:131 and :138) are independent: both
    persistent sampler kernels are enqueued on their own CUDA streams before either is awaited, so that a K = 20
    ensemble (10 proposals = 10 busy SMs per half-step) shares the GPU with the other instead of running after it."""
    if not torch.cuda.is_available():          # kernel emulator (tests): one after the other
        return (sample_hypers(Xf, y, K, burn_in, iterations), sample_hypers(X, c, K, burn_in, iterations))
    pending = []
    for XX, yy in ((Xf, y), (X, c)):
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            eng = runtime.new_engine()         # binds the engine to `stream`
            pending.append((eng, stream, _enqueue_chain(eng, XX, yy, K, burn_in + iterations, FAB_FAMILY_M52_BASIS,
                                                        FAB_PRIOR_FABOLAS)))
    out = []
    for eng, stream, (coords, lp) in pending:
        with torch.cuda.stream(stream):
            out.append(_collect_chain(coords, lp))
        eng.close()
    return tuple(out)


def _model_theta(hypers, ndim):
    """Q2: models are rebuilt as cov * Matern52(metric=e**lamb) * (Linear(lamb[-2]) + Constant(lamb[-1]))."""
    h = np.atleast_2d(hypers)
    with np.errstate(all="ignore"):
        return np.column_stack([np.log(h[:, 0] / ndim), h[:, 1:-3], h[:, -3:-1]]), h[:, -1]


def prepare_candidate_search(dataset, bounds, n_hyper_samples=20, n_gen_samples=50, n_innovations=20, burn_in=100,
                             iterations=500):
    """fabolas.py:120-254, everything before the optimiser: hyper-posterior samples of the function and the cost
    model, the resident model batches, representers, EI, p_min, innovations.  Returns (models, cost_models,
    starting_point); `models` carries the Entropy-Search state (prepare_entropy_search)."""
    if len(dataset["size"].shape) == 1:
        dataset["size"] = dataset["size"].reshape(-1, 1)
    X = np.concatenate([dataset["X"], dataset["size"]], axis=1)
    Xf = np.concatenate((X[:, :-1], ((1 - X[:, -1]) ** 2).reshape(-1, 1)), axis=1)    # Q4
    D = X.shape[1]

    logging.info("Sampling hyperparameters for function and cost models...")
    hypers, cost_hypers = sample_hypers_pair(Xf, dataset["y"], X, dataset["c"], K=n_hyper_samples, burn_in=burn_in,
                                             iterations=iterations)

    theta, noise = _model_theta(hypers, D)
    models = ModelBatch(FAB_FAMILY_M52_BASIS, Xf, dataset["y"], theta, noise, amp_mult=D)   # yerr = sqrt(noise)
    eng = models.engine
    K = len(models)
    reps = np.random.uniform(low=[b[0] for b in bounds], high=[b[1] for b in bounds], size=(K, n_gen_samples, D))
    mean, _, cov = eng.predict(reps, want_var=True, want_cov=True)
    eps = np.finfo(np.float64).eps
    cov_clip = torch.clamp(cov, min=eps)                                               # Q6
    # incumbent of EI = max of a PRIOR draw at the representers (Q5: regressor.sample(reps), fabolas.py:198-199), one
    # launch for all K models; numpy's global state (which the reference draws from) picks the seed of the device stream
    y_max, _ = eng.prior_draw_max(reps, seed=int(np.random.randint(0, 2 ** 31 - 1)))
    var_reps = torch.diagonal(cov_clip, dim1=1, dim2=2).contiguous()
    ei, _, _ = eng.expected_improvement(mean, var_reps, y_max)
    U = torch.clamp(ei, min=eps)
    logging.debug("Computing pMin")
    p_min = epmgp.joint_min_batch(mean, cov_clip, with_derivatives=True, engine=eng)
    Omega = np.random.normal(size=(K, n_innovations, n_gen_samples))
    prepare_entropy_search(models, reps, cov, p_min, U, Omega)

    ctheta, cnoise = _model_theta(cost_hypers, D)
    cost_models = ModelBatch(FAB_FAMILY_M52_BASIS, X, dataset["c"], ctheta, cnoise, amp_mult=D)
    imax = np.argmax(dataset["y"])
    starting_point = np.concatenate([dataset["X"][imax], dataset["size"][imax]])
    return models, cost_models, starting_point


def batched_fd_objective(cost_models, models, h=1e-8):
    """value and forward-difference gradient of -information_gain_cost from ONE batched
    call (x and the D points x + h e_k), the numbers L-BFGS-B otherwise collects with D + 1 separate objective calls
    (scipy's default `2-point` scheme with the same absolute step)."""
    def fun(x):
        pts = np.vstack([x] + [x + h * e for e in np.eye(len(x))])
        v = -information_gain_cost_batch(pts, cost_models, models)
        return v[0], (v[1:] - v[0]) / h
    return fun


def get_candidate(dataset, bounds, n_hyper_samples=20, n_gen_samples=50, n_innovations=20, batched_fd=False,
                  burn_in=100, iterations=500, maxiter=None):
    """fabolas.py:120-264."""
    models, cost_models, starting_point = prepare_candidate_search(
        dataset, bounds, n_hyper_samples, n_gen_samples, n_innovations, burn_in, iterations)
    logging.info("Ready to optimize Information Gain by Cost")
    opts = {} if maxiter is None else {"maxiter": maxiter}
    if batched_fd:
        return minimize(fun=batched_fd_objective(cost_models, models), x0=starting_point, jac=True, method="L-BFGS-B",
                        bounds=bounds, options=opts)
    return minimize(fun=lambda x: -information_gain_cost(x, cost_models, models), x0=starting_point,
                    method="L-BFGS-B", bounds=bounds, options=opts)


def make_gp(X, mean):
    """The george kernel/GP pair of fabolas.py:81-99 (single-model facade; used by log_likelihood)."""
    nd = X.shape[1]
    kernel = 1 * Matern52Kernel(metric=np.ones(nd - 1) * 0.1, ndim=nd, axes=list(range(nd - 1))) * (
        LinearKernel(log_gamma2=np.log(1), order=1, ndim=nd, axes=[nd - 1])
        + ConstantKernel(log_constant=np.log(1), ndim=nd, axes=[nd - 1]))
    return GP(kernel=kernel, mean=mean)
