"""The reference's acquisition callables (acquisitions.py) on the CUDA engine, same names and
argument meaning.  `models` is a fabolas_b200.models.ModelBatch prepared with `prepare_entropy_search`
(it replaces the reference's parallel Python lists models / p_min / representers / U / Omega)."""
import logging

import numpy as np

from . import runtime
from .models import ModelBatch


def expected_improvement(mean, covariance, y_values, exploration=0):
    """acquisitions.py:40-75: closed-form EI from the diagonal of `covariance`."""
    mean = np.asarray(mean, dtype=np.float64).reshape(-1)
    var = np.ascontiguousarray(np.diag(np.asarray(covariance, dtype=np.float64)))
    assert var.shape == (mean.shape[0],)
    ei, _, _ = runtime.shared_engine().expected_improvement(mean[None], var[None], np.array([np.max(y_values)]),
                                                            xi=float(exploration))
    return ei[0].cpu().numpy()


def predict_testpoint_george(models, y, test_point):
    """acquisitions.py:99-116: mixture mean / variance over the hyper-samples at one point."""
    mean, var = models.mixture(np.asarray(test_point, dtype=np.float64).reshape(1, -1))
    return float(mean[0]), float(var[0])


def prepare_entropy_search(models: ModelBatch, representers, cov_representers, p_min, U, Omega):
    """Make the Entropy-Search state resident next to the models (see fab_es_setup)."""
    models.engine.es_setup(representers, cov_representers, p_min, U, Omega)
    models.es_ready = True
    return models


def information_gain_batch(test_points, models):
    """Information gain of every row of `test_points` (M, D) -> numpy (M,).  Raises ArithmeticError
    where the reference does (acquisitions.py:160-164)."""
    ig, flag = models.engine.information_gain(np.atleast_2d(test_points))
    if int(flag.max()) != 0:
        logging.warning("Cannot compute Information Gain with this model")
        raise ArithmeticError("Cannot compute Information Gain with this model")
    return ig.cpu().numpy()


def information_gain(test_point, models, p_min=None, representers=None, U=None, Omega=None, dataset=None,
                     n_innovations=20, enable_log=True):
    """acquisitions.py:119-167.  The list arguments of the reference are accepted for signature
    compatibility; the state they describe is the one made resident by prepare_entropy_search."""
    if not getattr(models, "es_ready", False):
        if p_min is None:
            raise ValueError("models carry no Entropy-Search state: call prepare_entropy_search first")
        cov = models.engine.predict(np.asarray(representers), want_var=True, want_cov=True)[2]
        stack = lambda k: np.stack([np.asarray(p[k]) for p in p_min])
        prepare_entropy_search(models, np.asarray(representers), cov, tuple(stack(k) for k in range(4)),
                               np.asarray(U), np.asarray(Omega).reshape(len(models), n_innovations, -1))
    a = float(information_gain_batch(np.asarray(test_point, dtype=np.float64).reshape(1, -1), models)[0])
    if enable_log:
        logging.info(f"IG: {a} for test point: {test_point}")
    return a


def information_gain_cost(test_point, cost_models, models, dataset=None, p_min=None, representers=None, U=None,
                          Omega=None, n_innovations=20):
    """acquisitions.py:170-183."""
    overhead_cost = 0.0001
    predicted_cost, _ = predict_testpoint_george(cost_models, None, test_point)
    cost_factor = 1 / (predicted_cost + overhead_cost)
    ig = information_gain(test_point, models, p_min, representers, U, Omega, dataset, n_innovations, enable_log=False)
    ig_cost = cost_factor * ig
    logging.info(f"IG: {ig}, cost_f {cost_factor}, ig_cost {ig_cost}. x = {test_point}")
    return ig_cost


def information_gain_cost_batch(test_points, cost_models, models):
    """Batched information_gain_cost for (M, D) test points."""
    T = np.atleast_2d(test_points)
    cost, _ = cost_models.mixture(T)
    return information_gain_batch(T, models) / (cost.cpu().numpy() + 0.0001)
