"""ModelBatch: the K per-hyper-sample GP models of one candidate search, resident on one engine.

The reference keeps a Python list of K george GPs (fabolas.py:140-176, 223-248; entropy_search.py:
100-120) and loops over it; here the K models share one observation matrix and live in one context,
so fitting, prediction and the acquisition are single batched kernel launches.  The object still
behaves like that list (len, indexing, iteration yield single-model views with george's
`predict(y, t, return_cov=True)`), so reference-style code keeps working."""
from __future__ import annotations

import numpy as np

from . import runtime


class ModelView:
    def __init__(self, batch, index):
        self.batch, self.index = batch, index

    def predict(self, y, t, return_cov=True, return_var=False, **kwargs):
        t = np.atleast_2d(np.asarray(t, dtype=np.float64))
        if return_cov and not return_var:
            mu, _, cov = self.batch.engine.predict(t, want_var=True, want_cov=True)
            return mu[self.index].cpu().numpy(), cov[self.index].cpu().numpy()
        mu, var = self.batch.engine.predict(t, want_var=True)
        mu, var = mu[self.index].cpu().numpy(), var[self.index].cpu().numpy()
        return (mu, var) if return_var else mu


class ModelBatch:
    def __init__(self, family, X, y, theta, yerr2, mean=None, amp_mult=None, engine=None):
        self.engine = engine or runtime.new_engine()
        self.X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        self.y = np.asarray(y, dtype=np.float64).reshape(-1)
        self.mean = float(np.mean(self.y) if mean is None else mean)
        self.family = family
        self.theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        self.yerr2 = np.asarray(yerr2, dtype=np.float64).reshape(-1)
        self.engine.set_data(family, self.X, self.y, self.mean, amp_mult)
        ll, info = self.engine.fit(self.theta, self.yerr2)
        self.log_likelihoods = ll.cpu().numpy()
        self.info = info.cpu().numpy()
        if np.any(self.info != 0):
            bad = np.nonzero(self.info)[0]
            raise np.linalg.LinAlgError(f"covariance of hyper-sample(s) {bad.tolist()} is not positive definite")

    def __len__(self):
        return self.theta.shape[0]

    def __getitem__(self, i):
        return ModelView(self, range(len(self))[i])

    def __iter__(self):
        return (ModelView(self, i) for i in range(len(self)))

    def predict(self, T, want_var=True, want_cov=False):
        return self.engine.predict(T, want_var=want_var, want_cov=want_cov)

    def mixture(self, T):
        """predict_testpoint_george for every row of T (acquisitions.py:99-116): (mean, variance)."""
        mu, var = self.engine.predict(T, want_var=True)
        mean = mu.mean(0)
        return mean, (var + mu * mu).mean(0) - mean * mean
