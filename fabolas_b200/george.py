"""george-shaped facade over the CUDA engine: exactly the objects the reference's GP path touches.

  kernels  Matern52Kernel(metric, ndim, axes), LinearKernel(log_gamma2, order, ndim, axes),
           ConstantKernel(log_constant, ndim, axes), scalar*kernel, kernel*kernel, kernel+kernel,
           len(kernel), kernel.ndim, set_parameter_vector        (fabolas.py:81-97, 153-169; entropy_search.py:32-36)
  GP       GP(kernel, mean), compute(x, yerr), log_likelihood(y, quiet), grad_log_likelihood(y),
           predict(y, t, return_cov, return_var), sample(t)      (fabolas.py:62-66, 171-199; acquisitions.py:31, 103)

The expression tree is only a description: `compile_kernel` maps the two shapes the reference builds
onto the engine's kernel families and every number is produced by the sm_100a kernels.
george semantics kept: non-radial kernels are summed over their axes, so
`b * kernel` stores log(b / ndim) and writing v into that slot yields the amplitude ndim * exp(v).
"""
from __future__ import annotations

import numpy as np

from . import runtime
from ._capi import FAB_FAMILY_M52, FAB_FAMILY_M52_BASIS

TINY = 1.25e-12


class Kernel:
    ndim = 1

    def _wrap(self, other):
        if isinstance(other, Kernel):
            return other
        return ConstantKernel(log_constant=np.log(float(other) / self.ndim), ndim=self.ndim)

    def __mul__(self, other):
        return Product(self, self._wrap(other))

    def __rmul__(self, other):
        return Product(self._wrap(other), self)

    def __add__(self, other):
        return Sum(self, self._wrap(other))

    def __radd__(self, other):
        return Sum(self._wrap(other), self)

    def __len__(self):
        return len(self.get_parameter_vector())

    def get_parameter_vector(self):
        raise NotImplementedError

    def set_parameter_vector(self, v):
        raise NotImplementedError

    def get_value(self, x1, x2=None):
        family, amp_mult, theta = compile_kernel(self)
        K = runtime.shared_engine().kernel_matrix(family, theta, np.atleast_2d(x1),
                                                  None if x2 is None else np.atleast_2d(x2), amp_mult)
        return K[0].cpu().numpy()

    def get_gradient(self, x1):
        family, amp_mult, theta = compile_kernel(self)
        _, dK = runtime.shared_engine().kernel_matrix(family, theta, np.atleast_2d(x1), None, amp_mult, grad=True)
        return dK[0].cpu().numpy()


def _axes(axes, ndim):
    return np.arange(ndim) if axes is None else np.atleast_1d(np.asarray(axes, dtype=int))


class ConstantKernel(Kernel):
    def __init__(self, log_constant, ndim=1, axes=None):
        self.ndim, self.axes, self.log_constant = ndim, _axes(axes, ndim), float(log_constant)

    def get_parameter_vector(self):
        return np.array([self.log_constant])

    def set_parameter_vector(self, v):
        self.log_constant = float(v[0])


class LinearKernel(Kernel):
    def __init__(self, log_gamma2, order=1, ndim=1, axes=None):
        if order != 1:
            raise NotImplementedError("only order=1 is on the reference's path (fabolas.py:86-91)")
        self.ndim, self.axes, self.log_gamma2, self.order = ndim, _axes(axes, ndim), float(log_gamma2), order

    def get_parameter_vector(self):
        return np.array([self.log_gamma2])

    def set_parameter_vector(self, v):
        self.log_gamma2 = float(v[0])


class Matern52Kernel(Kernel):
    def __init__(self, metric, ndim=1, axes=None):
        self.ndim, self.axes = ndim, _axes(axes, ndim)
        m = np.atleast_1d(np.asarray(metric, dtype=np.float64))
        if m.size != len(self.axes):
            raise NotImplementedError("axis-aligned metric with one entry per axis expected")
        with np.errstate(all="ignore"):
            self.log_M = np.log(m)

    def get_parameter_vector(self):
        return self.log_M.copy()

    def set_parameter_vector(self, v):
        self.log_M = np.asarray(v, dtype=np.float64).copy()


class _Binary(Kernel):
    def __init__(self, k1, k2):
        if k1.ndim != k2.ndim:
            raise ValueError("Dimension mismatch")
        self.k1, self.k2, self.ndim = k1, k2, k1.ndim

    def get_parameter_vector(self):
        return np.concatenate([self.k1.get_parameter_vector(), self.k2.get_parameter_vector()])

    def set_parameter_vector(self, v):
        v = np.asarray(v, dtype=np.float64)
        n1 = len(self.k1)
        if v.size != n1 + len(self.k2):
            raise ValueError("dimension mismatch")
        self.k1.set_parameter_vector(v[:n1])
        self.k2.set_parameter_vector(v[n1:])


class Sum(_Binary):
    pass


class Product(_Binary):
    pass


def compile_kernel(kernel):
    """(family, amp_mult, theta) of one of the two kernel shapes the reference builds; theta is the
    kernel's own parameter vector (george order, log space)."""
    def is_amp_m52(k, d):
        return (isinstance(k, Product) and isinstance(k.k1, ConstantKernel) and isinstance(k.k2, Matern52Kernel)
                and np.array_equal(k.k2.axes, np.arange(d)))

    nd = kernel.ndim
    if is_amp_m52(kernel, nd):
        return FAB_FAMILY_M52, float(len(kernel.k1.axes)), kernel.get_parameter_vector()
    if (isinstance(kernel, Product) and is_amp_m52(kernel.k1, nd - 1) and isinstance(kernel.k2, Sum)
            and isinstance(kernel.k2.k1, LinearKernel) and isinstance(kernel.k2.k2, ConstantKernel)
            and np.array_equal(kernel.k2.k1.axes, [nd - 1]) and np.array_equal(kernel.k2.k2.axes, [nd - 1])):
        return FAB_FAMILY_M52_BASIS, float(len(kernel.k1.k1.axes)), kernel.get_parameter_vector()
    raise NotImplementedError("kernel expression outside the FABOLAS / Entropy-Search path: expected "
                              "c*Matern52(all axes) or c*Matern52(axes 0..d-1)*(Linear(axis d)+Constant(axis d))")


class GP:
    """One Gaussian process on one engine context (single-model path; batches go through
    fabolas_b200.models.ModelBatch)."""

    def __init__(self, kernel, mean=0.0, engine=None):
        self.kernel = kernel
        self.mean = float(mean)
        self._engine = engine
        self._x = None
        self._y = None
        self._yerr2 = 0.0
        self._ll = None
        self.computed = False

    @property
    def engine(self):
        if self._engine is None:
            self._engine = runtime.new_engine()
        return self._engine

    def _fit(self, y):
        family, amp_mult, theta = compile_kernel(self.kernel)
        self.engine.set_data(family, self._x, y, self.mean, amp_mult)
        ll, info = self.engine.fit(theta[None], np.array([self._yerr2]))
        self._ll = float(ll[0])
        self._y = np.array(y, dtype=np.float64, copy=True)
        if int(info[0]) != 0:
            self.computed = False
            raise np.linalg.LinAlgError(f"{int(info[0])}-th leading minor of the array is not positive definite")
        self.computed = True

    def compute(self, x, yerr=0.0, **kwargs):
        self._x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        if self._x.shape[1] != self.kernel.ndim:
            raise ValueError("Dimension mismatch")
        self._yerr2 = float(np.asarray(yerr, dtype=np.float64)) ** 2
        y = self._y if (self._y is not None and self._y.shape[0] == self._x.shape[0]) else np.zeros(self._x.shape[0])
        self._fit(y)

    def _ensure(self, y):
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        if not self.computed:
            raise RuntimeError("You need to compute the model first")
        if self._y is None or not np.array_equal(y, self._y):
            self._fit(y)

    def log_likelihood(self, y, quiet=False):
        try:
            self._ensure(y)
        except np.linalg.LinAlgError:
            if quiet:
                return -np.inf
            raise
        return self._ll if np.isfinite(self._ll) else -np.inf

    lnlikelihood = log_likelihood

    def grad_log_likelihood(self, y, quiet=False):
        self._ensure(y)
        _, theta = compile_kernel(self.kernel)[1:]
        _, g, _ = self.engine.loglik(theta[None], np.array([self._yerr2]), grad=True)
        self.engine.fit(theta[None], np.array([self._yerr2]))          # loglik's gradient pass reuses the workspace
        return g[0, :len(theta)].cpu().numpy()

    def predict(self, y, t, return_cov=True, return_var=False, **kwargs):
        self._ensure(y)
        t = np.atleast_2d(np.asarray(t, dtype=np.float64))
        if return_var or not return_cov:
            mu, var = self.engine.predict(t, want_var=True)
            mu, var = mu[0].cpu().numpy(), var[0].cpu().numpy()
            return (mu, var) if return_var else mu
        mu, _, cov = self.engine.predict(t, want_var=True, want_cov=True)
        return mu[0].cpu().numpy(), cov[0].cpu().numpy()

    def sample(self, t, size=1):
        """Draw from the PRIOR N(mean, k(t,t) + 1.25e-12 I) (george's GP.sample; fabolas.py:199)."""
        t = np.atleast_2d(np.asarray(t, dtype=np.float64))
        cov = self.kernel.get_value(t)
        cov[np.diag_indices_from(cov)] += TINY
        out = np.random.multivariate_normal(np.full(t.shape[0], self.mean), cov, size)
        return out[0] if size == 1 else out
