"""numpy/scipy restatement of the george 0.4.0 objects the reference's GP path touches.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED against george itself:
george 0.4.0 (``/root/reference/requirements.txt:6``) is a third-party dependency that is not in
``/root/reference`` and cannot be installed here, so this file restates its published algorithm and
is anchored on the reference's call sites:

  kernel construction   fabolas.py:81-97, 153-169, 227-243 ; entropy_search.py:32-36, 113-117
  set_parameter_vector  fabolas.py:57 ; entropy_search.py:71
  GP.compute            fabolas.py:62, 176, 246 ; entropy_search.py:38, 73, 119
  GP.log_likelihood     fabolas.py:66 ; entropy_search.py:77
  GP.predict            fabolas.py:187 ; entropy_search.py:130 ; acquisitions.py:31, 103, 109
  GP.sample             fabolas.py:199 ; entropy_search.py:137

Semantics restated:
  * non-radial kernels (Constant, Linear) are SUMMED over their axes; ``b * kernel`` becomes
    ``Product(ConstantKernel(log(b / ndim), ndim), kernel)`` so that the amplitude is ``b``;
  * ``Matern52Kernel(metric=vec)``: r2 = sum_k (x_k - x'_k)^2 / M_k, parameters log M_k,
    k = (1 + sqrt(5 r2) + 5 r2 / 3) exp(-sqrt(5 r2));
  * ``LinearKernel(order=1)``: x x' / gamma^2 with gamma^2 = exp(log_gamma2);
  * parameter vector = depth-first concatenation over the expression tree;
  * ``compute``: K += yerr^2 + 1.25e-12 on the diagonal, upper Cholesky, logdet = 2 sum log diag;
  * ``log_likelihood`` = -(N log 2pi + logdet)/2 - r' K^-1 r / 2, r = y - mean, non-finite -> -inf;
  * ``predict``: alpha = K^-1 r ; mu = K* alpha + mean ; cov = k(t,t) - K* K^-1 K*' (latent);
  * ``sample(t)``: draw from N(mean, k(t,t) + 1.25e-12 I)  (a PRIOR draw);
  * ``grad_log_likelihood`` = 0.5 * sum_ij (alpha alpha' - K^-1)_ij dK_ij/dtheta_k.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import cho_solve, cholesky

TINY = 1.25e-12


class Kernel:
    ndim: int

    # -- expression-tree algebra (george.kernels.Kernel.__mul__/__add__) ------------------------
    def __mul__(self, other):
        if not isinstance(other, Kernel):
            other = ConstantKernel(log_constant=np.log(float(other) / self.ndim), ndim=self.ndim)
        return Product(self, other)

    def __rmul__(self, other):
        if not isinstance(other, Kernel):
            other = ConstantKernel(log_constant=np.log(float(other) / self.ndim), ndim=self.ndim)
        return Product(other, self)

    def __add__(self, other):
        if not isinstance(other, Kernel):
            other = ConstantKernel(log_constant=np.log(float(other) / self.ndim), ndim=self.ndim)
        return Sum(self, other)

    def __radd__(self, other):
        if not isinstance(other, Kernel):
            other = ConstantKernel(log_constant=np.log(float(other) / self.ndim), ndim=self.ndim)
        return Sum(other, self)

    def __len__(self):
        return len(self.get_parameter_vector())

    # leaves override these three -----------------------------------------------------------------
    def get_parameter_vector(self):
        raise NotImplementedError

    def set_parameter_vector(self, v):
        raise NotImplementedError

    def _value_and_grads(self, A, B):
        """Return (K, [dK/dtheta_k ...]) for row sets A (n x ndim), B (m x ndim)."""
        raise NotImplementedError

    def get_value(self, x1, x2=None, diag=False):
        x1 = np.atleast_2d(np.asarray(x1, dtype=np.float64))
        if diag:
            return np.array([self._value_and_grads(r[None], r[None])[0][0, 0] for r in x1])
        x2 = x1 if x2 is None else np.atleast_2d(np.asarray(x2, dtype=np.float64))
        return self._value_and_grads(x1, x2)[0]

    def get_gradient(self, x1):
        x1 = np.atleast_2d(np.asarray(x1, dtype=np.float64))
        g = self._value_and_grads(x1, x1)[1]
        return np.stack(g, axis=-1) if g else np.zeros(x1.shape[:1] * 2 + (0,))


def _axes(axes, ndim):
    if axes is None:
        return np.arange(ndim)
    return np.atleast_1d(np.asarray(axes, dtype=int))


class ConstantKernel(Kernel):
    def __init__(self, log_constant, ndim=1, axes=None):
        self.ndim = ndim
        self.axes = _axes(axes, ndim)
        self.log_constant = float(log_constant)

    def get_parameter_vector(self):
        return np.array([self.log_constant])

    def set_parameter_vector(self, v):
        (self.log_constant,) = (float(v[0]),)

    def _value_and_grads(self, A, B):
        K = np.full((A.shape[0], B.shape[0]), len(self.axes) * np.exp(self.log_constant))
        return K, [K.copy()]


class LinearKernel(Kernel):
    def __init__(self, log_gamma2, order=1, ndim=1, axes=None):
        if order != 1:
            raise NotImplementedError("reference only uses order=1 (fabolas.py:86-91)")
        self.ndim = ndim
        self.axes = _axes(axes, ndim)
        self.log_gamma2 = float(log_gamma2)

    def get_parameter_vector(self):
        return np.array([self.log_gamma2])

    def set_parameter_vector(self, v):
        self.log_gamma2 = float(v[0])

    def _value_and_grads(self, A, B):
        K = np.zeros((A.shape[0], B.shape[0]))
        for a in self.axes:
            K += np.outer(A[:, a], B[:, a])
        K *= np.exp(-self.log_gamma2)
        return K, [-K]


class Matern52Kernel(Kernel):
    def __init__(self, metric, ndim=1, axes=None):
        self.ndim = ndim
        self.axes = _axes(axes, ndim)
        m = np.atleast_1d(np.asarray(metric, dtype=np.float64))
        if m.size == 1 and len(self.axes) > 1:
            raise NotImplementedError("isotropic metric not used by the reference")
        assert m.size == len(self.axes)
        with np.errstate(all="ignore"):
            self.log_M = np.log(m)

    def get_parameter_vector(self):
        return self.log_M.copy()

    def set_parameter_vector(self, v):
        self.log_M = np.asarray(v, dtype=np.float64).copy()

    def _value_and_grads(self, A, B):
        with np.errstate(all="ignore"):
            invM = np.exp(-self.log_M)
            d2 = [(A[:, a][:, None] - B[:, a][None, :]) ** 2 * invM[i] for i, a in enumerate(self.axes)]
            r2 = np.zeros((A.shape[0], B.shape[0]))
            for t in d2:
                r2 = r2 + t
            r = np.sqrt(5.0 * r2)
            e = np.exp(-r)
            K = (1.0 + r + 5.0 * r2 / 3.0) * e
            g = (5.0 / 6.0) * (1.0 + r) * e  # = -dk/dr2
        return K, [g * t for t in d2]


class _Binary(Kernel):
    def __init__(self, k1, k2):
        assert k1.ndim == k2.ndim
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
    def _value_and_grads(self, A, B):
        K1, g1 = self.k1._value_and_grads(A, B)
        K2, g2 = self.k2._value_and_grads(A, B)
        return K1 + K2, g1 + g2


class Product(_Binary):
    def _value_and_grads(self, A, B):
        K1, g1 = self.k1._value_and_grads(A, B)
        K2, g2 = self.k2._value_and_grads(A, B)
        return K1 * K2, [g * K2 for g in g1] + [K1 * g for g in g2]


class GP:
    """george.gp.GP with the default BasicSolver, constant mean, frozen white noise log(TINY)."""

    def __init__(self, kernel, mean=0.0):
        self.kernel = kernel
        self.mean = float(mean)
        self.computed = False

    def compute(self, x, yerr=0.0):
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        self._x = x
        yerr2 = np.broadcast_to(np.asarray(yerr, dtype=np.float64) ** 2, (x.shape[0],))
        K = self.kernel.get_value(x)
        K[np.diag_indices_from(K)] += yerr2 + TINY
        self._U = cholesky(K, lower=False, overwrite_a=True, check_finite=True)
        self.log_determinant = 2.0 * np.sum(np.log(np.diag(self._U)))
        self._const = -0.5 * (x.shape[0] * np.log(2.0 * np.pi) + self.log_determinant)
        self._alpha = None
        self._y = None
        self.computed = True

    def apply_inverse(self, b):
        return cho_solve((self._U, False), b)

    def _compute_alpha(self, y):
        y = np.asarray(y, dtype=np.float64)
        if self._alpha is None or self._y is None or not np.array_equal(y, self._y):
            self._y = y.copy()
            self._alpha = self.apply_inverse(y - self.mean)
        return self._alpha

    def log_likelihood(self, y, quiet=False):
        r = np.asarray(y, dtype=np.float64) - self.mean
        ll = self._const - 0.5 * np.dot(r, self.apply_inverse(r))
        return ll if np.isfinite(ll) else -np.inf

    def grad_log_likelihood(self, y, with_yerr2=False):
        alpha = self._compute_alpha(y)
        Kinv = self.apply_inverse(np.eye(self._x.shape[0]))
        W = np.outer(alpha, alpha) - Kinv
        dK = self.kernel._value_and_grads(self._x, self._x)[1]          # list of N x N arrays
        g = 0.5 * np.array([np.vdot(dk, W) for dk in dK])               # = 0.5 einsum('ijk,ij->k')
        if with_yerr2:
            g = np.append(g, 0.5 * np.trace(W))
        return g

    def predict(self, y, t, return_cov=True, return_var=False):
        alpha = self._compute_alpha(y)
        xs = np.atleast_2d(np.asarray(t, dtype=np.float64))
        Kxs = self.kernel.get_value(xs, self._x)
        mu = Kxs @ alpha + self.mean
        if return_var:
            var = self.kernel.get_value(xs, diag=True)
            var = var - np.sum(Kxs.T * self.apply_inverse(Kxs.T), axis=0)
            return mu, var
        if not return_cov:
            return mu
        cov = self.kernel.get_value(xs)
        cov = cov - Kxs @ self.apply_inverse(Kxs.T)
        return mu, cov

    def sample(self, t, size=1, rng=None):
        xs = np.atleast_2d(np.asarray(t, dtype=np.float64))
        cov = self.kernel.get_value(xs)
        cov[np.diag_indices_from(cov)] += TINY
        rng = np.random if rng is None else rng
        out = rng.multivariate_normal(np.full(xs.shape[0], self.mean), cov, size)
        return out[0] if size == 1 else out


# ---------------------------------------------------------------------------------------------------
# Flat ("engine") parameterisation used by the C ABI: one function that evaluates a whole batch of
# hyper-parameter rows the slow way, one GP object per row.  family 0 = amp*M52 ; 1 = amp*M52*(lin+const)
# ---------------------------------------------------------------------------------------------------
def build_kernel(family, theta, ndim, amp_axes=None):
    """theta in george order / log space: [log_c0, log_M_0..log_M_{d-1}(, log_gamma2, log_cb)]."""
    theta = np.asarray(theta, dtype=np.float64)
    d = ndim - 1 if family == 1 else ndim
    amp_axes = ndim if amp_axes is None else amp_axes
    k = Product(ConstantKernel(theta[0], ndim=ndim, axes=np.arange(amp_axes)),
                Matern52Kernel(np.exp(theta[1:1 + d]), ndim=ndim, axes=np.arange(d)))
    if family == 1:
        k = Product(k, Sum(LinearKernel(theta[1 + d], order=1, ndim=ndim, axes=[d]),
                           ConstantKernel(theta[2 + d], ndim=ndim, axes=[d])))
    k.set_parameter_vector(theta)  # exact log-space values (avoid exp/log round trip)
    return k


def loglik_batch(family, X, y, mean, theta, yerr2, grad=False, amp_axes=None):
    """ll[b] (and grad[b, :Pk+1], last column d/d yerr2) for every row of theta."""
    X = np.asarray(X, dtype=np.float64)
    B = theta.shape[0]
    ll = np.empty(B)
    g = np.zeros((B, theta.shape[1] + 1)) if grad else None
    for b in range(B):
        gp = GP(build_kernel(family, theta[b], X.shape[1], amp_axes), mean=mean)
        try:
            with np.errstate(all="ignore"):
                gp.compute(X, np.sqrt(yerr2[b]))
                ll[b] = gp.log_likelihood(y, quiet=True)
                if grad:
                    g[b] = gp.grad_log_likelihood(y, with_yerr2=True)
        except (np.linalg.LinAlgError, ValueError):
            ll[b] = -np.inf
            if grad:
                g[b] = np.nan
    return (ll, g) if grad else ll
