"""ard.AutomaticRelevanceDetermination (reference: ard.py) on the CUDA kernel-matrix builder.

Same class surface: AutomaticRelevanceDetermination(length_scale, length_scale_bounds, nu),
__call__(X, Y=None, eval_gradient=False), repr.  The reference passes an N x N `VI` to scipy's
Mahalanobis distance (ard.py:113,117), which is only well defined when N == D; this class implements
the INTENDED metric VI = 0.1 * I_D (distance^2 = 0.1 * sum_k ((x_k - y_k) / l_k)^2) and agrees with the
reference wherever the reference is well defined (tests/golden/ard_ref.npz).  As in ard.py:142-178 the
gradient w.r.t. log length scales is that of the kernel on the UNSCALED distances.  Every branch of ard.py:119-135
runs on the device: nu in {0.5, 1.5, 2.5, inf} in closed form, any other nu through the Bessel function K_nu
(csrc/kmat.cuh: bessel_k), whose gradient the reference itself takes by forward differences (ard.py:168-173)."""
import numpy as np
from sklearn.gaussian_process.kernels import RBF

from . import runtime
from ._capi import FAB_FAMILY_M52

_NU_CODE = {2.5: 0, 1.5: 1, 0.5: 2, np.inf: 3}


def _check_length_scale(X, length_scale):
    length_scale = np.squeeze(length_scale).astype(float)
    if np.ndim(length_scale) > 1:
        raise ValueError("length_scale cannot be of dimension greater than 1")
    if np.ndim(length_scale) == 1 and X.shape[1] != length_scale.shape[0]:
        raise ValueError("Anisotropic kernel must have the same number of dimensions as data (%d!=%d)"
                         % (length_scale.shape[0], X.shape[1]))
    return length_scale


class AutomaticRelevanceDetermination(RBF):
    def __init__(self, length_scale=1.0, length_scale_bounds=(1e-5, 1e5), nu=1.5):
        super().__init__(length_scale, length_scale_bounds)
        self.nu = nu

    def _kernel(self, eng, X, Y, ls, metric, grad=False):
        """k (and analytic d/dlog M) on r2 = metric * sum_k ((x_k - y_k) / l_k)^2 from the device builder."""
        D = X.shape[1]
        theta = np.concatenate([[-np.log(D)], np.log(ls ** 2 / metric)])     # amp_mult * exp(log_c0) = 1 ; M_k = l_k^2 / metric
        code = _NU_CODE.get(self.nu, 4)
        return eng.kernel_matrix(FAB_FAMILY_M52, theta, X, Y, D, nu_code=code, grad=grad, nu=float(self.nu) if code == 4 else 0.0)

    def __call__(self, X, Y=None, eval_gradient=False):
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        ls = np.broadcast_to(_check_length_scale(X, self.length_scale), (X.shape[1],)).astype(np.float64)
        if Y is not None and eval_gradient:
            raise ValueError("Gradient can only be evaluated when Y is None.")
        eng = runtime.shared_engine()
        Y = None if Y is None else np.atleast_2d(np.asarray(Y, dtype=np.float64))
        K = self._kernel(eng, X, Y, ls, 0.1)[0].cpu().numpy()               # VI = 0.1 I (see the module docstring)
        if Y is None:
            np.fill_diagonal(K, 1)
        if not eval_gradient:
            return K
        if self.hyperparameter_length_scale.fixed:
            return K, np.empty((X.shape[0], X.shape[0], 0))
        if self.nu not in _NU_CODE:
            # ard.py:168-173: forward differences of the kernel in theta = log length_scale, step 1e-10
            th = np.log(np.atleast_1d(self.length_scale).astype(np.float64))
            G = np.empty(K.shape + (th.shape[0],))
            for k in range(th.shape[0]):
                tk = th.copy(); tk[k] += 1e-10
                lk = np.broadcast_to(np.exp(tk), (X.shape[1],)).astype(np.float64)
                Kk = self._kernel(eng, X, None, lk, 0.1)[0].cpu().numpy()
                np.fill_diagonal(Kk, 1)
                G[..., k] = (Kk - K) / 1e-10
            return K, G
        # ard.py:150-155: D = squared scaled differences WITHOUT the 0.1 of the metric when anisotropic, dists**2 (metric
        # kept) otherwise
        Kg, dK = self._kernel(eng, X, None, ls, 1.0 if self.anisotropic else 0.1, grad=True)
        G = 2.0 * dK[0, :, :, 1:].cpu().numpy()                     # d/dlog l = 2 d/dlog M
        if self.nu in (np.inf, 0.5):
            # ard.py:156-160, 166-167: these two branches multiply by the 0.1-metric K itself; the diagonal of the
            # nu = 0.5 gradient is 0/0 in the reference (np.divide(..., where=) leaves it unset): 0 here, its limit
            with np.errstate(all="ignore"):
                G = np.where(Kg[0].cpu().numpy()[..., None] > 0, G / Kg[0].cpu().numpy()[..., None] * K[..., None], 0.0)
        if not self.anisotropic:
            return K, G.sum(-1)[:, :, np.newaxis]
        return K, G

    def __repr__(self):
        if self.anisotropic:
            return "{0}(length_scale=[{1}], nu={2:.3g})".format(
                self.__class__.__name__, ", ".join(map("{0:.3g}".format, self.length_scale)), self.nu)
        return "{0}(length_scale={1:.3g}, nu={2:.3g})".format(
            self.__class__.__name__, np.ravel(self.length_scale)[0], self.nu)
