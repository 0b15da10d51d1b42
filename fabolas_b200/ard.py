"""ard.AutomaticRelevanceDetermination (reference: ard.py) on the CUDA kernel-matrix builder.

Same class surface: AutomaticRelevanceDetermination(length_scale, length_scale_bounds, nu),
__call__(X, Y=None, eval_gradient=False), repr.  The reference passes an N x N `VI` to scipy's
Mahalanobis distance (ard.py:113,117), which is only well defined when N == D; this class implements
the INTENDED metric VI = 0.1 * I_D (distance^2 = 0.1 * sum_k ((x_k - y_k) / l_k)^2) and agrees with the
reference wherever the reference is well defined (tests/golden/ard_ref.npz).  As in ard.py:142-178 the
gradient w.r.t. log length scales is that of the kernel on the UNSCALED distances."""
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

    def __call__(self, X, Y=None, eval_gradient=False):
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        ls = np.broadcast_to(_check_length_scale(X, self.length_scale), (X.shape[1],))
        if self.nu not in _NU_CODE:
            raise NotImplementedError("general nu (Bessel functions) is outside the accelerated path")
        if Y is not None and eval_gradient:
            raise ValueError("Gradient can only be evaluated when Y is None.")
        code, D = _NU_CODE[self.nu], X.shape[1]
        eng = runtime.shared_engine()
        amp = -np.log(D)                                            # amp_mult * exp(log_c0) = 1
        theta = np.concatenate([[amp], np.log(ls ** 2 / 0.1)])      # M_k = l_k^2 / 0.1
        K = eng.kernel_matrix(FAB_FAMILY_M52, theta, X, None if Y is None else np.atleast_2d(Y), D, nu_code=code)
        K = K[0].cpu().numpy()
        if Y is None:
            np.fill_diagonal(K, 1)
        if not eval_gradient:
            return K
        if self.hyperparameter_length_scale.fixed:
            return K, np.empty((X.shape[0], X.shape[0], 0))
        if self.nu == 0.5:
            raise NotImplementedError("nu=0.5 gradient: the reference's np.divide(..., where=) leaves "
                                      "uninitialised memory on the diagonal (ard.py:156-160)")
        if self.anisotropic:
            theta_g = np.concatenate([[amp], np.log(ls ** 2)])      # ard.py:150-153: D without the 0.1
        else:
            theta_g = theta                                         # ard.py:155: D = dists**2 (0.1 metric kept)
        Kg, dK = eng.kernel_matrix(FAB_FAMILY_M52, theta_g, X, None, D, nu_code=code, grad=True)
        G = 2.0 * dK[0, :, :, 1:].cpu().numpy()                     # d/dlog l = 2 d/dlog M
        if self.nu == np.inf:                                       # ard.py:166-167 multiplies by the 0.1-metric K
            with np.errstate(all="ignore"):
                G = G / Kg[0].cpu().numpy()[..., None] * K[..., None]
        if not self.anisotropic:
            return K, G.sum(-1)[:, :, np.newaxis]
        return K, G

    def __repr__(self):
        if self.anisotropic:
            return "{0}(length_scale=[{1}], nu={2:.3g})".format(
                self.__class__.__name__, ", ".join(map("{0:.3g}".format, self.length_scale)), self.nu)
        return "{0}(length_scale={1:.3g}, nu={2:.3g})".format(
            self.__class__.__name__, np.ravel(self.length_scale)[0], self.nu)
