"""epmgp.joint_min with the reference's signature (epmgp.py:52-129), evaluated by the CUDA EPMGP
kernels.  `joint_min_batch` is the batched entry the candidate search uses."""
import numpy as np

from . import runtime


def joint_min_batch(mu, var, with_derivatives=False, engine=None):
    """mu (B, Z), var (B, Z, Z) tensors or arrays -> device tensors (logP[, dMu, dSigma, dMuMu])."""
    eng = engine or runtime.shared_engine()
    out = eng.joint_min(mu, var, with_derivatives=with_derivatives)
    info = out[-1].cpu().numpy()
    if np.any(info == 1):
        raise Exception("an error occurs while running expectation propagation in entropy search. "
                        "Resulting variance contains NaN")
    if np.any(info == 2):
        raise np.linalg.LinAlgError("Matrix is not positive definite")
    return out[:-1] if with_derivatives else out[0]


def joint_min(mu, var, with_derivatives=False, **kwargs):
    mu = np.asarray(mu, dtype=np.float64).reshape(-1)
    var = np.asarray(var, dtype=np.float64)
    out = joint_min_batch(mu[None], var[None], with_derivatives)
    if not with_derivatives:
        return out[0].cpu().numpy()
    return tuple(t[0].cpu().numpy() for t in out)
