"""numpy restatement of the reference's ard.AutomaticRelevanceDetermination.__call__ (ard.py:82-180) with
the INTENDED metric VI = 0.1 * I_D (the reference passes an N x N VI, ard.py:113,117, which scipy only
interprets sensibly when N == D).  TEST INFRASTRUCTURE ONLY.
PINNED on N == D inputs against ard.py itself (tests/golden/ard_*.npz)."""
import numpy as np


def ard_kernel(X, Y=None, length_scale=1.0, nu=2.5, eval_gradient=False):
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    ls = np.broadcast_to(np.asarray(length_scale, dtype=np.float64), (X.shape[1],))
    A = X / ls
    Bm = A if Y is None else np.atleast_2d(np.asarray(Y, dtype=np.float64)) / ls
    d2 = ((A[:, None, :] - Bm[None, :, :]) ** 2)
    dist = np.sqrt(0.1 * d2.sum(-1))
    if nu == 0.5:
        K = np.exp(-dist)
    elif nu == 1.5:
        t = np.sqrt(3.0) * dist
        K = (1.0 + t) * np.exp(-t)
    elif nu == 2.5:
        t = np.sqrt(5.0) * dist
        K = (1.0 + t + t * t / 3.0) * np.exp(-t)
    elif nu == np.inf:
        K = np.exp(-dist ** 2 / 2.0)
    else:
        raise NotImplementedError("general nu needs Bessel functions; not on the hot path")
    if Y is None:
        np.fill_diagonal(K, 1.0)
    if not eval_gradient:
        return K
    # ard.py:142-178: gradient w.r.t. log length scales uses the UNSCALED (no 0.1) per-dimension
    # squared distances D = (x_i - x_j)^2 / l^2 (ard.py:150-153), reproduced as such.
    if nu == 2.5:
        t = np.sqrt(5.0 * d2.sum(-1))[..., None]
        G = 5.0 / 3.0 * d2 * (t + 1.0) * np.exp(-t)
    elif nu == 1.5:
        G = 3.0 * d2 * np.exp(-np.sqrt(3.0 * d2.sum(-1)))[..., None]
    elif nu == np.inf:
        G = d2 * K[..., None]
    else:
        raise NotImplementedError
    if np.ndim(length_scale) == 0:
        G = G.sum(-1)[:, :, None]
    return K, G
