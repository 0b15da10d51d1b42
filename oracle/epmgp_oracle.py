"""ctypes wrapper of oracle/epmgp_oracle.c (C restatement of the reference's epmgp.joint_min,
epmgp.py:52-129).  TEST INFRASTRUCTURE; pinned against the reference by tests/golden/epmgp_*.npz."""
import ctypes

import numpy as np

from .build_oracle import build_oracle

_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_oracle())
        _lib.epmgp_joint_min.restype = ctypes.c_int
        _lib.epmgp_joint_min.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 2 + [ctypes.c_int] + [ctypes.c_void_p] * 5
    return _lib


def joint_min(mu, var, with_derivatives=False, return_sweeps=False):
    """Same contract as epmgp.joint_min: logP (D,) [, dlogPdMu (D,D), dlogPdSigma (D,D(D+1)/2),
    dlogPdMudMu (D,D,D)].  Raises Exception where the reference does."""
    lib = _load()
    mu = np.ascontiguousarray(mu, dtype=np.float64).reshape(-1)
    var = np.ascontiguousarray(var, dtype=np.float64)
    D = mu.shape[0]
    T = D * (D + 1) // 2
    logP = np.zeros(D)
    sweeps = np.zeros(D, dtype=np.int32)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    if with_derivatives:
        dMu, dSg, dMM = np.zeros((D, D)), np.zeros((D, T)), np.zeros((D, D, D))
        rc = lib.epmgp_joint_min(D, p(mu), p(var), 1, p(logP), p(dMu), p(dSg), p(dMM), p(sweeps))
    else:
        rc = lib.epmgp_joint_min(D, p(mu), p(var), 0, p(logP), None, None, None, p(sweeps))
    if rc == 1:
        raise Exception("an error occurs while running expectation propagation in entropy search. "
                        "Resulting variance contains NaN")
    if rc == 2:
        raise np.linalg.LinAlgError("Matrix is not positive definite")
    out = (logP, dMu, dSg, dMM) if with_derivatives else logP
    return (out, sweeps) if return_sweeps else out
