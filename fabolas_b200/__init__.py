"""fabolas_b200 -- B200-native (sm_100a) engine for the Gaussian-process inner loop of FABOLAS.

Only the hot path of rickie95/fabolas is re-built here (batched GP marginal likelihood + gradient,
prediction, Entropy Search / EPMGP acquisition) behind the reference's Python-facing surface.  The
numerical work is hand-written CUDA reached through the C ABI in ``include/fabolas_b200.h``;
PyTorch tensors are the buffer carrier only.  There is no CPU fallback.
"""
from ._capi import Engine, FabolasError, FAB_FAMILY_M52, FAB_FAMILY_M52_BASIS, load_library  # noqa: F401

__version__ = "0.1.0"
