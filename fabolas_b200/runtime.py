"""Process-wide engine plumbing for the reference-shaped Python surface."""
from __future__ import annotations

from typing import Callable, Optional

from ._capi import Engine

_factory: Optional[Callable[[], Engine]] = None
_shared: Optional[Engine] = None


def set_engine_factory(factory: Optional[Callable[[], Engine]]):
    """Override how engines are created (tests hand in emulator-backed engines).  None restores the
    default: one CUDA engine per call on the current device."""
    global _factory, _shared
    _factory = factory
    _shared = None


def new_engine() -> Engine:
    return _factory() if _factory is not None else Engine()


def shared_engine() -> Engine:
    """Engine for stateless calls (joint_min, kernel_matrix, expected_improvement)."""
    global _shared
    if _shared is None:
        _shared = new_engine()
    return _shared
