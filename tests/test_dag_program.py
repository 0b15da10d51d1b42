"""Host-side static scheduler of the fused GP kernel (fabolas_b200/csrc/dag_program.h): for every mode,
slot budget and problem size the emitted two-stream tile program must pass the independent verifier
(slot contents at every access, every cross-stream hazard ordered by an explicit wait, no deadlock).
No device needed: the scheduler is plain C++ inside the C-ABI library."""
import ctypes

import pytest

from fabolas_b200 import load_library

MODES = {0: "ll", 1: "factor", 2: "fit", 3: "grad"}


@pytest.mark.parametrize("mode", sorted(MODES))
@pytest.mark.parametrize("nslot", [4, 5, 6])
def test_programs_verify(mode, nslot):
    lib = load_library()
    msg = ctypes.create_string_buffer(512)
    stats = (ctypes.c_int * 4)()
    for NT in (1, 2, 3, 4, 5, 6, 8, 11, 16, 32):
        rc = lib.fab_debug_dag_verify(NT, nslot, mode, msg, 512, stats)
        assert rc == 0, f"mode {MODES[mode]} slots {nslot} NT {NT}: {msg.value.decode()}"
        ntiles = NT * (NT + 1) // 2
        assert stats[1] == NT                      # one chain op per diagonal tile
        if mode == 2:
            assert stats[3] >= ntiles              # fit: every W tile reaches the workspace
    # N <= 256 at the full slot budget: the log-likelihood streams (almost) nothing through L2
    if nslot == 6 and mode == 0:
        lib.fab_debug_dag_verify(4, 6, 0, msg, 512, stats)
        assert stats[2] <= 4


def test_bad_arguments():
    lib = load_library()
    msg = ctypes.create_string_buffer(64)
    assert lib.fab_debug_dag_verify(4, 3, 0, msg, 64, None) == -1      # too few slots
    assert lib.fab_debug_dag_verify(4, 6, 7, msg, 64, None) != 0       # unknown mode
