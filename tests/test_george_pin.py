"""The george layer against george itself -- active the moment george is importable or
tests/golden/george_ref.npz has been committed from a machine where it is (tests/golden/make_george_golden.py).
Until then these tests SKIP and say why: the george oracle is PARITY UNPINNED.

What always runs: the generating script against a shim that presents oracle.george_oracle under george's module
layout, so that the pin mechanism itself is known to work (it proves nothing about george)."""
import importlib.util
import os
import types

import numpy as np
import pytest

from oracle import george_oracle as go
from tests.conftest import GOLDEN, rel_err

spec = importlib.util.spec_from_file_location("make_george_golden", os.path.join(GOLDEN, "make_george_golden.py"))
mgg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mgg)

UNPINNED = ("george 0.4.0 is not importable and tests/golden/george_ref.npz has not been generated: "
            "the george oracle stays PARITY UNPINNED")


def _reference_vectors():
    """george's own numbers: live when george imports, else the committed file, else None."""
    try:
        import george
        from george import kernels
        return mgg.generate(george, kernels)
    except ImportError:
        path = os.path.join(GOLDEN, "george_ref.npz")
        return dict(np.load(path, allow_pickle=False)) if os.path.exists(path) else None


def _oracle_shim():
    g = types.SimpleNamespace(GP=go.GP, __version__="oracle-shim")
    k = types.SimpleNamespace(Matern52Kernel=go.Matern52Kernel, LinearKernel=go.LinearKernel, ConstantKernel=go.ConstantKernel)
    return g, k


def _compare_oracle(ref):
    mine = mgg.generate(*_oracle_shim())
    for tag in "AB":
        assert int(ref[f"{tag}_len"]) == int(mine[f"{tag}_len"])                                   # convention (1)
        assert np.allclose(ref[f"{tag}_initial_vector"], mine[f"{tag}_initial_vector"], rtol=0, atol=1e-14)
        assert rel_err(mine[f"{tag}_K"], ref[f"{tag}_K"]) < 1e-12                                  # (2)
        assert np.max(np.abs(mine[f"{tag}_ll"] - ref[f"{tag}_ll"]) / np.abs(ref[f"{tag}_ll"])) < 1e-9   # (3)
        assert rel_err(mine[f"{tag}_grad"], ref[f"{tag}_grad"], axis=1) < 1e-9                     # (4)
        assert rel_err(mine[f"{tag}_mu"], ref[f"{tag}_mu"]) < 1e-9                                 # (5)
        scale = np.abs(ref[f"{tag}_Kt"]).max()
        assert np.max(np.abs(mine[f"{tag}_cov"] - ref[f"{tag}_cov"])) < 1e-9 * scale
        assert rel_err(mine[f"{tag}_Kt"], ref[f"{tag}_Kt"]) < 1e-12                                # (6)


def test_pin_script_runs_against_oracle_shim():
    out = mgg.generate(*_oracle_shim())
    assert int(out["A_len"]) == 5 and int(out["B_len"]) == 4          # D = 3: [c0, M0, M1, g2, cb] and [c0, M0, M1, M2]
    assert out["A_K"].shape == (4, 100, 100) and out["A_grad"].shape == (4, 5) and out["B_cov"].shape == (4, 9, 9)
    assert np.all(np.isfinite(out["A_ll"])) and np.all(np.isfinite(out["B_ll"]))
    _compare_oracle(out)                                              # self-consistency of the comparison code


def test_oracle_matches_george():
    ref = _reference_vectors()
    if ref is None:
        pytest.skip(UNPINNED)
    _compare_oracle(ref)


@pytest.mark.gpu
def test_engine_matches_george_gpu(gpu_engine):
    """CUDA path against george's own ll / gradient / prediction (george order, log space)."""
    ref = _reference_vectors()
    if ref is None:
        pytest.skip(UNPINNED)
    X, y, T = ref["X"], ref["y"], ref["T"]
    nd = X.shape[1]
    for tag, family in (("A", 1), ("B", 0)):
        gpu_engine.set_data(family, X, y, float(np.mean(y)), nd)
        V, yerr = ref[f"{tag}_vectors"], ref[f"{tag}_yerr"]
        ll, g, info = gpu_engine.loglik(V, yerr ** 2, grad=True)
        assert int(info.abs().sum()) == 0
        assert np.max(np.abs(ll.cpu().numpy() - ref[f"{tag}_ll"]) / np.abs(ref[f"{tag}_ll"])) < 1e-9
        assert rel_err(g.cpu().numpy()[:, :-1], ref[f"{tag}_grad"], axis=1) < 1e-9
        gpu_engine.fit(V, yerr ** 2)
        mu, var, cov = gpu_engine.predict(T, want_cov=True)
        assert rel_err(mu.cpu().numpy(), ref[f"{tag}_mu"]) < 1e-9
        assert np.max(np.abs(cov.cpu().numpy() - ref[f"{tag}_cov"])) < 1e-9 * np.abs(ref[f"{tag}_Kt"]).max()
        K = gpu_engine.kernel_matrix(family, V, X, amp_mult=nd).cpu().numpy()
        assert rel_err(K, ref[f"{tag}_K"]) < 1e-12
