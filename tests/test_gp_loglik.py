"""Batched GP log-likelihood + gradient: CUDA path (C ABI) against the CPU oracle.

`-m "not gpu"`: the same CUDA sources replayed by the fiber emulator (tests/cusim) -- kernel logic.
`-m gpu`      : the real sm_100a kernels through the C ABI on a B200 -- the parity tests proper.
Tolerance (north star): ll and gradients within 1e-9 relative in FP64.
"""
import numpy as np
import pytest

from fabolas_b200.workloads import gp_workload
from oracle import george_oracle as go
from tests.conftest import golden, rel_err

TOL = 1e-9


def _problem(N, D, family, B, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (N, D))
    if family == 1:
        X[:, -1] = rng.uniform(0, 1, N)
    y = np.sin(X.sum(1)) + 0.05 * rng.standard_normal(N)
    d = D - 1 if family == 1 else D
    Pk = d + 3 if family == 1 else d + 1
    theta = rng.uniform(-1, 1, (B, Pk))
    theta[:, 1:1 + d] = rng.uniform(-1, 2, (B, d))
    return X, y, theta, rng.uniform(1e-3, 0.1, B)


def _check(eng, X, y, theta, yerr2, family, grad=True, nref=None, amp_mult=None):
    D = X.shape[1]
    amp_mult = D if amp_mult is None else amp_mult
    eng.set_data(family, X, y, float(y.mean()), amp_mult)
    out = eng.loglik(theta, yerr2, grad=grad)
    ll = out[0].cpu().numpy()
    info = out[-1].cpu().numpy()
    nref = len(theta) if nref is None else nref
    ref = go.loglik_batch(family, X, y, y.mean(), theta[:nref], yerr2[:nref], grad=grad, amp_axes=int(amp_mult))
    if grad:
        ref, gref = ref
        g = out[1].cpu().numpy()[:nref]
        assert rel_err(g, gref, axis=1) < TOL
    assert np.all(info[:nref] == 0)
    assert np.max(np.abs(ll[:nref] - ref) / np.abs(ref)) < TOL
    return ll


SIM_CASES = [(1, 1, 0, 1), (7, 2, 0, 2), (40, 3, 1, 2), (64, 3, 0, 1), (65, 4, 1, 1), (130, 6, 1, 1)]


@pytest.mark.parametrize("N,D,family,B", SIM_CASES)
def test_kernel_logic_emulated(sim_engine, N, D, family, B):
    X, y, theta, yerr2 = _problem(N, D, family, B, seed=N)
    _check(sim_engine, X, y, theta, yerr2, family, grad=True)
    _check(sim_engine, X, y, theta, yerr2, family, grad=False)


def test_tile_streaming_path_emulated(sim_engine):
    """Fewer shared-memory slots than tiles: factor and gradient kernels stream tiles through TMA."""
    X, y, theta, yerr2 = _problem(200, 4, 1, 1, seed=5)          # NT = 4 -> 10 tiles
    try:
        sim_engine.debug_set_max_slots(4)
        _check(sim_engine, X, y, theta, yerr2, 1, grad=True)
        sim_engine.debug_set_max_slots(5)
        sim_engine.debug_force_slices(True)                       # large-N mode of the gradient kernel
        _check(sim_engine, X, y, theta, yerr2, 1, grad=True)
    finally:
        sim_engine.debug_set_max_slots(6)
        sim_engine.debug_force_slices(False)


def test_not_positive_definite_is_data_emulated(sim_engine):
    X, y, theta, yerr2 = _problem(20, 2, 0, 3, seed=1)
    X[5] = X[4]                                   # duplicate observation + negative jitter => indefinite
    yerr2[:] = [1e-3, -1e-2, 1e-3]
    sim_engine.set_data(0, X, y, float(y.mean()), 2)
    ll, g, info = sim_engine.loglik(theta, yerr2, grad=True)
    ll, info = ll.numpy(), info.numpy()
    assert np.isfinite(ll[0]) and np.isfinite(ll[2]) and ll[1] == -np.inf
    assert info[0] == 0 and info[2] == 0 and info[1] > 0
    assert np.all(np.isnan(g.numpy()[1]))
    ref = go.loglik_batch(0, X, y, y.mean(), theta, yerr2, amp_axes=2)
    assert ref[1] == -np.inf and abs(ll[0] - ref[0]) < TOL * abs(ref[0])


def test_host_buffer_entry_emulated(sim_engine):
    """fab_gp_loglik_batch_host on the emulator: the kernel's host-input mode (one warp stages the walker's record in
    shared memory) gives exactly what the device-buffer entry gives."""
    import torch
    rng = np.random.default_rng(5)
    X = rng.uniform(-1, 1, (70, 3)); y = rng.standard_normal(70)
    theta = np.column_stack([rng.uniform(-1, 1, 3), rng.uniform(0, 2, (3, 3))]); yerr2 = rng.uniform(1e-3, 0.1, 3)
    sim_engine.set_data(0, X, y, float(y.mean()), 3)
    ll, g, info = sim_engine.loglik(theta, yerr2, grad=True)
    ll_h, g_h, info_h = sim_engine.loglik_host(torch.from_numpy(theta), torch.from_numpy(yerr2), grad=True)
    sim_engine.synchronize()
    assert torch.equal(ll_h, ll) and torch.equal(g_h, g) and torch.equal(info_h, info)


def test_argument_errors_emulated(sim_engine):
    from fabolas_b200 import FabolasError
    X, y, theta, yerr2 = _problem(10, 2, 0, 2)
    sim_engine.set_data(0, X, y, 0.0, 2)
    with pytest.raises(ValueError):
        sim_engine.loglik(theta[:, :2], yerr2)
    with pytest.raises(FabolasError):
        sim_engine.set_data(0, np.zeros((4, 12)), np.zeros(4), 0.0, 12)        # > 8 Matern axes


# ------------------------------------------------------------------------------------------- GPU
GPU_CASES = [(1, 1, 0, 2), (7, 2, 0, 3), (63, 3, 1, 4), (64, 3, 0, 4), (65, 4, 1, 4), (100, 3, 1, 20),
             (128, 6, 1, 8), (200, 5, 0, 6), (256, 6, 1, 16), (300, 6, 1, 4), (512, 6, 1, 3)]


@pytest.mark.gpu
@pytest.mark.parametrize("N,D,family,B", GPU_CASES)
def test_loglik_grad_parity_gpu(gpu_engine, N, D, family, B):
    X, y, theta, yerr2 = _problem(N, D, family, B, seed=N)
    _check(gpu_engine, X, y, theta, yerr2, family, grad=True, nref=min(B, 6))
    _check(gpu_engine, X, y, theta, yerr2, family, grad=False, nref=min(B, 6))


@pytest.mark.gpu
@pytest.mark.parametrize("config", [1, 2])
def test_baseline_configs_gpu(gpu_engine, config):
    """BASELINE.json configs[0] (N=100, 20 walkers) and configs[1] (N=256, 128 walkers) at full size."""
    w = gp_workload(config)
    ll = _check(gpu_engine, w["X"], w["y"], w["theta"], w["yerr2"], w["family"], grad=True, nref=8,
                amp_mult=w["amp_mult"])
    assert ll.shape[0] == {1: 20, 2: 128}[config] and np.all(np.isfinite(ll))
    # size-independent property on the whole batch: permuting the observations leaves ll and grad unchanged
    perm = np.random.default_rng(0).permutation(w["N"])
    gpu_engine.set_data(w["family"], w["X"][perm], w["y"][perm], w["mean"], w["amp_mult"])
    ll2, g2, _ = gpu_engine.loglik(w["theta"], w["yerr2"], grad=True)
    gpu_engine.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])
    ll1, g1, _ = gpu_engine.loglik(w["theta"], w["yerr2"], grad=True)
    assert np.max(np.abs(ll2.cpu().numpy() - ll) / np.abs(ll)) < TOL
    assert rel_err(g2.cpu().numpy(), g1.cpu().numpy(), axis=1) < 1e-8


@pytest.mark.gpu
def test_tile_streaming_path_gpu(gpu_engine):
    X, y, theta, yerr2 = _problem(256, 6, 1, 5, seed=6)
    try:
        for n in (4, 5):
            gpu_engine.debug_set_max_slots(n)
            _check(gpu_engine, X, y, theta, yerr2, 1, grad=True)
    finally:
        gpu_engine.debug_set_max_slots(6)


@pytest.mark.gpu
@pytest.mark.parametrize("N", [1000, 2048])
def test_large_history_loglik_gpu(gpu_engine, N):
    """BASELINE configs[4] shape (N = 2048 observations): the factor streams its tiles through TMA."""
    w = gp_workload(5, n_walkers=3) if N == 2048 else None
    if w is None:
        X, y, theta, yerr2 = _problem(N, 6, 1, 3, seed=N)
        _check(gpu_engine, X, y, theta, yerr2, 1, grad=False)
    else:
        _check(gpu_engine, w["X"], w["y"], w["theta"], w["yerr2"], w["family"], grad=False, amp_mult=w["amp_mult"])


@pytest.mark.gpu
@pytest.mark.parametrize("N", [1000, 2048])
def test_large_history_gradient_gpu(gpu_engine, N):
    """Gradient parity where the tiles no longer fit the shared-memory slots (NT = 16 / 32 tile rows cycling through
    six slots; W = L^-1 and the dK contraction read spilled tiles back through TMA): two walkers against the oracle."""
    if N == 2048:
        w = gp_workload(5, n_walkers=2)
        _check(gpu_engine, w["X"], w["y"], w["theta"], w["yerr2"], w["family"], grad=True, amp_mult=w["amp_mult"])
    else:
        X, y, theta, yerr2 = _problem(N, 6, 1, 2, seed=N)
        _check(gpu_engine, X, y, theta, yerr2, 1, grad=True)


@pytest.mark.gpu
@pytest.mark.parametrize("nslot,slices", [(4, False), (5, False), (4, True)])
def test_tile_streaming_path_n512_gpu(gpu_engine, nslot, slices):
    """N = 512 (36 tiles) squeezed through 4 / 5 slots, and with the coordinate slices staged per tile (the large-N
    mode of the covariance generation and the contraction): every tile is evicted and reloaded several times."""
    X, y, theta, yerr2 = _problem(512, 6, 1, 3, seed=77)
    try:
        gpu_engine.debug_set_max_slots(nslot)
        gpu_engine.debug_force_slices(slices)
        _check(gpu_engine, X, y, theta, yerr2, 1, grad=True)
        _check(gpu_engine, X, y, theta, yerr2, 1, grad=False)
    finally:
        gpu_engine.debug_set_max_slots(6)
        gpu_engine.debug_force_slices(False)


@pytest.mark.gpu
def test_golden_svm_mnist_prior_gpu(gpu_engine):
    g = golden("gp_oracle_svm_mnist.npz")                      # the reference's real prior CSV (N=100, D=3)
    gpu_engine.set_data(1, g["X"], g["y"], float(g["y"].mean()), 3)
    ll, gr, info = gpu_engine.loglik(g["theta"], g["yerr2"], grad=True)
    assert np.max(np.abs(ll.cpu().numpy() - g["ll"]) / np.abs(g["ll"])) < TOL
    assert rel_err(gr.cpu().numpy(), g["grad"], axis=1) < TOL


@pytest.mark.gpu
def test_gradient_vs_finite_differences_gpu(gpu_engine):
    """Property that does not involve the oracle: central differences of the GPU ll itself."""
    X, y, theta, yerr2 = _problem(256, 6, 1, 1, seed=9)
    gpu_engine.set_data(1, X, y, float(y.mean()), 6)
    _, g, _ = gpu_engine.loglik(theta, yerr2, grad=True)
    g = g.cpu().numpy()[0]
    P = theta.shape[1]
    h = 1e-5
    tt = np.repeat(theta, 2 * P, axis=0)
    for k in range(P):
        tt[2 * k, k] += h
        tt[2 * k + 1, k] -= h
    ll, _ = gpu_engine.loglik(tt, np.repeat(yerr2, 2 * P))
    ll = ll.cpu().numpy()
    fd = (ll[0::2] - ll[1::2]) / (2 * h)
    assert np.max(np.abs(fd - g[:P]) / np.maximum(1.0, np.abs(g[:P]))) < 1e-5


@pytest.mark.gpu
def test_not_positive_definite_is_data_gpu(gpu_engine):
    X, y, theta, yerr2 = _problem(150, 3, 0, 4, seed=2)
    X[7] = X[3]
    yerr2[1] = -5e-2
    gpu_engine.set_data(0, X, y, float(y.mean()), 3)
    ll, g, info = gpu_engine.loglik(theta, yerr2, grad=True)
    ll, info = ll.cpu().numpy(), info.cpu().numpy()
    assert ll[1] == -np.inf and info[1] > 0 and np.all(np.isnan(g.cpu().numpy()[1]))
    ok = [0, 2, 3]
    ref = go.loglik_batch(0, X, y, y.mean(), theta[ok], yerr2[ok], amp_axes=3)
    assert np.all(info[ok] == 0) and np.max(np.abs(ll[ok] - ref) / np.abs(ref)) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("pinned", [True, False])
def test_loglik_host_buffers_gpu(pinned):
    """fab_gp_loglik_batch_host (host buffers, the reference's calling pattern fabolas.py:103-117): zero-copy for pinned
    memory, staged for pageable memory -- bit-identical to the device-buffer entry either way, ll and ll+grad."""
    import torch
    from fabolas_b200 import Engine
    from fabolas_b200.workloads import gp_workload
    w = gp_workload(2, n_walkers=37)
    eng = Engine()
    eng.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])
    th = torch.from_numpy(w["theta"].copy())
    ye = torch.from_numpy(w["yerr2"].copy())
    ye[3] = -5.0                                   # a walker whose covariance is not positive definite
    if pinned:
        th, ye = th.pin_memory(), ye.pin_memory()
    ll_d, g_d, info_d = eng.loglik(th.cuda(), ye.cuda(), grad=True)
    ll_h, g_h, info_h = eng.loglik_host(th, ye, grad=True)
    eng.synchronize()
    assert ll_h.is_pinned() == pinned
    assert torch.equal(ll_h, ll_d.cpu()) and torch.equal(info_h, info_d.cpu())
    assert torch.equal(torch.nan_to_num(g_h, nan=123.0), torch.nan_to_num(g_d.cpu(), nan=123.0))
    assert info_h[3] > 0 and ll_h[3] == -float("inf")
    ll2, info2 = eng.loglik_host(th, ye)
    eng.synchronize()
    ll2_d, _ = eng.loglik(th.cuda(), ye.cuda())
    assert torch.equal(ll2, ll2_d.cpu()) and torch.equal(info2, info_d.cpu())
    with pytest.raises(ValueError):
        eng.loglik_host(th.cuda(), ye.cuda())
