"""Multi-process sharding (world_size 2, gloo, CPU): the N > 1 host logic of fabolas_b200.dist with
emulator-backed engines.  The GPU box runs the same code over NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, size, port, lib, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from fabolas_b200 import Engine
        from fabolas_b200 import dist as fd
        rng = np.random.default_rng(0)                     # identical inputs on every rank
        N, D, B = 20, 3, 5                                 # 5 walkers over 2 ranks: uneven shards
        X = rng.uniform(-1, 1, (N, D)); X[:, -1] = rng.uniform(0, 1, N)
        y = np.sin(X.sum(1))
        theta = rng.uniform(-1, 1, (B, D + 2)); yerr2 = rng.uniform(1e-3, 0.1, B)
        eng = Engine(device="cpu", lib_path=lib)
        eng.set_data(1, X, y, float(y.mean()), D)
        ll, g, info = fd.sharded_loglik(eng, theta, yerr2, grad=True)
        ll1, g1, _ = eng.loglik(theta, yerr2, grad=True)
        assert ll.shape == (B,) and g.shape == (B, D + 3) and int(info.abs().sum()) == 0
        assert torch.equal(ll, ll1) and torch.equal(g, g1)
        # candidate sharding + argmax all-gather
        eng.fit(theta[:2], yerr2[:2])
        T = rng.uniform(0, 1, (37, D))
        val, idx = fd.sharded_ei_argmax(eng, T, float(y.max()))
        mu, var = eng.predict(T)
        ei, bv, bi = eng.expected_improvement(mu[:1], var[:1], np.array([y.max()]))
        assert idx == int(bi[0]) and abs(val - float(bv[0])) < 1e-15
        assert fd.shard_range(5, 0, 2) == (0, 3) and fd.shard_range(5, 1, 2) == (3, 5)
        # np.argmax semantics across shards: first occurrence of the maximum; the first NaN is the maximum; a rank
        # without candidates does not take part
        vals = torch.tensor([0.5, 2.0, 1.0, 2.0, 0.1, 2.0, 0.3])
        lo, hi = fd.shard_range(7, rank, size)
        assert fd.sharded_argmax(vals[lo:hi], lo) == (2.0, 1)
        vals[5] = float("nan")
        v, i = fd.sharded_argmax(vals[lo:hi], lo)
        assert np.isnan(v) and i == 5 == int(np.argmax(vals.numpy()))
        vals[0] = float("nan")
        v, i = fd.sharded_argmax(vals[lo:hi], lo)
        assert np.isnan(v) and i == 0
        one = torch.tensor([3.0])                          # one candidate in all: rank 1 holds nothing
        lo, hi = fd.shard_range(1, rank, size)
        assert fd.sharded_argmax(one[lo:hi], lo) == (3.0, 0)
        ig_ref = None
        # candidate-sharded information gain needs the Entropy-Search state: small setup, same on both ranks
        Z, I = 5, 3
        reps = rng.uniform(0, 1, (2, Z, D)); Omega = rng.standard_normal((2, I, Z))
        mu_r, _, cov_r = eng.predict(reps, want_var=True, want_cov=True)
        cov_c = torch.clamp(cov_r, min=2.2e-16)
        pm = eng.joint_min(mu_r, cov_c)
        U = torch.full((2, Z), 0.1, dtype=torch.float64)
        eng.es_setup(reps, cov_r, pm[:4], U, Omega)
        Xc = rng.uniform(0, 1, (9, D))
        ig, flag = fd.sharded_information_gain(eng, Xc)
        ig_ref, flag_ref = eng.information_gain(Xc)
        assert torch.equal(ig, ig_ref) and torch.equal(flag, flag_ref)
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_sharded_paths_world2_gloo():
    from tests.cusim.build_sim import build_sim
    lib = build_sim()
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lib, out)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs) and len(out) == 2


def _check_exchange_world1(eng):
    """The fused exchange with a world of one: the kernel's peer stores land in the local gather buffer."""
    import torch
    from fabolas_b200.workloads import gp_workload
    w = gp_workload(1, n_walkers=6)
    eng.set_data(w["family"], w["X"][:40], w["y"][:40], w["mean"], w["amp_mult"])
    eng.peer_setup(6)
    try:
        for step, grad in enumerate((True, False, True)):
            th = w["theta"] + 0.01 * step
            ref = eng.loglik(th, w["yerr2"], grad=True)
            out = eng.loglik_exchange(th, w["yerr2"], grad=grad)
            rows = eng.peer_wait()
            assert not eng.peer_error()
            rows = rows.cpu().numpy() if rows.is_cuda else rows.numpy()
            assert rows.shape == (6, eng.Pk + 3)
            assert np.array_equal(rows[:, -2], ref[0].cpu().numpy()) and np.array_equal(rows[:, -1], ref[2].cpu().numpy())
            assert np.array_equal(out[0].cpu().numpy(), ref[0].cpu().numpy())
            if grad:
                assert np.array_equal(rows[:, :-2], ref[1].cpu().numpy())
    finally:
        eng.peer_free()


def test_fused_exchange_world1_emulated(sim_engine):
    _check_exchange_world1(sim_engine)


@pytest.mark.gpu
def test_fused_exchange_world1_gpu(gpu_engine):
    _check_exchange_world1(gpu_engine)
