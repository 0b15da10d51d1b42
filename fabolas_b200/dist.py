"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed (NCCL over NVLink 5 /
NVSwitch on the B200 box; gloo in the CPU tests) for the plumbing.

The path shards without any data-path exchange: walkers / hyper-samples are
split contiguously across ranks for the likelihood, candidates for prediction / EI / information
gain; X, y and the factorised models are replicated.  The only collectives are an all-gather of
the per-walker log-likelihoods (+ gradients), 8 B (+ 8 B P) bytes per walker, once per stretch
half-step, and an all-gather of one (value, index) pair per rank for the acquisition argmax."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, size: int):
    """Contiguous split of range(n): first n % size ranks get one extra item."""
    base, extra = divmod(n, size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _allgather_rows(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather of the row shards produced by shard_range (uneven shards are padded)."""
    rank, size = world()
    if size == 1:
        return local
    per = -(-n_total // size)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((size * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    parts = []
    for r in range(size):
        lo, hi = shard_range(n_total, r, size)
        parts.append(out[r * per: r * per + (hi - lo)])
    return torch.cat(parts, 0)


def sharded_loglik(engine, theta, yerr2, grad: bool = False):
    """ll (B,), [grad (B, Pk+1),] info (B,) for the FULL batch on every rank; each rank evaluates its
    contiguous shard of the walkers on its own GPU."""
    rank, size = world()
    theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
    yerr2 = np.asarray(yerr2, dtype=np.float64).reshape(-1)
    B = theta.shape[0]
    lo, hi = shard_range(B, rank, size)
    if size > 1 and getattr(engine, "peer_world", 0) == size and getattr(engine, "peer_rows", 0) == hi - lo \
            and B == size * (hi - lo):
        # exchange fused into the kernel (Engine.peer_setup was called): each CTA stores its walker's row into every
        # rank's gather buffer over NVLink peer memory; no collective
        engine.loglik_exchange(theta[lo:hi], yerr2[lo:hi], grad=grad)
        rows = engine.peer_wait()
        ll, info = rows[:, -2].contiguous(), rows[:, -1].to(torch.int32)
        # a peer that never delivered leaves stale rows behind: every rank learns about it and raises (the flag is
        # read after the kernel of the step; the callers of this function consume the rows on the host anyway)
        bad = torch.tensor([1.0 if engine.peer_error() else 0.0], device=rows.device)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if bad.item() != 0.0:
            from ._capi import FabolasError
            raise FabolasError("fused likelihood exchange: a rank did not deliver its rows within the timeout")
        return (ll, rows[:, :-2].contiguous(), info) if grad else (ll, info)
    if hi > lo:
        out = engine.loglik(theta[lo:hi], yerr2[lo:hi], grad=grad)
    else:
        out = (engine.empty(0), engine.empty(0, engine.Pk + 1), engine.empty(0, dtype=torch.int32)) if grad else \
              (engine.empty(0), engine.empty(0, dtype=torch.int32))
    if size == 1:
        return out
    # one collective per step: rows packed as [ll | grad | info] (info is a small integer, exact in FP64)
    cols = [out[0][:, None]] + ([out[1]] if grad else []) + [out[-1].double()[:, None]]
    packed = _allgather_rows(torch.cat(cols, 1), B)
    ll, info = packed[:, 0].contiguous(), packed[:, -1].to(torch.int32)
    return (ll, packed[:, 1:-1].contiguous(), info) if grad else (ll, info)


def _argmax_pairs(pairs):
    """np.argmax over the concatenation, from one (value, global index) pair per shard: a NaN is the maximum, the
    first occurrence wins, shards without candidates (index < 0) do not take part."""
    best = None
    for v, i in pairs:
        if i < 0:
            continue
        if best is None:
            best = (v, i)
            continue
        bv, bi = best
        if np.isnan(v) != np.isnan(bv):
            take = bool(np.isnan(v))
        elif np.isnan(v):
            take = i < bi
        else:
            take = v > bv or (v == bv and i < bi)
        if take:
            best = (v, i)
    if best is None:
        raise ValueError("argmax of an empty candidate set")
    return float(best[0]), int(best[1])


def sharded_argmax_pair(value: float, index: int, device):
    """np.argmax over all shards from this rank's (value, GLOBAL index) pair (index < 0: this rank holds no
    candidates): one all-gather of a pair per rank."""
    rank, size = world()
    if size == 1:
        return _argmax_pairs([(value, index)])
    pair = torch.tensor([value, float(index)], dtype=torch.float64, device=device)
    allp = torch.empty(size * 2, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(allp, pair)
    return _argmax_pairs(allp.view(size, 2).cpu().numpy())


def sharded_argmax(values: torch.Tensor, offset: int):
    """Global (max value, global index) of per-rank candidate shards held as tensors (np.argmax semantics: first
    occurrence, NaN is the maximum)."""
    if values.numel():
        nan = torch.isnan(values)
        if bool(nan.any()):
            i = torch.nonzero(nan)[0, 0]
            v = values[i]
        else:
            v = values.max()
            i = torch.nonzero(values == v)[0, 0]            # first occurrence
        return sharded_argmax_pair(float(v), int(i) + offset, values.device)
    return sharded_argmax_pair(-float("inf"), -1, values.device)


def sharded_ei_argmax(engine, T, y_max, xi: float = 0.0):
    """expected_improvement.get_candidate's sweep (predict -> EI -> argmax, expected_improvement.py:85-93) with the M
    candidates of T sharded across ranks; the models are resident (replicated) on every rank.  Model 0 decides.  The
    local argmax is the engine's own (ei_argmax_kernel); only one (value, index) pair per rank is exchanged."""
    rank, size = world()
    T = np.atleast_2d(np.asarray(T, dtype=np.float64)) if not isinstance(T, torch.Tensor) else T
    lo, hi = shard_range(T.shape[0], rank, size)
    if hi > lo:
        mu, var = engine.predict(T[lo:hi], want_var=True)
        _, bv, bi = engine.expected_improvement(mu[:1], var[:1], np.array([y_max]), xi=xi, want_ei=False)
        return sharded_argmax_pair(float(bv[0]), int(bi[0]) + lo, engine.device)
    return sharded_argmax_pair(-float("inf"), -1, engine.device)


def sharded_information_gain(engine, Xc):
    """Information gain of every row of Xc on every rank; candidates sharded across ranks."""
    rank, size = world()
    Xc = np.atleast_2d(np.asarray(Xc, dtype=np.float64))
    lo, hi = shard_range(Xc.shape[0], rank, size)
    if hi > lo:
        ig, flag = engine.information_gain(Xc[lo:hi])
    else:
        ig, flag = engine.empty(0), engine.empty(0, dtype=torch.int32)
    return _allgather_rows(ig, Xc.shape[0]), _allgather_rows(flag, Xc.shape[0])
