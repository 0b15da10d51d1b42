"""Seeded synthetic workloads of the BASELINE.json configurations.

Shared by bench.py, the tests and __graft_entry__.smoke() so that every leg (GPU path, CPU oracle,
reference arm) sees identical inputs.  All arrays are numpy FP64 on the host.
"""
from __future__ import annotations

import numpy as np

FAMILY_M52 = 0
FAMILY_M52_BASIS = 1


def _objective(X):
    z = (X - X.mean(0)) / (X.std(0) + 1e-12)
    return np.sin(z[:, 0]) + 0.5 * np.cos(1.3 * z[:, 1:].sum(1)) + 0.1 * (z ** 2).sum(1)


def gp_workload(config: int, n_walkers: int | None = None, seed_offset: int = 0):
    """Return dict(X, y, mean, family, theta, yerr2, D, N) for BASELINE config 1, 2, 4 or 5.

    X's last column holds the dataset-size basis input u = (1 - s)^2 exactly as the reference feeds
    it to the function model (fabolas.py:133).  theta is in george order / log space.
    """
    rng = np.random.default_rng(20260000 + config + 1000 * seed_offset)
    if config == 1:      # svm_mnist shape: 2 hyper-parameters + s, N = 100, 20 walkers
        N, d, B = 100, 2, 20
        Xh = rng.uniform(-10, 10, (N, d))
        s = rng.choice([1 / 128, 1 / 64, 1 / 32, 1 / 4], N)
        logM = lambda b: rng.uniform(0, 4, (b, d))
    else:                # cnn_cifar10 search space (cnn_cifar10.py:160-166) + s
        N = {2: 256, 4: 512, 5: 2048}[config]
        d, B = 5, {2: 128, 4: 128, 5: 1024}[config]
        Xh = np.column_stack([rng.uniform(4, 9, (N, 3)), rng.uniform(32, 512, N), rng.uniform(-6, -1, N)])
        s = rng.uniform(1 / 256, 1, N)

        def logM(b):
            m = rng.uniform(0, 4, (b, d))
            m[:, 3] = rng.uniform(6, 12, b)     # raw batch-size column needs a long length scale to stay PD
            return m
    B = n_walkers or B
    X = np.column_stack([Xh, (1.0 - s) ** 2])
    y = _objective(X) + 0.05 * rng.standard_normal(N)
    theta = np.column_stack([rng.uniform(-1, 1, B), logM(B), rng.uniform(-1, 1, B), rng.uniform(-1, 1, B)])
    yerr2 = rng.uniform(1e-3, 0.1, B)
    return dict(X=X, y=y, mean=float(y.mean()), family=FAMILY_M52_BASIS, theta=theta, yerr2=yerr2,
                N=N, D=d + 1, amp_mult=float(d + 1))


# algorithmic work per evaluation, flops
def flops_ll(N, d, basis=True):
    ck = 3 * d + 30 + (4 if basis else 0)
    return N ** 3 / 3 + 2 * N ** 2 + 0.5 * N * (N + 1) * ck


def flops_llg(N, d, P, basis=True):
    ck = 3 * d + 30 + (4 if basis else 0)
    return N ** 3 + 2 * N ** 2 + 0.5 * N * (N + 1) * (ck + 4 * (P - 1) + 2)
