"""Latency of one batched log-likelihood (20 walkers) as a function of N: the cost of one 8-column step of the
in-tile Cholesky pipeline (chain_ops.cuh) is the slope between N = 8 ... 64.
    python tools/time_small_n.py [--grad]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from fabolas_b200 import Engine
GRAD = "--grad" in sys.argv
eng = Engine()
rng = np.random.default_rng(0)
for N in (8, 16, 24, 32, 40, 48, 56, 64, 72, 100, 128, 192, 256):
    X = np.column_stack([rng.uniform(-10, 10, (N, 2)), rng.uniform(0, 1, N)]); y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(N)
    B = 20
    theta = np.column_stack([rng.uniform(-1, 1, B), rng.uniform(0, 4, (B, 2)), rng.uniform(-1, 1, B), rng.uniform(-1, 1, B)])
    eng.set_data(1, X, y, float(y.mean()), 3)
    th, ye = eng.tensor(theta), eng.tensor(np.full(B, 0.01))
    for _ in range(5):
        eng.loglik(th, ye, grad=GRAD)
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.loglik(th, ye, grad=GRAD); b.record(); b.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    print(f"N={N:4d} {'ll+grad' if GRAD else 'll'}: median {np.median(ts):7.1f} us  min {np.min(ts):7.1f} us", flush=True)
