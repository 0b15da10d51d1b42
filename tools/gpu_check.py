"""Ad-hoc GPU check used during development: parity vs the oracle + CUDA-event timing."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
from fabolas_b200 import Engine
from oracle import george_oracle as go

def problem(N, D, family, B, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (N, D))
    if family == 1: X[:, -1] = rng.uniform(0, 1, N)
    y = np.sin(X.sum(1)) + 0.05 * rng.standard_normal(N)
    d = D - 1 if family == 1 else D
    Pk = d + 3 if family == 1 else d + 1
    theta = rng.uniform(-1, 1, (B, Pk)); theta[:, 1:1 + d] = rng.uniform(-1, 2, (B, d))
    yerr2 = rng.uniform(1e-3, 0.1, B)
    return X, y, theta, yerr2

eng = Engine()
for (N, D, fam, B) in [(40, 3, 1, 4), (100, 3, 1, 20), (256, 6, 1, 128), (200, 5, 0, 16), (512, 6, 1, 8)]:
    X, y, theta, yerr2 = problem(N, D, fam, B)
    eng.set_data(fam, X, y, float(y.mean()), D)
    grad = "--grad" in sys.argv
    out = eng.loglik(theta, yerr2, grad=grad)
    torch.cuda.synchronize()
    ll = out[0].cpu().numpy()
    nref = min(B, 8)
    ref = go.loglik_batch(fam, X, y, y.mean(), theta[:nref], yerr2[:nref], grad=grad, amp_axes=D)
    if grad:
        ref, gref = ref
        g = out[1].cpu().numpy()[:nref]
        print("  grad relerr", np.max(np.abs(g - gref) / (np.abs(gref).max(axis=1, keepdims=True))))
    print(N, D, fam, B, "ll relerr", np.max(np.abs(ll[:nref] - ref) / np.abs(ref)), "info", out[-1].cpu().numpy()[:4])
    th = eng.tensor(theta); ye = eng.tensor(yerr2)
    for _ in range(3): eng.loglik(th, ye, grad=grad)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps): eng.loglik(th, ye, grad=grad)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("   %.1f us per batch -> %.0f evals/s" % (ms * 1e3, B / ms * 1e3))
