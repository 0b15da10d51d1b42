"""A/B of two builds of the library on one GPU box: results of B against A (ll, gradient) and kernel times.
  python tools/ab_check.py libA.so libB.so [config ...]     (names relative to fabolas_b200/_lib)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from fabolas_b200 import Engine
from fabolas_b200.workloads import gp_workload
libs = [a for a in sys.argv[1:] if a.endswith(".so")]
cfgs = [int(a) for a in sys.argv[1:] if a.isdigit()] or [2]
WALKERS = {5: 128}      # N = 2048: one wave
def timeit(f, n=30):
    for _ in range(5): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]
for cfg in cfgs:
    w = gp_workload(cfg, n_walkers=WALKERS.get(cfg))
    ref = None
    for name in libs:
        eng = Engine(device="cuda:0", lib_path=os.path.join(ROOT, "fabolas_b200", "_lib", name))
        eng.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])
        th, ye = eng.tensor(w["theta"]), eng.tensor(w["yerr2"])
        out = eng.loglik(th, ye, grad=True)
        ll, g = out[0].cpu().numpy().copy(), out[1].cpu().numpy().copy()
        tg = timeit(lambda: eng.loglik(th, ye, grad=True), 30 if cfg != 5 else 5)
        tl = timeit(lambda: eng.loglik(th, ye, grad=False), 30 if cfg != 5 else 5)
        msg = f"config {cfg} {name}: ll+grad {tg:.1f} us, ll {tl:.1f} us"
        if ref is None:
            ref = (ll, g)
        else:
            dl = np.max(np.abs(ll - ref[0]) / np.maximum(1.0, np.abs(ref[0])))
            dg = np.max(np.abs(g - ref[1]) / np.maximum(1e-300, np.max(np.abs(ref[1]), axis=1, keepdims=True)))
            msg += f"   vs {libs[0]}: ll {dl:.2e}, grad {dg:.2e}"
        print(msg, flush=True)
        del eng
