"""Small driver for ncu captures of the prediction sweep (config 4 shape: N = 512, one model) and of the
persistent sampler kernel (config 2 shape, a few steps)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from fabolas_b200 import Engine
from fabolas_b200._capi import FAB_PRIOR_FABOLAS
from fabolas_b200.workloads import gp_workload
what = sys.argv[1] if len(sys.argv) > 1 else "predict"
if what == "predict":
    w = gp_workload(4, n_walkers=1)
    eng = Engine()
    eng.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])
    eng.fit(w["theta"], w["yerr2"])
    T = eng.tensor(np.random.default_rng(7).uniform(w["X"].min(0), w["X"].max(0), (148 * 32 * 8, w["D"])))
    for _ in range(3):
        mu, var = eng.predict(T, want_var=True)
    torch.cuda.synchronize()
    print("ok", float(mu[0, 0]))
else:
    w = gp_workload(2)
    eng = Engine()
    eng.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])
    p0 = np.random.default_rng(1).uniform(0, 1, (128, 9))
    for _ in range(2):
        c, lp, _, _, nacc = eng.mcmc_run(FAB_PRIOR_FABOLAS, p0.copy(), 10, seed=1, store_chain=False)
    torch.cuda.synchronize()
    print("ok", float(lp[0]))
