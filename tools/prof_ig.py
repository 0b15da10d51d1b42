"""Driver for ncu launch lists of the Entropy-Search pipeline (config 3 shape)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from fabolas_b200 import Engine
from fabolas_b200.workloads import gp_workload
B, Z, I_, M = 128, 50, 20, 2048
w = gp_workload(2)
eng = Engine(); eng.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])
rng = np.random.default_rng(3)
lo = np.array([4, 4, 4, 32, -6, 1 / 256]); hi = np.array([9, 9, 9, 512, -1, 1.0])
reps = rng.uniform(lo, hi, (B, Z, 6)); Omega = rng.standard_normal((B, I_, Z)); cand = eng.tensor(rng.uniform(lo, hi, (M, 6)))
eng.fit(w["theta"], w["yerr2"])
mu, _, cov = eng.predict(reps, want_var=True, want_cov=True)
covc = torch.clamp(cov, min=2.2e-16)
ei, _, _ = eng.expected_improvement(mu, torch.diagonal(covc, dim1=1, dim2=2).contiguous(), mu.max(dim=1).values)
pm = eng.joint_min(mu, covc)
eng.es_setup(reps, cov, pm[:4], torch.clamp(ei, min=2.2e-16), Omega)
for _ in range(2):
    ig, flag = eng.information_gain(cand)
torch.cuda.synchronize()
print("ok", float(ig[0]), int(flag.sum()))
