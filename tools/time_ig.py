"""Per-kernel times of one information-gain sweep (config 3 shape) and of the EPMGP setup, CUDA events inside the library.
    python tools/time_ig.py            (FABOLAS_B200_LIB=<variant .so> to time a kernel variant)"""
import os, sys
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
eng.set_timing(True)
pm = eng.joint_min(mu, covc); ep = eng.get_timing()
eng.set_timing(False)
eng.es_setup(reps, cov, pm[:4], torch.clamp(ei, min=2.2e-16), Omega)
for _ in range(2):
    ig, flag = eng.information_gain(cand)
eng.set_timing(True)
ts = []
for _ in range(5):
    ig, flag = eng.information_gain(cand); ts.append(eng.get_timing())
ts = np.median(np.array(ts), axis=0)
print(os.environ.get("FABOLAS_B200_LIB", "default"), "predict/mix/entropy/reduce ms:", np.round(ts, 3), "epmgp ep/normalize ms:", np.round(ep, 3),
      "ig[0]", float(ig[0]), "flags", int(flag.sum()), flush=True)
