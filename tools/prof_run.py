"""Small driver for ncu captures: config-2 ll+grad batch, a few launches."""
import sys
import torch
sys.path.insert(0, ".")
from fabolas_b200 import Engine
from fabolas_b200.workloads import gp_workload
w = gp_workload(2)
eng = Engine()
eng.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])
th, ye = eng.tensor(w["theta"]), eng.tensor(w["yerr2"])
for _ in range(4):
    ll, g, info = eng.loglik(th, ye, grad=True)
torch.cuda.synchronize()
print("ok", float(ll[0]))
