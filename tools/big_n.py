import sys, time, torch
sys.path.insert(0,".")
from fabolas_b200 import Engine
from fabolas_b200.workloads import gp_workload
w=gp_workload(5, n_walkers=128)
eng=Engine(); eng.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])
th,ye=eng.tensor(w["theta"]),eng.tensor(w["yerr2"])
for _ in range(2): ll,info=eng.loglik(th,ye)
torch.cuda.synchronize(); t=time.time()
for _ in range(3): ll,info=eng.loglik(th,ye)
torch.cuda.synchronize(); dt=(time.time()-t)/3
print("N=2048 B=128: %.1f ms per batch -> %.0f evals/s, finite %d" % (dt*1e3, 128/dt, int(torch.isfinite(ll).sum())))
