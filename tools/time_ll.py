"""Time the fused GP kernel (config 2 by default) with CUDA events: python tools/time_ll.py [config] [--prof] [--ll]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fabolas_b200 import Engine
from fabolas_b200.workloads import gp_workload
args = [a for a in sys.argv[1:] if not a.startswith("--")]
cfg = int(args[0]) if args else 2
lib = os.path.join(ROOT, "fabolas_b200", "_lib", "libfabolas_b200_prof.so" if "--prof" in sys.argv else "libfabolas_b200.so")
for a in sys.argv:
    if a.startswith("--lib="):
        lib = os.path.join(ROOT, "fabolas_b200", "_lib", a[6:])
w = gp_workload(cfg)
eng = Engine(device="cuda:0", lib_path=lib)
eng.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])
th, ye = eng.tensor(w["theta"]), eng.tensor(w["yerr2"])
GRAD = "--ll" not in sys.argv
for _ in range(5):
    eng.loglik(th, ye, grad=GRAD)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.loglik(th, ye, grad=GRAD); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
print(f"config {cfg} {'grad' if GRAD else 'll'} {os.path.basename(lib)}: min {ts[0]:.1f} us, median {ts[len(ts)//2]:.1f} us, max {ts[-1]:.1f} us")
