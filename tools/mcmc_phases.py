"""Where a half-step of the device sampler goes (thread 0 of CTA 0, -DFAB_PHASE_PROF build):
  python tools/phase_prof.py --build --lib=libfabolas_b200_prof.so [-DFAB_DEV_D=2]    then
  python tools/mcmc_phases.py [config] [--lib=...]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from fabolas_b200 import Engine, fabolas
from fabolas_b200.workloads import gp_workload
LIB = os.path.join(ROOT, "fabolas_b200", "_lib", "libfabolas_b200_prof.so")
for a in sys.argv:
    if a.startswith("--lib="):
        LIB = os.path.join(ROOT, "fabolas_b200", "_lib", a[6:])
args = [a for a in sys.argv[1:] if not a.startswith("--")]
cfg = int(args[0]) if args else 1
K = {1: 20, 2: 128}[cfg]
w = gp_workload(cfg)
eng = Engine(device="cuda:0", lib_path=LIB)
lib = ctypes.CDLL(LIB)
np.random.seed(1)
fabolas.sample_hypers(w["X"], w["y"], K=K, engine=eng, on_device=True)
torch.cuda.synchronize()
lib.fab_debug_phase_reset()
np.random.seed(1)
fabolas.sample_hypers(w["X"], w["y"], K=K, engine=eng, on_device=True)
buf = (ctypes.c_longlong * 64)()
lib.fab_debug_phase_read(buf)
names = {30: "whole chain (CTA 0)", 31: "split ranking", 32: "proposal (thread 0) + barrier", 33: "prior (thread 0) + barrier",
         34: "likelihood (dag_eval) + barrier", 35: "accept / store", 36: "grid barrier", 40: "  dag bulk: whole", 51: "  dag bulk: end (wait chain)",
         55: "  dag chain: whole", 56: "  dag chain: wait for bulk", 57: "  dag chain: potrf64 + inverse", 59: "  dag chain: vectors"}
hs = 1200 + 1
print(f"config {cfg}, K = {K}: us per half-step at 1965 MHz (thread 0 of CTA 0, {hs} evaluations)")
for k in sorted(names):
    print(f"{k:3d} {names[k]:36s} {buf[k] / 1965.0 / hs:9.2f}")
