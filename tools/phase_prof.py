"""Per-phase cycle breakdown of the fused GP kernel (CTA 0; thread 0 = bulk stream, thread 256 = chain warp 0).
  build (CPU box):  python tools/phase_prof.py --build
  run   (GPU box):  python tools/phase_prof.py [config_id]
The instrumented library is a separate development artefact (-DFAB_PHASE_PROF); the shipped
library carries no counters."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "fabolas_b200", "_lib", "libfabolas_b200_prof.so")
NAMES = {40: "dag bulk: whole", 41: "  op start (barrier, chain waits, TMA issue)", 42: "  TMA load waits", 43: "  BUILD", 44: "  MMA_NT", 45: "  PUT",
         46: "  TRSM (gemm form) + accw", 47: "  P1_MMA", 48: "  P1_FIN + alpha", 49: "  P2_MMA (incl. contraction)", 50: "    contraction", 51: "  end (wait chain, reduce)",
         55: "dag chain: whole", 56: "  wait for bulk", 57: "  potrf64 + inverse (pipeline)", 59: "  vectors + publish",
         60: "    warp 0: potrf8 (one lane)", 61: "    warp 0: wait workers + next row block + diagonal update"}
if "--build" in sys.argv:
    src = os.path.join(ROOT, "fabolas_b200", "csrc", "api.cu")
    cmd = ["nvcc", "-shared", "-Xcompiler", "-fPIC", "-O3", "-std=c++17", "-lineinfo", "-DFAB_PHASE_PROF",
           "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, src]
    subprocess.check_call(cmd)
    print("built", LIB)
    sys.exit(0)
import torch
from fabolas_b200 import Engine
from fabolas_b200.workloads import gp_workload
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
GRAD = "--ll" not in sys.argv
w = gp_workload(cfg)
eng = Engine(device="cuda:0", lib_path=LIB)
lib = ctypes.CDLL(LIB)
eng.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])
th, ye = eng.tensor(w["theta"]), eng.tensor(w["yerr2"])
for _ in range(3):
    eng.loglik(th, ye, grad=GRAD)
torch.cuda.synchronize()
lib.fab_debug_phase_reset()
REP = 5
for _ in range(REP):
    eng.loglik(th, ye, grad=GRAD)
buf = (ctypes.c_longlong * 64)()
lib.fab_debug_phase_read(buf)
mhz = 1965.0
print(f"config {cfg}: cycles per launch (CTA 0), us at {mhz:.0f} MHz")
for k in sorted(NAMES):
    c = buf[k] / REP
    print(f"{k:3d} {NAMES[k]:40s} {c:12.0f} cyc {c / mhz:9.2f} us")
