"""Per-phase cycle breakdown of the factor / gradient kernels (CTA 0, thread 0's view).
  build (CPU box):  python tools/phase_prof.py --build
  run   (GPU box):  python tools/phase_prof.py [config_id]
The instrumented library is a separate development artefact (-DFAB_PHASE_PROF); the shipped
library carries no counters."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "fabolas_b200", "_lib", "libfabolas_b200_prof.so")
NAMES = {0: "factor: whole kernel", 1: "  diag tile: alloc+stage+build", 2: "  diag tile: syrk update", 3: "  panel group0: reserve+stage",
         4: "  tile_potrf (total)", 5: "    potrf8 serial (thread 0)", 6: "    row solves + trailing (b,c)", 7: "  post-potrf (log, store)",
         8: "  panel: build (not shadowed)", 9: "  panel: gemm update", 10: "  panel: trsm", 11: "  panel: accw matvec", 12: "  panel: store",
         13: "  epilogue", 20: "grad: whole kernel", 21: "  prologue", 22: "  op overhead (barrier, loads, waits)", 23: "  INV_DIAG trsm",
         40: "dag bulk: whole", 41: "  op start (barrier, chain waits, TMA issue)", 42: "  TMA load waits", 43: "  BUILD", 44: "  MMA_NT", 45: "  PUT",
         46: "  TRSM (gemm form) + accw", 47: "  P1_MMA", 48: "  P1_FIN + alpha", 49: "  P2_MMA (incl. contraction)", 50: "    contraction", 51: "  end (wait chain, reduce)",
         55: "dag chain: whole", 56: "  wait for bulk", 57: "  potrf64", 58: "  inv64", 60: "    potrf8 (lane 0)", 61: "    8x8 inverse + row solves", 62: "    trailing update", 59: "  vectors + publish",
         24: "  matvec_t", 25: "  store", 26: "  P1 mma", 27: "  P1 fin (acc_store+trsm)", 28: "  P2 mma", 29: "  P2 contraction", 30: "  epilogue"}
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
