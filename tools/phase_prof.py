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

if "--trace" in sys.argv:
    tr = (ctypes.c_longlong * (2 * 128 * 4))()
    lib.fab_debug_trace_read(tr)
    prog = (ctypes.c_int * (4 * 128))()
    NT = (w["X"].shape[0] + 63) // 64
    n = lib.fab_debug_program_read(NT, 6, 3 if GRAD else 0, prog, 128)
    names = ["NOP", "BUILD", "MMA_NT", "PUT", "TRSM", "P1_MMA", "P1_FIN", "P2_MMA", "CH", "GET"]
    t0 = tr[0]
    us = lambda c: (c - t0) / mhz
    print("bulk ops (CTA 0, last launch): index, op, arrive, start, end [us]; wait = start - arrive")
    for i in range(n):
        a, b, c = tr[4 * i], tr[4 * i + 1], tr[4 * i + 2]
        print(f"B{i:3d} {names[prog[4*i]]:7s}({prog[4*i+1]},{prog[4*i+2]}) k={prog[4*i+3]}  {us(a):8.2f} {us(b):8.2f} {us(c):8.2f}  wait {us(b)-us(a):6.2f}  run {us(c)-us(b):6.2f}")
    print("chain ops: top, after wait, tiles published, vectors published [us]")
    for j in range(NT):
        o = 4 * 128 + 4 * j
        print(f"C{j}  {us(tr[o]):8.2f} {us(tr[o+1]):8.2f} {us(tr[o+2]):8.2f} {us(tr[o+3]):8.2f}")
