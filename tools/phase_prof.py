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
for _a in sys.argv:
    if _a.startswith("--lib="):
        LIB = os.path.join(ROOT, "fabolas_b200", "_lib", _a[6:])
if "--build" in sys.argv:
    src = os.path.join(ROOT, "fabolas_b200", "csrc", "api.cu")
    cmd = ["nvcc", "-shared", "-Xcompiler", "-fPIC", "-O3", "-std=c++17", "-lineinfo", "-DFAB_PHASE_PROF",
           ] + [a for a in sys.argv if a.startswith("-D")] + [
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
        a, b, c, x = tr[4 * i], tr[4 * i + 1], tr[4 * i + 2], tr[4 * i + 3]
        extra = f"  mark {us(x)-us(b):6.2f} / {us(c)-us(x):6.2f}" if x else ""
        print(f"B{i:3d} {names[prog[4*i]]:7s}({prog[4*i+1]},{prog[4*i+2]}) k={prog[4*i+3]}  {us(a):8.2f} {us(b):8.2f} {us(c):8.2f}  wait {us(b)-us(a):6.2f}  run {us(c)-us(b):6.2f}{extra}")
    print("chain ops: top, after wait, tiles published, vectors published [us]")
    for j in range(NT):
        o = 4 * 128 + 4 * j
        print(f"C{j}  {us(tr[o]):8.2f} {us(tr[o+1]):8.2f} {us(tr[o+2]):8.2f} {us(tr[o+3]):8.2f}")

if "--fine" in sys.argv:
    NBW = 16 if "16" in os.path.basename(LIB) else 8
    tw = (ctypes.c_longlong * (16 * 128 * 6))()
    lib.fab_debug_trace_w_read(tw)
    import numpy as np
    T = np.array(tw[:], dtype=np.int64).reshape(16, 128, 6)[:NBW, :n, :].astype(float)
    t0 = T[:, 0, 0].min()
    T = (T - t0) / mhz
    print("per bulk op over the bulk warps [us]: first/last arrival, barrier exit (last), chain wait done, TMA issued (warp 0), loads landed (last), body: shortest / longest, first / last end")
    tot = np.zeros(6)
    for i in range(n):
        a = T[:, i, :]
        body = a[:, 5] - a[:, 4]
        prev_end = T[:, i - 1, 5].max() if i else 0.0
        print(f"W{i:3d} {names[prog[4*i]]:7s}({prog[4*i+1]},{prog[4*i+2]}) k={prog[4*i+3]}  arrive {a[:,0].min():7.2f}..{a[:,0].max():7.2f}  bar {a[:,1].max():7.2f}  chain {a[:,2].max():7.2f}  tma {a[0,3]:7.2f}  landed {a[:,4].max():7.2f}  body {body.min():5.2f}/{body.max():5.2f}  end {a[:,5].min():7.2f}..{a[:,5].max():7.2f}")
        tot += [a[:, 0].max() - a[:, 0].min(), a[:, 1].max() - a[:, 0].max(), a[:, 2].max() - a[:, 1].max(), a[:, 4].max() - a[:, 2].max(), body.max(), a[:, 0].min() - prev_end]
    print("totals [us]: arrival spread %.1f, barrier %.1f, chain wait %.1f, TMA issue + landing %.1f, longest bodies %.1f, end -> next arrival %.1f" % tuple(tot))
