"""Multi-GPU check of the fused likelihood exchange (run under torchrun on N GPUs of one box):
every rank's gathered rows must equal the NCCL all-gather of the plain per-rank results, over several steps
(alternating step parity), and the per-step time of both paths is printed.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/peer_check.py"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
from fabolas_b200 import Engine
from fabolas_b200.workloads import gp_workload

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
Bg = 128
w = gp_workload(2, n_walkers=Bg * world)
eng = Engine(f"cuda:{local}")
eng.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])
P = eng.Pk + 1
th = eng.tensor(w["theta"][rank * Bg:(rank + 1) * Bg]); ye = eng.tensor(w["yerr2"][rank * Bg:(rank + 1) * Bg])
eng.peer_setup(Bg)
ok = True
for step in range(6):
    ths = th + 0.001 * step
    ll, g, info = eng.loglik(ths, ye, grad=True)
    packed = torch.cat([g, ll[:, None], info.double()[:, None]], 1).contiguous()
    ref = torch.empty(world * Bg, P + 2, dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(ref, packed)
    eng.loglik_exchange(ths, ye, grad=True)
    rows = eng.peer_wait()
    torch.cuda.synchronize()
    same = bool(torch.equal(rows, ref)) and not eng.peer_error()
    ok = ok and same
    dist.barrier()
flag = torch.tensor([1.0 if ok else 0.0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)

def timed(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
packed = torch.empty(Bg * (P + 1), dtype=torch.float64, device="cuda")
g_loc, ll_loc = packed[:Bg * P].view(Bg, P), packed[Bg * P:]
info_loc = torch.empty(Bg, dtype=torch.int32, device="cuda")
allp = torch.empty(world * Bg * (P + 1), dtype=torch.float64, device="cuda")
def nccl_step():
    eng.loglik(th, ye, grad=True, out=(ll_loc, g_loc, info_loc)); dist.all_gather_into_tensor(allp, packed)
def peer_step():
    eng.loglik_exchange(th, ye, grad=True, out=(ll_loc, g_loc, info_loc)); eng.peer_wait()
def local_step():
    eng.loglik(th, ye, grad=True, out=(ll_loc, g_loc, info_loc))
t_local, t_nccl, t_peer = timed(local_step), timed(nccl_step), timed(peer_step)
if rank == 0:
    print(f"world {world}: fused exchange == NCCL all-gather on every rank: {bool(flag.item())}; per step: kernel only {t_local*1e3:.1f} us, "
          f"+ NCCL all-gather {t_nccl*1e3:.1f} us, fused peer stores + wait {t_peer*1e3:.1f} us; peer_error={eng.peer_error()}", flush=True)
eng.peer_free()
dist.destroy_process_group()
