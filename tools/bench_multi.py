"""BASELINE configs[3] and [4] on the N GPUs of one box (run under torchrun, one process per GPU):
  [4] N = 2048, 128 walkers per GPU (1024 on 8 GPUs): ll and ll+grad per batch with the exchange fused into the kernel;
  [3] 1M candidates at N = 512 sharded over the GPUs: predict + EI + argmax, one (value, index) all-gather.
One JSON line per measurement (rank 0).  CUDA events, max over ranks.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29516 tools/bench_multi.py"""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
from fabolas_b200 import Engine, dist as fd
from fabolas_b200.workloads import gp_workload, flops_ll, flops_llg

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
PEAK = 36.954


def timed(fn, warm, reps):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t))
    return float(np.median(ts))


# ---- configs[4]: N = 2048, 128 walkers per GPU ----
Bg = 128
w = gp_workload(5, n_walkers=Bg * world)
eng = Engine(f"cuda:{local}")
eng.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])
th, ye = eng.tensor(w["theta"][rank * Bg:(rank + 1) * Bg]), eng.tensor(w["yerr2"][rank * Bg:(rank + 1) * Bg])
eng.peer_setup(Bg)
N, d, P = w["N"], w["D"] - 1, w["theta"].shape[1] + 1
for grad in (False, True):
    def step():
        eng.loglik_exchange(th, ye, grad=grad); eng.peer_wait()
    ms = timed(step, 1, 3)
    rows = eng.peer_wait(); torch.cuda.synchronize()
    F = flops_llg(N, d, P) if grad else flops_ll(N, d)
    tf = F * Bg * world / (ms * 1e-3) / 1e12
    if rank == 0:
        print(json.dumps({"config": 5, "what": "ll+grad" if grad else "ll", "N": N, "n_gpus": world, "walkers": Bg * world,
                          "ms_per_batch": ms, "evals_per_s": Bg * world / (ms * 1e-3), "tflops_aggregate": tf,
                          "frac_fp64_tensor_peak": tf / (PEAK * world), "exchange": "fused into the kernel (NVLink peer stores)",
                          "finite_ll": int(torch.isfinite(rows[:, -2]).sum()), "peer_error": eng.peer_error()}), flush=True)
eng.peer_free(); eng.close()

# ---- configs[3]: 1M candidates at N = 512, candidates sharded, NCCL all-gather of one (value, index) pair per rank ----
w = gp_workload(4, n_walkers=1)
eng = Engine(f"cuda:{local}")
eng.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])
eng.fit(w["theta"], w["yerr2"])
M = 1 << 20
T = np.random.default_rng(7).uniform(w["X"].min(0), w["X"].max(0), (M, w["D"]))
lo, hi = fd.shard_range(M, rank, world)
Td = eng.tensor(T[lo:hi]); ymax = eng.tensor(np.array([float(w["y"].max())]))
best = {}
def sweep():
    mu, var = eng.predict(Td, want_var=True)
    _, bv, bi = eng.expected_improvement(mu, var, ymax, want_ei=False)
    pair = torch.stack([bv[0], (bi[0] + lo).double()])
    allp = torch.empty(world * 2, dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(allp, pair)
    best["pairs"] = allp
ms = timed(sweep, 2, 5)
pairs = best["pairs"].view(world, 2).cpu().numpy()
g = int(np.lexsort((pairs[:, 1], -pairs[:, 0]))[0])         # largest value, first index among ties (np.argmax)
F = w["N"] ** 2 + w["N"] * (3 * (w["D"] - 1) + 38)
if rank == 0:
    print(json.dumps({"config": 4, "what": "predict mean+var + EI + argmax over 1M candidates, sharded", "N": w["N"], "n_gpus": world,
                      "ms_per_sweep": ms, "candidates_per_s": M / (ms * 1e-3), "tflops_aggregate": F * M / (ms * 1e-3) / 1e12,
                      "frac_fp64_tensor_peak": F * M / (ms * 1e-3) / 1e12 / (PEAK * world), "argmax_index": int(pairs[g, 1]),
                      "argmax_value": float(pairs[g, 0])}), flush=True)
dist.destroy_process_group()
