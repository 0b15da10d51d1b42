"""Multi-GPU check of the sharded device sampler (run under torchrun on N GPUs of one box): the chain with the
walkers dealt to N ranks must be BIT-IDENTICAL on every rank to the single-GPU chain with the same seed; then the
time per half-step of both.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29515 tools/mcmc_peer_check.py [K]"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
from fabolas_b200 import Engine
from fabolas_b200._capi import FAB_PRIOR_FABOLAS
from fabolas_b200.workloads import gp_workload

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
w = gp_workload(2)
eng = Engine(f"cuda:{local}")
eng.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])
P = eng.Pk + 1
p0 = np.random.default_rng(11).uniform(0, 1, (K, P))
nsteps = 12
ref_c, ref_lp, _, _, ref_n = eng.mcmc_run(FAB_PRIOR_FABOLAS, p0.copy(), nsteps, seed=5, store_chain=False)
torch.cuda.synchronize()
eng.mcmc_peer_setup(K, P)
c, lp, n = eng.mcmc_run_sharded(FAB_PRIOR_FABOLAS, p0.copy(), 5, seed=5)
c, lp, n = eng.mcmc_run_sharded(FAB_PRIOR_FABOLAS, c, nsteps - 5, log_prob=lp, seed=5, iter0=5, naccepted=n)
torch.cuda.synchronize()
err = eng.mcmc_peer_error()
dist.all_reduce(n)                                   # accepted moves are counted on the owning rank
same = bool(torch.equal(c, ref_c) and torch.equal(lp, ref_lp) and torch.equal(n, ref_n)) and not err
flag = torch.tensor([1.0 if same else 0.0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)

def timed(fn, steps):
    fn(3); torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(steps); b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) / (2 * steps) * 1e3
steps = 60
t_single = timed(lambda s: eng.mcmc_run(FAB_PRIOR_FABOLAS, ref_c.clone(), s, log_prob=ref_lp.clone(), seed=6, store_chain=False), steps)
t_multi = timed(lambda s: eng.mcmc_run_sharded(FAB_PRIOR_FABOLAS, c, s, log_prob=lp, seed=6), steps)
if rank == 0:
    print(f"world {world}, {K} walkers, N=256: sharded chain bit-identical to the single-GPU chain on every rank: {bool(flag.item())} "
          f"(peer_error={err}); per half-step: one GPU {t_single:.1f} us, {world} GPUs {t_multi:.1f} us "
          f"({K // 2} proposals per half-step -> {(K // 2 + world - 1) // world} per GPU)", flush=True)
eng.mcmc_peer_free()
dist.destroy_process_group()
