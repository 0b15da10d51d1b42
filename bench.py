#!/usr/bin/env python
"""bench.py -- headline benchmark: batched GP log-likelihood + gradient evaluations per second.

Workload (BASELINE.json configs[1]): CNN-CIFAR10 search space (5 hyper-parameters) + dataset size s,
N = 256 synthetic observations, 128 MCMC walkers per GPU, FP64.  A "step" is one batched evaluation
of ll and d ll / d theta for every walker (weak scaling: 128 walkers on every GPU; the per-walker
ll vector and gradients are exchanged with an NCCL all-gather when --gpus > 1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` = evaluations/s with inputs resident in HBM; `e2e` = the
same through the public Python API with HOST (pinned) buffers, host<->device copies inside the timed
region; `roofline` = the dominant kernel against the measured FP64 peak; `cpu_baseline` = the CPU
oracle (numpy/scipy restatement of george) timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GP loglik+grad evals/sec (N=256, 128 walkers/GPU)"
# dram__bytes_read.sum + dram__bytes_write.sum of one gp_dag_kernel launch from the committed ncu --set full
# capture (profiles/); None until a capture of the current kernel exists.
TRAFFIC_BYTES_PER_LAUNCH = 186880 + 1583360
TRAFFIC_SOURCE = ("profiles/r01aj_ncu_full_raw.csv: dram__bytes_read.sum 187 KB + dram__bytes_write.sum 1.58 MB per "
                  "launch of 128 evals (algorithmic bytes 128 x 14.4 KB = 1.84 MB; tiles stream through L2 only)")
UNIT = "evals/s"
CONFIG_ID = 2


# ------------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    cfg, lo, hi = args                      # evaluates walkers lo .. hi-1 (indices taken modulo the batch)
    from threadpoolctl import threadpool_limits
    from fabolas_b200.workloads import gp_workload
    from oracle import george_oracle as go
    w = gp_workload(cfg)
    B = w["theta"].shape[0]
    idx = np.arange(lo, hi) % B
    with threadpool_limits(1):
        t0 = time.perf_counter()
        go.loglik_batch(w["family"], w["X"], w["y"], w["mean"], w["theta"][idx], w["yerr2"][idx], grad=True,
                        amp_axes=int(w["amp_mult"]))
        return time.perf_counter() - t0


def cpu_oracle_rate(per_core: int, cores: int | None = None, pool=None):
    """evals/s of the numpy/scipy oracle with one single-threaded process per host core."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    from fabolas_b200.workloads import gp_workload
    B = gp_workload(CONFIG_ID)["theta"].shape[0]
    jobs = [(CONFIG_ID, (i * per_core) % B, (i * per_core) % B + per_core) for i in range(cores)]
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(cores)
        pool.map(_cpu_worker, [(CONFIG_ID, 0, 1)] * cores)      # import + BLAS warm-up, untimed
    t0 = time.perf_counter()
    pool.map(_cpu_worker, jobs)
    dt = time.perf_counter() - t0
    if own:
        pool.close(); pool.join()
    n = sum(hi - lo for _, lo, hi in jobs)
    return n / dt, cores, n, dt


def run_reference(args):
    """--impl reference: the reference path's CPU implementation (oracle port; george itself is not
    installable offline) with every host core, same metric/config, bounded sample per step."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_core = 8
    pool = mp.get_context("spawn").Pool(cores)
    pool.map(_cpu_worker, [(CONFIG_ID, 0, 1)] * cores)
    for _ in range(args.warmup):
        cpu_oracle_rate(per_core, cores, pool)
    t_total, n_total = 0.0, 0
    for _ in range(args.steps):
        _, _, n, dt = cpu_oracle_rate(per_core, cores, pool)
        t_total += dt; n_total += n
    pool.close(); pool.join()
    value = n_total / t_total
    sample = f"{per_core} walkers x {cores} single-threaded processes per step ({n_total} evals total)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: N=256, D=6 (5 hyper-parameters + s), P=9, 128 walkers",
                   "note": "george 0.4.0 cannot be installed offline; numpy/scipy oracle port of its GP path"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from fabolas_b200 import Engine
    from fabolas_b200.workloads import gp_workload, flops_ll, flops_llg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    Bg = 128                                   # walkers per GPU (weak scaling)
    w = gp_workload(CONFIG_ID, n_walkers=Bg * world)
    N, D = w["X"].shape
    d, P = D - 1, D + 3                        # P = kernel parameters + noise
    eng = Engine(f"cuda:{local}")
    eng.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])
    theta_h = torch.from_numpy(w["theta"][rank * Bg:(rank + 1) * Bg].copy()).pin_memory()
    yerr2_h = torch.from_numpy(w["yerr2"][rank * Bg:(rank + 1) * Bg].copy()).pin_memory()
    theta_d, yerr2_d = theta_h.cuda(), yerr2_h.cuda()
    # results of one rank packed as [grad (Bg x P) | ll (Bg)] so that ONE all-gather ships both
    packed = torch.empty(Bg * P + Bg, dtype=torch.float64, device="cuda")
    g_loc, ll_loc = packed[:Bg * P].view(Bg, P), packed[Bg * P:]
    info_loc = torch.empty(Bg, dtype=torch.int32, device="cuda")
    packed_all = torch.empty(world * (Bg * P + Bg), dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")   # 256 MB > 126 MB L2

    # N > 1: the exchange of the step (every rank ends up with every walker's ll + gradient) is fused into the kernel:
    # each CTA stores its walker's row into every rank's gather buffer over NVLink peer memory and a one-thread
    # kernel waits for all rows (fab_gp_loglik_exchange / fab_peer_wait).  --exchange nccl keeps the kernel +
    # NCCL all-gather pair; it is also the fallback when the peers cannot be mapped (CUDA IPC unavailable).
    exchange = "none"
    if world > 1:
        exchange = args.exchange
        if exchange == "peer":
            try:
                eng.peer_setup(Bg)
            except Exception as e:                       # noqa: BLE001 -- reported in the JSON line
                exchange = f"nccl (peer mapping failed: {str(e)[:80]})"
            ok_all = torch.tensor([1.0 if exchange == "peer" else 0.0], device="cuda")
            dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
            if ok_all.item() == 0.0 and exchange == "peer":
                exchange = "nccl (a peer could not map)"
        if exchange == "peer":                           # one-time check against the collective it replaces
            ll, g, info = eng.loglik(theta_d, yerr2_d, grad=True)
            mine = torch.cat([g, ll[:, None], info.double()[:, None]], 1).contiguous()
            ref = torch.empty(world * Bg, P + 2, dtype=torch.float64, device="cuda")
            dist.all_gather_into_tensor(ref, mine)
            eng.loglik_exchange(theta_d, yerr2_d, grad=True)
            rows = eng.peer_wait()
            torch.cuda.synchronize()
            assert torch.equal(rows, ref) and not eng.peer_error(), "fused exchange differs from the NCCL all-gather"

    def step_resident():
        if exchange == "peer":
            ll, g, info = eng.loglik_exchange(theta_d, yerr2_d, grad=True, out=(ll_loc, g_loc, info_loc))
            eng.peer_wait()
            return ll, g
        ll, g, info = eng.loglik(theta_d, yerr2_d, grad=True, out=(ll_loc, g_loc, info_loc))
        if world > 1:
            dist.all_gather_into_tensor(packed_all, packed)
        return ll, g

    # e2e: the walkers' parameters arrive in ONE pinned host block [theta (Bg x Pk) | yerr2 (Bg)] and the results leave
    # in one [grad (Bg x P) | ll (Bg)]: one H2D and one D2H copy per step around the public Engine.loglik call
    Pk = P - 1
    in_host = torch.cat([theta_h.reshape(-1), yerr2_h]).pin_memory()
    in_dev = torch.empty_like(in_host, device="cuda")
    th_view, ye_view = in_dev[:Bg * Pk].view(Bg, Pk), in_dev[Bg * Pk:]
    out_host = torch.empty(Bg * P + Bg, dtype=torch.float64).pin_memory()
    g_host, ll_host = out_host[:Bg * P].view(Bg, P), out_host[Bg * P:]

    def step_e2e():
        in_dev.copy_(in_host, non_blocking=True)
        eng.loglik(th_view, ye_view, grad=True, out=(ll_loc, g_loc, info_loc))
        out_host.copy_(packed, non_blocking=True)

    def timed(step, K, W):
        for _ in range(W):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for a, b in ev:
            flush.zero_()                      # evict L2 between timed iterations
            a.record(); step(); b.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall

    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total, wall = timed(step_resident, K, W)
    clocks = sampler.stop() if sampler else None
    ms_e2e, _ = timed(step_e2e, K, W)

    # per-kernel durations (CUDA events recorded inside the library on the launching stream)
    eng.set_timing(True)
    kt = []
    for _ in range(max(10, K // 2)):
        flush.zero_()
        eng.loglik(theta_d, yerr2_d, grad=True)
        kt.append(eng.get_timing())
    eng.set_timing(False)
    ms_kernel = float(np.mean(np.array(kt)[:, 0]))        # one fused kernel per step (gp_dag_kernel, mode GRAD)
    kt = []
    eng.set_timing(True)
    for _ in range(max(10, K // 2)):
        flush.zero_()
        eng.loglik(theta_d, yerr2_d, grad=False)
        kt.append(eng.get_timing())
    eng.set_timing(False)
    ms_ll_only = float(np.mean(np.array(kt)[:, 0]))       # the MCMC inner loop proper: log-likelihood without gradient

    # ---- second headline number: Entropy-Search acquisition evaluations per second (BASELINE configs[2]) ----
    # Z = 50 representers, I = 20 innovations, 2048 candidates per GPU x 128 hyper-samples at N = 256.
    # Setup (fit -> predict at representers -> EI -> EPMGP -> es_setup) is timed separately.
    Z, I_, M = 50, 20, 2048
    rng = np.random.default_rng(20260003 + rank)
    lo = np.array([4, 4, 4, 32, -6, 1 / 256]); hi = np.array([9, 9, 9, 512, -1, 1.0])
    reps = rng.uniform(lo, hi, (Bg, Z, D)); Omega = rng.standard_normal((Bg, I_, Z)); cand = rng.uniform(lo, hi, (M, D))
    cand_d = eng.tensor(cand)

    def es_setup():
        eng.fit(theta_d, yerr2_d)
        mu_r, _, cov_r = eng.predict(reps, want_var=True, want_cov=True)
        cov_c = torch.clamp(cov_r, min=np.finfo(np.float64).eps)                  # fabolas.py:192
        y_inc = mu_r.max(dim=1).values
        ei_r, _, _ = eng.expected_improvement(mu_r, torch.diagonal(cov_c, dim1=1, dim2=2).contiguous(), y_inc)
        U = torch.clamp(ei_r, min=np.finfo(np.float64).eps)
        pm = eng.joint_min(mu_r, cov_c)
        eng.es_setup(reps, cov_r, pm[:4], U, Omega)
        return pm[4]
    es_setup(); torch.cuda.synchronize()
    t0 = time.perf_counter(); ep_info = es_setup(); torch.cuda.synchronize()
    es_setup_ms = (time.perf_counter() - t0) * 1e3
    for _ in range(2):
        eng.information_gain(cand_d)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n_ig = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_ig):
        ig, flag = eng.information_gain(cand_d)
    e1.record(); torch.cuda.synchronize()
    t_ig = torch.tensor([e0.elapsed_time(e1) / n_ig], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ig, op=dist.ReduceOp.MAX)
    ms_ig = float(t_ig.item())
    es_out = {"value": world * M / (ms_ig * 1e-3), "unit": "information-gain evals/s (each: 128 hyper-samples x 20 innovations, Z=50)",
              "candidates_per_gpu": M, "ms_per_sweep": ms_ig, "setup_ms_fit_predict_ei_epmgp": es_setup_ms,
              "nan_flags": int(flag.sum()), "epmgp_info_max": int(ep_info.max())}
    eng.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])       # back to the ll+grad problem state

    # ---- the caller of the path: fabolas.sample_hypers' chain (fabolas.py:103-117: 100 burn-in + 500
    # steps, 128 walkers) as ONE persistent kernel -- prior, stretch proposals, likelihood and acceptance on the device
    from fabolas_b200._capi import FAB_PRIOR_FABOLAS
    p0 = eng.tensor(np.random.default_rng(20260009 + rank).uniform(0, 1, (Bg, P)))
    c_, lp_, _, _, _ = eng.mcmc_run(FAB_PRIOR_FABOLAS, p0.clone(), 20, seed=1, store_chain=False)     # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    c_, lp_, _, _, nacc_ = eng.mcmc_run(FAB_PRIOR_FABOLAS, p0.clone(), 600, seed=2, store_chain=False)
    e1.record(); torch.cuda.synchronize()
    ms_mcmc = e0.elapsed_time(e1)
    mcmc_out = {"value": (Bg + 600 * Bg) / (ms_mcmc * 1e-3), "unit": "log-prob evals/s inside one persistent sampler kernel",
                "walkers": Bg, "steps": 600, "ms_per_chain": ms_mcmc, "us_per_half_step": ms_mcmc / 1200 * 1e3,
                "acceptance_fraction": float(nacc_.double().mean().item() / 600),
                "finite_log_prob": int(torch.isfinite(lp_).sum().item())}
    eng.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])

    # parity spot check of the timed path against the oracle (2 walkers; cheap)
    ll, g = step_resident()
    torch.cuda.synchronize()
    if rank == 0:
        from oracle import george_oracle as go
        lo = rank * Bg
        ref_ll, ref_g = go.loglik_batch(w["family"], w["X"], w["y"], w["mean"], w["theta"][lo:lo + 2],
                                        w["yerr2"][lo:lo + 2], grad=True, amp_axes=int(w["amp_mult"]))
        err_ll = float(np.max(np.abs(ll[:2].cpu().numpy() - ref_ll) / np.abs(ref_ll)))
        err_g = float(np.max(np.abs(g[:2].cpu().numpy() - ref_g) / np.abs(ref_g).max(axis=1, keepdims=True)))
        assert err_ll < 1e-9 and err_g < 1e-9, (err_ll, err_g)

    if rank == 0:
        # FP64 peak: DMMA/DFMA pipe peak measured on this pool (profiles/r01_fp64_peaks.json, 36.96 TF/s),
        # cross-checked live against cuBLAS DGEMM (library call used only as a denominator)
        n = 4096
        a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        (a @ b); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); (a @ b); (a @ b); e1.record(); torch.cuda.synchronize()
        dgemm_tf = 2 * 2 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
        peak = 36.96
        try:
            with open(os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")) as f:
                peak = max(v for k, v in json.load(f).items() if k.startswith("dmma_m8n8k4"))
        except Exception:
            pass
        f_ll, f_llg = flops_ll(N, d), flops_llg(N, d, P)
        kern_tf = Bg * f_llg / (ms_kernel * 1e-3) / 1e12
        ll_tf = Bg * f_ll / (ms_ll_only * 1e-3) / 1e12
        cpu_val, cores, n_evals, cpu_dt = (None, None, None, None)
        if world == 1 and not args.no_cpu:
            cpu_val, cores, n_evals, cpu_dt = cpu_oracle_rate(per_core=args.cpu_per_core)
        ms_step = ms_total / K
        out = {
            "metric": METRIC, "value": world * Bg / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: CNN-CIFAR10 space (5 hyper-parameters) + s, N=256, D=6, P=9, "
                                   "128 walkers per GPU, Matern52 x (linear+const) basis kernel",
                       "walkers_per_gpu": Bg, "N": N, "D": D, "l2_flush_between_steps": True,
                       "parallelism": f"walkers sharded x{world}; exchange of the per-walker (grad | ll | info) rows per step: "
                                      + {"none": "none (one GPU)", "peer": "fused into the kernel (NVLink peer stores into every "
                                         "rank's gather buffer + arrival counter; one-thread wait kernel)"}.get(
                                             exchange, "NCCL all-gather after the kernel [" + exchange + "]")},
            "e2e": {"value": world * Bg / (ms_e2e / K * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(in_host.numel() * 8),
                    "d2h_bytes_per_step": int(out_host.numel() * 8)},
            "gpu_launches": K * (2 if exchange == "peer" else 1),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "pipe": "FP64 tensor cores (DMMA mma.sync m8n8k4.f64; tcgen05 has no FP64 kind)",
                         "kernel": "gp_dag_kernel", "achieved": kern_tf, "peak": peak,
                         "unit": "TFLOP/s", "frac": kern_tf / peak, "traffic": TRAFFIC_BYTES_PER_LAUNCH,
                         "traffic_source": TRAFFIC_SOURCE,
                         "peak_source": "measured DMMA m8n8k4 peak on this pool (profiles/r01_fp64_peaks.json); "
                                        "MEASURED_PEAKS.json carries no FP64 figure; "
                                        f"live cuBLAS DGEMM 4096^3 = {dgemm_tf:.2f} TFLOP/s",
                         "flops_per_eval": f_llg, "evals_per_launch": Bg, "ms_per_launch": ms_kernel,
                         "ll_only": {"evals_per_s": Bg / (ms_ll_only * 1e-3), "ms_per_launch": ms_ll_only,
                                     "flops_per_eval": f_ll, "achieved": ll_tf, "frac": ll_tf / peak},
                         "whole_step_frac": Bg * f_llg / (ms_step * 1e-3) / 1e12 / peak},
            "es_acquisition": es_out,
            "mcmc_device_sampler": mcmc_out,
            "cpu_baseline": None if cpu_val is None else {
                "value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{n_evals} ll+grad evals of the same workload, one single-threaded numpy/scipy oracle "
                          f"process per core, {cpu_dt:.1f} s wall"},
            "wall_s_timed_region": wall,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how the per-walker results reach every rank (fused peer stores, or NCCL all-gather)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU-oracle baseline leg (profiling runs)")
    ap.add_argument("--cpu-per-core", type=int, default=640,
                    help="ll+grad evaluations per host core of the bounded CPU-oracle sample (~10-30 s of CPU work)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
