"""Throughput and roofline fraction of the BASELINE.json configurations that bench.py does not headline
(configs[0], [3], [4]; [1] and [2] are bench.py's `value` and `es_acquisition`).  One JSON line per measurement.
CUDA events on the engine's stream, warm-up first, inputs resident; FP64 tensor-pipe peak = profiles/r01_fp64_peaks.json.
    python tools/bench_configs.py [--quick]
"""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from fabolas_b200 import Engine
from fabolas_b200.workloads import gp_workload, flops_ll, flops_llg

PEAK = 36.954          # TFLOP/s, DMMA m8n8k4 measured on this pool
quick = "--quick" in sys.argv


def timed(fn, warm, reps):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def emit(**kw):
    print(json.dumps(kw), flush=True)


def loglik_config(cfg, walkers, warm, reps):
    w = gp_workload(cfg, n_walkers=walkers)
    eng = Engine()
    eng.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])
    th, ye = eng.tensor(w["theta"]), eng.tensor(w["yerr2"])
    N, d, P = w["N"], w["D"] - 1, w["theta"].shape[1] + 1
    for grad in (False, True):
        ms, best = timed(lambda: eng.loglik(th, ye, grad=grad), warm, reps)
        out = eng.loglik(th, ye, grad=grad)
        F = flops_llg(N, d, P) if grad else flops_ll(N, d)
        tf = F * walkers / (ms * 1e-3) / 1e12
        emit(config=cfg, what="ll+grad" if grad else "ll", N=N, walkers=walkers, ms_per_launch=ms, ms_best=best,
             evals_per_s=walkers / (ms * 1e-3), flops_per_eval=F, tflops=tf, frac_fp64_tensor_peak=tf / PEAK,
             sms_busy=min(walkers, 148), frac_of_busy_sms=tf / (PEAK * min(walkers, 148) / 148),
             finite=int(torch.isfinite(out[0]).sum()))
    eng.close()


def predict_config(N_cfg, B, M, warm, reps):
    w = gp_workload(N_cfg, n_walkers=B)
    eng = Engine()
    eng.set_data(w["family"], w["X"], w["y"], w["mean"], w["amp_mult"])
    ll, info = eng.fit(w["theta"], w["yerr2"])
    rng = np.random.default_rng(7)
    lo, hi = w["X"].min(0), w["X"].max(0)
    T = eng.tensor(rng.uniform(lo, hi, (M, w["D"])))
    ymax = eng.tensor(np.full(B, float(w["y"].max())))
    N, d = w["N"], w["D"] - 1

    def sweep():
        mu, var = eng.predict(T, want_var=True)
        return eng.expected_improvement(mu, var, ymax, want_ei=False)

    ms, best = timed(sweep, warm, reps)
    ms_p, _ = timed(lambda: eng.predict(T, want_var=True), 1, reps)
    _, bv, bi = sweep()
    F = N * N + N * (3 * d + 38)
    tf = F * M * B / (ms_p * 1e-3) / 1e12
    emit(config=4, what="predict mean+var + EI argmax sweep", N=N, models=B, candidates=M, ms_per_sweep=ms,
         ms_predict_kernel=ms_p, candidate_model_pairs_per_s=M * B / (ms * 1e-3), flops_per_pair=F, tflops_predict=tf,
         frac_fp64_tensor_peak=tf / PEAK, best_idx0=int(bi[0]), best_val0=float(bv[0]), info_max=int(info.abs().max()))
    eng.close()


if __name__ == "__main__":
    loglik_config(1, 20, 5, 50)                                   # configs[0]: svm_mnist shape
    predict_config(4, 1, 1 << 20, 2, 5 if quick else 10)          # configs[3]: 1M candidates, one model (EI path)
    predict_config(4, 128, 1 << 14, 2, 5 if quick else 10)        #             128 hyper-samples x 16k candidates
    loglik_config(5, 128, 1, 2 if quick else 3)                   # configs[4]: N = 2048, 128 walkers (one GPU's share of 1024)
