"""sample_hypers (100 burn-in + 500 steps, fabolas.py:69-117) end to end: the emcee-shaped host loop over the
batched GPU likelihood against the persistent device sampler.  Wall clock around the public call, synchronised."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from fabolas_b200 import Engine, fabolas
from fabolas_b200.workloads import gp_workload

for cfg, K in ((1, 20), (2, 128)):
    w = gp_workload(cfg)
    X, y = w["X"], w["y"]
    eng = Engine()
    for on_device in (False, True):
        for rep in range(2):                      # first repetition warms up (program upload, allocations)
            np.random.seed(1)
            torch.cuda.synchronize(); t = time.perf_counter()
            h = fabolas.sample_hypers(X, y, K=K, engine=eng, on_device=on_device)
            torch.cuda.synchronize(); dt = time.perf_counter() - t
        evals = K + 600 * K
        print("config %d (N=%d, %d walkers) %s: %.1f ms per sample_hypers -> %.0f log-prob evals/s, %.1f us per half-step"
              % (cfg, len(y), K, "device" if on_device else "host  ", dt * 1e3, evals / dt, dt / 1200 * 1e6), flush=True)
    eng.close()
