"""Pin the george layer the moment george is importable.

george 0.4.0 (/root/reference/requirements.txt:6) holds the GP arithmetic of the reference's hot path and is NOT
installable in the authoring container (no wheel, no sdist, no index), so oracle/george_oracle.py restates it from
its published behaviour: PARITY UNPINNED.  This script is the committed mechanism that pins it: wherever
`import george` works (any machine with the reference's requirements installed), run

    python tests/golden/make_george_golden.py            # writes tests/golden/george_ref.npz

and commit the file; tests/test_george_pin.py then compares the oracle (CPU tier) and the CUDA engine (GPU tier)
with george's own numbers, convention by convention.  When george is importable at test time the test generates the
vectors on the fly as well, so no commit is needed for the pin to become active.

What is recorded, and the reference line that depends on each convention:

  tree A (fabolas.py:81-97, 153-169, 227-243)   1 * Matern52(metric, axes 0..d-1) * (Linear(axis d) + Constant(axis d))
  tree B (entropy_search.py:32-36, 113-117)     1 * Matern52(metric, axes 0..D-1)
  (1) len(kernel), get_parameter_names(), get_parameter_vector() of both trees      -> fabolas.py:57, entropy_search.py:71
      (depth-first order [log_constant, log_M_0.., log_gamma2, log_constant]; `1 * k` -> log(1/ndim))
  (2) kernel.get_value(X) after set_parameter_vector(v)                             -> fabolas.py:57-62 (Q1: amplitude
      ndim * exp(v[0]); Constant/Linear summed over their axes; x x'/gamma^2)
  (3) GP.compute(X, yerr) + log_likelihood(y, quiet=True)                           -> fabolas.py:62-66, entropy_search.py:73-77
      (yerr^2 + 1.25e-12 on the diagonal)
  (4) grad_log_likelihood(y)                                                        -> SURVEY a6 (no reference call site)
  (5) predict(y, t, return_cov=True)                                                -> fabolas.py:187, acquisitions.py:31,103
  (6) kernel.get_value(t) + 1.25e-12 I = covariance of GP.sample(t)                 -> fabolas.py:199, entropy_search.py:137
on the reference's own data prior/prior_svm_mnist.csv (N = 100, D = 3; loaded as svm_mnist.load_prior does) when
/root/reference is present, else on the copy of it inside tests/golden/gp_oracle_svm_mnist.npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def _data():
    csv = "/root/reference/prior/prior_svm_mnist.csv"
    if os.path.exists(csv):
        import pandas as pd
        df = pd.read_csv(csv)
        Xh = np.column_stack([np.log10(df["C"].to_numpy()), np.log10(df["gamma"].to_numpy())])
        s = df["size"].to_numpy(dtype=float)
        return np.column_stack([Xh, (1 - s) ** 2]), df["validation_score"].to_numpy(dtype=float)
    g = np.load(os.path.join(HERE, "gp_oracle_svm_mnist.npz"))
    return g["X"], g["y"]


def trees(kernels, nd):
    """The two kernel expressions exactly as the reference writes them."""
    a = 1 * kernels.Matern52Kernel(metric=np.ones(nd - 1) * 0.1, ndim=nd, axes=list(range(nd - 1))) * (
        kernels.LinearKernel(log_gamma2=np.log(1), order=1, ndim=nd, axes=[nd - 1])
        + kernels.ConstantKernel(log_constant=np.log(1), ndim=nd, axes=[nd - 1]))            # fabolas.py:81-97
    b = 1 * kernels.Matern52Kernel(metric=np.ones(nd) * 0.1, ndim=nd, axes=list(range(nd)))   # entropy_search.py:32-36
    return a, b


def generate(george_module, kernels_module):
    """All pinned quantities from `george_module` (george itself; tests also pass a shim of the oracle to check that
    this script runs)."""
    X, y = _data()
    N, nd = X.shape
    rng = np.random.default_rng(41)
    out = {"X": X, "y": y}
    T = np.column_stack([rng.uniform(-10, 10, (9, nd - 1)), rng.uniform(0, 1, 9)])
    out["T"] = T
    for tag, k in zip("AB", trees(kernels_module, nd)):
        P = len(k)
        out[f"{tag}_len"] = np.array(P)
        names = getattr(k, "get_parameter_names", lambda: ())()
        out[f"{tag}_names"] = np.array([str(n) for n in names])
        out[f"{tag}_initial_vector"] = np.asarray(k.get_parameter_vector(), dtype=np.float64)
        V = np.column_stack([rng.uniform(-1, 1, 4), rng.uniform(0, 4, (4, P - 1))])
        if tag == "A":
            V[:, -2:] = rng.uniform(-1, 1, (4, 2))
        yerr = rng.uniform(0.03, 0.3, 4)
        out[f"{tag}_vectors"], out[f"{tag}_yerr"] = V, yerr
        Ks, lls, grads, mus, covs, Kts = [], [], [], [], [], []
        for v, ye in zip(V, yerr):
            gp = george_module.GP(kernel=k, mean=float(np.mean(y)))
            gp.kernel.set_parameter_vector(v)
            Ks.append(np.array(gp.kernel.get_value(X)))
            gp.compute(X, yerr=ye)
            lls.append(float(gp.log_likelihood(y, quiet=True)))
            grads.append(np.asarray(gp.grad_log_likelihood(y), dtype=np.float64))
            mu, cov = gp.predict(y, T, return_cov=True)
            mus.append(np.array(mu)); covs.append(np.array(cov))
            Kts.append(np.array(gp.kernel.get_value(T)))
        out[f"{tag}_K"], out[f"{tag}_ll"], out[f"{tag}_grad"] = np.array(Ks), np.array(lls), np.array(grads)
        out[f"{tag}_mu"], out[f"{tag}_cov"], out[f"{tag}_Kt"] = np.array(mus), np.array(covs), np.array(Kts)
    out["george_version"] = np.array(str(getattr(george_module, "__version__", "unknown")))
    return out


if __name__ == "__main__":
    try:
        import george
        from george import kernels
    except ImportError as e:
        sys.exit(f"george is not importable here ({e}); the george layer stays PARITY UNPINNED")
    res = generate(george, kernels)
    np.savez_compressed(os.path.join(HERE, "george_ref.npz"), **res)
    print("george_ref.npz written with george", res["george_version"])
