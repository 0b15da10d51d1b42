"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own Python code
(/root/reference/{epmgp,acquisitions,horseshoe,ard}.py, imported unmodified; numpy-2 alias shim for
np.NaN / np.NAN / np.Infinity only) in the authoring container.

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so the outputs are committed as small .npz files.
george/emcee are not installable, hence the GP models handed to the reference's duck-typed
acquisition functions are oracle.george_oracle.GP objects, and gp_*.npz (log-likelihood, gradient,
prediction) is ORACLE-generated (parity unpinned against george; see oracle/__init__.py).
"""
import os
import sys
import warnings

import numpy as np

np.NaN = np.nan          # numpy-2 aliases the reference's epmgp.py still uses (epmgp.py:145,165,262,267)
np.NAN = np.nan
np.Infinity = np.inf
warnings.filterwarnings("ignore")

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import acquisitions as ref_acq          # noqa: E402
import ard as ref_ard                   # noqa: E402
import epmgp as ref_epmgp               # noqa: E402
import scipy.stats as sts               # noqa: E402
from horseshoe import Horseshoe         # noqa: E402

from oracle import george_oracle as go  # noqa: E402


def epmgp_cases():
    rng = np.random.default_rng(7)
    out = {}
    specs = [("z5", 5, 1.0), ("z12", 12, 1.0), ("z20_spread", 20, 4.0), ("z50", 50, 1.0), ("z8_tight", 8, 0.05)]
    for name, Z, spread in specs:
        A = rng.standard_normal((Z, Z))
        S = A @ A.T / Z + 0.05 * np.eye(Z)
        mu = spread * rng.standard_normal(Z)
        logP, dMu, dSg, dMM = ref_epmgp.joint_min(mu, S, with_derivatives=True)
        out[name + "_mu"], out[name + "_S"] = mu, S
        out[name + "_logP"], out[name + "_dMu"], out[name + "_dSg"], out[name + "_dMM"] = logP, dMu, dSg, dMM
    # posterior-like covariance from a GP (what the hot path really feeds it), clipped as fabolas.py:192
    X = rng.uniform(0, 1, (25, 2)); y = np.sin(3 * X[:, 0]) + X[:, 1]
    gp = go.GP(3.0 * go.Matern52Kernel([0.3, 0.5], ndim=2, axes=[0, 1]), mean=y.mean())
    gp.compute(X, 0.05)
    reps = rng.uniform(0, 1, (15, 2))
    mu, S = gp.predict(y, reps, return_cov=True)
    S = np.clip(S, np.finfo(S.dtype).eps, np.inf)
    logP, dMu, dSg, dMM = ref_epmgp.joint_min(mu, S, with_derivatives=True)
    out.update(gp15_mu=mu, gp15_S=S, gp15_logP=logP, gp15_dMu=dMu, gp15_dSg=dSg, gp15_dMM=dMM)
    out["names"] = np.array([s[0] for s in specs] + ["gp15"])
    np.savez_compressed(os.path.join(HERE, "epmgp_ref.npz"), **out)
    print("epmgp_ref.npz:", {k: v.shape for k, v in out.items() if k.endswith("logP")})


def acquisition_cases():
    rng = np.random.default_rng(11)
    N, d, B, Z, I = 30, 2, 3, 10, 4
    Xh = rng.uniform(-1, 1, (N, d)); s = rng.uniform(1 / 256, 1, N)
    Xraw = np.column_stack([Xh, s])
    Xtrain = np.column_stack([Xh, (1 - s) ** 2])                    # fabolas.py:133, 173-174
    y = np.sin(2 * Xh[:, 0]) * (0.5 + s) + 0.3 * Xh[:, 1]
    c = 1.0 + 2.0 * s + 0.1 * rng.standard_normal(N)
    bounds = [(-1, 1), (-1, 1), (1 / 256, 1)]
    theta = np.column_stack([rng.uniform(-1, 0.5, B), rng.uniform(-1, 1, (B, d)), rng.uniform(-1, 1, B),
                             rng.uniform(-1, 1, B)])
    yerr2 = rng.uniform(1e-3, 1e-2, B)
    ctheta = theta + 0.1
    models, cost_models, reps, pmin, U, Omega = [], [], [], [], [], []
    for b in range(B):
        m = go.GP(go.build_kernel(1, theta[b], d + 1), mean=y.mean()); m.compute(Xtrain, np.sqrt(yerr2[b]))
        cm = go.GP(go.build_kernel(1, ctheta[b], d + 1), mean=c.mean()); cm.compute(Xraw, np.sqrt(yerr2[b]))
        models.append(m); cost_models.append(cm)
        R = rng.uniform([lo for lo, _ in bounds], [hi for _, hi in bounds], (Z, d + 1))
        reps.append(R)
        mu, cov = m.predict(y, R, return_cov=True)
        cov = np.clip(cov, np.finfo(cov.dtype).eps, np.inf)         # fabolas.py:192
        ei = ref_acq.expected_improvement(mu, cov, m.sample(R, rng=rng))
        U.append(np.clip(ei, np.finfo(ei.dtype).eps, np.inf))       # fabolas.py:201-205
        pmin.append(ref_epmgp.joint_min(mu, cov, with_derivatives=True))
        Omega.append([rng.normal(size=Z).reshape(-1, 1) for _ in range(I)])
    dataset = {"y": y, "c": c}
    tests = rng.uniform([lo for lo, _ in bounds], [hi for _, hi in bounds], (6, d + 1))
    ig = np.array([ref_acq.information_gain(t, models, pmin, reps, U, Omega, dataset, n_innovations=I, enable_log=False)
                   for t in tests])
    igc = np.array([ref_acq.information_gain_cost(t, cost_models, models, dataset, pmin, reps, U, Omega, n_innovations=I)
                    for t in tests])
    pm = np.array([ref_acq.predict_testpoint_george(models, y, t) for t in tests])
    pc = np.array([ref_acq.predict_testpoint_george(cost_models, c, t) for t in tests])
    dmu, dsg = ref_acq.compute_innovations(tests[0].reshape(1, -1), y, models[0], reps[0], np.array([[pm[0, 1]]]))
    mu0, cov0 = models[0].predict(y, reps[0], return_cov=True)
    ei0 = ref_acq.expected_improvement(mu0, cov0, y)
    np.savez_compressed(
        os.path.join(HERE, "acq_ref.npz"), Xraw=Xraw, Xtrain=Xtrain, y=y, c=c, theta=theta, ctheta=ctheta, yerr2=yerr2,
        reps=np.array(reps), U=np.array(U), Omega=np.array(Omega)[..., 0], tests=tests,
        pmin_logP=np.array([p[0] for p in pmin]), pmin_dMu=np.array([p[1] for p in pmin]),
        pmin_dSg=np.array([p[2] for p in pmin]), pmin_dMM=np.array([p[3] for p in pmin]),
        ig=ig, igc=igc, pred_model=pm, pred_cost=pc, innov_dmu=dmu, innov_dsigma=dsg, ei_mu=mu0, ei_cov=cov0, ei=ei0)
    print("acq_ref.npz: ig", ig, "igc", igc)


def acquisition_case_full():
    """The reference's own sizes (fabolas.py:120: n_gen_samples=50, n_innovations=20) on its own data: the svm_mnist
    prior (N=100, D=3), B=8 hyper-samples, Z=50, I=20.  p_min (1.5 MB per model) is NOT stored: the CUDA path computes
    its own joint_min from the same (mu, cov) and is compared with the reference's information gain end to end, with
    the reference's logP and strided samples of its derivative tensors stored for the EPMGP leg."""
    import pandas as pd
    df = pd.read_csv("/root/reference/prior/prior_svm_mnist.csv")
    Xh = np.column_stack([np.log10(df["C"].to_numpy()), np.log10(df["gamma"].to_numpy())])
    s = df["size"].to_numpy(dtype=float)
    y = df["validation_score"].to_numpy(dtype=float)
    c = np.log10(df["time_s"].to_numpy(dtype=float))
    c = c - min(c.min(), 0.0)                                        # fabolas.py:270-273 (costs shifted to >= 0)
    N, d = Xh.shape
    B, Z, I = 8, 50, 20
    rng = np.random.default_rng(23)
    Xraw = np.column_stack([Xh, s])
    Xtrain = np.column_stack([Xh, (1 - s) ** 2])
    bounds = [(-10, 10), (-10, 10), (1 / 256, 1)]
    theta = np.column_stack([rng.uniform(-1, 0.5, B), rng.uniform(1, 4, (B, d)), rng.uniform(-1, 1, B),
                             rng.uniform(-1, 1, B)])
    yerr2 = rng.uniform(1e-3, 1e-2, B)
    ctheta = theta + 0.1
    models, cost_models, reps, pmin, U, Omega, mus, covs = [], [], [], [], [], [], [], []
    for b in range(B):
        m = go.GP(go.build_kernel(1, theta[b], d + 1), mean=y.mean()); m.compute(Xtrain, np.sqrt(yerr2[b]))
        cm = go.GP(go.build_kernel(1, ctheta[b], d + 1), mean=c.mean()); cm.compute(Xraw, np.sqrt(yerr2[b]))
        models.append(m); cost_models.append(cm)
        R = rng.uniform([lo for lo, _ in bounds], [hi for _, hi in bounds], (Z, d + 1))
        reps.append(R)
        mu, cov = m.predict(y, R, return_cov=True)
        cov = np.clip(cov, np.finfo(cov.dtype).eps, np.inf)         # fabolas.py:192
        mus.append(mu); covs.append(cov)
        ei = ref_acq.expected_improvement(mu, cov, m.sample(R, rng=rng))
        U.append(np.clip(ei, np.finfo(ei.dtype).eps, np.inf))       # fabolas.py:201-205
        pmin.append(ref_epmgp.joint_min(mu, cov, with_derivatives=True))
        Omega.append([rng.normal(size=Z).reshape(-1, 1) for _ in range(I)])
        print("  model", b, "joint_min done", flush=True)
    dataset = {"y": y, "c": c}
    tests = rng.uniform([lo for lo, _ in bounds], [hi for _, hi in bounds], (8, d + 1))
    tests[0] = Xraw[np.argmax(y)]                                    # the optimiser's starting point (fabolas.py:252-254)
    ig = np.array([ref_acq.information_gain(t, models, pmin, reps, U, Omega, dataset, n_innovations=I, enable_log=False)
                   for t in tests])
    igc = np.array([ref_acq.information_gain_cost(t, cost_models, models, dataset, pmin, reps, U, Omega, n_innovations=I)
                    for t in tests])
    pm = np.array([ref_acq.predict_testpoint_george(models, y, t) for t in tests])
    np.savez_compressed(
        os.path.join(HERE, "acq_ref_full.npz"), Xraw=Xraw, Xtrain=Xtrain, y=y, c=c, theta=theta, ctheta=ctheta,
        yerr2=yerr2, reps=np.array(reps), U=np.array(U), Omega=np.array(Omega)[..., 0], tests=tests,
        mu_reps=np.array(mus), cov_reps_clipped=np.array(covs),
        pmin_logP=np.array([p[0] for p in pmin]), pmin_dMu=np.array([p[1] for p in pmin]),
        pmin_dSg_s7=np.array([p[2][::7, ::7] for p in pmin]), pmin_dMM_s7=np.array([p[3][::7, ::7, ::7] for p in pmin]),
        ig=ig, igc=igc, pred_model=pm)
    print("acq_ref_full.npz: ig", ig, "igc", igc)


def prior_cases():
    xs = np.concatenate([np.linspace(-0.5, 4.5, 41), [1e-8, 1e-3, 3.85, 3.87]])
    hs = np.array([Horseshoe(scale=0.1).logpdf(x) for x in xs])
    hs1 = np.array([Horseshoe(scale=1).logpdf(x) for x in xs])
    cs = np.array([1e-3, 0.1, 0.5, 1.0, 2.0, 7.5, 19.9])
    ln = np.array([sts.lognorm.logpdf(cv, 1, 0) for cv in cs])
    np.savez_compressed(os.path.join(HERE, "priors_ref.npz"), xs=xs, horseshoe01=hs, horseshoe1=hs1, covs=cs, lognorm=ln)
    print("priors_ref.npz")


def ard_cases():
    rng = np.random.default_rng(3)
    out = {}
    for n in (2, 4, 6):
        X = rng.standard_normal((n, n)); Y = rng.standard_normal((n, n)); ls = rng.uniform(0.5, 2.0, n)
        for nu in (0.5, 1.5, 2.5, np.inf, 1.0, 2.0, 3.3):          # the last three: general branch (Bessel K_nu, ard.py:129-135)
            k = ref_ard.AutomaticRelevanceDetermination(length_scale=ls, nu=nu)
            tag = f"n{n}_nu{nu}"
            out[tag + "_Kxy"] = k(X, Y)
            if nu in (1.0, 2.0, 3.3):                              # the reference's gradient of this branch raises NameError
                out[tag + "_K"] = k(X)                             # (ard.py:173 calls _approx_fprime, never imported)
                continue
            K, G = k(X, eval_gradient=True)
            if nu == 0.5:                                          # ard.py:156-160 leaves the diagonal unset (0/0)
                G = G.copy(); G[np.arange(n), np.arange(n), :] = 0.0
            out[tag + "_K"], out[tag + "_G"] = K, G
        # a pair of coinciding points in k(X, Y): the strict-zero distance of the general branch (ard.py:131)
        Yd = Y.copy(); Yd[0] = X[0]
        out[f"n{n}_nu2.0_Kxy_dup"] = ref_ard.AutomaticRelevanceDetermination(length_scale=ls, nu=2.0)(X, Yd)
        out[f"n{n}_X"], out[f"n{n}_Y"], out[f"n{n}_ls"] = X, Y, ls
    np.savez_compressed(os.path.join(HERE, "ard_ref.npz"), **out)
    print("ard_ref.npz")


def gp_cases():
    """ORACLE-generated (unpinned vs george): real prior CSV of svm_mnist (config 1 shape)."""
    import pandas as pd
    df = pd.read_csv("/root/reference/prior/prior_svm_mnist.csv")
    # svm_mnist.load_prior(with_size=True): log10 C, log10 gamma ; size ; y ; c = log10 time  (svm_mnist.py:42-60)
    X = np.column_stack([np.log10(df["C"].to_numpy()), np.log10(df["gamma"].to_numpy())])
    s = df["size"].to_numpy(dtype=float)
    y = df["validation_score"].to_numpy(dtype=float)
    Xtrain = np.column_stack([X, (1 - s) ** 2])
    rng = np.random.default_rng(5)
    B = 6
    theta = np.column_stack([rng.uniform(-1, 1, B), rng.uniform(0, 4, (B, 2)), rng.uniform(-1, 1, B), rng.uniform(-1, 1, B)])
    yerr2 = rng.uniform(1e-3, 0.1, B)
    ll, g = go.loglik_batch(1, Xtrain, y, y.mean(), theta, yerr2, grad=True)
    T = np.column_stack([rng.uniform(-10, 10, (12, 2)), rng.uniform(1 / 256, 1, 12)])
    mus, covs = [], []
    for b in range(B):
        gp = go.GP(go.build_kernel(1, theta[b], 3), mean=y.mean()); gp.compute(Xtrain, np.sqrt(yerr2[b]))
        mu, cov = gp.predict(y, T, return_cov=True)
        mus.append(mu); covs.append(cov)
    np.savez_compressed(os.path.join(HERE, "gp_oracle_svm_mnist.npz"), X=Xtrain, y=y, theta=theta, yerr2=yerr2, ll=ll, grad=g,
                        T=T, mu=np.array(mus), cov=np.array(covs))
    print("gp_oracle_svm_mnist.npz: ll", ll)


if __name__ == "__main__":
    if "--ard-only" in sys.argv:
        ard_cases()
        sys.exit(0)
    if "--full-only" in sys.argv:          # the ~2 min full-size information-gain case alone
        acquisition_case_full()
        sys.exit(0)
    epmgp_cases()
    acquisition_cases()
    acquisition_case_full()
    prior_cases()
    ard_cases()
    gp_cases()
