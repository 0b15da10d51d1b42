"""Device-resident ensemble sampler (fab_mcmc_run, one persistent kernel) -- SURVEY.md section 8(f) row 1.

* the device log-prior against the REFERENCE's own horseshoe / lognormal values (tests/golden/priors_ref.npz);
* a device chain against the host replay of the same Philox stream with the CPU oracle's log-probability
  (exact up to the 1e-9 likelihood tolerance; `-m "not gpu"` on the emulator with ONE CTA looping over the
  proposals, `-m gpu` with one CTA per proposal and the grid barrier, and with a capped grid);
* distributional agreement with the emcee-shaped host sampler driven by the batched GPU likelihood.
"""
import numpy as np
import pytest

from fabolas_b200._capi import FAB_PRIOR_ES, FAB_PRIOR_FABOLAS
from oracle import acq_oracle as ao
from tests.conftest import golden
from tests.mcmc_replay import replay


def _problem(N, D, family, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (N, D))
    if family == 1:
        X[:, -1] = rng.uniform(0, 1, N)
    y = np.sin(X.sum(1)) + 0.05 * rng.standard_normal(N)
    return X, y


def _check_prior(eng):
    g = golden("priors_ref.npz")
    xs = np.array([x for x in g["xs"] if 0 <= x < 20])
    hs = np.array([h for x, h in zip(g["xs"], g["horseshoe01"]) if 0 <= x < 20])
    covs, ln = g["covs"], g["lognorm"]
    rng = np.random.default_rng(3)
    for prior, nl in ((FAB_PRIOR_FABOLAS, 5), (FAB_PRIOR_ES, 3)):
        rows, want = [], []
        for i, (x, h) in enumerate(zip(xs, hs)):
            c, l = covs[i % len(covs)], ln[i % len(covs)]
            inside = 0 < c < (20 if prior == FAB_PRIOR_FABOLAS else 100)
            rows.append(np.concatenate([[c], rng.uniform(-8.9, 1.9, nl), [x]]))
            want.append(l + h if inside else -np.inf)
        rows, want = np.array(rows), np.array(want)
        out, theta, yerr2 = eng.log_prior(prior, rows, want_theta=True)
        out, theta, yerr2 = out.cpu().numpy(), theta.cpu().numpy(), yerr2.cpu().numpy()
        fin = np.isfinite(want)
        assert fin.sum() > 5
        assert np.array_equal(np.isfinite(out), fin) or np.all(np.isposinf(out[~fin & ~np.isneginf(want)]))
        assert np.max(np.abs(out[fin] - want[fin]) / np.maximum(1.0, np.abs(want[fin]))) < 1e-12
        # the decoded george vector (quirk Q1) where the walker is inside the support
        for r, o, t, ye in zip(rows, out, theta, yerr2):
            if np.isfinite(o):
                tr, nr = ao.fabolas_theta(r) if prior == FAB_PRIOR_FABOLAS else ao.es_theta(r)
                assert np.allclose(t, tr, rtol=1e-14, atol=0) and ye == nr
        # outside the box: every boundary of fabolas.py:27-45 / entropy_search.py:49-60
        base = np.concatenate([[1.0], np.zeros(nl), [0.01]])
        bad = []
        for k, v in ((0, 0.0), (0, -1.0), (0, 20.0 if prior == FAB_PRIOR_FABOLAS else 100.0), (1, -9.0), (nl, 2.0),
                     (nl + 1, -1e-9), (nl + 1, 20.0), (0, np.nan), (2, np.nan)):
            b = base.copy(); b[k] = v; bad.append(b)
        assert np.all(np.isneginf(eng.log_prior(prior, np.array(bad)).cpu().numpy()))
        ref = ao.fabolas_log_prior(base) if prior == FAB_PRIOR_FABOLAS else ao.es_log_prior(base)
        assert abs(float(eng.log_prior(prior, base[None, :]).cpu().numpy()[0]) - ref) < 1e-12 * abs(ref)


def test_device_prior_matches_reference_emulated(sim_engine):
    _check_prior(sim_engine)


@pytest.mark.gpu
def test_device_prior_matches_reference(gpu_engine):
    _check_prior(gpu_engine)


def _check_replay(eng, N, D, family, K, nsteps, seed, p0_seed=1):
    X, y = _problem(N, D, family, seed=N + D)
    prior = FAB_PRIOR_FABOLAS if family == 1 else FAB_PRIOR_ES
    eng.set_data(family, X, y, float(y.mean()), D)
    P = D + 2 + family
    p0 = np.random.default_rng(p0_seed).uniform(0, 1, (K, P))          # the reference starts from np.random.rand(K, ndim)
    coords, lp, chain, chain_lp, nacc = eng.mcmc_run(prior, p0.copy(), nsteps, seed=seed)
    coords, lp, chain, chain_lp, nacc = (t.cpu().numpy() for t in (coords, lp, chain, chain_lp, nacc))
    fn = lambda p: ao.sampler_log_prob(prior, p, family, X, y, y.mean(), D)
    rchain, rlp, rnacc, margins = replay(p0, fn, nsteps, seed)
    assert margins.min() > 1e-6, "borderline acceptance in the test chain: pick another seed"
    assert np.array_equal(nacc, rnacc)
    assert 0 < nacc.sum() < 2 * nsteps * K // 2
    assert np.allclose(chain, rchain, rtol=1e-9, atol=1e-12)
    fin = np.isfinite(rlp)
    assert np.array_equal(np.isfinite(chain_lp), fin)
    assert np.allclose(chain_lp[fin], rlp[fin], rtol=1e-9, atol=1e-9)
    assert np.array_equal(coords, chain[-1]) and np.array_equal(lp, chain_lp[-1])
    return chain


def test_chain_replay_emulated(sim_engine):
    """One CTA runs every proposal of every half-step (7 evaluations per step): state carried across evaluations."""
    _check_replay(sim_engine, N=24, D=3, family=1, K=6, nsteps=4, seed=11)
    _check_replay(sim_engine, N=18, D=2, family=0, K=4, nsteps=3, seed=12)


def test_chain_replay_streaming_emulated(sim_engine):
    """Tiles streamed through TMA (4 slots, N > 128): mbarrier parities must survive from one evaluation to the next."""
    sim_engine.debug_set_max_slots(4)
    try:
        _check_replay(sim_engine, N=150, D=3, family=1, K=4, nsteps=2, seed=13)
    finally:
        sim_engine.debug_set_max_slots(6)


def test_continuing_a_chain_emulated(sim_engine):
    """run_mcmc(state, n) after a burn-in: iter0 keeps the stream going, log_prob is carried."""
    X, y = _problem(20, 3, 1, seed=4)
    sim_engine.set_data(1, X, y, float(y.mean()), 3)
    p0 = np.random.default_rng(2).uniform(0, 1, (4, 6))
    _, _, whole, _, _ = sim_engine.mcmc_run(FAB_PRIOR_FABOLAS, p0.copy(), 4, seed=5)
    c, lp, _, _, _ = sim_engine.mcmc_run(FAB_PRIOR_FABOLAS, p0.copy(), 2, seed=5)
    _, _, rest, _, _ = sim_engine.mcmc_run(FAB_PRIOR_FABOLAS, c, 2, log_prob=lp, seed=5, iter0=2)
    assert np.array_equal(whole[2:].cpu().numpy(), rest.cpu().numpy())


def test_argument_errors_emulated(sim_engine):
    from fabolas_b200 import FabolasError
    X, y = _problem(10, 3, 1)
    sim_engine.set_data(1, X, y, 0.0, 3)
    with pytest.raises(FabolasError):
        sim_engine.mcmc_run(FAB_PRIOR_FABOLAS, np.zeros((5, 6)), 1)          # odd number of walkers
    with pytest.raises(FabolasError):
        sim_engine.mcmc_run(FAB_PRIOR_FABOLAS, np.zeros((4, 5)), 1)          # wrong walker dimension
    with pytest.raises(FabolasError):
        sim_engine.mcmc_run(FAB_PRIOR_ES, np.zeros((4, 6)), 1)               # prior / kernel family mismatch


@pytest.mark.gpu
def test_chain_replay_gpu(gpu_engine):
    """svm_mnist shape (BASELINE configs[0]: N = 100, 20 walkers): one CTA per proposal, grid barrier per half-step."""
    _check_replay(gpu_engine, N=100, D=3, family=1, K=20, nsteps=12, seed=21)
    _check_replay(gpu_engine, N=60, D=4, family=0, K=8, nsteps=10, seed=22)


@pytest.mark.gpu
def test_chain_replay_capped_grid_gpu(gpu_engine):
    """Fewer CTAs than proposals: each CTA loops; the chain must not depend on the grid."""
    X, y = _problem(100, 3, 1, seed=103)
    gpu_engine.set_data(1, X, y, float(y.mean()), 3)
    p0 = np.random.default_rng(1).uniform(0, 1, (20, 6))
    full = gpu_engine.mcmc_run(FAB_PRIOR_FABOLAS, p0.copy(), 12, seed=21)[2].cpu().numpy()
    gpu_engine.debug_set_mcmc_ctas(3)
    try:
        capped = gpu_engine.mcmc_run(FAB_PRIOR_FABOLAS, p0.copy(), 12, seed=21)[2].cpu().numpy()
    finally:
        gpu_engine.debug_set_mcmc_ctas(0)
    assert np.array_equal(full, capped)


@pytest.mark.gpu
def test_streaming_many_walkers_gpu(gpu_engine):
    """N = 300 (15 tiles, streamed) with more proposals (160) than SMs: parities and workspaces across evaluations."""
    X, y = _problem(300, 4, 1, seed=9)
    gpu_engine.set_data(1, X, y, float(y.mean()), 4)
    K = 320
    p0 = np.random.default_rng(5).uniform(0, 1, (K, 7))
    coords, lp, chain, chain_lp, nacc = gpu_engine.mcmc_run(FAB_PRIOR_FABOLAS, p0.copy(), 3, seed=3)
    coords, lp = coords.cpu().numpy(), lp.cpu().numpy()
    idx = np.random.default_rng(0).choice(K, 12, replace=False)
    ref = np.array([ao.sampler_log_prob(0, coords[i], 1, X, y, y.mean(), 4) for i in idx])
    assert np.allclose(lp[idx], ref, rtol=1e-9, atol=1e-9)
    assert 0 < int(nacc.sum()) < 3 * K


@pytest.mark.gpu
def test_distribution_matches_host_sampler_gpu(gpu_engine):
    """Same target, two samplers (host emcee-shaped stretch move with the batched GPU likelihood vs the persistent
    device kernel): posterior means of every coordinate and of the log-probability agree within Monte-Carlo error."""
    from fabolas_b200.fabolas import log_likelihood_batch
    from fabolas_b200.sampler import EnsembleSampler
    X, y = _problem(40, 3, 1, seed=8)
    K, P, burn, n = 64, 6, 800, 1500
    gpu_engine.set_data(1, X, y, float(y.mean()), 3)
    p0 = np.random.default_rng(4).uniform(0, 1, (K, P))
    host = EnsembleSampler(K, P, log_likelihood_batch, args=[gpu_engine], vectorize=True)
    host._random = np.random.RandomState(7)
    st = host.run_mcmc(p0, burn)
    host.reset()
    host.run_mcmc(st, n)
    hc, hl = host.get_chain().reshape(-1, P), host.get_log_prob().reshape(-1)
    gpu_engine.set_data(1, X, y, float(y.mean()), 3)
    c, lp, _, _, _ = gpu_engine.mcmc_run(FAB_PRIOR_FABOLAS, p0.copy(), burn, seed=99, store_chain=False)
    _, _, chain, chain_lp, nacc = gpu_engine.mcmc_run(FAB_PRIOR_FABOLAS, c, n, log_prob=lp, seed=99, iter0=burn)
    dc, dl = chain.cpu().numpy().reshape(-1, P), chain_lp.cpu().numpy().reshape(-1)
    acc_dev, acc_host = nacc.cpu().numpy().mean() / n, host.acceptance_fraction.mean()
    assert abs(acc_dev - acc_host) < 0.05
    # effective sample size is far below K * n (autocorrelated): compare in units of the posterior spread
    sd = hc.std(axis=0)
    assert np.all(np.abs(dc.mean(axis=0) - hc.mean(axis=0)) < 0.3 * sd), (dc.mean(axis=0), hc.mean(axis=0), sd)
    assert np.all(np.abs(dc.std(axis=0) / sd - 1.0) < 0.25)
    assert abs(dl.mean() - hl.mean()) < 0.3 * hl.std()


def test_sample_hypers_surface_emulated(sim_engine):
    """fabolas.sample_hypers / entropy_search.sample_hypers return the last positions (K, P), inside the prior box."""
    from fabolas_b200 import entropy_search, fabolas
    X, y = _problem(16, 3, 1, seed=2)
    for on_device in (True, False):
        np.random.seed(3)
        h = fabolas.sample_hypers(X, y, K=4, burn_in=1, iterations=2, engine=sim_engine, on_device=on_device)
        assert h.shape == (4, 6) and np.all(np.isfinite(h)) and np.all((h[:, 0] > 0) & (h[:, 0] < 20))
        h = entropy_search.sample_hypers(X, y, K=4, burn_in=1, iterations=2, engine=sim_engine, on_device=on_device)
        assert h.shape == (4, 5) and np.all((h[:, -1] >= 0) & (h[:, -1] < 20))


def _check_sharded_world1(eng):
    """The sharded entry points with a world of one: same chain as fab_mcmc_run, state carried in the replica."""
    X, y = _problem(22, 3, 1, seed=6)
    eng.set_data(1, X, y, float(y.mean()), 3)
    p0 = np.random.default_rng(3).uniform(0, 1, (6, 6))
    ref_c, ref_lp, _, _, ref_n = eng.mcmc_run(FAB_PRIOR_FABOLAS, p0.copy(), 5, seed=8, store_chain=False)
    eng.mcmc_peer_setup(6, 6)
    try:
        c, lp, n = eng.mcmc_run_sharded(FAB_PRIOR_FABOLAS, p0.copy(), 2, seed=8)
        c, lp, n = eng.mcmc_run_sharded(FAB_PRIOR_FABOLAS, c, 3, log_prob=lp, seed=8, iter0=2, naccepted=n)
        assert not eng.mcmc_peer_error()
        assert np.array_equal(c.cpu().numpy(), ref_c.cpu().numpy()) and np.array_equal(lp.cpu().numpy(), ref_lp.cpu().numpy())
        assert np.array_equal(n.cpu().numpy(), ref_n.cpu().numpy())
    finally:
        eng.mcmc_peer_free()


def test_sharded_sampler_world1_emulated(sim_engine):
    _check_sharded_world1(sim_engine)


@pytest.mark.gpu
def test_sharded_sampler_world1_gpu(gpu_engine):
    _check_sharded_world1(gpu_engine)
