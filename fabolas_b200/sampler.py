"""emcee-shaped affine-invariant ensemble sampler (stretch move, a = 2, red/blue halves) whose
log-probability is evaluated for a whole half-ensemble per call -- the batch the CUDA engine
consumes.  Mirrors the emcee 3.1.2 surface the reference uses
(fabolas.py:103-117, entropy_search.py:80-85): EnsembleSampler(nwalkers, ndim, log_prob_fn, args),
run_mcmc(p0, n, rstate0=...), reset(), .chain.  Distributional parity only (MCMC is stochastic)."""
from __future__ import annotations

import numpy as np


class State:
    def __init__(self, coords, log_prob=None):
        self.coords = np.array(coords, dtype=np.float64, copy=True)
        self.log_prob = None if log_prob is None else np.array(log_prob, dtype=np.float64, copy=True)

    def __iter__(self):          # emcee's State unpacks as (coords, log_prob, random_state)
        return iter((self.coords, self.log_prob, None))


class EnsembleSampler:
    def __init__(self, nwalkers, ndim, log_prob_fn, args=None, kwargs=None, a=2.0, vectorize=False):
        if nwalkers < 2 or nwalkers % 2:
            raise ValueError("an even number of walkers >= 2 is required for the red/blue stretch move")
        self.nwalkers, self.ndim, self.a, self.vectorize = nwalkers, ndim, float(a), vectorize
        self.log_prob_fn, self.args, self.kwargs = log_prob_fn, list(args or []), dict(kwargs or {})
        self._random = np.random.RandomState()
        self.reset()

    def reset(self):
        self._chain = []
        self._log_prob = []
        self.naccepted = np.zeros(self.nwalkers)
        self.iteration = 0

    # ----------------------------------------------------------------------------------------
    def compute_log_prob(self, coords):
        coords = np.atleast_2d(coords)
        if self.vectorize:
            lp = np.asarray(self.log_prob_fn(coords, *self.args, **self.kwargs), dtype=np.float64).reshape(-1)
        else:
            lp = np.array([float(self.log_prob_fn(c, *self.args, **self.kwargs)) for c in coords])
        if np.any(np.isnan(lp)):
            raise ValueError("Probability function returned NaN")
        return lp

    def run_mcmc(self, initial_state, nsteps, rstate0=None, **kwargs):
        if rstate0 is not None:
            try:
                self._random.set_state(rstate0)
            except Exception:
                pass
        state = initial_state if isinstance(initial_state, State) else State(initial_state)
        if state.coords.shape != (self.nwalkers, self.ndim):
            raise ValueError("incompatible input dimensions")
        if state.log_prob is None:
            state.log_prob = self.compute_log_prob(state.coords)
        rng = self._random
        half = self.nwalkers // 2
        for _ in range(nsteps):
            inds = np.arange(self.nwalkers) % 2
            rng.shuffle(inds)
            for split in (0, 1):
                S = np.where(inds == split)[0]
                C = np.where(inds != split)[0]
                s, c = state.coords[S], state.coords[C]
                zz = ((self.a - 1.0) * rng.rand(half) + 1.0) ** 2 / self.a
                factors = (self.ndim - 1.0) * np.log(zz)
                rint = rng.randint(len(C), size=half)
                q = c[rint] - (c[rint] - s) * zz[:, None]
                new_lp = self.compute_log_prob(q)
                with np.errstate(invalid="ignore"):
                    lnpdiff = factors + new_lp - state.log_prob[S]
                accepted = lnpdiff > np.log(rng.rand(half))
                state.coords[S[accepted]] = q[accepted]
                state.log_prob[S[accepted]] = new_lp[accepted]
                self.naccepted[S[accepted]] += 1
            self._chain.append(state.coords.copy())
            self._log_prob.append(state.log_prob.copy())
            self.iteration += 1
        return State(state.coords, state.log_prob)

    # ----------------------------------------------------------------------------------------
    def get_chain(self):
        return np.array(self._chain).reshape(-1, self.nwalkers, self.ndim)

    def get_log_prob(self):
        return np.array(self._log_prob).reshape(-1, self.nwalkers)

    @property
    def chain(self):            # (nwalkers, nsteps, ndim) like emcee's deprecated property
        return np.swapaxes(self.get_chain(), 0, 1)

    @property
    def acceptance_fraction(self):
        return self.naccepted / max(self.iteration, 1)


class DeviceEnsembleSampler:
    """The same emcee surface with the WHOLE chain on the GPU: `fab_mcmc_run` runs every half-step's proposals,
    log-probabilities (the reference's own: prior 0 = fabolas.py:16-66, prior 1 = entropy_search.py:40-77, on
    the engine's data) and acceptance tests inside one persistent kernel.
    `log_prob_fn` is therefore not a callable but the prior id; the random stream is Philox keyed by `seed`
    (numpy's global state, which the reference seeds through rstate0, only picks the seed)."""

    def __init__(self, nwalkers, ndim, engine, prior, a=2.0, seed=None):
        if nwalkers < 2 or nwalkers % 2:
            raise ValueError("an even number of walkers >= 2 is required for the red/blue stretch move")
        self.nwalkers, self.ndim, self.engine, self.prior, self.a = nwalkers, ndim, engine, int(prior), float(a)
        self.seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self._steps_taken = 0           # position in the Philox stream; reset() keeps it (a new chain, fresh numbers)
        self.reset()

    def reset(self):
        self._chain, self._log_prob = None, None
        self._nacc = None
        self.iteration = 0

    def run_mcmc(self, initial_state, nsteps, rstate0=None, store=True, **kwargs):
        if isinstance(initial_state, State):
            coords, lp = initial_state.coords, initial_state.log_prob
        else:
            coords, lp = initial_state, None
        import torch
        coords = self.engine.tensor(coords.clone() if isinstance(coords, torch.Tensor)
                                    else np.array(coords, dtype=np.float64, copy=True))
        if tuple(coords.shape) != (self.nwalkers, self.ndim):
            raise ValueError("incompatible input dimensions")
        sharded = (getattr(self.engine, "mcmc_world", 0) > 1 and getattr(self.engine, "mcmc_K", 0) == self.nwalkers
                   and getattr(self.engine, "mcmc_P", 0) == self.ndim)
        if sharded:
            # walkers dealt to the GPUs of the box (Engine.mcmc_peer_setup was called on every rank; every rank passes
            # the same state and the same seed); the chain itself is not recorded in this mode
            if store:
                raise ValueError("the sharded device sampler does not record the chain: call run_mcmc(..., store=False)")
            coords, lp, nacc = self.engine.mcmc_run_sharded(self.prior, coords, nsteps, log_prob=lp, seed=self.seed,
                                                            iter0=self._steps_taken, a=self.a, naccepted=self._nacc)
            chain = chain_lp = None
            if self.engine.mcmc_peer_error():     # a box-wide barrier gave up: the replicas may have diverged
                from ._capi import FabolasError
                raise FabolasError("sharded device sampler: a rank did not reach a half-step barrier within the timeout")
        else:
            coords, lp, chain, chain_lp, nacc = self.engine.mcmc_run(
                self.prior, coords, nsteps, log_prob=lp, seed=self.seed, iter0=self._steps_taken, a=self.a,
                store_chain=store, naccepted=self._nacc)
        self._steps_taken += nsteps
        self.iteration += nsteps
        self._nacc = nacc
        if store:
            self._chain = chain if self._chain is None else torch.cat([self._chain, chain])
            self._log_prob = chain_lp if self._log_prob is None else torch.cat([self._log_prob, chain_lp])
        lp_host = lp.cpu().numpy()
        if np.any(np.isnan(lp_host)):
            raise ValueError("Probability function returned NaN")
        st = State(coords.cpu().numpy(), lp_host)
        st.device_coords, st.device_log_prob = coords, lp
        return st

    def get_chain(self):
        return np.empty((0, self.nwalkers, self.ndim)) if self._chain is None else self._chain.cpu().numpy()

    def get_log_prob(self):
        return np.empty((0, self.nwalkers)) if self._log_prob is None else self._log_prob.cpu().numpy()

    @property
    def chain(self):
        return np.swapaxes(self.get_chain(), 0, 1)

    @property
    def naccepted(self):
        return np.zeros(self.nwalkers) if self._nacc is None else self._nacc.cpu().numpy().astype(np.float64)

    @property
    def acceptance_fraction(self):
        return self.naccepted / max(self.iteration, 1)
