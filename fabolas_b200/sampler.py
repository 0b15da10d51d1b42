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
