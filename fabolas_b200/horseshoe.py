"""Horseshoe prior on the noise (reference: horseshoe.py:10-33), scalar and vectorised over walkers.
Host-side on purpose: B scalars per MCMC step, E1 from scipy.special."""
import numpy as np
from scipy.special import exp1

LN_PI = 1.1447298858494001
LN_2 = 0.6931471805599453


class Horseshoe:
    def __init__(self, scale=1):
        self.scale = scale

    def logpdf(self, x):
        x = np.asarray(x, dtype=np.float64)
        with np.errstate(all="ignore"):
            t = (x / self.scale) ** 2 / 2
            e1 = exp1(t)
            val = t - (LN_2 + 3 * LN_PI + 2 * np.log(self.scale)) / 2 + np.log(e1)
        out = np.where(e1 > 0, val, -np.inf)       # E1 underflow -> -inf (horseshoe.py:30-33); x = 0 -> +inf
        return float(out) if out.ndim == 0 else out
