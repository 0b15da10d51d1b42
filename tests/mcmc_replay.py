"""Host replay of the device-resident ensemble sampler (fabolas_b200/csrc/mcmc.cuh): the same Philox4x32-10
stream, split rule, stretch proposal and acceptance test in numpy, with the log-probability supplied by the
caller (the CPU oracle in the tests).  TEST INFRASTRUCTURE: restates the kernel's bookkeeping so that a device
chain can be checked step by step; the sampling law itself is emcee's stretch move (fabolas.py:103-117)."""
import numpy as np

M32 = 0xFFFFFFFF


def philox4x32_10(key, c0, c1, c2, c3):
    k0, k1 = key & M32, (key >> 32) & M32
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0, k1 = (k0 + 0x9E3779B9) & M32, (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


def uniform53(hi, lo):
    return (float((hi << 21) | (lo >> 11)) + 0.5) * 2.0 ** -53


def mcmc_rand(seed, it, purpose, idx):
    return philox4x32_10(seed, it & M32, (it >> 32) & M32, purpose, idx)


def replay(coords, log_prob_fn, nsteps, seed, iter0=0, a=2.0, lp=None):
    """Returns (chain (nsteps, K, P), chain_lp (nsteps, K), naccepted (K,), margins): `margins` are the
    |lnpdiff - log u| of every decision, so that a test can tell a borderline flip from a logic error."""
    coords = np.array(coords, dtype=np.float64)
    K, P = coords.shape
    half = K // 2
    lp = np.array([log_prob_fn(c) for c in coords]) if lp is None else np.array(lp, dtype=np.float64)
    chain, chain_lp, nacc, margins = [], [], np.zeros(K, dtype=np.int64), []
    for step in range(nsteps):
        it = iter0 + step
        keys = [(mcmc_rand(seed, it, 0, i)[0] << 32) | i for i in range(K)]
        order = np.argsort(np.array(keys, dtype=np.uint64), kind="stable")
        for split in (0, 1):
            new = []
            for m in range(half):
                s = order[split * half + m]
                r = mcmc_rand(seed, it, 1 + split, m)
                u = uniform53(r[0], r[1])
                zz = ((a - 1.0) * u + 1.0) * ((a - 1.0) * u + 1.0) / a
                c = order[(1 - split) * half + ((r[2] * half) >> 32)]
                q = coords[c] - (coords[c] - coords[s]) * zz
                new_lp = log_prob_fn(q)
                with np.errstate(invalid="ignore"):
                    lnpdiff = (P - 1.0) * np.log(zz) + new_lp - lp[s]
                r2 = mcmc_rand(seed, it, 3 + split, m)
                lu = np.log(uniform53(r2[0], r2[1]))
                margins.append(abs(lnpdiff - lu) if np.isfinite(lnpdiff) else np.inf)
                new.append((s, q, new_lp, bool(lnpdiff > lu)))
            for s, q, new_lp, acc in new:          # the active half never reads its own members: order-independent
                if acc:
                    coords[s], lp[s] = q, new_lp
                    nacc[s] += 1
        chain.append(coords.copy())
        chain_lp.append(lp.copy())
    return np.array(chain), np.array(chain_lp), nacc, np.array(margins)
