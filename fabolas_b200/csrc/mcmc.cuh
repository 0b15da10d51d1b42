// mcmc.cuh -- device-resident affine-invariant ensemble sampler (stretch move, red/blue halves): the whole
// chain of emcee's EnsembleSampler.run_mcmc as ONE persistent cooperative kernel.
//
// Replaces the host loop of fabolas.sample_hypers (fabolas.py:103-117) and entropy_search.sample_hypers
// (entropy_search.py:80-85): 100 burn-in + 500 steps x 2 half-steps, each half-step evaluating the
// log-probability (fabolas.py:16-66 / entropy_search.py:40-77: box + lognormal + horseshoe prior, then the
// GP marginal likelihood) of the K/2 proposals.  SURVEY.md section 8(f) row 1.
//
// Mapping: one CTA per proposal of a half-step (gridDim.x = min(K/2, CTAs that are co-resident); a CTA loops
// when K/2 is larger).  A walker of the active half only needs the CURRENT positions of the other half, which
// do not change during the half-step, so every CTA draws its own proposal, evaluates it with dag_eval (the
// same code as gp_dag_kernel) and accepts or rejects it on its own; the only global synchronisation is one
// grid barrier per half-step (a release/acquire counter in global memory), after which the updated half is
// visible to everybody.  No host round trip, no launch gap, X / exp table / mbarriers set up once.
//
// Random numbers: Philox4x32-10 keyed by the caller's seed, counter = (iteration, purpose, index): the split
// of a step is the ranking of K 32-bit keys (walkers ranked < K/2 form the first half -- the same law as emcee's
// shuffle of arange(K) % 2), computed redundantly by every CTA.  The stream is a pure function of
// (seed, iteration, index), so tests/ replays a device chain step by step on the host with the oracle's
// log-probability; parity with emcee itself is distributional (MCMC is stochastic).
#pragma once
#include "gp_dag.cuh"

namespace fab {

enum { MCMC_PRIOR_FABOLAS = 0, MCMC_PRIOR_ES = 1 };
constexpr int MCMC_MAX_P = MAX_PK + 1;
constexpr int MCMC_STATIC_SMEM = 1024;   // static shared memory of the sampler kernel (a flag), rounded up generously

struct McmcArgs {
    DagArgs dag;                 // theta / yerr2 / ll / info: one scratch record per CTA (gridDim.x records)
    double* coords;              // K x P   walker positions (in/out)
    double* lp;                  // K       log-probabilities (in/out; recomputed first when init_lp)
    double* q;                   // gridDim.x x P   proposal scratch
    unsigned long long* keys;    // gridDim.x x K   split keys of the current step
    int* order;                  // gridDim.x x K   walkers by rank: [0, K/2) first half, [K/2, K) second half
    int K, P, prior, nsteps, init_lp;
    unsigned long long seed;
    long long iter0;             // iteration number of the first step (continuing a chain keeps the stream going)
    double a;                    // stretch scale (emcee default 2)
    double* chain;               // nsteps x K x P or null
    double* chain_lp;            // nsteps x K or null
    long long* nacc;             // K accepted-move counters (accumulated) or null
    unsigned* gbar;              // grid barrier counter, zero at launch
    // ---- walkers sharded over the GPUs of one box (one process per GPU; world <= 1: single GPU) ----
    // Every rank holds a replica of coords / lp inside a CUDA-IPC exported buffer; `coords` / `lp` above are this
    // rank's replica.  Proposal positions are dealt round-robin to the ranks; the owner of an accepted move stores
    // the new row into EVERY replica over NVLink peer memory, and the per-half-step barrier spans the box: local
    // grid barrier, one flag store per peer after a system-scope fence, wait for every peer's flag, release.
    double* peer_coords[FAB_MAX_PEERS];
    double* peer_lp[FAB_MAX_PEERS];
    unsigned* peer_flag[FAB_MAX_PEERS];   // flag block of rank r: generation reached by source rank q at [16 q]
    int* peer_err;               // local: set when a peer did not show up within ~2 s
    unsigned xgen0;              // cross-GPU barrier generations used by earlier launches
    int world, rank;
    int cluster;                 // 1: the whole grid is ONE thread-block cluster (<= 16 CTAs): hardware cluster barrier
};

struct Philox4 { unsigned x, y, z, w; };
__host__ __device__ inline Philox4 philox4x32_10(unsigned long long key, unsigned c0, unsigned c1, unsigned c2,
                                                 unsigned c3) {
    unsigned k0 = (unsigned)key, k1 = (unsigned)(key >> 32);
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1;
        const unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}
// 53 random bits -> (0, 1]
__host__ __device__ inline double uniform53(unsigned hi, unsigned lo) {
    const unsigned long long v = ((unsigned long long)hi << 21) | (lo >> 11);
    return ((double)v + 0.5) * 1.1102230246251565e-16;       // 2^-53
}
// purposes (counter word c2): 0 split key of walker `idx`, 1 + split: proposal of position `idx`,
// 3 + split: acceptance draw of position `idx`
__host__ __device__ inline Philox4 mcmc_rand(unsigned long long seed, long long iter, int purpose, int idx) {
    return philox4x32_10(seed, (unsigned)iter, (unsigned)((unsigned long long)iter >> 32), (unsigned)purpose, (unsigned)idx);
}

// Exponential integral E1(x), x >= 0: the algorithm of scipy.special.exp1 (specfun E1XB: power series for
// x <= 1, continued fraction of 20 + 80/x terms above), which horseshoe.py:25 calls.
__host__ __device__ inline double exp1_e1xb(double x) {
    if (x == 0.0) return INFINITY;
    if (x <= 1.0) {
        double e1 = 1.0, r = 1.0;
        for (int k = 1; k < 26; ++k) {
            r = -r * k * x / ((k + 1.0) * (k + 1.0));
            e1 += r;
            if (fabs(r) <= fabs(e1) * 1e-15) break;
        }
        return -0.5772156649015328 - log(x) + x * e1;
    }
    const int m = 20 + (int)(80.0 / x);
    double t0 = 0.0;
    for (int k = m; k > 0; --k) t0 = k / (1.0 + k / (x + t0));
    return exp(-x) * (1.0 / (x + t0));
}

// Horseshoe(scale).logpdf(x) as horseshoe.py:10-33 evaluates it (lower bound of the density; -inf once E1
// underflows) and scipy.stats.lognorm.logpdf(x, 1) (fabolas.py:47, entropy_search.py:62).
__host__ __device__ inline double horseshoe_logpdf(double x, double scale) {
    const double t = (x / scale) * (x / scale) / 2.0;
    const double e1 = exp1_e1xb(t);
    if (!(e1 > 0.0)) return -INFINITY;
    return t - (0.6931471805599453 + 3.0 * 1.1447298858494001 + 2.0 * log(scale)) / 2.0 + log(e1);
}
__host__ __device__ inline double lognorm1_logpdf(double x) {
    const double lx = log(x);
    return -log(x * 2.5066282746310002) - lx * lx / 2.0;
}

// Log-prior of one walker and the george parameter vector / yerr^2 its likelihood is evaluated at.
//   FABOLAS (fabolas.py:27-62):  p = [cov, lamb_0..lamb_{d+1}, noise];  0 < cov < 20, -9 < lamb < 2, 0 <= noise < 20;
//       theta = [cov, e^lamb_0..e^lamb_{d-1}, lamb_d, lamb_{d+1}]  (quirk Q1: raw values into log slots), yerr2 = noise
//   ES (entropy_search.py:49-75): p = [cov, lamb_0..lamb_{d-1}, noise];  0 < cov < 100, same boxes;
//       theta = [cov, lamb...], yerr2 = noise
// Returns -inf outside the support (theta is then left untouched and no likelihood is evaluated).
// Two halves, so that the sampler can evaluate the density (two logs and an exponential integral) next to the
// likelihood instead of in front of it: the box test + decode, and the density inside the box.
__host__ __device__ inline bool mcmc_prior_decode(int prior, const double* p, int P, int d, double* theta, double* yerr2) {
    const double cov = p[0], noise = p[P - 1];
    const double cov_hi = prior == MCMC_PRIOR_FABOLAS ? 20.0 : 100.0;
    bool ok = (cov > 0.0) && (cov < cov_hi) && (noise > -20.0) && (noise < 20.0) && !(noise < 0.0);
    for (int k = 1; k < P - 1; ++k) ok = ok && (p[k] > -9.0) && (p[k] < 2.0);
    if (!ok) return false;
    theta[0] = cov;
    for (int k = 0; k < d; ++k) theta[1 + k] = prior == MCMC_PRIOR_FABOLAS ? exp(p[1 + k]) : p[1 + k];
    if (prior == MCMC_PRIOR_FABOLAS) { theta[1 + d] = p[1 + d]; theta[2 + d] = p[2 + d]; }
    *yerr2 = noise;
    return true;
}
__host__ __device__ inline double mcmc_prior_density(const double* p, int P) {
    return lognorm1_logpdf(p[0]) + horseshoe_logpdf(p[P - 1], 0.1);
}
__host__ __device__ inline double mcmc_log_prior(int prior, const double* p, int P, int d, double* theta, double* yerr2) {
    if (!mcmc_prior_decode(prior, p, P, d, theta, yerr2)) return -INFINITY;
    return mcmc_prior_density(p, P);
}

// Grid-wide barrier number `gen` (1, 2, ...) of a cooperative launch: every CTA's thread 0 arrives with
// release semantics and spins with acquire loads; the other threads wait at the block barrier.
__device__ __forceinline__ void grid_barrier(unsigned* gbar, unsigned gen, int cluster = 0) {
#ifndef FAB_CUSIM
    if (cluster) {
        // the grid is one cluster (the reference's own ensemble, 20 walkers = 10 CTAs): the hardware barrier of the
        // cluster with release / acquire semantics replaces the counter in global memory and its polling round trips
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        return;
    }
#endif
    __syncthreads();
#ifndef FAB_CUSIM
    if (threadIdx.x == 0) {
        const unsigned target = gen * gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gbar) : "memory");
        unsigned v;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gbar) : "memory");
            if (v >= target) break;
            __nanosleep(64);
        }
    }
    __syncthreads();
#else
    (void)gbar; (void)gen;       // the emulator runs one CTA (api.cu launches the sampler with a single CTA there)
#endif
}
// Positions and log-probabilities are written by other CTAs between grid barriers: read them past L1.
__device__ __forceinline__ double ld_cg(const double* p) {
#ifdef FAB_CUSIM
    return *p;
#else
    return __ldcg(p);
#endif
}

// Barrier over every CTA of every rank.  Remote stores issued by any thread of this rank before the barrier are
// visible to every thread of every rank after it (block barrier -> gpu-scope release/acquire -> system-scope
// fence and flag -> acquire on the peer -> its grid barrier).
__device__ __forceinline__ void box_barrier(const McmcArgs& m, unsigned& gen, unsigned& xgen) {
    grid_barrier(m.gbar, ++gen, m.cluster);
    if (m.world > 1) {
        ++xgen;
#ifndef FAB_CUSIM
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            __threadfence_system();
            for (int r = 0; r < m.world; ++r)
                asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(m.peer_flag[r] + 16 * m.rank), "r"(xgen) : "memory");
            const long long t0 = clock64();
            for (int r = 0; r < m.world; ++r) {
                const unsigned* f = m.peer_flag[m.rank] + 16 * r;
                for (;;) {
                    unsigned v;
                    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                    if ((int)(v - xgen) >= 0) break;
                    if (clock64() - t0 > 4000000000LL) { *m.peer_err = 1; break; }
                    __nanosleep(64);
                }
            }
        }
#endif
        grid_barrier(m.gbar, ++gen);
    }
}

// Evaluate the log-probability of the P-vector in m.q (this CTA's record): thread 0 decodes the prior, the whole
// CTA runs the likelihood.  Result valid in thread 0.
// `extra`: more scalar work of thread 0 that can run next to the likelihood (the acceptance draw of the move).
template <int DD, class Extra>
__device__ __forceinline__ double mcmc_log_prob(const McmcArgs& m, unsigned char* smem_raw, unsigned& phases, bool& first,
                                                int* flag, Extra extra) {
    const int cta = blockIdx.x;
    double prior_lp = 0.0;
    PHB(33);
    if (threadIdx.x == 0) {
        // box test and decode only: the density is evaluated by this thread inside dag_eval, while the bulk stream waits
        // for the chain's last diagonal tile (NaN cannot occur: every term is finite inside the box, +inf at noise == 0)
        *flag = mcmc_prior_decode(m.prior, m.q + (size_t)cta * m.P, m.P, m.dag.gd.d,
                                  const_cast<double*>(m.dag.theta) + (size_t)cta * m.dag.gd.Pk,
                                  const_cast<double*>(m.dag.yerr2) + cta) ? 1 : 0;
    }
    __syncthreads();
    PHE(33);
    const bool run = *flag != 0;                  // uniform over the CTA
    PHB(34);
    if (run) {
        auto side = [&]() { prior_lp = mcmc_prior_density(m.q + (size_t)cta * m.P, m.P); extra(); };
        dag_eval<DD>(m.dag, cta, smem_raw, phases, first, side);
        first = false;
    }
    __syncthreads();
    PHE(34);
    if (threadIdx.x == 0) {
        if (!run) return -INFINITY;
        const double ll = m.dag.ll[cta];
        return (m.dag.info[cta] != 0) ? -INFINITY : prior_lp + ll;
    }
    return 0.0;
}

template <int DD>
__global__ void __launch_bounds__(DAG_THREADS, 1) mcmc_stretch_kernel(McmcArgs m) {
    FAB_DYN_SMEM(smem_raw);
    __shared__ int flag[4];
    const int tid = threadIdx.x, cta = blockIdx.x;
    const int K = m.K, P = m.P, half = K / 2;
    unsigned phases = 0u;
    bool first = true;
    unsigned gen = 0, xgen = m.xgen0;
    const int W = m.world > 1 ? m.world : 1, R = m.world > 1 ? m.rank : 0;
    double* q = m.q + (size_t)cta * P;
    unsigned long long* keys = m.keys + (size_t)cta * K;
    int* order = m.order + (size_t)cta * K;

    // every rank's kernel is running: whatever its stream did before (initial state, copies of the last chain) is done
    if (m.world > 1) box_barrier(m, gen, xgen);

    // One loop over "rounds" so that the evaluation (the bulk of the code) is instantiated ONCE: round -1 computes the
    // log-probability of the initial positions (emcee: state.log_prob = compute_log_prob(coords) when not given),
    // rounds 2 * step + split are the half-steps.
    PHB(30);
    for (int round = m.init_lp ? -1 : 0; round < 2 * m.nsteps; ++round) {
        const bool init = round < 0;
        const int step = init ? 0 : round >> 1, split = init ? 0 : round & 1;
        const long long iter = m.iter0 + step;
        PHB(31);
        if (!init && split == 0) {
            // ---- split of this step: rank the K keys; order[r] = walker of rank r ----
            for (int i = tid; i < K; i += DAG_THREADS)
                keys[i] = ((unsigned long long)mcmc_rand(m.seed, iter, 0, i).x << 32) | (unsigned)i;
            __syncthreads();
            for (int i = tid; i < K; i += DAG_THREADS) {
                const unsigned long long ki = keys[i];
                int r = 0;
                for (int j = 0; j < K; ++j) r += keys[j] < ki;
                order[r] = i;
            }
            __syncthreads();
        }
        PHE(31);
        const int items = init ? K : half;
        for (int mm = cta * W + R; mm < items; mm += gridDim.x * W) {
            PHB(32);
            const int s = init ? mm : order[split * half + mm];
            double zfac = 0.0;
            if (init) {
                if (tid < P) q[tid] = ld_cg(&m.coords[(size_t)s * P + tid]);
            } else if (tid == 0) {
                const Philox4 r = mcmc_rand(m.seed, iter, 1 + split, mm);
                const double u = uniform53(r.x, r.y);
                const double zz = ((m.a - 1.0) * u + 1.0) * ((m.a - 1.0) * u + 1.0) / m.a;
                int rint = (int)(((unsigned long long)r.z * (unsigned)half) >> 32);
                const int c = order[(1 - split) * half + rint];
                for (int k = 0; k < P; ++k) {
                    const double ck = ld_cg(&m.coords[(size_t)c * P + k]);
                    const double sk = ld_cg(&m.coords[(size_t)s * P + k]);
                    q[k] = ck - (ck - sk) * zz;
                }
                zfac = (P - 1.0) * log(zz);
            }
            __syncthreads();
            PHE(32);
            // (thread 0) the acceptance draw does not depend on the likelihood either
            double logu = 0.0;
            auto draw = [&]() {
                if (!init) { const Philox4 r = mcmc_rand(m.seed, iter, 3 + split, mm); logu = log(uniform53(r.x, r.y)); }
            };
            const double new_lp = mcmc_log_prob<DD>(m, smem_raw, phases, first, flag, draw);
            PHB(35);
            if (tid == 0 && init) {
                if (m.world > 1) { for (int r = 0; r < W; ++r) m.peer_lp[r][s] = new_lp; }
                else m.lp[s] = new_lp;
            } else if (tid == 0) {
                const double old_lp = ld_cg(&m.lp[s]);
                const double lnpdiff = zfac + new_lp - old_lp;          // NaN (-inf - -inf) compares false: rejected
                if (!(new_lp > -INFINITY)) draw();                       // (no likelihood was evaluated: draw now)
                const bool acc = lnpdiff > logu;
                if (acc) {
                    if (m.world > 1) {                                  // every replica, this rank's included
                        for (int r = 0; r < W; ++r) {
                            for (int k = 0; k < P; ++k) m.peer_coords[r][(size_t)s * P + k] = q[k];
                            m.peer_lp[r][s] = new_lp;
                        }
                    } else {
                        for (int k = 0; k < P; ++k) m.coords[(size_t)s * P + k] = q[k];
                        m.lp[s] = new_lp;
                    }
                    if (m.nacc) m.nacc[s] += 1;                         // counted on the owning rank
                }
                if (split == 1 && (m.chain || m.chain_lp)) {
                    // after the second half-step: record this walker and the first-half walker of the same rank
                    const int s0 = order[mm];
                    const size_t row = (size_t)step * K;
                    if (m.chain)
                        for (int k = 0; k < P; ++k) {
                            m.chain[(row + s) * P + k] = acc ? q[k] : ld_cg(&m.coords[(size_t)s * P + k]);
                            m.chain[(row + s0) * P + k] = ld_cg(&m.coords[(size_t)s0 * P + k]);
                        }
                    if (m.chain_lp) {
                        m.chain_lp[row + s] = acc ? new_lp : old_lp;
                        m.chain_lp[row + s0] = ld_cg(&m.lp[s0]);
                    }
                }
            }
            PHE(35);
        }
        PHB(36);
        box_barrier(m, gen, xgen);
        PHE(36);
    }
    PHE(30);
}

// Log-priors (and the theta / yerr2 they decode to) for K rows: the test hook that pins the device prior
// arithmetic against the reference's horseshoe.py / fabolas.py / entropy_search.py functions.
__global__ void mcmc_prior_kernel(const double* p, int K, int P, int prior, int d, int Pk, double* out, double* theta,
                                  double* yerr2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    double th[MAX_PK], ye = 0.0;
    for (int k = 0; k < MAX_PK; ++k) th[k] = 0.0;
    out[i] = mcmc_log_prior(prior, p + (size_t)i * P, P, d, th, &ye);
    if (theta) for (int k = 0; k < Pk; ++k) theta[(size_t)i * Pk + k] = th[k];
    if (yerr2) yerr2[i] = ye;
}

}  // namespace fab
