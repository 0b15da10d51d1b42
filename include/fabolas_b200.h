/* fabolas_b200.h -- C ABI of the B200-native FABOLAS GP inner-loop engine (libfabolas_b200.so).
 *
 * The reference (rickie95/fabolas) is pure Python and has no FFI: the seam this library plugs
 * into is the set of Python objects its hot path calls (george's GP/kernels, emcee's log-prob
 * callback, epmgp.joint_min, acquisitions.information_gain).  Each entry point below names the
 * reference interface it replaces (file:line in /root/reference); the ctypes binding a maintainer
 * would add is shown in INTEGRATION.md and implemented in fabolas_b200/_capi.py.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to FP64 (double) / int32 data unless stated otherwise;
 *     the caller (PyTorch in this repo) owns all input and output buffers;
 *   - all work is enqueued on the stream given at fab_ctx_create; no hidden synchronisation;
 *   - functions return 0 on success or a negative FAB_E* code and never throw or abort;
 *     numerical failure of one item is DATA (info[b] > 0, ll[b] = -inf), mirroring the reference's
 *     `try: gp.compute(...) except: return -np.inf` (fabolas.py:58-64, entropy_search.py:72-75);
 *   - a fab_ctx owns only scratch workspaces; it is bound to one device and is not thread-safe.
 *
 * Kernel families (george parameter order, log space)
 *   FAB_FAMILY_M52       theta = [log_c0, log_M_0 .. log_M_{d-1}]                       d = D
 *       k = amp * m52(r2),  amp = amp_mult * exp(log_c0),  r2 = sum_k (x_k - x'_k)^2 / M_k
 *       (entropy_search.py:32-36, 113-117)
 *   FAB_FAMILY_M52_BASIS theta = [log_c0, log_M_0 .. log_M_{d-1}, log_gamma2, log_cb]   d = D - 1
 *       k = amp * m52(r2) * (u u' / gamma2 + cb),  u = column d of X
 *       (fabolas.py:81-97, 153-169, 227-243)
 */
#ifndef FABOLAS_B200_H
#define FABOLAS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fab_ctx fab_ctx;

enum {
    FAB_OK = 0,
    FAB_EINVAL = -1,       /* bad argument (shape, family, null pointer) */
    FAB_ENOMEM = -2,       /* workspace allocation failed */
    FAB_ECUDA = -3,        /* CUDA runtime error; see fab_last_error */
    FAB_ESTATE = -4,       /* call order violated (e.g. predict before factorize) */
    FAB_EUNSUPPORTED = -5  /* shape outside what this build handles */
};
enum { FAB_FAMILY_M52 = 0, FAB_FAMILY_M52_BASIS = 1 };

/* library / context ------------------------------------------------------------------------- */
int fab_version(void);
int fab_ctx_create(fab_ctx** out, int device, void* cuda_stream);
int fab_ctx_destroy(fab_ctx* ctx);
const char* fab_last_error(const fab_ctx* ctx);

/* Optional per-kernel timing for benchmarks: when enabled, every batched call records CUDA events
 * around each of its kernel launches on the context's stream.  fab_ctx_get_timing synchronises on
 * the last event and writes the durations (ms) of the kernels of the most recent call, in launch
 * order (loglik / fit / mcmc: [the one fused kernel]; joint_min: [ep, normalize]; information gain:
 * [predict, mix, entropy, reduce]); returns the number of intervals (<= 7) or a negative code.     */
int fab_ctx_set_timing(fab_ctx* ctx, int enable);
int fab_ctx_get_timing(fab_ctx* ctx, double* ms_out, int n);

/* Observations of one GP problem: X is N x D row-major, y has N entries, `mean` is the constant
 * mean george's GP(kernel, mean=np.mean(y)) subtracts (fabolas.py:99,171; entropy_search.py:37,118).
 * `amp_mult` is the number of axes george sums the amplitude ConstantKernel over (ndim).
 * The data are copied into the context.                                                        */
int fab_gp_set_data(fab_ctx* ctx, int family, const double* X, int N, int D, const double* y,
                    double mean, double amp_mult);

/* Batched marginal log-likelihood (+ gradient): replaces, for B hyper-parameter rows at once,
 *   gp.kernel.set_parameter_vector(theta); gp.compute(X, sqrt(yerr2)); gp.log_likelihood(y, quiet=True)
 * (fabolas.py:57-66, entropy_search.py:71-77) and george's grad_log_likelihood convention.
 *   theta  B x Pk   (Pk = d+1 or d+3)          yerr2  B   (added to the diagonal with 1.25e-12)
 *   ll     B        (-inf where not positive definite)
 *   grad   B x (Pk+1) or NULL: d ll / d theta_k, last column d ll / d yerr2
 *   info   B        0, or the 1-based index of the failing leading minor                        */
int fab_gp_loglik_batch(fab_ctx* ctx, const double* theta, const double* yerr2, int B, double* ll,
                        double* grad, int* info);

/* The same call with HOST buffers -- the reference's own calling pattern: fabolas.log_likelihood is handed numpy
 * arrays by emcee and returns floats (fabolas.py:16-66, 103-117).  Buffers the CUDA driver knows (cudaHostAlloc,
 * cudaHostRegister, torch pin_memory) are read and written by the kernel itself over the bus (zero-copy: one warp per
 * walker fetches the Pk + 1 inputs, the results are posted stores -- no separate copy engine transfers around a
 * 0.24 ms kernel); pageable memory is staged through buffers of the context.  Stream-ordered: the outputs are valid
 * once the context's stream has been synchronised (fab_ctx_synchronize, or the caller's own stream sync).            */
int fab_gp_loglik_batch_host(fab_ctx* ctx, const double* theta, const double* yerr2, int B, double* ll,
                             double* grad, int* info);
/* Blocks until everything enqueued on the context's stream has completed. */
int fab_ctx_synchronize(fab_ctx* ctx);

/* Cholesky factors of B models into the context workspace (tiled, swizzled; test / debug use).  */
int fab_gp_factorize_batch(fab_ctx* ctx, const double* theta, const double* yerr2, int B, double* ll,
                           int* info);

/* Fit B models and keep them RESIDENT for prediction / acquisition calls: factorise, invert the
 * factor (W = L^-1), alpha = K^-1 (y - mean), copy the hyper-parameters.  Replaces the
 * per-hyper-sample `GP(kernel, mean); regressor.compute(x, yerr)` of fabolas.py:171-176, 245-248 and
 * entropy_search.py:118-119.  Also returns ll / info like fab_gp_loglik_batch.                    */
int fab_gp_fit_batch(fab_ctx* ctx, const double* theta, const double* yerr2, int B, double* ll, int* info);

/* GP.predict for every resident model (fabolas.py:187; entropy_search.py:130; acquisitions.py:31,
 * 103,109): T is M x D row-major, shared by all models (per_model = 0) or B x M x D with each model
 * predicting at its own points (per_model = 1: the per-hyper-sample representer points of
 * fabolas.py:180-188).  mu, var: B x M (latent mean / variance, no noise term) or NULL;
 * cov: B x M x M or NULL (full predictive covariance, M <= 64: the reference uses M in {1,50,51}).
 * Any N: up to N = 512 the cross-covariance panel of a candidate block lives in shared memory whole; beyond, it is
 * generated a few tile rows at a time (acquisitions.py:31 predicts at whatever size the history has grown to). */
int fab_gp_predict_batch(fab_ctx* ctx, const double* T, int per_model, int M, double* mu, double* var,
                         double* cov);

/* GP.sample(t).max() of every resident model at its own points: the incumbent the reference hands to
 * expected_improvement is the maximum of ONE draw from the PRIOR N(mean, k(t,t) + 1.25e-12 I) at the representer
 * points (fabolas.py:198-199, entropy_search.py:137).  reps B x Z x D (Z <= 64); out_max B; sample B x Z or NULL (the
 * draw itself); info B or NULL (pivots of the covariance that were not positive and were dropped, normally 0).
 * The normals are Box-Muller pairs of the Philox4x32-10 stream keyed by (seed, model, pair): a pure function of the
 * seed, reproducible on the host (tests/test_prior_draw.py); parity with numpy's generator is distributional.   */
int fab_gp_prior_draw_batch(fab_ctx* ctx, const double* reps, int Z, unsigned long long seed, double* out_max,
                            double* sample, int* info);

/* acquisitions.expected_improvement (acquisitions.py:40-75) on B x M predictive moments, plus the
 * np.argmax of expected_improvement.get_candidate (expected_improvement.py:92-93) per model:
 * ei (B x M), best_val (B), best_idx (B, int64) may each be NULL.  y_max: B incumbents.           */
int fab_ei_batch(fab_ctx* ctx, const double* mu, const double* var, int B, int M, const double* y_max,
                 double xi, double* ei, double* best_val, long long* best_idx);

/* epmgp.joint_min(mu, Sigma, with_derivatives) for B (mu, Sigma) pairs at once (epmgp.py:52-129;
 * call sites fabolas.py:209, entropy_search.py:141):  mu B x Z, Sigma B x Z x Z  ->
 *   logP B x Z ; dMu B x Z x Z ; dSigma B x Z x Z(Z+1)/2 (row-major lower order) ; dMuMu B x Z x Z x Z
 * (pass all three derivative buffers, or NULL for with_derivatives=False).  Z <= 64.
 *   info B : 0 ok; 1 = EP produced a NaN variance (the reference raises Exception, epmgp.py:250-253);
 *            2 = Cholesky of I + R^T Sigma R failed after the jitter ladder (reference raises LinAlgError).
 *   sweeps B x Z or NULL: EP sweeps used per "k is the minimum" problem (diagnostic).             */
int fab_epmgp_joint_min_batch(fab_ctx* ctx, const double* mu, const double* Sigma, int B, int Z,
                              double* logP, double* dMu, double* dSigma, double* dMuMu, int* info,
                              int* sweeps);

/* Entropy-Search state for the resident models (the per-hyper-sample lists built by
 * fabolas.get_candidate, fabolas.py:149-221, and entropy_search.entropy_search, entropy_search.py:
 * 108-153): representer points reps (B x Z x D), their UNCLIPPED posterior covariance cov_reps
 * (B x Z x Z, what compute_innovations re-predicts, acquisitions.py:31), p_min = joint_min outputs
 * (logP, dMu, dSigma, dMuMu), U (B x Z, EI at the representers) and Omega (B x I x Z innovation
 * vectors).  Everything the information gain needs is reduced to O(B Z) coefficients inside the
 * context; the caller's buffers are not referenced after the call returns (stream order).         */
int fab_es_setup(fab_ctx* ctx, const double* reps, int Z, const double* cov_reps, const double* logP,
                 const double* dMu, const double* dSigma, const double* dMuMu, const double* U,
                 const double* Omega, int I);

/* acquisitions.information_gain (acquisitions.py:119-167) for M test points at once: ig (M),
 * nan_flag (M or NULL; 1 where the reference raises ArithmeticError), and optionally the mixture
 * predictive mean / variance of predict_testpoint_george (acquisitions.py:99-116), M each.        */
int fab_es_information_gain_batch(fab_ctx* ctx, const double* Xc, int M, double* ig, int* nan_flag,
                                  double* mix_mean, double* mix_var);

/* Materialised covariance matrices k(Xa, Xb) for B hyper-parameter rows (and, optionally, their
 * derivatives with respect to every theta_k): the kernel object surface of ard.py
 * (AutomaticRelevanceDetermination.__call__, ard.py:82-180: nu_code 0 = 2.5, 1 = 1.5, 2 = 0.5, 3 = inf, 4 = the
 * general Matern kernel of smoothness `nu` with the Bessel function K_nu evaluated on the device (ard.py:129-135; no
 * analytic gradient, as in the reference), on r2 = sum_k (x_k - x'_k)^2 / M_k) and george's kernel.get_value /
 * get_gradient.
 *   Xa Na x D, Xb Nb x D, theta B x Pk  ->  K B x Na x Nb ; dK B x Na x Nb x Pk or NULL.           */
int fab_kernel_matrix(fab_ctx* ctx, int family, int nu_code, double nu, const double* theta, int B, const double* Xa,
                      int Na, const double* Xb, int Nb, int D, double amp_mult, double* K, double* dK);

/* Multi-GPU likelihood step with the exchange fused into the kernel (walkers sharded across
 * the GPUs of one box, one process per GPU; the only exchange of the path is the all-gather of the per-walker
 * log-likelihood (+ gradient) after every step -- the reference itself is single-process, fabolas.py:103-117).
 * Instead of a collective after the kernel, the CTA that finishes a walker stores its result row straight into the
 * gather buffer of EVERY rank through NVLink / NVSwitch peer memory; the last CTA of a rank publishes the rank's step
 * number in its own arrival slot on every rank and then waits (inside the same kernel, bounded) until every rank's
 * slot on this rank has reached the step: ONE launch per step, complete gather buffer when the kernel completes.
 *   fab_peer_alloc    allocates this rank's buffers (two gather buffers by step parity, one arrival slot per source
 *                     rank, error flag) and writes the 64-byte CUDA IPC handle to handle_out (host); exchange the
 *                     handles with any transport
 *   fab_peer_connect  handles: world x 64 bytes (host, rank order); maps the peers' buffers
 *   fab_gp_loglik_exchange  = fab_gp_loglik_batch on this rank's B = rows_per_rank walkers, plus the peer stores and
 *                     the wait for every rank's rows of this step; theta / yerr2 may be pinned HOST memory (both or
 *                     neither): the kernel then fetches them over the bus itself, as fab_gp_loglik_batch_host does
 *   fab_peer_wait     enqueues nothing (kept for the call order of the first version of this interface): returns the
 *                     device pointer of this rank's gathered rows of the most recent exchanged step: rows x
 *                     row_doubles, row of walker b of rank r at (r * B + b): [grad (Pk+1; zero when grad was not
 *                     requested) | ll | info]; valid for stream-ordered work until the next-but-one exchanged step
 *   fab_peer_error    synchronises the stream; 1 if a wait gave up after ~2 s (a peer never delivered): the rows of
 *                     that step are then stale and the caller must not use them                                  */
int fab_peer_alloc(fab_ctx* ctx, int rank, int world, int rows_per_rank, void* handle_out);
int fab_peer_connect(fab_ctx* ctx, const void* handles);
int fab_gp_loglik_exchange(fab_ctx* ctx, const double* theta, const double* yerr2, int B, double* ll, double* grad,
                           int* info);
int fab_peer_wait(fab_ctx* ctx, const double** gathered, int* rows, int* row_doubles);
int fab_peer_error(fab_ctx* ctx);
int fab_peer_free(fab_ctx* ctx);

/* emcee's EnsembleSampler(K, P, log_prob).run_mcmc(p0, nsteps) with the stretch move (a = 2, red/blue halves)
 * as the reference drives it (fabolas.py:103-117, entropy_search.py:80-85), with the reference's own
 * log-probability -- prior 0: fabolas.log_likelihood (fabolas.py:16-66, goes with FAB_FAMILY_M52_BASIS),
 * prior 1: entropy_search's log_prob (entropy_search.py:40-77, FAB_FAMILY_M52) -- evaluated on the data of
 * fab_gp_set_data.  The whole chain runs as ONE persistent kernel: every half-step's K/2 proposals are drawn,
 * evaluated and accepted on the device; nothing returns to the host between steps.
 *   coords    K x P   walker positions, in/out (P = len(kernel) + 1: [cov, lamb..., noise])
 *   log_prob  K       in/out; computed from coords first when init_lp != 0 (emcee: state.log_prob is None)
 *   seed, iter0       Philox4x32-10 stream: a pure function of (seed, iter0 + step, walker); continue a chain
 *                     with iter0 = steps already taken
 *   chain     nsteps x K x P or NULL (emcee's get_chain()); chain_lp  nsteps x K or NULL (get_log_prob())
 *   naccepted K int64 or NULL: accepted moves, accumulated
 * K even, 2 <= K; any K runs (CTAs loop over proposals beyond the number of SMs).  MCMC is stochastic: parity with
 * emcee is distributional, and exact against a host replay of the same Philox stream (tests/test_mcmc.py).       */
int fab_mcmc_run(fab_ctx* ctx, int prior, double* coords, double* log_prob, int K, int P, int nsteps, int init_lp,
                 unsigned long long seed, long long iter0, double a, double* chain, double* chain_lp,
                 long long* naccepted);

/* The same chain with the walkers sharded over the GPUs of one box (one process per GPU; SURVEY.md section 8e,
 * BASELINE configs[4]).  Every rank holds a replica of the ensemble (coords K x P, log_prob K) in a CUDA-IPC exported
 * buffer: fab_mcmc_peer_alloc writes the 64-byte handle to handle_out (host), fab_mcmc_peer_connect takes all
 * world x 64 bytes in rank order.  fab_mcmc_peer_state copies the ensemble into (to_replica = 1: the SAME state on
 * every rank) or out of (0) this rank's replica.  fab_mcmc_run_sharded is fab_mcmc_run on the replicas (every rank
 * calls it with the same arguments): the proposals of a half-step are dealt round-robin to the ranks, the owner of
 * an accepted move stores the new row into every replica over NVLink / NVSwitch peer memory, and the per-half-step
 * barrier spans the box (flag stores after a system-scope fence; no collective, no host round trip).  The chain is
 * bit-identical to fab_mcmc_run with the same seed.  naccepted counts on the owning rank only.
 * fab_mcmc_peer_error synchronises: 1 if a barrier gave up after ~2 s (a peer never arrived).                   */
int fab_mcmc_peer_alloc(fab_ctx* ctx, int rank, int world, int K, int P, void* handle_out);
int fab_mcmc_peer_connect(fab_ctx* ctx, const void* handles);
int fab_mcmc_peer_state(fab_ctx* ctx, double* coords, double* log_prob, int to_replica);
int fab_mcmc_run_sharded(fab_ctx* ctx, int prior, int nsteps, int init_lp, unsigned long long seed, long long iter0,
                         double a, long long* naccepted);
int fab_mcmc_peer_error(fab_ctx* ctx);
int fab_mcmc_peer_free(fab_ctx* ctx);

/* The log-prior part of those two log-probabilities for K rows p (K x P), and the george parameter vector
 * (K x (P-1), may be NULL) and yerr^2 (K, may be NULL) their likelihood is evaluated at; -inf outside the
 * support (theta row then zero).  horseshoe.py:10-33 with scipy.special.exp1's algorithm, lognorm(s=1).       */
int fab_mcmc_log_prior(fab_ctx* ctx, int prior, const double* p, int K, int P, double* out, double* theta,
                       double* yerr2);

/* Debug / test knob: cap the sampler's grid (0 = default: one CTA per proposal up to the SM count).           */
int fab_ctx_debug_set_mcmc_ctas(fab_ctx* ctx, int n);

/* Debug / test knob: cap the number of shared-memory tile slots (4..6) so that small problems
 * exercise the L2/HBM tile-streaming path of the fused GP kernel.                                  */
int fab_ctx_debug_set_max_slots(fab_ctx* ctx, int n);
/* Debug / test knob: cap the tile rows of the cross-covariance panel the prediction kernel keeps in shared memory
 * (0 = whatever fits), so that small problems exercise the chunked-panel path of N > 512.              */
int fab_ctx_debug_set_predict_chunk(fab_ctx* ctx, int jc);
/* Debug / test knob: make the fused GP kernel stage 64-row coordinate slices per tile (its large-N
 * mode) even when all coordinates would fit in shared memory.                                      */
int fab_ctx_debug_force_slices(fab_ctx* ctx, int on);
/* Debug / test access to the host-side static scheduler of the fused GP kernel (no device needed):
 * builds the two-stream tile program for NT tile rows, `nslot` shared-memory slots and `mode`
 * (0 log-likelihood, 1 + factor tiles stored, 2 fit, 3 gradient) and runs the independent verifier
 * (slot contents at every access, every cross-stream hazard ordered by an explicit wait).
 * Returns 0 when the program is valid; otherwise -1 and the reason in msg.  stats (may be NULL)
 * receives {bulk ops, chain ops, tile loads, tile stores}.                                           */
int fab_debug_dag_verify(int NT, int nslot, int mode, char* msg, int msg_len, int* stats);

/* Debug / test access: copy the dense lower Cholesky factor of walker b (N x N row-major, device)
 * out of the tiled workspace left by the last fab_gp_loglik_batch(grad != NULL) / factorize call. */
int fab_gp_debug_get_factor(fab_ctx* ctx, int b, double* L_out);

#ifdef __cplusplus
}
#endif
#endif /* FABOLAS_B200_H */
