// ig.cuh -- kernels E: Entropy-Search information gain for M candidates x B resident models
// (replaces acquisitions.information_gain / compute_innovations / predict_testpoint_george,
// acquisitions.py:7-37, 99-167, evaluated for a whole batch of test points without host round trips).
//
// Algebra used (exact rewrites of the reference's expressions; only summation order changes):
//   c = C[-1, :-1] (acquisitions.py:32) = [c0(x), ct_1 .. ct_{Z-1}],  ct_j = cov(rep_{Z-1}, rep_{j-1})
//       is x-independent; c0(x) = cov(rep_{Z-1}, x) = k(rep_{Z-1}, x) - vrl . V_x ,  vrl = W k(X, rep_{Z-1})
//   d_mu = c s ,  s = sqrt(v + 1e-4) / v ;   d_sigma = c c^T / v            (v = mixture variance)
//   trace_k = s^2 (c0^2 At_k + c0 Bt_k + Ct_k) ,  sum_t dSigma_k[t] d_sigma[tril t] = (c0^2 As_k + c0 Bs_k + Cs_k) / v
//   (dMu d_mu)_k = s (c0 E_k + F_k)                       -> the O(Z^3) contractions are per-model setup
//   q[a, bb] - norm[bb] = t_a - lse_a(t_a) ,  t_a = det_a + sdm_a * omega_bb   (logP[bb] cancels, :142-151)
//   dH[bb] = base - (sum_a e_a t_a / sum_a e_a - lse + logU_bb) ,  e_a = exp(t_a - max_a t)   (one exp / element;
//            np.add(q_min, np.log(U)) broadcasts log U along the COLUMN index bb, acquisitions.py:155-158)
// Non-finite columns surface as a NaN information gain and set nan_flag (the reference raises
// ArithmeticError, acquisitions.py:160-164; its "any isinf -> subtract the max" branch, :150, only
// differs from this when the result is already non-finite).
#pragma once
#include "gp_common.cuh"

namespace fab {

#ifndef FAB_ES_THREADS
#define FAB_ES_THREADS 128
#endif
#ifndef FAB_ES_NC
#define FAB_ES_NC 4
#endif
constexpr int ES_THREADS = FAB_ES_THREADS;
constexpr int ES_XPB = 8;           // candidates per CTA in the entropy kernel

struct EsState {
    int Z, I;
    double* rl;      // B x D    last representer point of each model
    double* vrl;     // B x Np   W k(X, rl)
    double* coef;    // B x 8 x Z : At, Bt, Ct, As, Bs, Cs, E, F
    double* logP;    // B x Z
    double* logU;    // B x Z
    double* base;    // B
    double* omega;   // B x I x Z
};

struct EsSetupArgs {
    EsState st;
    const double* cov_reps;   // B x Z x Z   posterior covariance among the representers (unclipped)
    const double* logP;       // p_min[0]    B x Z
    const double* dMu;        // p_min[1]    B x Z x Z
    const double* dSigma;     // p_min[2]    B x Z x T
    const double* dMuMu;      // p_min[3]    B x Z x Z x Z
    const double* U;          // B x Z
};

// grid (B); warp w handles k = w, w + nwarps, ...; lanes sweep the contracted index (coalesced).
__global__ void __launch_bounds__(ES_THREADS) es_setup_kernel(EsSetupArgs a) {
    __shared__ double ct[EP_MAXZ];
    const int Z = a.st.Z, T = Z * (Z + 1) / 2, b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = ES_THREADS / 32;
    if (tid < Z) ct[tid] = (tid == 0) ? 0.0 : a.cov_reps[((size_t)b * Z + (Z - 1)) * Z + (tid - 1)];
    __syncthreads();
    double* coef = a.st.coef + (size_t)b * 8 * Z;
    for (int k = w; k < Z; k += nw) {
        const double* H = a.dMuMu + ((size_t)b * Z + k) * Z * Z;
        const double* S = a.dSigma + ((size_t)b * Z + k) * T;
        const double* dm = a.dMu + ((size_t)b * Z + k) * Z;
        double bt = 0.0, ctt = 0.0, bs = 0.0, cs = 0.0, f = 0.0;
        for (int j = 0; j < Z; ++j) {
            for (int l = lane; l < Z; l += 32) {
                const double h = H[j * Z + l];
                if (j == 0 && l >= 1) bt = fma(h, ct[l], bt);
                if (l == 0 && j >= 1) bt = fma(h, ct[j], bt);
                if (j >= 1 && l >= 1) ctt = fma(h * ct[j], ct[l], ctt);
                if (l <= j) {
                    const double sv = S[j * (j + 1) / 2 + l];
                    if (l == 0 && j >= 1) bs = fma(sv, ct[j], bs);
                    if (l >= 1) cs = fma(sv * ct[j], ct[l], cs);
                }
            }
        }
        for (int j = 1 + lane; j < Z; j += 32) f = fma(dm[j], ct[j], f);
        bt = warp_sum(bt); ctt = warp_sum(ctt); bs = warp_sum(bs); cs = warp_sum(cs); f = warp_sum(f);
        if (lane == 0) {
            coef[0 * Z + k] = H[0];
            coef[1 * Z + k] = bt;
            coef[2 * Z + k] = ctt;
            coef[3 * Z + k] = S[0];
            coef[4 * Z + k] = bs;
            coef[5 * Z + k] = cs;
            coef[6 * Z + k] = dm[0];
            coef[7 * Z + k] = f;
        }
    }
    if (tid < Z) {
        a.st.logP[(size_t)b * Z + tid] = a.logP[(size_t)b * Z + tid];
        a.st.logU[(size_t)b * Z + tid] = log(a.U[(size_t)b * Z + tid]);
    }
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int k = 0; k < Z; ++k) {
            const double lp = a.logP[(size_t)b * Z + k];
            s += exp(lp) * (lp + log(a.U[(size_t)b * Z + k]));
        }
        a.st.base[b] = s;
    }
}

struct EsMixArgs {
    GpData gd;
    const double* theta; const double* yerr2;   // resident model hyper-parameters
    EsState st;
    const double* Xc; int M; int B;
    const double* mu; const double* var; const double* dot;   // B x M each (predict kernel outputs)
    double* vmix;    // M      mixture predictive variance (predict_testpoint_george, acquisitions.py:113-114)
    double* mmix;    // M      mixture mean
    double* c0;      // B x M  cov(last representer, x)
};

// grid (M): threads sweep the models; fixed-order block reduction => deterministic.
__global__ void __launch_bounds__(ES_THREADS) es_mix_kernel(EsMixArgs a) {
    __shared__ double red[2 * ES_THREADS / 32];
    const int x = blockIdx.x, tid = threadIdx.x, D = a.gd.D, d = a.gd.d;
    double part[2] = {0.0, 0.0};
    for (int b = tid; b < a.B; b += ES_THREADS) {
        const Hyper h = decode_hyper(a.gd, a.theta + (size_t)b * a.gd.Pk, a.yerr2[b]);
        const double* rl = a.st.rl + (size_t)b * D;
        const double* xc = a.Xc + (size_t)x * D;
        double r2 = 0.0;
        for (int k = 0; k < d; ++k) {
            const double dz = rl[k] * h.scale[k] - xc[k] * h.scale[k];
            r2 = fma(dz, dz, r2);
        }
        double m, gf;
        matern52(r2, m, gf);
        double kv = h.amp * m;
        if (a.gd.family == 1) kv *= fma(rl[d] * xc[d], h.inv_g2, h.cb);
        const size_t o = (size_t)b * a.M + x;
        a.c0[o] = kv - a.dot[o];
        const double mu = a.mu[o];
        part[0] += mu;
        part[1] += a.var[o] + mu * mu;
    }
    block_sum<2>(part, red);
    if (tid == 0) {
        const double mean = part[0] / a.B;
        a.mmix[x] = mean;
        a.vmix[x] = part[1] / a.B - mean * mean;
    }
}

struct EsEntropyArgs {
    EsState st;
    int M, B;
    const double* vmix;   // M
    const double* c0;     // B x M
    double* partial;      // M x B   per-model information gain (already averaged over innovations and columns)
};

// exp() of the entropy kernel: e^t / 2^qref for any finite t, exponent handled with integer arithmetic.
//   n = round(t * 256 / ln2) = 256 q + j ;  e^t = 2^q * 2^(j/256) * e^rr ,  rr = t - n ln2/256 , |rr| <= ln2/512
//   2^(j/256) from a 256-entry table, e^rr by a degree-4 Taylor polynomial (remainder 3.8e-17), and the power of two
//   2^(q - qref) added straight into the exponent field -- the reference shift of the log-sum-exp costs no FP64
//   instruction.  11 FP64-pipe instructions per element: t, T, nd, rr, four of the polynomial, table * p, and the two
//   accumulations (17 with the former max pass + exp(t - max), profiles/r02a_es_entropy_before_rewrite_ncu_raw.csv).
// The table is replicated over the 16 bank pairs of shared memory (entry j of lane l at tab[16 j + (l & 15)]): lanes
// pick arbitrary entries, and without the replication the look-ups serialise ~5-fold on bank conflicts.
constexpr int ES_TAB_N = 256;
#ifndef FAB_ES_REPL
#define FAB_ES_REPL 16
#endif
#ifndef FAB_ES_UNROLL
#define FAB_ES_UNROLL 4
#endif
#ifndef FAB_ES_MINB
#define FAB_ES_MINB 1
#endif
constexpr int ES_TAB_REPL = FAB_ES_REPL;
constexpr int ES_UNROLL = FAB_ES_UNROLL;    // rows of the innermost loop unrolled
constexpr int ES_NC = FAB_ES_NC;    // columns per thread and pass: independent chains, broadcasts of det / sdm amortised
__device__ __forceinline__ double es_exp_scaled(double t, int qref, const unsigned char* __restrict__ tab_lane) {
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52
    const double T = fma(t, 369.32993046757462, MAGIC);      // 256 / ln2
    const double nd = T - MAGIC;
    const int n = __double2loint(T);
    // one FMA with the full-precision constant: |error| <= 3e-19 |nd| = 1.1e-16 |t|, the size of the rounding error t
    // itself carries (in this kernel and in the reference's q = logP + det + sto alike); the exact two-constant
    // reduction costs a twelfth FP64 instruction per element and 4.5 % of the kernel (profiles/r02e_es_variants.txt)
    const double rr = fma(nd, -2.7076061740622863e-03, t);   // ln2 / 256
    double p = 4.1666666666666664e-02;                       // 1/4!
    p = fma(p, rr, 1.6666666666666666e-01);
    p = fma(p, rr, 0.5);
    p = fma(p, rr, 1.0);
    p = fma(p, rr, 1.0);
    // entry (n & 255) of this lane's copy: byte offset (n & 255) * 128 from tab_lane
    const double v = *reinterpret_cast<const double*>(tab_lane + ((unsigned)(n & (ES_TAB_N - 1)) * (ES_TAB_REPL * 8))) * p;
    // 2^(q - qref) goes straight into the exponent field (high word).  q - qref <= 0 by the choice of qref; anything
    // below 2^-1020 of the reference is pinned there instead of being flushed: the fast path is only taken when the
    // sum is above 1e-240, so that such terms stay below 2^-200 of it.
    const int dq = max((n >> 8) - qref, -1020);
    return __hiloint2double(__double2hiint(v) + (dq << 20), __double2loint(v));
}

// One column the careful way (exact maximum first, polynomial exp without tables): the path of columns whose sums
// left the range of the fast path, and of inputs that are not finite -- those must surface as NaN (nan_flag).
__device__ __noinline__ double es_column_exact(double o, const double2* __restrict__ sd, int Z) {
    double mx = -INFINITY;
    for (int q = 0; q < Z; ++q) mx = fmax(mx, fma(sd[q].x, o, sd[q].y));
    double se = 0.0, sw = 0.0;
    for (int q = 0; q < Z; ++q) {
        const double t = fma(sd[q].x, o, sd[q].y);
        const double d = t - mx;
        const double e = (d == d) ? exp_nonpos(d) : d;       // t <= mx; a non-finite t makes the sums NaN
        se += e; sw = fma(e, t, sw);
    }
    return sw / se - (mx + log(se));
}

// grid (ceil(M / ES_XPB), B): one model, ES_XPB candidates per CTA; the I * Z columns (innovation p, column bb) of a
// candidate are dealt to the threads, ES_NC per thread and pass.
__global__ void __launch_bounds__(ES_THREADS, FAB_ES_MINB) es_entropy_kernel(EsEntropyArgs a) {
    FAB_DYN_SMEM(smem_raw);
    const int Z = a.st.Z, I = a.st.I, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = ES_THREADS / 32;
    double* etab = reinterpret_cast<double*>(smem_raw);     // ES_TAB_N x ES_TAB_REPL (first: 128-byte rows stay aligned)
    double* coef = etab + ES_TAB_N * ES_TAB_REPL;           // 8 x Z
    double* logU = coef + 8 * Z;                            // Z
    double* om = logU + Z;                                  // I x Z
    double2* sd = reinterpret_cast<double2*>(om + (size_t)((I * Z + 1) & ~1));   // Z x (sdm, det)
    double* red = reinterpret_cast<double*>(sd + Z);        // nw, then [8] smax, [9] dmax, [10] inputs finite, [11] max |det|
    for (int j = tid; j < ES_TAB_N; j += ES_THREADS) {
        const double v = exp2((double)j / ES_TAB_N);
#pragma unroll
        for (int r = 0; r < ES_TAB_REPL; ++r) etab[j * ES_TAB_REPL + r] = v;
    }
    for (int e = tid; e < 8 * Z; e += ES_THREADS) coef[e] = a.st.coef[(size_t)b * 8 * Z + e];
    for (int e = tid; e < I * Z; e += ES_THREADS) om[e] = a.st.omega[(size_t)b * I * Z + e];
    const double base = a.st.base[b];
    double logU_sum = 0.0;
    for (int e = 0; e < Z; ++e) logU_sum += a.st.logU[(size_t)b * Z + e];
    const unsigned char* tab_lane = reinterpret_cast<const unsigned char*>(etab + (lane & (ES_TAB_REPL - 1)));
    __syncthreads();
    for (int xi = 0; xi < ES_XPB; ++xi) {
        const int x = blockIdx.x * ES_XPB + xi;
        if (x >= a.M) break;                               // uniform
        const double v = a.vmix[x];
        const double c0 = a.c0[(size_t)b * a.M + x];
        const double s = sqrt(v + 1e-4) / v;               // chol(v + var_noise) / v   (acquisitions.py:27,33-34)
        if (w == 0) {                                      // Z <= 64: two rows per lane; bounds of the columns' maxima
            double smax = 0.0, dmax = -INFINITY, dabs = 0.0;
            bool fin = true;
            for (int q = lane; q < Z; q += 32) {
                const double tr = s * s * (c0 * c0 * coef[0 * Z + q] + c0 * coef[1 * Z + q] + coef[2 * Z + q]);
                const double ds = (c0 * c0 * coef[3 * Z + q] + c0 * coef[4 * Z + q] + coef[5 * Z + q]) / v;
                const double dt = ds + 0.5 * tr;
                const double sm = s * (c0 * coef[6 * Z + q] + coef[7 * Z + q]);
                sd[q] = make_double2(sm, dt);
                smax = fmax(smax, fabs(sm)); dmax = fmax(dmax, dt); dabs = fmax(dabs, fabs(dt));
                fin = fin && fabs(sm) < 1e100 && fabs(dt) < 1e100;      // false for NaN / inf as well
            }
            smax = warp_max(smax); dmax = warp_max(dmax); dabs = warp_max(dabs);
            const unsigned allfin = __ballot_sync(0xffffffffu, fin);
            if (lane == 0) { red[8] = smax; red[9] = dmax; red[10] = (allfin == 0xffffffffu) ? 1.0 : 0.0; red[11] = dabs; }
        }
        __syncthreads();
        const double smax = red[8], dmax = red[9], dabs = red[11];
        const bool fin = red[10] != 0.0;
        double acc = 0.0;
        const int ncol = I * Z;
        for (int c = tid; c < ncol; c += ES_NC * ES_THREADS) {
            double o[ES_NC], se[ES_NC], sw[ES_NC];
            int qref[ES_NC];
#pragma unroll
            for (int k = 0; k < ES_NC; ++k) {
                const int ck = c + k * ES_THREADS;
                o[k] = (ck < ncol) ? om[ck] : 0.0;          // om[p * Z + bb]
                // every t_a = det_a + sdm_a * o of the column is <= dmax + smax |o|: reference exponent of its sums
                const double mb = fma(fabs(o[k]), smax, dmax);
                qref[k] = (int)floor(fmin(fmax(mb, -1e6), 1e6) * 1.4426950408889634) + 2;
                se[k] = 0.0; sw[k] = 0.0;
            }
#pragma unroll ES_UNROLL
            for (int q = 0; q < Z; ++q) {
                const double2 sq = sd[q];
#pragma unroll
                for (int k = 0; k < ES_NC; ++k) {
                    const double t = fma(sq.x, o[k], sq.y);
                    const double e = es_exp_scaled(t, qref[k], tab_lane);
                    se[k] += e; sw[k] = fma(e, t, sw[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < ES_NC; ++k) {
                const int ck = c + k * ES_THREADS;
                if (ck < ncol) {
                    double h;           // sum_a softmax_a t_a - logsumexp_a t_a of the column
                    // fast path: every |t_a| <= dabs + smax |o| small enough for the 32-bit n of es_exp_scaled, and the
                    // sums inside the exponent window around the reference (bound - maximum < ~800 binades)
                    if (fin && fma(fabs(o[k]), smax, dabs) < 9.0e5 && se[k] > 1e-240 && se[k] < 1e300)
                        h = sw[k] / se[k] - fma((double)qref[k], 0.69314718055994530942, log(se[k]));
                    else
                        h = es_column_exact(o[k], sd, Z);
                    acc -= h;
                }
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) red[w] = acc;
        __syncthreads();
        if (tid == 0) {
            // dH of column (p, bb) = base - (h + logU[bb]); every bb occurs once per innovation: the logU terms and the
            // base sum up to constants of the model (acquisitions.py:155-158)
            double t = 0.0;
            for (int q = 0; q < nw; ++q) t += red[q];
            a.partial[(size_t)x * a.B + b] = (t + (double)ncol * base - (double)I * logU_sum) / ((double)Z * I);
        }
        __syncthreads();
    }
}

__host__ __device__ inline size_t es_entropy_smem_bytes(int Z, int I) {
    return 8 * (size_t)(ES_TAB_N * ES_TAB_REPL + 8 * Z + Z + ((I * Z + 1) & ~1) + 2 * Z + 16) + 64;
}

// grid (ceil(M / 128)): ig[x] = mean_b partial[x][b] in fixed order; NaN -> flag.
__global__ void es_reduce_kernel(const double* partial, int M, int B, double* ig, int* nan_flag) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= M) return;
    double s = 0.0;
    for (int b = 0; b < B; ++b) s += partial[(size_t)x * B + b];
    s /= B;
    ig[x] = s;
    if (nan_flag) nan_flag[x] = isnan(s) ? 1 : 0;
}

}  // namespace fab
