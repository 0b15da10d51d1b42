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

constexpr int ES_THREADS = 128;
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

// grid (ceil(M / ES_XPB), B): one model, ES_XPB candidates per CTA; warp w takes innovations w, w+4, ...
__global__ void __launch_bounds__(ES_THREADS) es_entropy_kernel(EsEntropyArgs a) {
    FAB_DYN_SMEM(smem_raw);
    const int Z = a.st.Z, I = a.st.I, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = ES_THREADS / 32;
    double* coef = reinterpret_cast<double*>(smem_raw);     // 8 x Z
    double* logU = coef + 8 * Z;                            // Z
    double* om = logU + Z;                                  // I x Z
    double* det = om + (size_t)I * Z;                       // Z
    double* sdm = det + Z;                                  // Z
    double* red = sdm + Z;                                  // nw (8 reserved)
    double* etab = red + 8;                                 // exp table of exp_neg_fast
    exp_table_fill(etab);
    for (int e = tid; e < 8 * Z; e += ES_THREADS) coef[e] = a.st.coef[(size_t)b * 8 * Z + e];
    for (int e = tid; e < Z; e += ES_THREADS) logU[e] = a.st.logU[(size_t)b * Z + e];
    for (int e = tid; e < I * Z; e += ES_THREADS) om[e] = a.st.omega[(size_t)b * I * Z + e];
    const double base = a.st.base[b];
    __syncthreads();
    for (int xi = 0; xi < ES_XPB; ++xi) {
        const int x = blockIdx.x * ES_XPB + xi;
        if (x >= a.M) break;                               // uniform
        const double v = a.vmix[x];
        const double c0 = a.c0[(size_t)b * a.M + x];
        const double s = sqrt(v + 1e-4) / v;               // chol(v + var_noise) / v   (acquisitions.py:27,33-34)
        if (tid < Z) {
            const double tr = s * s * (c0 * c0 * coef[0 * Z + tid] + c0 * coef[1 * Z + tid] + coef[2 * Z + tid]);
            const double ds = (c0 * c0 * coef[3 * Z + tid] + c0 * coef[4 * Z + tid] + coef[5 * Z + tid]) / v;
            det[tid] = ds + 0.5 * tr;
            sdm[tid] = s * (c0 * coef[6 * Z + tid] + coef[7 * Z + tid]);
        }
        __syncthreads();
        // two (innovation p, column bb) pairs per thread and pass (they share the det / sdm broadcasts and give the
        // exp chains a partner to overlap with): I * Z columns over the whole CTA
        double acc = 0.0;
        const int ncol = I * Z;
        for (int c = tid; c < ncol; c += 2 * ES_THREADS) {
            const int c2 = c + ES_THREADS;
            const bool two = c2 < ncol;
            const double o1 = om[c], o2 = two ? om[c2] : 0.0;     // om[p * Z + bb]
            double mx1 = -INFINITY, mx2 = -INFINITY;
            for (int q = 0; q < Z; ++q) {
                const double sq = sdm[q], dq = det[q];
                mx1 = fmax(mx1, fma(sq, o1, dq));
                mx2 = fmax(mx2, fma(sq, o2, dq));
            }
            double se1 = 0.0, sw1 = 0.0, se2 = 0.0, sw2 = 0.0;
            for (int q = 0; q < Z; ++q) {
                const double sq = sdm[q], dq = det[q];
                const double t1 = fma(sq, o1, dq), t2 = fma(sq, o2, dq);
                const double e1 = exp_neg_fast(mx1 - t1, etab);   // t <= mx; non-finite t makes the sums NaN -> nan_flag
                const double e2 = exp_neg_fast(mx2 - t2, etab);
                se1 += e1; sw1 = fma(e1, t1, sw1);
                se2 += e2; sw2 = fma(e2, t2, sw2);
            }
            acc += base - (sw1 / se1 - (mx1 + log(se1)) + logU[c % Z]);
            if (two) acc += base - (sw2 / se2 - (mx2 + log(se2)) + logU[c2 % Z]);
        }
        acc = warp_sum(acc);
        if (lane == 0) red[w] = acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int q = 0; q < nw; ++q) t += red[q];
            a.partial[(size_t)x * a.B + b] = t / ((double)Z * I);
        }
        __syncthreads();
    }
}

__host__ __device__ inline size_t es_entropy_smem_bytes(int Z, int I) {
    return 8 * (size_t)(8 * Z + Z + I * Z + 2 * Z + 8 + EXP_TAB_N) + 64;
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
