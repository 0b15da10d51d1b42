// epmgp.cuh -- kernel D: EPMGP, the log-probability of each of Z jointly Gaussian points being the
// minimum, with its derivatives (replaces epmgp.joint_min, epmgp.py:52-129, for B (mu, Sigma) pairs).
//
//   epmgp_ep_kernel        grid (Z, B): one CTA per "k is the minimum" event.  Expectation propagation
//                          over the Z-1 half-space factors (min_faktor / lt_factor, epmgp.py:132-282):
//                          V (Z x Z) and M live in shared memory; the scalar site update is computed
//                          redundantly by every thread (identical instruction stream => identical
//                          values => uniform control flow, no broadcast); the rank-1 downdate is spread
//                          over the CTA.  Then logZ and its derivatives (epmgp.py:176-218) from a
//                          shared-memory Cholesky of I + R^T Sigma R.
//   epmgp_normalize_kernel grid (B): renormalisation of logP and of the three derivative tensors
//                          across k (epmgp.py:102-128), in place.
// Constants follow the reference exactly: float32 eps, <= 50 sweeps, |sum d| < 1e-3,
// z < -6 / z > 6 exits, 1e-25, -500, jitter ladder 0 / 1e-10 / 1e-6.
#pragma once
#include "prims.cuh"

namespace fab {

constexpr int EP_THREADS = 128;
constexpr int EP_MAXZ = 64;

struct EpArgs {
    const double* mu;      // B x Z
    const double* Sigma;   // B x Z x Z
    int Z;
    double* logP;          // B x Z      (un-normalised logZ_k after the EP kernel)
    double* dMu;           // B x Z x Z
    double* dSigma;        // B x Z x T,  T = Z (Z + 1) / 2, row-major lower-triangle order
    double* dMuMu;         // B x Z x Z x Z
    int* info;             // B : 0 ok, 1 = "variance contains NaN" (reference raises), 2 = Cholesky failed
    int* sweeps;           // B x Z or null (diagnostic: EP sweeps used)
    int with_derivatives;
};

__host__ __device__ inline size_t ep_smem_bytes(int Z) {
    return 8 * (size_t)(2 * Z * (Z + 1) + 8 * Z + 32) + 64;
}

__device__ __forceinline__ double np_max2(double a, double b) {     // np.max([a, b]) propagates NaN
    if (isnan(a) || isnan(b)) return NAN;
    return a > b ? a : b;
}

__global__ void __launch_bounds__(EP_THREADS) epmgp_ep_kernel(EpArgs a) {
    FAB_DYN_SMEM(smem_raw);
    const int Z = a.Z, n = Z - 1, LD = Z + 1, T = Z * (Z + 1) / 2;
    const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const double SQ2 = 1.4142135623730951, EPS32 = 1.1920928955078125e-07, L2P = 1.8378770664093453;

    double* V = reinterpret_cast<double*>(smem_raw);      // Z x LD   (later: IRSR / its Cholesky factor)
    double* Y = V + Z * LD;                               // Z x LD   (later: L^-1, then H)
    double* M = Y + Z * LD;                               // Z
    double* Vc = M + Z;                                   // Z
    double* P = Vc + Z;                                   // n
    double* MP = P + Z;                                   // n
    double* lS = MP + Z;                                  // n
    double* rv = lS + Z;                                  // Z : r
    double* bv = rv + Z;                                  // Z : b, then A b
    double* sc = bv + Z;                                  // scratch (32)

    const double* Mu = a.mu + (size_t)b * Z;
    const double* Sg = a.Sigma + (size_t)b * Z * Z;
    for (int e = tid; e < Z * Z; e += EP_THREADS) V[(e / Z) * LD + (e % Z)] = Sg[e];
    for (int i = tid; i < Z; i += EP_THREADS) { M[i] = Mu[i]; P[i] = 0.0; MP[i] = 0.0; lS[i] = 0.0; }
    __syncthreads();

    // ------------------------------ EP sweeps (epmgp.py:146-163) ------------------------------
    double d = NAN;
    int err = 0, nsweep = 0;
    for (int count = 0; count < 50; ++count) {
        double diff = 0.0;
        ++nsweep;
        for (int i = 0; i < n; ++i) {
            const int l = i < k ? i : i + 1;
            // ---- lt_factor(k, l, M, V, MP[i], P[i], 1) : scalar part, every thread identically ----
            const double p = P[i], mp = MP[i];
            const double cVc = (V[l * LD + l] - 2 * V[k * LD + l] + V[k * LD + k]) / 2.0;
            const double cM = (M[l] - M[k]) / SQ2;
            const double cVnic = np_max2(cVc / (1 - p * cVc), 0.0);
            const double cmni = cM + cVnic * (p * cM - mp);
            double z = cmni / sqrt(cVnic + 1e-25);
            if (isnan(z)) z = -INFINITY;
            if (tid < Z) Vc[tid] = (V[tid * LD + l] - V[tid * LD + k]) / SQ2;
            double f = 0.0, fm = 0.0, pnew = 0.0, mpnew = 0.0, logS = 0.0;
            int flag;
            if (z < -6) {
                flag = -1;
            } else if (z > 6) {
                flag = 1;
                const double dp = -p, dmp = -mp;
                d = np_max2(dmp, dp);
                f = dp / (1 + dp * cVc);
                fm = (dmp - cM * dp) / (1 + dp * cVc);
            } else {
                flag = 0;
                const double logphi = -0.5 * (z * z + L2P);
                const double lP = log(0.5 * erfc(-z / SQ2));
                const double e = exp(logphi - lP);
                const double alpha = e / sqrt(cVnic);
                const double beta = alpha * (alpha * cVnic + cmni);
                const double r = beta / (1 - beta);
                pnew = r / cVnic;
                mpnew = r * (alpha + cmni / cVnic) + alpha;
                const double dp = np_max2(-p + EPS32, pnew - p);
                const double dmp = np_max2(-mp + EPS32, mpnew - mp);
                d = np_max2(dmp, dp);
                pnew = p + dp;
                mpnew = mp + dmp;
                f = dp / (1 + dp * cVc);
                fm = (dmp - cM * dp) / (1 + dp * cVc);
                logS = lP - 0.5 * (log(beta) - log(pnew) - log(cVnic)) + (alpha * alpha) / (2 * beta) * cVnic;
            }
            __syncthreads();                     // Vc complete; every thread has read its scalars
            if (flag == -1) { d = NAN; break; }
            int bad = 0;
            for (int e = tid; e < Z * Z; e += EP_THREADS) {
                const int r_ = e / Z, c_ = e - r_ * Z;
                const double v = V[r_ * LD + c_] - f * (Vc[r_] * Vc[c_]);
                V[r_ * LD + c_] = v;
                if (isnan(v)) bad = 1;
            }
            if (tid < Z) M[tid] = M[tid] + fm * Vc[tid];
            if (tid == 0) { P[i] = pnew; MP[i] = mpnew; lS[i] = logS; }
            bad = __syncthreads_or(bad && flag == 0);
            if (bad) { err = 1; break; }
            if (isnan(d)) break;                 // epmgp.py:158-159
            diff += fabs(d);
        }
        if (err || isnan(d)) break;
        if (fabs(diff) < 0.001) break;
    }
    if (a.sweeps && tid == 0) a.sweeps[(size_t)b * Z + k] = nsweep;
    double* o_dMu = a.dMu ? a.dMu + ((size_t)b * Z + k) * Z : nullptr;
    double* o_dMM = a.dMuMu ? a.dMuMu + ((size_t)b * Z + k) * Z * Z : nullptr;
    double* o_dSg = a.dSigma ? a.dSigma + ((size_t)b * Z + k) * T : nullptr;
    if (err) {
        if (tid == 0) { atomicMax(&a.info[b], 1); a.logP[(size_t)b * Z + k] = NAN; }
        return;
    }
    if (isnan(d)) {                              // epmgp.py:164-175
        if (tid == 0) a.logP[(size_t)b * Z + k] = -INFINITY;
        if (a.with_derivatives) {
            for (int e = tid; e < Z; e += EP_THREADS) o_dMu[e] = 0.0;
            for (int e = tid; e < Z * Z; e += EP_THREADS) o_dMM[e] = 0.0;
            for (int e = tid; e < T; e += EP_THREADS) o_dSg[e] = 0.0;
        }
        return;
    }

    // ------------------------------ logZ and derivatives (epmgp.py:176-218) ------------------------------
    // r = C MP : r[l_j] = MP_j / sq2, r[k] = -sum_j MP_j / sq2
    if (tid < Z) {
        double v;
        if (tid == k) {
            v = 0.0;
            for (int j = 0; j < n; ++j) v += MP[j] * (-1 / SQ2);
        } else {
            const int j = tid < k ? tid : tid - 1;
            v = MP[j] * (1 / SQ2);
        }
        rv[tid] = v;
    }
    __syncthreads();
    // b = Mu + Sigma r ; rSr ; Mu.r ; sum logS ; mpm
    if (tid < Z) {
        double s = 0.0;
        for (int q = 0; q < Z; ++q) s += Sg[tid * Z + q] * rv[q];
        Vc[tid] = s;                              // Sigma r
        bv[tid] = Mu[tid] + s;
    }
    __syncthreads();
    if (tid == 0) {
        double rSr = 0.0, Mur = 0.0, s = 0.0, mpm = 0.0;
        for (int q = 0; q < Z; ++q) { rSr += rv[q] * Vc[q]; Mur += Mu[q] * rv[q]; }
        for (int j = 0; j < n; ++j) { s += lS[j]; if (MP[j] != 0.0) mpm += MP[j] * MP[j] / P[j]; }
        sc[0] = rSr; sc[1] = Mur; sc[2] = s; sc[3] = mpm;
    }
    // IRSR = I + R^T Sigma R with the jitter ladder; Cholesky in V (n x n, leading dimension LD)
    int chol_fail = 1;
    for (int attempt = 0; attempt < 3 && chol_fail; ++attempt) {
        const double jitter = attempt == 0 ? 0.0 : (attempt == 1 ? 1e-10 : 1e-6);
        __syncthreads();
        for (int e = tid; e < n * n; e += EP_THREADS) {
            const int i = e / n, j = e - i * n;
            const int li = i < k ? i : i + 1, lj = j < k ? j : j + 1;
            const double q = 0.5 * (Sg[li * Z + lj] - Sg[li * Z + k] - Sg[k * Z + lj] + Sg[k * Z + k]);
            V[i * LD + j] = sqrt(P[i]) * sqrt(P[j]) * q + (i == j ? 1.0 + jitter : 0.0);
        }
        __syncthreads();
        int failed = 0;
        for (int j = 0; j < n; ++j) {            // right-looking, column j
            const double piv = V[j * LD + j];
            if (!(piv > 0.0)) { failed = 1; break; }   // uniform: every thread reads the same value
            const double ljj = sqrt(piv);
            __syncthreads();
            for (int i = j + tid; i < n; i += EP_THREADS) V[i * LD + j] = (i == j) ? ljj : V[i * LD + j] / ljj;
            __syncthreads();
            for (int e = tid; e < (n - j - 1) * (n - j - 1); e += EP_THREADS) {
                const int i = j + 1 + e / (n - j - 1), c = j + 1 + e % (n - j - 1);
                if (c <= i) V[i * LD + c] -= V[i * LD + j] * V[c * LD + j];
            }
            __syncthreads();
        }
        chol_fail = failed;
    }
    if (chol_fail) {
        if (tid == 0) { atomicMax(&a.info[b], 2); a.logP[(size_t)b * Z + k] = NAN; }
        return;
    }
    // Y = L^-1 (lower), one column per thread
    for (int c = tid; c < n; c += EP_THREADS) {
        for (int i = 0; i < c; ++i) Y[i * LD + c] = 0.0;
        Y[c * LD + c] = 1.0 / V[c * LD + c];
        for (int i = c + 1; i < n; ++i) {
            double s = 0.0;
            for (int q = c; q < i; ++q) s += V[i * LD + q] * Y[q * LD + c];
            Y[i * LD + c] = -s / V[i * LD + i];
        }
    }
    if (tid == 0) {
        double dts = 0.0;
        for (int j = 0; j < n; ++j) dts += log(V[j * LD + j]);
        sc[4] = 2.0 * dts;
    }
    __syncthreads();
    // H[i][j] = sp_i G[i][j] sp_j / 2, G = Y^T Y = IRSR^-1 ; H overwrites V (L no longer needed after Y)
    for (int e = tid; e < n * n; e += EP_THREADS) {
        const int i = e / n, j = e - i * n;
        if (j > i) continue;
        double s = 0.0;
        for (int q = i; q < n; ++q) s += Y[q * LD + i] * Y[q * LD + j];
        V[i * LD + j] = 0.5 * sqrt(P[i]) * sqrt(P[j]) * s;      // safe: this pass only reads Y
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += EP_THREADS) {   // mirror to the upper triangle
        const int i = e / n, j = e - i * n;
        if (j > i) V[i * LD + j] = V[j * LD + i];
    }
    __syncthreads();
    // A (Z x Z) from H: A[l_i][l_j] = H[i][j] ; A[k][l_j] = A[l_j][k] = -sum_i H[i][j] ; A[k][k] = sum_ij H
    // stored in Y with leading dimension LD
    if (tid < n) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += V[i * LD + tid];
        Vc[tid] = s;                              // column sums of H
    }
    __syncthreads();
    for (int e = tid; e < Z * Z; e += EP_THREADS) {
        const int r_ = e / Z, c_ = e - r_ * Z;
        double v;
        if (r_ == k && c_ == k) {
            v = 0.0;
            for (int j = 0; j < n; ++j) v += Vc[j];
        } else if (r_ == k) {
            v = -Vc[c_ < k ? c_ : c_ - 1];
        } else if (c_ == k) {
            v = -Vc[r_ < k ? r_ : r_ - 1];
        } else {
            v = V[(r_ < k ? r_ : r_ - 1) * LD + (c_ < k ? c_ : c_ - 1)];
        }
        Y[r_ * LD + c_] = v;
    }
    __syncthreads();
    // Ab
    if (tid < Z) {
        double s = 0.0;
        for (int q = 0; q < Z; ++q) s += Y[tid * LD + q] * bv[q];
        M[tid] = s;                               // Ab (M is free now)
    }
    __syncthreads();
    if (tid == 0) {
        double bAb = 0.0;
        for (int q = 0; q < Z; ++q) bAb += bv[q] * M[q];
        a.logP[(size_t)b * Z + k] = 0.5 * (sc[0] - bAb - sc[4]) + sc[1] + sc[2] - 0.5 * sc[3];
    }
    if (!a.with_derivatives) return;
    for (int e = tid; e < Z; e += EP_THREADS) o_dMu[e] = rv[e] - M[e];
    for (int e = tid; e < Z * Z; e += EP_THREADS) o_dMM[e] = -Y[(e / Z) * LD + (e % Z)];
    // dSigma: S = -A - 2 r Ab^T + r r^T + (b^T A) Ab^T, b^T A = Ab ; 0.5 (S + S^T - diag S), lower order
    for (int e = tid; e < T; e += EP_THREADS) {
        int r_ = 0;
        while ((r_ + 1) * (r_ + 2) / 2 <= e) ++r_;
        const int c_ = e - r_ * (r_ + 1) / 2;
        const double s_rc = -Y[r_ * LD + c_] - 2 * rv[r_] * M[c_] + rv[r_] * rv[c_] + M[r_] * M[c_];
        const double s_cr = -Y[c_ * LD + r_] - 2 * rv[c_] * M[r_] + rv[c_] * rv[r_] + M[c_] * M[r_];
        o_dSg[e] = (r_ == c_) ? 0.5 * s_rc : 0.5 * (s_rc + s_cr);
    }
}

// Renormalisation across k (epmgp.py:102-128).  One CTA per hyper-sample.
__global__ void __launch_bounds__(256) epmgp_normalize_kernel(EpArgs a) {
    FAB_DYN_SMEM(smem_raw);
    const int Z = a.Z, T = Z * (Z + 1) / 2, b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    double* ep = reinterpret_cast<double*>(smem_raw);     // Z
    double* Zm = ep + Z;                                  // Z
    double* Zs = Zm + Z;                                  // T
    double* gg = Zs + T;                                  // Z x Z
    double* sc = gg + Z * Z;                              // 4
    double* lp = a.logP + (size_t)b * Z;
    if (a.info[b] != 0) return;                           // reference raised: outputs undefined
    if (tid == 0) {
        double Zsum = 0.0, mx = -INFINITY;
        for (int k = 0; k < Z; ++k) {
            double v = lp[k];
            if (isinf(v)) v = -500.0;
            lp[k] = v;
            ep[k] = exp(v);
            Zsum += ep[k];
            if (v > mx) mx = v;
        }
        double se = 0.0;
        for (int k = 0; k < Z; ++k) se += exp(lp[k] - mx);
        double s = mx + log(se);
        if (isinf(s)) s = mx;
        sc[0] = Zsum; sc[1] = s;
    }
    __syncthreads();
    const double Zsum = sc[0], s = sc[1];
    if (a.with_derivatives) {
        double* dMu = a.dMu + (size_t)b * Z * Z;
        double* dSg = a.dSigma + (size_t)b * Z * T;
        double* dMM = a.dMuMu + (size_t)b * Z * Z * Z;
        for (int j = tid; j < Z; j += nt) {
            double v = 0.0;
            for (int k = 0; k < Z; ++k) v += ep[k] * dMu[k * Z + j];
            Zm[j] = v / Zsum;
        }
        for (int t = tid; t < T; t += nt) {
            double v = 0.0;
            for (int k = 0; k < Z; ++k) v += ep[k] * dSg[(size_t)k * T + t];
            Zs[t] = v / Zsum;
        }
        for (int e = tid; e < Z * Z; e += nt) {
            const int i = e / Z, j = e - i * Z;
            double v = 0.0;
            for (int k = 0; k < Z; ++k) v += (dMM[((size_t)k * Z + i) * Z + j] + dMu[k * Z + i] * dMu[k * Z + j]) * ep[k];
            gg[e] = v;
        }
        __syncthreads();
        for (int e = tid; e < Z * Z; e += nt) dMu[e] -= Zm[e % Z];
        for (size_t e = tid; e < (size_t)Z * T; e += nt) dSg[e] -= Zs[e % T];
        for (size_t e = tid; e < (size_t)Z * Z * Z; e += nt) {
            const int ij = (int)(e % ((size_t)Z * Z)), j = ij % Z;
            dMM[e] += -gg[ij] / Zsum + Zm[j] * Zm[j];      // elementwise Zm.T * Zm (epmgp.py:126)
        }
    }
    __syncthreads();
    for (int k = tid; k < Z; k += nt) lp[k] = lp[k] - s;
}

__host__ __device__ inline size_t ep_norm_smem_bytes(int Z) {
    return 8 * (size_t)(2 * Z + Z * (Z + 1) / 2 + Z * Z + 8) + 64;
}

}  // namespace fab
