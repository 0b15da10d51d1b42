/* epmgp_oracle.c -- plain-C restatement of the reference's EPMGP "probability of being the minimum"
 * (expectation propagation over half-space factors).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows /root/reference/epmgp.py:
 *   joint_min            epmgp.py:52-129   (loop over k, -inf -> -500, renormalisation of logP and of
 *                                           the three derivative tensors, incl. the elementwise
 *                                           `Zm.T * Zm` quirk at :126)
 *   min_faktor           epmgp.py:132-218  (<= 50 EP sweeps, |sum d| < 1e-3; logZ and derivatives)
 *   lt_factor            epmgp.py:221-282  (one site update)
 *   log_relative_gauss   epmgp.py:286-298  (+-6 cut-offs)
 * PINNED: tests/golden/epmgp_*.npz hold outputs of the reference's own epmgp.joint_min run in the
 * authoring container (tests/golden/make_golden.py); tests/test_oracle_epmgp.py compares.
 *
 * Differences from the reference that do not change results beyond round-off: np.linalg.solve (LU)
 * is restated as Gaussian elimination with partial pivoting; np.linalg.cholesky as an unblocked
 * lower Cholesky with the same "pivot <= 0 or NaN -> failure" rule and the same jitter ladder.
 *
 * Build: gcc -O2 -fPIC -shared -o _build/libepmgp_oracle.so epmgp_oracle.c -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static const double SQ2 = 1.4142135623730951;
static const double EPS32 = 1.1920928955078125e-07; /* np.finfo(np.float32).eps */
static const double L2P = 1.8378770664093453;       /* log(2) + log(pi) */

static double np_max2(double a, double b) { /* np.max([a, b]) propagates NaN */
    if (isnan(a) || isnan(b)) return NAN;
    return a > b ? a : b;
}

/* epmgp.py:286-298 */
static int log_relative_gauss(double z, double* e, double* logPhi) {
    if (z < -6) { *e = 1; *logPhi = -1.0e12; return -1; }
    if (z > 6) { *e = 0; *logPhi = 0; return 1; }
    double logphi = -0.5 * (z * z + L2P);
    *logPhi = log(0.5 * erfc(-z / SQ2));
    *e = exp(logphi - *logPhi);
    return 0;
}

/* epmgp.py:221-282.  Updates M (D), V (D x D) in place; returns d (NaN = abort).  *err = 1 when the
 * reference would raise ("Resulting variance contains NaN"). */
static double lt_factor(int s, int l, int D, double* M, double* V, double* mp, double* p, double* logS,
                        double gamma, double* Vc, int* err) {
    double cVc = (V[l * D + l] - 2 * V[s * D + l] + V[s * D + s]) / 2.0;
    for (int i = 0; i < D; ++i) Vc[i] = (V[i * D + l] - V[i * D + s]) / SQ2;
    double cM = (M[l] - M[s]) / SQ2;
    double cVnic = np_max2(cVc / (1 - (*p) * cVc), 0);
    double cmni = cM + cVnic * ((*p) * cM - (*mp));
    double z = cmni / sqrt(cVnic + 1e-25);
    if (isnan(z)) z = -INFINITY;
    double e, lP;
    int flag = log_relative_gauss(z, &e, &lP);
    double d;
    if (flag == 0) {
        double alpha = e / sqrt(cVnic);
        double beta = alpha * (alpha * cVnic + cmni);
        double r = beta / (1 - beta);
        double pnew = r / cVnic;
        double mpnew = r * (alpha + cmni / cVnic) + alpha;
        double dp = np_max2(-(*p) + EPS32, gamma * (pnew - (*p)));
        double dmp = np_max2(-(*mp) + EPS32, gamma * (mpnew - (*mp)));
        d = np_max2(dmp, dp);
        pnew = (*p) + dp;
        mpnew = (*mp) + dmp;
        double f = dp / (1 + dp * cVc);
        double fm = (dmp - cM * dp) / (1 + dp * cVc);
        int bad = 0;
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) {
                V[i * D + j] = V[i * D + j] - f * (Vc[i] * Vc[j]);
                if (isnan(V[i * D + j])) bad = 1;
            }
        for (int i = 0; i < D; ++i) M[i] = M[i] + fm * Vc[i];
        if (bad) { *err = 1; return NAN; }
        *logS = lP - 0.5 * (log(beta) - log(pnew) - log(cVnic)) + (alpha * alpha) / (2 * beta) * cVnic;
        *p = pnew; *mp = mpnew;
    } else if (flag == -1) {
        d = NAN;
        *p = 0; *mp = 0; *logS = -INFINITY;
    } else {
        double dp = -(*p), dmp = -(*mp);
        d = np_max2(dmp, dp);
        double f = dp / (1 + dp * cVc);
        double fm = (dmp - cM * dp) / (1 + dp * cVc);
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) V[i * D + j] = V[i * D + j] - f * (Vc[i] * Vc[j]);
        for (int i = 0; i < D; ++i) M[i] = M[i] + fm * Vc[i];
        *p = 0; *mp = 0; *logS = 0;
    }
    return d;
}

static int cholesky_lower(int n, const double* A, double jitter, double* L) {
    for (int i = 0; i < n * n; ++i) L[i] = 0;
    for (int j = 0; j < n; ++j) {
        double s = A[j * n + j] + jitter;
        for (int k = 0; k < j; ++k) s -= L[j * n + k] * L[j * n + k];
        if (!(s > 0)) return j + 1;
        double ljj = sqrt(s);
        L[j * n + j] = ljj;
        for (int i = j + 1; i < n; ++i) {
            double t = A[i * n + j];
            for (int k = 0; k < j; ++k) t -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = t / ljj;
        }
    }
    return 0;
}

/* solve A X = B (n x n, n x m) by Gaussian elimination with partial pivoting; A, B overwritten */
static void lu_solve(int n, int m, double* A, double* B) {
    for (int k = 0; k < n; ++k) {
        int piv = k; double best = fabs(A[k * n + k]);
        for (int i = k + 1; i < n; ++i) if (fabs(A[i * n + k]) > best) { best = fabs(A[i * n + k]); piv = i; }
        if (piv != k) {
            for (int j = 0; j < n; ++j) { double t = A[k * n + j]; A[k * n + j] = A[piv * n + j]; A[piv * n + j] = t; }
            for (int j = 0; j < m; ++j) { double t = B[k * m + j]; B[k * m + j] = B[piv * m + j]; B[piv * m + j] = t; }
        }
        for (int i = k + 1; i < n; ++i) {
            double f = A[i * n + k] / A[k * n + k];
            if (f == 0) continue;
            for (int j = k; j < n; ++j) A[i * n + j] -= f * A[k * n + j];
            for (int j = 0; j < m; ++j) B[i * m + j] -= f * B[k * m + j];
        }
    }
    for (int i = n - 1; i >= 0; --i)
        for (int j = 0; j < m; ++j) {
            double s = B[i * m + j];
            for (int k = i + 1; k < n; ++k) s -= A[i * n + k] * B[k * m + j];
            B[i * m + j] = s / A[i * n + i];
        }
}

/* epmgp.py:132-218.  Outputs: *logZ, dMu (D), dMuMu (D x D), dSigma (D(D+1)/2, row-major lower order).
 * Returns 0, or 1 when the reference would raise an Exception. sweeps_out (optional) = EP sweeps run. */
static int min_faktor(int D, const double* Mu, const double* Sigma, int k, double* logZ, double* dMu,
                      double* dMuMu, double* dSigma, int* sweeps_out) {
    const int n = D - 1, T = D * (D + 1) / 2;
    double* logS = calloc(n > 0 ? n : 1, sizeof(double));
    double* MP = calloc(n > 0 ? n : 1, sizeof(double));
    double* P = calloc(n > 0 ? n : 1, sizeof(double));
    double* M = malloc(sizeof(double) * D);
    double* V = malloc(sizeof(double) * D * D);
    double* Vc = malloc(sizeof(double) * D);
    memcpy(M, Mu, sizeof(double) * D);
    memcpy(V, Sigma, sizeof(double) * D * D);
    double d = NAN;
    int err = 0, sweeps = 0;
    for (int count = 0; count < 50; ++count) {
        double diff = 0;
        ++sweeps;
        for (int i = 0; i < n; ++i) {
            int l = i < k ? i : i + 1;
            d = lt_factor(k, l, D, M, V, &MP[i], &P[i], &logS[i], 1.0, Vc, &err);
            if (err) goto done;
            if (isnan(d)) break;
            diff += fabs(d);
        }
        if (isnan(d)) break;
        if (fabs(diff) < 0.001) break;
    }
    if (sweeps_out) *sweeps_out = sweeps;
    if (isnan(d)) {
        *logZ = -INFINITY;
        for (int i = 0; i < D; ++i) dMu[i] = 0;
        for (int i = 0; i < D * D; ++i) dMuMu[i] = 0;
        for (int i = 0; i < T; ++i) dSigma[i] = 0;
        goto done;
    }
    {
        /* C: D x n, C[i][j] = 1/sq2 if i == l_j ; -1/sq2 if i == k */
        double* R = calloc((size_t)D * n, sizeof(double));
        double* r = calloc(D, sizeof(double));
        for (int j = 0; j < n; ++j) {
            int l = j < k ? j : j + 1;
            double sp = sqrt(P[j]);
            R[l * n + j] = sp * (1 / SQ2);
            R[k * n + j] = sp * (-1 / SQ2);
        }
        for (int i = 0; i < D; ++i) {
            double s = 0;
            for (int j = 0; j < n; ++j) {
                int l = j < k ? j : j + 1;
                double c = (i == l) ? 1 / SQ2 : ((i == k) ? -1 / SQ2 : 0.0);
                s += MP[j] * c;
            }
            r[i] = s;
        }
        double mpm = 0;
        for (int j = 0; j < n; ++j) if (MP[j] != 0) mpm += MP[j] * MP[j] / P[j];
        double s = 0;
        for (int j = 0; j < n; ++j) s += logS[j];
        /* IRSR = I + R^T Sigma R */
        double* SR = malloc(sizeof(double) * D * n);
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < n; ++j) {
                double t = 0;
                for (int q = 0; q < D; ++q) t += Sigma[i * D + q] * R[q * n + j];
                SR[i * n + j] = t;
            }
        double* IRSR = malloc(sizeof(double) * n * n);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double t = 0;
                for (int q = 0; q < D; ++q) t += R[q * n + i] * SR[q * n + j];
                IRSR[i * n + j] = t + (i == j ? 1.0 : 0.0);
            }
        double* Sr = malloc(sizeof(double) * D);
        double rSr = 0;
        for (int i = 0; i < D; ++i) {
            double t = 0;
            for (int q = 0; q < D; ++q) t += Sigma[i * D + q] * r[q];
            Sr[i] = t;
        }
        for (int i = 0; i < D; ++i) rSr += r[i] * Sr[i];
        /* A = R solve(IRSR, R^T) */
        double* lu = malloc(sizeof(double) * n * n);
        double* X = malloc(sizeof(double) * n * D);
        memcpy(lu, IRSR, sizeof(double) * n * n);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < D; ++j) X[i * D + j] = R[j * n + i];
        lu_solve(n, D, lu, X);
        double* A = malloc(sizeof(double) * D * D);
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) {
                double t = 0;
                for (int q = 0; q < n; ++q) t += R[i * n + q] * X[q * D + j];
                A[i * D + j] = t;
            }
        for (int i = 0; i < D; ++i)
            for (int j = i + 1; j < D; ++j) {
                double t = 0.5 * (A[i * D + j] + A[j * D + i]);
                A[i * D + j] = A[j * D + i] = t;
            }
        double* b = malloc(sizeof(double) * D);
        double* Ab = malloc(sizeof(double) * D);
        for (int i = 0; i < D; ++i) b[i] = Mu[i] + Sr[i];
        for (int i = 0; i < D; ++i) {
            double t = 0;
            for (int q = 0; q < D; ++q) t += A[i * D + q] * b[q];
            Ab[i] = t;
        }
        double* Lc = malloc(sizeof(double) * n * n);
        if (cholesky_lower(n, IRSR, 0.0, Lc) != 0)
            if (cholesky_lower(n, IRSR, 1e-10, Lc) != 0)
                if (cholesky_lower(n, IRSR, 1e-6, Lc) != 0) err = 2; /* reference raises LinAlgError */
        double dts = 0;
        for (int i = 0; i < n; ++i) dts += log(Lc[i * n + i]);
        dts *= 2;
        double bAb = 0, Mur = 0;
        for (int i = 0; i < D; ++i) { bAb += b[i] * Ab[i]; Mur += Mu[i] * r[i]; }
        *logZ = 0.5 * (rSr - bAb - dts) + Mur + s - 0.5 * mpm;
        /* btA = b^T A = Ab (A symmetric) -- keep the reference's expression */
        for (int i = 0; i < D; ++i) dMu[i] = r[i] - Ab[i];
        for (int i = 0; i < D * D; ++i) dMuMu[i] = -A[i];
        double* Sg = malloc(sizeof(double) * D * D);
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j) {
                double btA_i = 0;
                for (int q = 0; q < D; ++q) btA_i += b[q] * A[q * D + i];
                Sg[i * D + j] = -A[i * D + j] - 2 * r[i] * Ab[j] + r[i] * r[j] + btA_i * Ab[j];
            }
        int t = 0;
        for (int a = 0; a < D; ++a)
            for (int c = 0; c <= a; ++c) {
                double v = (a == c) ? 0.5 * Sg[a * D + a] : 0.5 * (Sg[a * D + c] + Sg[c * D + a]);
                dSigma[t++] = v;
            }
        free(R); free(r); free(SR); free(IRSR); free(Sr); free(lu); free(X); free(A); free(b); free(Ab);
        free(Lc); free(Sg);
    }
done:
    free(logS); free(MP); free(P); free(M); free(V); free(Vc);
    return err;
}

/* epmgp.py:52-129.  logP (D), dMu (D x D), dSigma (D x T), dMuMu (D x D x D); derivative pointers may
 * be NULL (with_derivatives=False).  Returns 0 or the error code of min_faktor. */
int epmgp_joint_min(int D, const double* mu, const double* var, int with_derivatives, double* logP,
                    double* dlogPdMu, double* dlogPdSigma, double* dlogPdMudMu, int* sweeps) {
    const int T = D * (D + 1) / 2;
    double* dMu = malloc(sizeof(double) * D * D);
    double* dSg = malloc(sizeof(double) * D * T);
    double* dMM = malloc(sizeof(double) * D * D * D);
    int err = 0;
    for (int k = 0; k < D; ++k) {
        int e = min_faktor(D, mu, var, k, &logP[k], dMu + (size_t)k * D, dMM + (size_t)k * D * D,
                           dSg + (size_t)k * T, sweeps ? &sweeps[k] : NULL);
        if (e) { err = e; break; }
    }
    if (err) { free(dMu); free(dSg); free(dMM); return err; }
    for (int k = 0; k < D; ++k) if (isinf(logP[k])) logP[k] = -500;
    double* ep = malloc(sizeof(double) * D);
    double Z = 0, mx = -INFINITY;
    for (int k = 0; k < D; ++k) { ep[k] = exp(logP[k]); Z += ep[k]; if (logP[k] > mx) mx = logP[k]; }
    double se = 0;
    for (int k = 0; k < D; ++k) se += exp(logP[k] - mx);
    double s = mx + log(se);
    if (isinf(s)) s = mx;
    for (int k = 0; k < D; ++k) logP[k] = logP[k] - s;
    if (with_derivatives) {
        double* Zm = calloc(D, sizeof(double));
        double* Zs = calloc(T, sizeof(double));
        for (int k = 0; k < D; ++k) {
            for (int j = 0; j < D; ++j) Zm[j] += ep[k] * dMu[k * D + j];
            for (int t = 0; t < T; ++t) Zs[t] += ep[k] * dSg[(size_t)k * T + t];
        }
        for (int j = 0; j < D; ++j) Zm[j] /= Z;
        for (int t = 0; t < T; ++t) Zs[t] /= Z;
        for (int k = 0; k < D; ++k) {
            for (int j = 0; j < D; ++j) dlogPdMu[k * D + j] = dMu[k * D + j] - Zm[j];
            for (int t = 0; t < T; ++t) dlogPdSigma[(size_t)k * T + t] = dSg[(size_t)k * T + t] - Zs[t];
        }
        double* gg = calloc((size_t)D * D, sizeof(double));
        for (int k = 0; k < D; ++k)
            for (int i = 0; i < D; ++i)
                for (int j = 0; j < D; ++j)
                    gg[i * D + j] += (dMM[((size_t)k * D + i) * D + j] + dMu[k * D + i] * dMu[k * D + j]) * ep[k];
        for (int k = 0; k < D; ++k)
            for (int i = 0; i < D; ++i)
                for (int j = 0; j < D; ++j)   /* adds = -gg/Z + (Zm.T * Zm)[j]  (elementwise, epmgp.py:126) */
                    dlogPdMudMu[((size_t)k * D + i) * D + j] =
                        dMM[((size_t)k * D + i) * D + j] + (-gg[i * D + j] / Z + Zm[j] * Zm[j]);
        free(Zm); free(Zs); free(gg);
    }
    free(ep); free(dMu); free(dSg); free(dMM);
    return 0;
}
