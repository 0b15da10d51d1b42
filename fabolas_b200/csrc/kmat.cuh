// kmat.cuh -- kernel F: materialised covariance matrices k(Xa, Xb) (+ d/dtheta) for B hyper-parameter
// rows at once.  This is the "kernel-matrix builder" in its HBM-write-bound form (8 * Na * Nb bytes per
// model and per gradient slice): used by the ard.py surface (AutomaticRelevanceDetermination.__call__,
// ard.py:82-180), by GP.sample's prior covariance (fabolas.py:199) and for inspection.  The
// likelihood / prediction kernels never call it -- they regenerate tiles in shared memory.
//   nu_code selects the radial function of ard.py:119-135 on r2 = sum_k (x_k - x'_k)^2 / M_k:
//     0: Matern-5/2 (george Matern52Kernel and ard nu=2.5)   1: nu=1.5   2: nu=0.5   3: nu=inf (RBF)
#pragma once
#include "gp_common.cuh"
#include "mcmc.cuh"

namespace fab {

struct KmatArgs {
    int family, D, d, Pk, nu_code;
    double nu;             // nu_code 4: the smoothness of the general Matern kernel
    double amp_mult;
    const double* theta;   // B x Pk
    const double* Xa; int Na;
    const double* Xb; int Nb;
    double* K;             // B x Na x Nb
    double* dK;            // B x Na x Nb x Pk or null
};

// Modified Bessel function of the second kind K_nu(x), x > 0, any real nu >= 0 (the reference's general-nu branch
// calls scipy.special.kv, ard.py:129-135): K_nu(x) = int_0^inf exp(-x cosh t) cosh(nu t) dt by the trapezoid rule, which
// converges geometrically for this entire integrand: step min(0.2, 0.7 / sqrt(x)) (error ~ exp(-2 pi^2 / (x h^2)) resp.
// exp(-pi^2 / h)), summed until the terms have dropped 1e-18 below the sum past the peak of the integrand.
__device__ inline double bessel_k(double nu, double x) {
    const double h = fmin(0.2, 0.7 / sqrt(x));
    const double t_peak = asinh(nu / x);                       // maximum of nu t - x cosh t
    double sum = 0.5 * exp(-x);                                 // t = 0 with the trapezoid weight 1/2
    for (int i = 1; i < 200000; ++i) {
        const double t = i * h, c = x * cosh(t);
        const double term = 0.5 * (exp(nu * t - c) + exp(-nu * t - c));
        sum += term;
        if (t > t_peak && term < 1e-18 * sum) break;
    }
    return h * sum;
}

__device__ __forceinline__ void radial(int nu_code, double r2, double& m, double& gfac, double nu = 0.0) {
    // m = k(r2), gfac = -dk/dr2
    if (nu_code == 4) {
        // ard.py:129-135: strict zeros get eps, tmp = sqrt(2 nu) dist, K = 2^(1-nu) / Gamma(nu) * tmp^nu * K_nu(tmp)
        double dist = sqrt(r2);
        if (dist == 0.0) dist += 2.220446049250313e-16;
        const double tmp = sqrt(2.0 * nu) * dist;
        m = (exp2(1.0 - nu) / tgamma(nu)) * pow(tmp, nu) * bessel_k(nu, tmp);
        gfac = 0.0;            // (the reference differentiates this branch numerically, ard.py:168-173)
    } else if (nu_code == 0) {
        matern52(r2, m, gfac);
    } else if (nu_code == 1) {
        const double r = sqrt(3.0 * r2), e = exp(-r);
        m = (1.0 + r) * e; gfac = 1.5 * e;
    } else if (nu_code == 2) {
        const double r = sqrt(r2), e = exp(-r);
        m = e; gfac = (r > 0.0) ? 0.5 * e / r : 0.0;
    } else {
        m = exp(-0.5 * r2); gfac = 0.5 * m;
    }
}

// grid (ceil(Na*Nb / 256), B); consecutive threads -> consecutive j: coalesced 8-byte stores.
__global__ void __launch_bounds__(256) kmat_kernel(KmatArgs a) {
    const int b = blockIdx.y;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)a.Na * a.Nb) return;
    const int i = (int)(e / a.Nb), j = (int)(e - (long long)i * a.Nb);
    const double* th = a.theta + (size_t)b * a.Pk;
    const double amp = a.amp_mult * exp(th[0]);
    double r2 = 0.0, tk[MAX_D];
#pragma unroll
    for (int k = 0; k < MAX_D; ++k) {
        tk[k] = 0.0;
        if (k < a.d) {
            const double s = exp(-0.5 * th[1 + k]);
            const double dz = a.Xa[(size_t)i * a.D + k] * s - a.Xb[(size_t)j * a.D + k] * s;
            tk[k] = dz * dz;
            r2 += tk[k];
        }
    }
    double m, gf;
    radial(a.nu_code, r2, m, gf, a.nu);
    double uu = 0.0, cb = 1.0, beta = 1.0;
    if (a.family == 1) {
        uu = a.Xa[(size_t)i * a.D + a.d] * a.Xb[(size_t)j * a.D + a.d] * exp(-th[1 + a.d]);
        cb = exp(th[2 + a.d]);
        beta = uu + cb;
    }
    const size_t o = ((size_t)b * a.Na + i) * a.Nb + j;
    a.K[o] = amp * m * beta;
    if (a.dK) {
        double* g = a.dK + o * a.Pk;
        g[0] = amp * m * beta;
        for (int k = 0; k < a.d; ++k) g[1 + k] = amp * beta * gf * tk[k];
        if (a.family == 1) {
            g[1 + a.d] = -amp * m * uu;
            g[2 + a.d] = amp * m * cb;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// GP.sample(t) of every resident model at its own points, reduced to the maximum: the incumbent that
// fabolas.get_candidate / entropy_search hand to expected_improvement is max of ONE draw from the PRIOR
// N(mean, k(t, t) + 1.25e-12 I) at the representer points (fabolas.py:198-199, entropy_search.py:137; quirk Q5).
// One CTA per model: covariance of the Z <= 64 points in shared memory (same evaluation as kmat_kernel), in-place
// Cholesky (a non-positive pivot -- points that coincide to rounding -- zeroes its column: the factor of the
// positive semi-definite part, what numpy's SVD-based multivariate_normal samples from), standard normals by
// Box-Muller on the Philox4x32-10 stream of mcmc.cuh keyed by (seed, model, pair), sample = mean + L z.
// ---------------------------------------------------------------------------------------------------
struct PriorDrawArgs {
    GpData gd;
    const double* theta;      // B x Pk resident hyper-parameters
    const double* reps;       // B x Z x D
    int Z;
    unsigned long long seed;
    double* out_max;          // B
    double* sample;           // B x Z or null
    int* info;                // B or null: number of pivots that had to be zeroed
};
constexpr int PRIOR_DRAW_PURPOSE = 7;     // Philox counter word c2 (0..4 belong to the sampler)

__host__ __device__ inline void prior_draw_normals(unsigned long long seed, int b, int pair, double& z0, double& z1) {
    const Philox4 r = philox4x32_10(seed, (unsigned)b, (unsigned)pair, (unsigned)PRIOR_DRAW_PURPOSE, 0u);
    const double u1 = uniform53(r.x, r.y), u2 = uniform53(r.z, r.w);
    const double rad = sqrt(-2.0 * log(u1));
    z0 = rad * cos(6.283185307179586 * u2);
    z1 = rad * sin(6.283185307179586 * u2);
}

__global__ void __launch_bounds__(128) prior_draw_kernel(PriorDrawArgs a) {
    __shared__ double Ks[64 * 65];
    __shared__ double zs[64], xs[64 * (MAX_D + 1)], red[4];
    __shared__ int nzero;
    const int b = blockIdx.x, tid = threadIdx.x, Z = a.Z, D = a.gd.D, d = a.gd.d;
    const Hyper h = decode_hyper(a.gd, a.theta + (size_t)b * a.gd.Pk, 0.0);
    for (int e = tid; e < Z * D; e += blockDim.x) {
        const int i = e / D, k = e - i * D;
        const double x = a.reps[((size_t)b * Z + i) * D + k];
        xs[i * (MAX_D + 1) + k] = k < d ? x * h.scale[k] : x;
    }
    if (tid == 0) nzero = 0;
    __syncthreads();
    for (int e = tid; e < Z * Z; e += blockDim.x) {
        const int i = e / Z, j = e - i * Z;
        if (j > i) continue;
        double r2 = 0.0;
        for (int k = 0; k < d; ++k) {
            const double dz = xs[i * (MAX_D + 1) + k] - xs[j * (MAX_D + 1) + k];
            r2 = fma(dz, dz, r2);
        }
        double v = h.amp * matern52_value(r2);
        if (a.gd.family == 1) v *= fma(xs[i * (MAX_D + 1) + d] * xs[j * (MAX_D + 1) + d], h.inv_g2, h.cb);
        Ks[i * 65 + j] = (i == j) ? v + kTiny : v;
    }
    if (tid < (Z + 1) / 2) {
        double z0, z1;
        prior_draw_normals(a.seed, b, tid, z0, z1);
        zs[2 * tid] = z0;
        if (2 * tid + 1 < Z) zs[2 * tid + 1] = z1;
    }
    __syncthreads();
    for (int j = 0; j < Z; ++j) {          // right-looking Cholesky, lower triangle in place
        const double p = Ks[j * 65 + j];
        const bool ok = p > 0.0;
        const double ri = ok ? 1.0 / sqrt(p) : 0.0;
        __syncthreads();
        if (tid == 0) { Ks[j * 65 + j] = ok ? sqrt(p) : 0.0; if (!ok) nzero += 1; }
        for (int i = j + 1 + tid; i < Z; i += blockDim.x) Ks[i * 65 + j] *= ri;
        __syncthreads();
        for (int e = tid; e < (Z - j - 1) * (Z - j - 1); e += blockDim.x) {
            const int i = j + 1 + e / (Z - j - 1), k = j + 1 + e % (Z - j - 1);
            if (k <= i) Ks[i * 65 + k] = fma(-Ks[i * 65 + j], Ks[k * 65 + j], Ks[i * 65 + k]);
        }
        __syncthreads();
    }
    double best = -INFINITY;
    for (int i = tid; i < Z; i += blockDim.x) {
        double sv = a.gd.mean;
        for (int k = 0; k <= i; ++k) sv = fma(Ks[i * 65 + k], zs[k], sv);
        if (a.sample) a.sample[(size_t)b * Z + i] = sv;
        best = fmax(best, sv);
    }
    best = warp_max(best);
    if ((tid & 31) == 0) red[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
        a.out_max[b] = fmax(fmax(red[0], red[1]), fmax(red[2], red[3]));
        if (a.info) a.info[b] = nzero;
    }
}

}  // namespace fab
