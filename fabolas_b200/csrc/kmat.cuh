// kmat.cuh -- kernel F: materialised covariance matrices k(Xa, Xb) (+ d/dtheta) for B hyper-parameter
// rows at once.  This is the "kernel-matrix builder" in its HBM-write-bound form (8 * Na * Nb bytes per
// model and per gradient slice): used by the ard.py surface (AutomaticRelevanceDetermination.__call__,
// ard.py:82-180), by GP.sample's prior covariance (fabolas.py:199) and for inspection.  The
// likelihood / prediction kernels never call it -- they regenerate tiles in shared memory.
//   nu_code selects the radial function of ard.py:119-135 on r2 = sum_k (x_k - x'_k)^2 / M_k:
//     0: Matern-5/2 (george Matern52Kernel and ard nu=2.5)   1: nu=1.5   2: nu=0.5   3: nu=inf (RBF)
#pragma once
#include "gp_common.cuh"

namespace fab {

struct KmatArgs {
    int family, D, d, Pk, nu_code;
    double amp_mult;
    const double* theta;   // B x Pk
    const double* Xa; int Na;
    const double* Xb; int Nb;
    double* K;             // B x Na x Nb
    double* dK;            // B x Na x Nb x Pk or null
};

__device__ __forceinline__ void radial(int nu_code, double r2, double& m, double& gfac) {
    // m = k(r2), gfac = -dk/dr2
    if (nu_code == 0) {
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
    radial(a.nu_code, r2, m, gf);
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

}  // namespace fab
