// gp_common.cuh -- problem description shared by the GP kernels and the covariance function
// (george Matern-5/2 x dataset-size basis).
#pragma once
#include "tiles.cuh"

namespace fab {

constexpr int MAX_D = 8;        // Matern axes
constexpr int MAX_PK = MAX_D + 3;
constexpr int MAX_SLOTS = 6;    // 6 x 32 KB tile slots + vectors fit the 227 KB of one SM

// family 0: amp * M52(r2)                        theta = [log_c0, log_M_0..log_M_{d-1}]
// family 1: amp * M52(r2) * (u u'/gamma2 + cb)   theta = [..., log_gamma2, log_cb], u = column d
// amp = amp_mult * exp(log_c0): george sums a ConstantKernel over its axes (fabolas.py:57 writes the
// raw parameter into that slot; scalar*kernel builds log(c/ndim) -- oracle/george_oracle.py).
struct GpData {
    const double* X;   // N x D row-major (device)
    const double* y;   // N
    double mean;
    double amp_mult;
    int N, D, d, family, Pk;
    int NT, Np;        // tiles per edge, padded size 64*NT
};

struct GpWork {        // per-walker workspaces in HBM (L2-resident at the benchmark sizes)
    double* Lw;        // B x ntiles x 4096   swizzled lower tiles of L (later W = L^-1 in place)
    double* zvec;      // B x Np              z = L^-1 (y - mean)
    double* alpha;     // B x Np              alpha = K^-1 (y - mean)
    double* accw;      // B x Np              running sum_J L_IJ z_J of the forward substitution
};

struct Hyper {         // decoded hyper-parameters of one walker
    double amp, inv_g2, cb, diag_add;
    double scale[MAX_D];      // 1/sqrt(M_k)
};

__device__ __forceinline__ Hyper decode_hyper(const GpData& gd, const double* __restrict__ th, double yerr2) {
    Hyper h;
    h.amp = gd.amp_mult * exp(th[0]);
#pragma unroll
    for (int k = 0; k < MAX_D; ++k) h.scale[k] = (k < gd.d) ? exp(-0.5 * th[1 + k]) : 0.0;
    if (gd.family == 1) {
        h.inv_g2 = exp(-th[1 + gd.d]);
        h.cb = exp(th[2 + gd.d]);
    } else {
        h.inv_g2 = 0.0;
        h.cb = 1.0;
    }
    h.diag_add = yerr2 + kTiny;
    return h;
}

// exp(x) for x <= 0 (the only range the Matern kernels need), ~1 ulp: round-to-nearest range
// reduction with a two-part ln2, degree-13 Taylor polynomial on |r| <= ln2/2 (remainder 5e-18),
// exponent patched in with an integer add.  21 FP64-pipe instructions instead of the ~46 of the
// general-purpose library exp() that profiles/r01a showed dominating the build and gradient kernels.
// Results below 2^-1000 are flushed to zero (they are < 1e-301 of the prior variance).
__device__ __forceinline__ double exp_nonpos(double x) {
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52
    const double t = fma(x, 1.4426950408889634, MAGIC);
    const double kd = t - MAGIC;
    const long long k = (long long)(int)(__double_as_longlong(t) & 0xffffffffLL);
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;                       // 1/13!
    p = fma(p, r, 2.08767569878681e-09);                     // 1/12!
    p = fma(p, r, 2.505210838544172e-08);                    // 1/11!
    p = fma(p, r, 2.755731922398589e-07);                    // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);                   // 1/9!
    p = fma(p, r, 2.48015873015873e-05);                     // 1/8!
    p = fma(p, r, 1.984126984126984e-04);                    // 1/7!
    p = fma(p, r, 1.388888888888889e-03);                    // 1/6!
    p = fma(p, r, 8.333333333333333e-03);                    // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);                   // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);                   // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double v = __longlong_as_double(__double_as_longlong(p) + (k << 52));
    return (x > -693.0) ? v : 0.0;                            // NaN input -> 0 never reached: callers pass finite r
}

// Matern-5/2 pieces from the squared scaled distance.
__device__ __forceinline__ void matern52(double r2, double& m, double& gfac) {
    const double r = sqrt(5.0 * r2);
    const double e = exp_nonpos(-r);
    m = (1.0 + r + (5.0 / 3.0) * r2) * e;
    gfac = (5.0 / 6.0) * (1.0 + r) * e;      // -dm/dr2
}
__device__ __forceinline__ double matern52_value(double r2) {
    const double r = sqrt(5.0 * r2);
    return (1.0 + r + (5.0 / 3.0) * r2) * exp_nonpos(-r);
}

// ---------------------------------------------------------------------------------------------
// Fast Matern-5/2 evaluation (the inner loop of the covariance build and of the dK contraction):
// 33 FP64-pipe instructions per entry instead of ~49, short dependent chains.
//   r = sqrt(5 r2): hardware rsqrt seed (~2^-20) + two coupled Newton steps on the root itself
//       (residual x - r^2 by FMA, correction e * y/2) => ~2^-60 relative; the +1e-300 (folded into
//       the FMA that forms 5 r2) makes r2 == 0 a regular input (r = 1e-150, m = 1 exactly).
//   exp(-r) = 2^-q * tab[j] * p(rr),  n = round(r * 128 / ln2) = 128 q + j,  |rr| <= ln2/256,
//       tab[j] = 2^(-j/128) from a 1 KB shared-memory table (exp_table_fill), degree-5 Taylor
//       polynomial (remainder 6e-19), exponent patched with an integer subtract.  ~1.5 ulp.
// r >= 690 (exp < 1e-299 of the prior variance) is flushed to zero.
// ---------------------------------------------------------------------------------------------
constexpr int EXP_TAB_N = 128;
__device__ __forceinline__ void exp_table_fill(double* tab) {   // all threads; caller syncs
    for (int j = threadIdx.x; j < EXP_TAB_N; j += blockDim.x) tab[j] = exp2(-(double)j / EXP_TAB_N);
}
__device__ __forceinline__ double sqrt5_fast(double r2) {       // sqrt(5 r2 + 1e-300), r2 >= 0 finite
    const double x = fma(5.0, r2, 1e-300);
#ifdef FAB_CUSIM
    return std::sqrt(x);
#else
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hy = 0.5 * y;
    double r = x * y;
    r = fma(fma(-r, r, x), hy, r);
    r = fma(fma(-r, r, x), hy, r);
    return r;
#endif
}
__device__ __forceinline__ double exp_neg_fast(double r, const double* __restrict__ tab) {   // exp(-r), r >= 0
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52
    const double t = fma(r, 184.66496523378731, MAGIC);      // 128 / ln2
    const double nd = t - MAGIC;
    const int n = (int)(__double_as_longlong(t) & 0xffffffffLL);
    double rr = fma(nd, 5.41521234663377981633e-03, -r);     // ln2_hi / 128 (32 significant bits: nd * hi exact)
    rr = fma(nd, 1.49079291349264664064e-12, rr);            // ln2_lo / 128
    double p = 8.333333333333333e-03;                        // 1/5!
    p = fma(p, rr, 4.1666666666666664e-02);
    p = fma(p, rr, 1.6666666666666666e-01);
    p = fma(p, rr, 0.5);
    p = fma(p, rr, 1.0);
    p = fma(p, rr, 1.0);
    const double v = tab[n & (EXP_TAB_N - 1)] * p;
    const double w = __longlong_as_double(__double_as_longlong(v) - ((long long)(n >> 7) << 52));
    return (r < 690.0) ? w : 0.0;
}
__device__ __forceinline__ double matern52_fast_value(double r2, const double* __restrict__ tab) {
    const double r = sqrt5_fast(r2);
    return fma(r2, 5.0 / 3.0, 1.0 + r) * exp_neg_fast(r, tab);
}
__device__ __forceinline__ void matern52_fast(double r2, double& m, double& gfac, const double* __restrict__ tab) {
    const double r = sqrt5_fast(r2);
    const double e = exp_neg_fast(r, tab);
    const double opr = 1.0 + r;
    m = fma(r2, 5.0 / 3.0, opr) * e;
    gfac = (5.0 / 6.0) * opr * e;            // -dm/dr2
}

// Scaled coordinates of the 64 rows and the 64 columns of the covariance tile being generated,
// staged in shared memory as structure-of-arrays slices: z[k * 64 + r] = x_rk / sqrt(M_k) (k < d)
// followed by u[r] (basis column, family 1; zero otherwise).  Slices are O(64 (d+1)) regardless of N.
__host__ __device__ constexpr int slice_doubles(int d) { return (d + 1) * TS; }

// Stage the slice of tile-row `I` (observations 64 I .. 64 I + 63) into `buf`; all threads, no barrier.
__device__ __forceinline__ void stage_slice(double* buf, const GpData& gd, const Hyper& h, int I, int nthr = NTHREADS) {
    const int D = gd.D, d = gd.d;
    for (int e = threadIdx.x; e < TS * D; e += nthr) {
        const int r = e / D, k = e - r * D;
        const int gi = TS * I + r;
        const double x = (gi < gd.N) ? __ldg(&gd.X[(size_t)gi * D + k]) : 0.0;
        if (k < d) buf[k * TS + r] = x * h.scale[k];
        else buf[d * TS + r] = x;                       // k == d only happens for family 1 (D = d + 1)
    }
    if (gd.family != 1)
        for (int r = threadIdx.x; r < TS; r += nthr) buf[d * TS + r] = 0.0;
}

}  // namespace fab
