// Matern-5/2 evaluation microbenchmark: how many covariance entries per clock one SM (256 threads,
// one CTA) generates, as a function of instruction-level parallelism per thread and of the sqrt / exp
// implementation; plus dependent-issue latencies of DFMA, DMMA and the rsqrt seed.  The covariance
// build and the dK contraction of the GP kernels are made of exactly this inner loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o matern_bench matern_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../fabolas_b200/csrc/gp_common.cuh"
using namespace fab;

// ---- variants ----------------------------------------------------------------------------------
struct V_lib {     // library sqrt + library exp
    static __device__ __forceinline__ double f(double r2, const double*) {
        const double r = sqrt(5.0 * r2);
        return (1.0 + r + (5.0 / 3.0) * r2) * exp(-r);
    }
};
struct V_cur {     // what the kernels use
    static __device__ __forceinline__ double f(double r2, const double*) { return matern52_value(r2); }
};
struct V_new {     // fab::matern52_fast
    static __device__ __forceinline__ double f(double r2, const double* tab) { return matern52_fast_value(r2, tab); }
};

template <typename V, int ILP>
__global__ void __launch_bounds__(256, 1) k_eval(double* out, int iters, double x0) {
    __shared__ double tab[EXP_TAB_N];
    for (int i = threadIdx.x; i < EXP_TAB_N; i += 256) tab[i] = exp2(-(double)i / EXP_TAB_N);
    __syncthreads();
    double x[ILP], acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = x0 * (1 + threadIdx.x + 256 * i) * 1e-3; acc[i] = 0.0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            acc[i] += V::f(x[i], tab);
            x[i] = fma(x[i], 1.0000001, 1e-7);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <typename V, int ILP>
void run(const char* name) {
    double* out; cudaMalloc(&out, 148 * 256 * 8);
    const int iters = 4000 / ILP * 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_eval<V, ILP><<<148, 256>>>(out, 10, 0.37);
    cudaEventRecord(e0); k_eval<V, ILP><<<148, 256>>>(out, iters, 0.37); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double evals_per_sm = (double)iters * ILP * 256;
    const double clk = ms * 1e-3 * 1965e6;
    printf("{\"variant\": \"%s\", \"ilp\": %d, \"cycles_per_eval_per_sm\": %.3f, \"us_per_4096_tile\": %.3f}\n", name, ILP,
           clk / evals_per_sm, ms * 1e3 / evals_per_sm * 4096);
    cudaFree(out);
}

// ---- latency probes (one warp) -----------------------------------------------------------------
__global__ void k_lat(double* out, long long* cyc, double s) {
    double a = threadIdx.x * 1e-3 + 1.0, b = s, c = 1e-9;
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) a = fma(a, b, c);
    long long t1 = clock64();
    double d0 = a, d1 = a;
#pragma unroll
    for (int i = 0; i < 256; ++i) dmma(d0, d1, b, c);
    long long t2 = clock64();
    double y = a + 2.0;
#pragma unroll
    for (int i = 0; i < 64; ++i) { double z; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(z) : "d"(y)); y = z + 1.5; }
    long long t3 = clock64();
    double q = y + 3.0;
#pragma unroll
    for (int i = 0; i < 64; ++i) q = rsqrt_fast(q) + 1.5;
    long long t4 = clock64();
    out[threadIdx.x] = a + d0 + d1 + y + q;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
}

// max relative error of the fast path against the library sqrt/exp over log-spaced r2 in [1e-30, 2e4]
__global__ void k_acc(double* err, int n) {
    __shared__ double tab[EXP_TAB_N];
    exp_table_fill(tab);
    __syncthreads();
    double emax = 0.0, gmax = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double r2 = (i == 0) ? 0.0 : exp(log(1e-30) + (log(2e4) - log(1e-30)) * i / n);
        const double r = sqrt(5.0 * r2), e = exp(-r);
        const double m_ref = (1.0 + r + (5.0 / 3.0) * r2) * e, g_ref = (5.0 / 6.0) * (1.0 + r) * e;
        double m, g;
        matern52_fast(r2, m, g, tab);
        if (m_ref > 1e-290) emax = fmax(emax, fabs(m - m_ref) / m_ref);
        if (g_ref > 1e-290) gmax = fmax(gmax, fabs(g - g_ref) / g_ref);
    }
    emax = warp_max(emax); gmax = warp_max(gmax);
    if ((threadIdx.x & 31) == 0) { atomicMax((unsigned long long*)&err[0], __double_as_longlong(emax)); atomicMax((unsigned long long*)&err[1], __double_as_longlong(gmax)); }
}

int main() {
    {
        double* err; cudaMalloc(&err, 16); cudaMemset(err, 0, 16);
        k_acc<<<148, 256>>>(err, 1 << 24);
        double h[2]; cudaMemcpy(h, err, 16, cudaMemcpyDeviceToHost);
        printf("{\"fast_matern_max_rel_err_value\": %.3e, \"fast_matern_max_rel_err_gfac\": %.3e}\n", h[0], h[1]);
    }
    double* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
    k_lat<<<1, 32>>>(out, cyc, 0.999); k_lat<<<1, 32>>>(out, cyc, 0.999);
    long long h[4]; cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
    printf("{\"dfma_dep_latency\": %.1f, \"dmma884_dep_latency\": %.1f, \"rsqrt_seed_plus_dadd\": %.1f, \"rsqrt_fast_plus_dadd\": %.1f}\n",
           h[0] / 256.0, h[1] / 256.0, h[2] / 64.0, h[3] / 64.0);
    run<V_lib, 1>("lib sqrt+exp"); run<V_lib, 4>("lib sqrt+exp");
    run<V_cur, 1>("current"); run<V_cur, 2>("current"); run<V_cur, 4>("current"); run<V_cur, 8>("current");
    run<V_new, 1>("fast"); run<V_new, 2>("fast"); run<V_new, 4>("fast"); run<V_new, 8>("fast");
    return 0;
}
