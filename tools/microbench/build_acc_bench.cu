// build_acc / contract_acc (gp_dag.cuh) in isolation: one CTA of 256 threads per SM, coordinates resident in
// shared memory, nothing else running -- separates the cost of the code itself from interference inside
// gp_dag_kernel (chain warps, instruction cache).
#include <cstdio>
#include "../../fabolas_b200/csrc/gp_dag.cuh"
using namespace fab;

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(double* out, int iters) {
    __shared__ double etab[EXP_TAB_N];
    __shared__ double zs[6 * 256];        // 5 axes + u, 256 rows
    __shared__ double al[256];
    __shared__ double gsum[8 * (MAX_PK + 1)];
    exp_table_fill(etab);
    for (int e = threadIdx.x; e < 6 * 256; e += 256) zs[e] = 0.01 * ((e * 37) % 101);
    al[threadIdx.x] = 0.001 * threadIdx.x;
    if (threadIdx.x < 8 * (MAX_PK + 1)) gsum[threadIdx.x] = 0.0;
    __syncthreads();
    Hyper h; h.amp = 1.3; h.inv_g2 = 0.7; h.cb = 0.4; h.diag_add = 1e-3;
    const WarpTile wt = warp_tile();
    double acc[4][2][2];
    acc_zero(acc);
    double s = 0.0;
    for (int it = 0; it < iters; ++it) {
        const int I = 1 + (it & 1), J = 0;
        if (MODE == 0) {
            build_acc<5>(acc, zs + 64 * I, zs + 64 * J, zs + 5 * 256 + 64 * I, zs + 5 * 256 + 64 * J, 256, h, 256, 1, I, J, wt, etab);
            s += acc[it & 3][0][0];
        } else {
            contract_acc<5>(acc, gsum + (threadIdx.x >> 5) * (MAX_PK + 1), zs + 64 * I, zs + 64 * J, zs + 5 * 256 + 64 * I,
                            zs + 5 * 256 + 64 * J, al + 64 * I, al + 64 * J, 256, h, 256, 1, I, J, wt, etab);
        }
    }
    out[blockIdx.x * 256 + threadIdx.x] = s + gsum[threadIdx.x & 63];
}
template <int MODE>
void run(const char* name) {
    double* out; cudaMalloc(&out, 148 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, 256>>>(out, 10);
    const int iters = 2000;
    cudaEventRecord(e0); k<MODE><<<148, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"op\": \"%s\", \"us_per_tile_per_sm\": %.3f}\n", name, ms * 1e3 / iters);
}
int main() { run<0>("build_acc<5>"); run<1>("contract_acc<5>"); return 0; }
