// Can the dependent FP64 chains of the covariance generation / the dK contraction hide behind a DMMA tile product
// when both are issued by the SAME eight warps (two accumulator sets, one fused op)?  One CTA of 384 threads per SM
// (same register budget as gp_dag_kernel, warps 8..11 idle), one named barrier per iteration like an op boundary.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fuse_bench tools/microbench/fuse_bench.cu
#define FAB_DAG_LAYOUT 0
#include <cstdio>
#include "../../fabolas_b200/csrc/gp_dag.cuh"
using namespace fab;

template <int MODE>
__global__ void __launch_bounds__(384, 1) k(double* out, int iters) {
    extern __shared__ __align__(1024) unsigned char smem[];
    double* A = reinterpret_cast<double*>(smem);
    double* B = A + TILE_ELEMS;
    double* etab = B + TILE_ELEMS;
    double* zs = etab + EXP_TAB_N;        // 5 axes + u, 256 rows
    double* al = zs + 6 * 256;
    double* gsum = al + 256;
    if (threadIdx.x >= 256) return;
    for (int e = threadIdx.x; e < TILE_ELEMS; e += 256) { A[e] = 1e-3 * (e % 13); B[e] = 1e-3 * (e % 7); }
    for (int j = threadIdx.x; j < EXP_TAB_N; j += 256) etab[j] = exp2(-(double)j / EXP_TAB_N);
    for (int e = threadIdx.x; e < 6 * 256; e += 256) zs[e] = 0.01 * ((e * 37) % 101);
    al[threadIdx.x] = 0.001 * threadIdx.x;
    if (threadIdx.x < 8 * (MAX_PK + 1)) gsum[threadIdx.x] = 0.0;
    named_bar_sync(1, 256);
    Hyper h; h.amp = 1.3; h.inv_g2 = 0.7; h.cb = 0.4; h.diag_add = 1e-3;
    const WarpTile wt = warp_tile();
    DagAcc acc, acc2;
    acc_zero(acc); acc_zero(acc2);
    double s = 0.0;
    for (int it = 0; it < iters; ++it) {
        const int I = 1 + (it & 1), J = 0;
        if (MODE == 0 || MODE == 3) {
            build_acc<5>(acc, zs + 64 * I, zs + 64 * J, zs + 5 * 256 + 64 * I, zs + 5 * 256 + 64 * J, 256, h, 256, 1, I, J, wt, etab);
        }
        if (MODE == 1 || MODE == 4) {
            contract_acc<5>(acc, gsum + (threadIdx.x >> 5) * (MAX_PK + 1), zs + 64 * I, zs + 64 * J, zs + 5 * 256 + 64 * I,
                            zs + 5 * 256 + 64 * J, al + 64 * I, al + 64 * J, 256, h, 256, 1, I, J, wt, etab);
        }
        if (MODE >= 2) tile_mma<true, false, false, 4, 2>(acc2, A, B, wt, 0, TS);
        if (MODE == 0 || MODE == 3) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) s += acc[i][j][0] + acc[i][j][1];
        } else {
            acc[it & 3][it & 1][0] += 1e-9;          // the contraction input changes with the iteration
        }
        named_bar_sync(1, 256);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) s += acc2[i][j][0] + acc2[i][j][1] + acc[i][j][0];
    out[blockIdx.x * 256 + threadIdx.x] = s + gsum[threadIdx.x & 63];
}
template <int MODE>
void run(const char* name) {
    double* out; cudaMalloc(&out, 148 * 256 * 8);
    const int smem = 2 * TILE_BYTES + 8 * (EXP_TAB_N + 6 * 256 + 256 + 128);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, 384, smem>>>(out, 10);
    const int iters = 2000;
    cudaEventRecord(e0); k<MODE><<<148, 384, smem>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"op\": \"%s\", \"us_per_iteration\": %.3f, \"err\": \"%s\"}\n", name, ms * 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    run<0>("build");
    run<1>("contract");
    run<2>("product (TN)");
    run<3>("build + product, one op");
    run<4>("contract + product, one op");
    return 0;
}
