// Isolated throughput of the 64x64x64 DMMA tile product (tiles.cuh tile_mma) per SM, all operands in
// shared memory, 8 warps -- tells how far the GEMM inner loop itself is from the DMMA pipe peak.
#include <cstdio>
#include "../../fabolas_b200/csrc/tiles.cuh"
using namespace fab;

template <bool TA, bool TB>
__global__ void __launch_bounds__(NTHREADS, 1) k_mma(double* out, int iters) {
    extern __shared__ __align__(1024) unsigned char smem[];
    double* A = reinterpret_cast<double*>(smem);
    double* B = A + TILE_ELEMS;
    for (int e = threadIdx.x; e < TILE_ELEMS; e += NTHREADS) { A[e] = 1e-3 * (e % 13); B[e] = 1e-3 * (e % 7); }
    __syncthreads();
    const WarpTile wt = warp_tile();
    double acc[4][2][2];
    acc_zero(acc);
    for (int it = 0; it < iters; ++it) tile_mma<TA, TB, false>(acc, A, B, wt, 0, TS);
    double s = 0;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 2; ++j) s += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * NTHREADS + threadIdx.x] = s;
}

template <typename K>
void run(const char* name, K kern) {
    double* out; cudaMalloc(&out, 148 * NTHREADS * 8);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE_BYTES);
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<148, NTHREADS, 2 * TILE_BYTES>>>(out, 10);
    cudaEventRecord(e0); kern<<<148, NTHREADS, 2 * TILE_BYTES>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 64 * 64 * 64 * iters * 148;
    printf("%s: %.2f us per tile product per SM, %.2f TFLOP/s chip (%.1f%% of 36.96)\n", name, ms * 1e3 / iters,
           flops / ms * 1e-9, flops / ms * 1e-9 / 36.96 * 100);
}

int main() {
    run("NT (A row-frag, B row-frag)", k_mma<false, true>);
    run("NN (A row-frag, B col-frag)", k_mma<false, false>);
    run("TN (A col-frag, B col-frag)", k_mma<true, false>);
    return 0;
}
