// Latency of the serial 8x8 register Cholesky (tiles.cuh potrf8) on one lane: alone, and while the eight
// throughput warps of the same SM stream DMMA tile products (FP64 pipe contention on the shared SMSP).
#include <cstdio>
#include "../../fabolas_b200/csrc/tiles.cuh"
using namespace fab;

__global__ void __launch_bounds__(288, 1) k(long long* out, int iters, int busy) {
    extern __shared__ __align__(1024) unsigned char smem[];
    double* A = reinterpret_cast<double*>(smem);
    double* B = A + TILE_ELEMS;
    double* T = B + TILE_ELEMS;
    __shared__ double rinv[8];
    __shared__ int stop;
    for (int e = threadIdx.x; e < TILE_ELEMS; e += blockDim.x) { A[e] = 1e-3 * (e % 13); B[e] = 1e-3 * (e % 7); }
    for (int e = threadIdx.x; e < TILE_ELEMS; e += blockDim.x) { const int r = e / 64, c = e % 64; T[tix(r, c)] = (r == c) ? 4.0 : 0.01 * ((r * 7 + c * 3) % 5); }
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    const int w = threadIdx.x >> 5;
    if (w == 8) {
        long long t0 = clock64();
        int f = 0;
        for (int it = 0; it < iters; ++it) {
            if ((threadIdx.x & 31) == 0) {
                // restore the diagonal block (cheap) then factor it
                for (int r = 0; r < 8; ++r) for (int c = 0; c <= r; ++c) T[tix(r, c)] = (r == c) ? 4.0 : 0.01 * ((r * 7 + c * 3) % 5);
                f += potrf8(T, 0, rinv);
            }
            __syncwarp();
        }
        long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) { out[0] = (t1 - t0) / iters; out[1] = f; atomicExch(&stop, 1); }
    } else if (busy) {
        const WarpTile wt = warp_tile();
        double acc[4][2][2];
        acc_zero(acc);
        while (*(volatile int*)&stop == 0) tile_mma<false, true, false>(acc, A, B, wt, 0, TS);
        double s = 0;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 2; ++j) s += acc[i][j][0] + acc[i][j][1];
        if (s == 12345.678) out[2] = 1;
    }
}
int main() {
    long long* out; cudaMalloc(&out, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * TILE_BYTES);
    for (int busy = 0; busy < 2; ++busy) {
        k<<<1, 288, 3 * TILE_BYTES>>>(out, 200, busy);
        k<<<1, 288, 3 * TILE_BYTES>>>(out, 2000, busy);
        long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("{\"potrf8_cycles_per_call_incl_restore\": %lld, \"bulk_warps_streaming_dmma\": %d}\n", h[0], busy);
    }
    return 0;
}
