// Latency of the serial 8x8 register Cholesky (tiles.cuh potrf8) on one lane: alone, and while the eight
// throughput warps of the same SM stream DMMA tile products (FP64 pipe contention on the shared SMSP).
#include <cstdio>
#include "../../fabolas_b200/csrc/tiles.cuh"
using namespace fab;

template <int CHAIN_WARP>
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
    if (w == CHAIN_WARP) {
        long long tsum = 0;
        int f = 0;
        for (int it = 0; it < iters; ++it) {
            if ((threadIdx.x & 31) == 0) {
                // restore the diagonal block (untimed) then factor it (timed)
                for (int r = 0; r < 8; ++r) for (int c = 0; c <= r; ++c) T[tix(r, c)] = (r == c) ? 4.0 : 0.01 * ((r * 7 + c * 3) % 5);
                __threadfence_block();
                const long long t0 = clock64();
                f += potrf8(T, 0, rinv);
                tsum += clock64() - t0;
            }
            __syncwarp();
        }
        if ((threadIdx.x & 31) == 0) { out[0] = tsum / iters; out[1] = f; atomicExch(&stop, 1); }
    } else if (busy == 2) {           // DFMA streams (covariance-build-like) on the throughput warps
        double x[8];
        for (int i = 0; i < 8; ++i) x[i] = 1.0 + 1e-3 * (threadIdx.x + i);
        while (*(volatile int*)&stop == 0) {
#pragma unroll
            for (int r = 0; r < 16; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fma(x[i], 0.999999, 1e-7);
        }
        double s = 0;
        for (int i = 0; i < 8; ++i) s += x[i];
        if (s == 12345.678) out[2] = 1;
    } else if (busy) {
        WarpTile wt;
        const int bw = w > CHAIN_WARP ? w - 1 : w;     // 8 throughput warps whatever the chain warp's index
        wt.m0 = (bw >> 2) * 32; wt.n0 = (bw & 3) * 16;
        double acc[4][2][2];
        acc_zero(acc);
        while (*(volatile int*)&stop == 0) tile_mma<false, true, false>(acc, A, B, wt, 0, TS);
        double s = 0;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 2; ++j) s += acc[i][j][0] + acc[i][j][1];
        if (s == 12345.678) out[2] = 1;
    }
}
template <int CW>
void run(long long* out) {
    cudaFuncSetAttribute(k<CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * TILE_BYTES);
    for (int busy = 0; busy < 3; ++busy) {
        k<CW><<<1, 288, 3 * TILE_BYTES>>>(out, 200, busy);
        k<CW><<<1, 288, 3 * TILE_BYTES>>>(out, 2000, busy);
        long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("{\"chain_warp_index\": %d, \"potrf8_cycles_per_call\": %lld, \"throughput_warps\": \"%s\"}\n", CW, h[0], busy == 0 ? "idle" : (busy == 1 ? "streaming DMMA tile products" : "streaming DFMA"));
    }
}
int main() {
    long long* out; cudaMalloc(&out, 64);
    run<8>(out);      // chain warp has the highest warp index of the CTA
    run<0>(out);      // ... the lowest
    return 0;
}
