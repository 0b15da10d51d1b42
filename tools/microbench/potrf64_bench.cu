// potrf64_chain (chain_ops.cuh: in-tile Cholesky + triangular inverse of one 64 x 64 diagonal tile by the four chain
// warps) in isolation: cycles per call with the SM otherwise idle -- the situation of the device sampler at the
// reference's own size (N = 100, 20 walkers: the bulk warps wait for the chain) -- and next to eight warps that stream
// DMMA tile products.  Checks L L^T = A and W L = I on the way.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o potrf64_bench tools/microbench/potrf64_bench.cu
#include <cstdio>
#include <cmath>
#include <vector>
__device__ long long g_cht[8][6];
#define CHT(jb, k) do { if ((threadIdx.x & 31) == 0) g_cht[jb][k] = clock64(); } while (0)
#include "../../fabolas_b200/csrc/chain_ops.cuh"
using namespace fab;

__device__ __forceinline__ double a0(int r, int c) { const double d = r - c; return exp(-0.02 * d * d) + (r == c ? 0.05 : 0.0); }

__global__ void __launch_bounds__(384, 1) k(long long* out, double* Lout, double* Wout, int iters, int busy, int nb) {
    extern __shared__ __align__(1024) unsigned char smem[];
    double* A = reinterpret_cast<double*>(smem);
    double* W = A + TILE_ELEMS;
    double* B0 = W + TILE_ELEMS;
    double* B1 = B0 + TILE_ELEMS;
    __shared__ double dinv[512];
    __shared__ double rinv[64];
    __shared__ int fail, stop;
    for (int e = threadIdx.x; e < TILE_ELEMS; e += blockDim.x) { B0[e] = 1e-3 * (e % 13); B1[e] = 1e-3 * (e % 7); }
    if (threadIdx.x == 0) { fail = 0; stop = 0; }
    __syncthreads();
    const int w = threadIdx.x >> 5;
    if (w >= 8) {
        const int cw = w - 8, ct = threadIdx.x - 256;
        long long tsum = 0;
        for (int it = 0; it < iters; ++it) {
            for (int e = ct; e < TILE_ELEMS; e += 128) { const int r = e >> 6, c = e & 63; A[tix(r, c)] = (r < 8 * nb && c < 8 * nb) ? a0(r, c) : (r == c ? 1.0 : 0.0); }
            chain_bar();
            const long long t0 = clock64();
            potrf64_chain(A, W, dinv, rinv, &fail, cw, nb);
            tsum += clock64() - t0;
        }
        if (ct == 0) { out[0] = tsum / iters; out[1] = fail; atomicExch(&stop, 1); }
        chain_bar();
        for (int e = ct; e < TILE_ELEMS; e += 128) { const int r = e >> 6, c = e & 63; Lout[e] = A[tix(r, c)]; Wout[e] = W[tix(r, c)]; }
    } else if (busy) {
        const WarpTile wt = warp_tile();
        double acc[4][2][2];
        acc_zero(acc);
        while (*(volatile int*)&stop == 0) tile_mma<false, true, false>(acc, B0, B1, wt, 0, TS);
        double s = 0;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 2; ++j) s += acc[i][j][0] + acc[i][j][1];
        if (s == 12345.678) out[2] = 1;
    }
}
int main() {
    long long* out; double *Ld, *Wd;
    cudaMalloc(&out, 64); cudaMalloc(&Ld, TILE_BYTES); cudaMalloc(&Wd, TILE_BYTES);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * TILE_BYTES);
    std::vector<double> L(TILE_ELEMS), W(TILE_ELEMS);
    for (int nb : {8, 5})
        for (int busy = 0; busy < 2; ++busy) {
            k<<<1, 384, 4 * TILE_BYTES>>>(out, Ld, Wd, 50, busy, nb);
            k<<<1, 384, 4 * TILE_BYTES>>>(out, Ld, Wd, 1000, busy, nb);
            long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            cudaMemcpy(L.data(), Ld, TILE_BYTES, cudaMemcpyDeviceToHost); cudaMemcpy(W.data(), Wd, TILE_BYTES, cudaMemcpyDeviceToHost);
            double e1 = 0, e2 = 0;
            const int n = 8 * nb;
            for (int r = 0; r < 64; ++r)
                for (int c = 0; c <= r; ++c) {
                    double s = 0, t = 0;
                    for (int q = 0; q <= c; ++q) s += L[r * 64 + q] * L[c * 64 + q];
                    for (int q = c; q <= r; ++q) t += W[r * 64 + q] * L[q * 64 + c];
                    const double d = r - c, a = (r < n && c < n) ? exp(-0.02 * d * d) + (r == c ? 0.05 : 0.0) : (r == c ? 1.0 : 0.0);
                    e1 = fmax(e1, fabs(s - a)); e2 = fmax(e2, fabs(t - (r == c ? 1.0 : 0.0)));
                }
            if (nb == 8 && !busy) {
                long long t[8][6]; cudaMemcpyFromSymbol(t, g_cht, sizeof t);
                for (int j = 0; j < 8; ++j) printf("  step %d: potrf8 %lld, wait workers %lld, solve %lld, update %lld, to next step %lld\n", j, t[j][1] - t[j][0], t[j][2] - t[j][1], t[j][4] - t[j][2], t[j][3] - t[j][4], j < 7 ? t[j + 1][0] - t[j][3] : 0LL);
            }
            printf("{\"nb\": %d, \"cycles_per_call\": %lld, \"fail\": %lld, \"throughput_warps\": \"%s\", \"max|LL^T-A|\": %.2e, \"max|WL-I|\": %.2e, \"err\": \"%s\"}\n",
                   nb, h[0], h[1], busy ? "streaming DMMA tile products" : "idle", e1, e2, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
