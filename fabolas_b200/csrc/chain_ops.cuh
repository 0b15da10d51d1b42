// chain_ops.cuh -- the latency-bound serial work of one diagonal tile, written for ONE warp (the
// "chain" warp of gp_dag_kernel) so that it runs concurrently with the eight throughput warps:
//   potrf64_warp : in-place Cholesky of a 64x64 tile (8x8 register Cholesky on one lane, row solves
//                  one lane per row, DMMA rank-8 trailing updates), producing the 8x8 block inverses
//   inv64_warp   : W = L^-1 of the factored tile into another tile, row-block by row-block; all
//                  column blocks of a row are independent DMMA chains (instruction-level parallelism)
//   matvec helpers for z = W rhs and alpha = W^T z.
// No block-level barriers: __syncwarp only.
#pragma once
#include "tiles.cuh"

namespace fab {

// Cholesky of tile A (lower part valid; strictly-upper 8x8 blocks are ignored and left untouched).
// dinv[8][64]: inverses of the eight 8x8 diagonal blocks of L (row-major).  rinv: 8 doubles scratch.
// Returns (in every lane) the 1-based tile-local index of the first non-positive pivot, or 0.
__device__ __forceinline__ int potrf64_warp(double* A, double* dinv, double* rinv) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    int fail = 0;
    for (int jb = 0; jb < 8; ++jb) {
        PHBC(60);
        if (lane == 0) {
            const int f = potrf8(A, jb, rinv);
            if (f && fail == 0) fail = f;
        }
        __syncwarp();
        PHEC(60);
        PHBC(61);
        // 8x8 inverse (every lane computes column lane & 7; identical values, lanes 0..7 store) and the
        // row solves x = a L8^-T of the rows below the block (lane and lane + 32), as one instruction stream
        {
            const int c = lane & 7;
            double xi[8];
            const int o = 8 * jb;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < r; ++k) s = fma(A[tix(o + r, o + k)], xi[k], s);
                const double ri = rinv[r];
                xi[r] = (r < c) ? 0.0 : ((r == c) ? ri : -s * ri);
            }
            if (lane < 8) {
#pragma unroll
                for (int r = 0; r < 8; ++r) dinv[64 * jb + r * 8 + c] = xi[r];
            }
            double l[8][8], ri[8];
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                ri[cc] = rinv[cc];
#pragma unroll
                for (int k = 0; k < 8; ++k) l[cc][k] = (k < cc) ? A[tix(o + cc, o + k)] : 0.0;
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int row = 8 * (jb + 1) + lane + 32 * half;
                if (row < TS) {
                    double x[8];
#pragma unroll
                    for (int cc = 0; cc < 8; cc += 2) {
                        const double2 v = *reinterpret_cast<const double2*>(&A[tix(row, o + cc)]);
                        x[cc] = v.x; x[cc + 1] = v.y;
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        x[k] *= ri[k];
#pragma unroll
                        for (int cc = k + 1; cc < 8; ++cc) x[cc] = fma(-x[k], l[cc][k], x[cc]);
                    }
#pragma unroll
                    for (int cc = 0; cc < 8; cc += 2) {
                        double2 v; v.x = x[cc]; v.y = x[cc + 1];
                        *reinterpret_cast<double2*>(&A[tix(row, o + cc)]) = v;
                    }
                }
            }
        }
        __syncwarp();
        PHEC(61);
        PHBC(62);
        // trailing update inside the tile: block (ib, kb) -= L(ib, jb) L(kb, jb)^T, jb < kb <= ib;
        // four column blocks of a row at a time (independent accumulators)
        for (int ib = jb + 1; ib < 8; ++ib) {
            const double a0 = -A[tix(8 * ib + g, 8 * jb + t)], a1 = -A[tix(8 * ib + g, 8 * jb + 4 + t)];
            for (int kb0 = jb + 1; kb0 <= ib; kb0 += 4) {
                double c0[4], c1[4], b0[4], b1[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (kb0 + u <= ib) {
                        blk_load(c0[u], c1[u], A, ib, kb0 + u);
                        b0[u] = A[tix(8 * (kb0 + u) + g, 8 * jb + t)];
                        b1[u] = A[tix(8 * (kb0 + u) + g, 8 * jb + 4 + t)];
                    }
#pragma unroll
                for (int u = 0; u < 4; ++u) if (kb0 + u <= ib) dmma(c0[u], c1[u], a0, b0[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) if (kb0 + u <= ib) dmma(c0[u], c1[u], a1, b1[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) if (kb0 + u <= ib) blk_store(c0[u], c1[u], A, ib, kb0 + u);
            }
            // (blocks of row ib read only column jb of rows ib and kb <= ib, which this step does not write)
        }
        __syncwarp();
        PHEC(62);
    }
    return __shfl_sync(0xffffffffu, fail, 0);
}

// W = L^-1 (64x64, lower triangular) from the factored tile L and the 8x8 diagonal-block inverses.
// The strictly-upper blocks of W are zero-filled (the tile products read whole tiles).
//   W(ib, ib) = inv8(ib);   W(ib, s) = -inv8(ib) * sum_{kb = s}^{ib - 1} L(ib, kb) W(kb, s),  s < ib
template <int IB>
__device__ __forceinline__ void inv64_row(double* W, const double* __restrict__ L, const double* __restrict__ dinv) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    // zero blocks right of the diagonal, inverse block on it
#pragma unroll
    for (int s = IB + 1; s < 8; ++s) blk_store(0.0, 0.0, W, IB, s);
    {
        const double2 v = *reinterpret_cast<const double2*>(&dinv[64 * IB + g * 8 + 2 * t]);
        blk_store(v.x, v.y, W, IB, IB);
    }
    if (IB == 0) return;
    constexpr int NS = IB > 0 ? IB : 1;
    double c0[NS], c1[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) c0[s] = c1[s] = 0.0;
#pragma unroll
    for (int kb = 0; kb < IB; ++kb) {
#pragma unroll
        for (int k0 = 0; k0 < 8; k0 += 4) {
            const double a = L[tix(8 * IB + g, 8 * kb + k0 + t)];
#pragma unroll
            for (int s = 0; s <= kb; ++s) dmma(c0[s], c1[s], a, W[tix(8 * kb + k0 + t, 8 * s + g)]);
        }
    }
    // T_s through the tile (accumulator -> B operand), then Y_s = -inv8(IB) T_s
#pragma unroll
    for (int s = 0; s < NS; ++s) blk_store(c0[s], c1[s], W, IB, s);
    __syncwarp();
    double y0[NS], y1[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) y0[s] = y1[s] = 0.0;
#pragma unroll
    for (int k0 = 0; k0 < 8; k0 += 4) {
        const double a = -dinv[64 * IB + g * 8 + k0 + t];
#pragma unroll
        for (int s = 0; s < NS; ++s) dmma(y0[s], y1[s], a, W[tix(8 * IB + k0 + t, 8 * s + g)]);
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < NS; ++s) blk_store(y0[s], y1[s], W, IB, s);
}

__device__ __forceinline__ void inv64_warp(double* W, const double* __restrict__ L, const double* __restrict__ dinv) {
    inv64_row<0>(W, L, dinv); __syncwarp();
    inv64_row<1>(W, L, dinv); __syncwarp();
    inv64_row<2>(W, L, dinv); __syncwarp();
    inv64_row<3>(W, L, dinv); __syncwarp();
    inv64_row<4>(W, L, dinv); __syncwarp();
    inv64_row<5>(W, L, dinv); __syncwarp();
    inv64_row<6>(W, L, dinv); __syncwarp();
    inv64_row<7>(W, L, dinv); __syncwarp();
}

// z = W v for a lower-triangular tile W and a 64-vector v in shared memory; lane l returns rows l and l + 32.
__device__ __forceinline__ void tile_matvec_warp(const double* __restrict__ W, const double* __restrict__ v, double& z0, double& z1) {
    const int lane = threadIdx.x & 31;
    const double v0 = v[lane], v1 = v[lane + 32];
    z0 = 0.0; z1 = 0.0;
#pragma unroll 4
    for (int r = 0; r < TS; ++r) {
        double s = W[tix(r, lane)] * v0;
        if (r >= 32) s = fma(W[tix(r, lane + 32)], v1, s);       // columns 32.. of rows < 32 are zero
        s = warp_sum(s);
        if ((r & 31) == lane) { if (r < 32) z0 = s; else z1 = s; }
    }
}
// a = W^T z: lane l returns columns l and l + 32.  z in shared memory.
__device__ __forceinline__ void tile_matvec_t_warp(const double* __restrict__ W, const double* __restrict__ z, double& a0, double& a1) {
    const int lane = threadIdx.x & 31;
    a0 = 0.0; a1 = 0.0;
#pragma unroll 8
    for (int r = 0; r < TS; ++r) {
        const double zr = z[r];
        a0 = fma(W[tix(r, lane)], zr, a0);
        a1 = fma(W[tix(r, lane + 32)], zr, a1);
    }
}

}  // namespace fab
