// tiles.cuh -- 64x64 FP64 tiles in shared memory and the DMMA tile algebra the GP kernels are
// built from (tile GEMM in NT/NN/TN form with k-range restriction for triangular operands, accumulator
// fragment load/store, the 8x8 register Cholesky of the serial chain, 8x8 block helpers).
//
// Layout: a tile is 4096 doubles (32 KB), row-major with an XOR swizzle of the column index,
//     offset(r, c) = r * 64 + (c ^ ((r & 3) << 2)),
// which makes both DMMA operand access patterns bank-conflict free for 8-byte loads
//   row fragment  (lane 4g+t reads (r0+g, k0+t)) and  column fragment (lane 4g+t reads (k0+t, c0+g)),
// keeps (c, c+1) pairs with even c adjacent (16-byte accumulator loads/stores), and keeps a tile a
// single contiguous 32 KB blob so that one cp.async.bulk (TMA) moves it between L2/HBM and smem.
// The same swizzled blob is the on-HBM format of the per-walker factor workspace.
#pragma once
#include "prims.cuh"

namespace fab {

constexpr int TS = 64;                 // tile edge
constexpr int TILE_ELEMS = TS * TS;    // doubles per tile
constexpr int TILE_BYTES = TILE_ELEMS * 8;
constexpr int NWARPS = 8;              // all tile routines assume 256 threads
constexpr int NTHREADS = NWARPS * 32;

__device__ __forceinline__ int tix(int r, int c) { return r * TS + (c ^ ((r & 3) << 2)); }

// lower-triangular tile index (I >= J)
__host__ __device__ __forceinline__ int tri_index(int I, int J) { return I * (I + 1) / 2 + J; }

// ---------------------------------------------------------------------------------------------
// Accumulator fragment of one warp: (8*MI) rows x (8*NJ) cols = MI x NJ DMMA blocks.
// acc[i][j][e] = C[m0 + 8i + g][n0 + 8j + 2t + e].  Default 4 x 2 (32 x 16): 8 warps tile 64 x 64.
// LD* = row stride (doubles) of the swizzled operand; any multiple of 16 keeps the swizzle
// conflict-free, so 64-wide tiles and 32-wide panels share the code.
// ---------------------------------------------------------------------------------------------
template <int LD>
__device__ __forceinline__ int tixw(int r, int c) { return r * LD + (c ^ ((r & 3) << 2)); }

struct WarpTile {
    int m0, n0;   // origin of this warp's block inside the output tile
};
// 8 warps as 2 x 4 blocks of 32 x 16 over a 64 x 64 output.  Warps w and w + 4 share an SM sub-partition (and
// its DMMA pipe); the lower row of blocks is laid out in reverse column order so that the two column blocks of
// a sub-partition are 16 s and 16 (3 - s): products whose k-range depends on the column block (triangular
// right-hand operand: k < n0 + 16 or k >= n0) then cost every sub-partition the same 5/8 of a full product
// instead of costing one of them all of it (profiles/r01y_*).
__device__ __forceinline__ WarpTile warp_tile() {
    const int w = threadIdx.x >> 5;
    WarpTile wt;
    wt.m0 = (w >> 2) * 32;
    wt.n0 = ((w >> 2) ? 3 - (w & 3) : (w & 3)) * 16;
    return wt;
}
// 8 warps as 4 x 2 blocks of 16 x 16 over a 64 x 32 output.  Warps w and w + 4 share an SM sub-partition: the upper
// four warps take the row blocks in reverse order, so that the two row blocks of a sub-partition are r and 3 - r and a
// product whose k-range grows with the row (lower-triangular left operand) loads every sub-partition alike.
__device__ __forceinline__ WarpTile warp_tile_32() {
    const int w = threadIdx.x >> 5;
    WarpTile wt;
    wt.m0 = ((w < 4) ? (w >> 1) : 3 - ((w - 4) >> 1)) * 16;
    wt.n0 = (w & 1) * 16;
    return wt;
}

template <int MI, int NJ>
__device__ __forceinline__ void acc_zero(double (&acc)[MI][NJ][2]) {
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}
template <int MI, int NJ, int LD = TS>
__device__ __forceinline__ void acc_load(double (&acc)[MI][NJ][2], const double* C, WarpTile wt) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const double2 v = *reinterpret_cast<const double2*>(&C[tixw<LD>(wt.m0 + 8 * i + g, wt.n0 + 8 * j + 2 * t)]);
            acc[i][j][0] = v.x; acc[i][j][1] = v.y;
        }
}
template <int MI, int NJ, int LD = TS>
__device__ __forceinline__ void acc_store(const double (&acc)[MI][NJ][2], double* C, WarpTile wt) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            double2 v; v.x = acc[i][j][0]; v.y = acc[i][j][1];
            *reinterpret_cast<double2*>(&C[tixw<LD>(wt.m0 + 8 * i + g, wt.n0 + 8 * j + 2 * t)]) = v;
        }
}

// acc += sign * opA(A) * opB(B) restricted to k in [k_lo, k_hi) (multiples of 4).
//   opA(A)[m][k] = TA ? A[k][m] : A[m][k]        opB(B)[k][n] = TB ? B[n][k] : B[k][n]
// SKIP01: the block (i, j) = (0, 1) is left alone (a 16 x 16 unit on the diagonal of a lower-triangular result).
template <bool TA, bool TB, bool NEG, int MI = 4, int NJ = 2, int LDA = TS, int LDB = TS, bool SKIP01 = false>
__device__ __forceinline__ void tile_mma(double (&acc)[MI][NJ][2], const double* __restrict__ A,
                                         const double* __restrict__ B, WarpTile wt, int k_lo, int k_hi) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    // row-fragment addressing: (row, k0 + t) -> row*LD + ((k0 ^ sg) | t),  sg = (g & 3) << 2
    // col-fragment addressing: (k0 + t, col) -> (k0 + t)*LD + (col ^ (t << 2))
    const int sg = (g & 3) << 2, st = t << 2;
    int a_off[MI], b_off[NJ];
#pragma unroll
    for (int i = 0; i < MI; ++i) a_off[i] = TA ? (t * LDA + ((wt.m0 + 8 * i + g) ^ st)) : ((wt.m0 + 8 * i + g) * LDA + t);
#pragma unroll
    for (int j = 0; j < NJ; ++j) b_off[j] = TB ? ((wt.n0 + 8 * j + g) * LDB + t) : (t * LDB + ((wt.n0 + 8 * j + g) ^ st));
#pragma unroll 4
    for (int k0 = k_lo; k0 < k_hi; k0 += 4) {
        const int krow = k0 ^ sg;      // swizzled column base for row fragments
        double a[MI], b[NJ];
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const double v = A[a_off[i] + (TA ? k0 * LDA : krow)];
            a[i] = NEG ? -v : v;
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) b[j] = B[b_off[j] + (TB ? krow : k0 * LDB)];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j)
                if (!(SKIP01 && i == 0 && j == 1)) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

// Two 16 x 16 units of one warp (origins w0, w1) over the same k-range in ONE loop: eight independent DMMA chains in
// flight instead of four after four (gp_dag.cuh, accumulator layout 2: full products).
template <bool TA, bool TB, bool NEG>
__device__ __forceinline__ void tile_mma_pair(double (&acc0)[2][2][2], double (&acc1)[2][2][2], const double* __restrict__ A,
                                              const double* __restrict__ B, WarpTile w0, WarpTile w1, int k_lo, int k_hi) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int sg = (g & 3) << 2, st = t << 2;
    int a_off[2][2], b_off[2][2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const WarpTile wt = u ? w1 : w0;
#pragma unroll
        for (int i = 0; i < 2; ++i) a_off[u][i] = TA ? (t * TS + ((wt.m0 + 8 * i + g) ^ st)) : ((wt.m0 + 8 * i + g) * TS + t);
#pragma unroll
        for (int j = 0; j < 2; ++j) b_off[u][j] = TB ? ((wt.n0 + 8 * j + g) * TS + t) : (t * TS + ((wt.n0 + 8 * j + g) ^ st));
    }
#pragma unroll 4
    for (int k0 = k_lo; k0 < k_hi; k0 += 4) {
        const int krow = k0 ^ sg;
        double a[2][2], b[2][2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double v = A[a_off[u][i] + (TA ? k0 * TS : krow)];
                a[u][i] = NEG ? -v : v;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) b[u][j] = B[b_off[u][j] + (TB ? krow : k0 * TS)];
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                dmma(acc0[i][j][0], acc0[i][j][1], a[0][i], b[0][j]);
                dmma(acc1[i][j][0], acc1[i][j][1], a[1][i], b[1][j]);
            }
    }
}

// The same product when only the blocks on and below the block diagonal of the 64 x 64 result are needed (a
// diagonal tile of the covariance update or of K^-1): a warp whose 32 x 16 block lies above the diagonal does
// nothing, one whose upper 16 rows do computes the lower 16 only (8-row granularity of the decision: rows
// [r, r + 16) are needed iff their last row block reaches the warp's first column block).
template <bool TA, bool TB, bool NEG, int I0>
__device__ __forceinline__ void tile_mma_from_row(double (&acc)[4][2][2], const double* __restrict__ A,
                                                  const double* __restrict__ B, WarpTile wt, int k_lo, int k_hi) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int sg = (g & 3) << 2, st = t << 2;
    int a_off[4], b_off[2];
#pragma unroll
    for (int i = I0; i < 4; ++i) a_off[i] = TA ? (t * TS + ((wt.m0 + 8 * i + g) ^ st)) : ((wt.m0 + 8 * i + g) * TS + t);
#pragma unroll
    for (int j = 0; j < 2; ++j) b_off[j] = TB ? ((wt.n0 + 8 * j + g) * TS + t) : (t * TS + ((wt.n0 + 8 * j + g) ^ st));
#pragma unroll 4
    for (int k0 = k_lo; k0 < k_hi; k0 += 4) {
        const int krow = k0 ^ sg;
        double a[4], b[2];
#pragma unroll
        for (int i = I0; i < 4; ++i) {
            const double v = A[a_off[i] + (TA ? k0 * TS : krow)];
            a[i] = NEG ? -v : v;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) b[j] = B[b_off[j] + (TB ? krow : k0 * TS)];
#pragma unroll
        for (int i = I0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}
template <bool TA, bool TB, bool NEG>
__device__ __forceinline__ void tile_mma_lower(double (&acc)[4][2][2], const double* __restrict__ A,
                                               const double* __restrict__ B, WarpTile wt, int k_lo) {
    if (wt.n0 <= wt.m0 + 8) tile_mma_from_row<TA, TB, NEG, 0>(acc, A, B, wt, k_lo, TS);
    else if (wt.n0 <= wt.m0 + 24) tile_mma_from_row<TA, TB, NEG, 2>(acc, A, B, wt, k_lo, TS);
}

// ---------------------------------------------------------------------------------------------
// 8x8 diagonal block, the latency-critical serial chain (one rsqrt per column, no cross-thread
// traffic): Cholesky by ONE thread entirely in registers.  Reads the lower part of block (jb, jb) of
// tile A, writes L back (upper part of the block zeroed) and the 8 reciprocal diagonal entries to
// rinv[8].  Returns the 1-based failing column (tile-local) or 0.
// The triangular inverse of the block is NOT on this chain: it is computed by a worker warp of the
// chain pipeline (chain_ops.cuh) one step behind.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int potrf8(double* A, int jb, double* rinv) {
    double a[8][8];
    const int o = 8 * jb;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
        for (int c = 0; c < 8; c += 2) {              // 16-byte loads; (c, c+1) are adjacent under the swizzle
            if (c <= r) {
                const double2 v = *reinterpret_cast<const double2*>(&A[tix(o + r, o + c)]);
                a[r][c] = v.x; a[r][c + 1] = v.y;
            } else {
                a[r][c] = 0.0; a[r][c + 1] = 0.0;
            }
        }
    }
    int fail = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double p = a[j][j];
        if (!(p > 0.0) && fail == 0) fail = j + 1;     // catches NaN too
        const double ri = rsqrt_fast(p);
        rinv[j] = ri;
        a[j][j] = p * ri;                              // sqrt(p) without a second serial chain (<= 1.5 ulp)
#pragma unroll
        for (int i = j + 1; i < 8; ++i) a[i][j] *= ri;
#pragma unroll
        for (int k = j + 1; k < 8; ++k)
#pragma unroll
            for (int i = k; i < 8; ++i) a[i][k] = fma(-a[i][j], a[k][j], a[i][k]);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
            double2 v; v.x = (c <= r) ? a[r][c] : 0.0; v.y = (c + 1 <= r) ? a[r][c + 1] : 0.0;
            *reinterpret_cast<double2*>(&A[tix(o + r, o + c)]) = v;
        }
    return fail ? o + fail : 0;
}

// One DMMA-block helper: 8x8 accumulator (c0,c1) at block (bi, bj) of a tile.
__device__ __forceinline__ void blk_load(double& c0, double& c1, const double* C, int bi, int bj) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const double2 v = *reinterpret_cast<const double2*>(&C[tix(8 * bi + g, 8 * bj + 2 * t)]);
    c0 = v.x; c1 = v.y;
}
__device__ __forceinline__ void blk_store(double c0, double c1, double* C, int bi, int bj) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double2 v; v.x = c0; v.y = c1;
    *reinterpret_cast<double2*>(&C[tix(8 * bi + g, 8 * bj + 2 * t)]) = v;
}

}  // namespace fab
