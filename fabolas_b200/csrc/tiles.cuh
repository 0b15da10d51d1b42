// tiles.cuh -- 64x64 FP64 tiles in shared memory and the DMMA tile algebra the GP kernels are
// built from (tile GEMM in NT/NN/TN form, 8x8 Cholesky + inverse, in-tile factorisation, right/left
// triangular solves against a factored diagonal tile).
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
// 8 warps as 2 x 4 blocks of 32 x 16 over a 64 x 64 output
__device__ __forceinline__ WarpTile warp_tile() {
    const int w = threadIdx.x >> 5;
    WarpTile wt;
    wt.m0 = (w >> 2) * 32;
    wt.n0 = (w & 3) * 16;
    return wt;
}
// 8 warps as 4 x 2 blocks of 16 x 16 over a 64 x 32 output
__device__ __forceinline__ WarpTile warp_tile_32() {
    const int w = threadIdx.x >> 5;
    WarpTile wt;
    wt.m0 = (w >> 1) * 16;
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
template <bool TA, bool TB, bool NEG, int MI = 4, int NJ = 2, int LDA = TS, int LDB = TS>
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
            for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

// ---------------------------------------------------------------------------------------------
// 8x8 diagonal block, the latency-critical serial chain (one rsqrt per column, no cross-thread
// traffic): Cholesky by ONE thread entirely in registers.  Reads the lower part of block (jb, jb) of
// tile A, writes L back (upper part of the block zeroed) and the 8 reciprocal diagonal entries to
// rinv[8].  Returns the 1-based failing column (tile-local) or 0.
// The triangular inverse of the block is NOT on this chain: inv8_block() runs on another warp while
// the CTA performs the sub-panel and trailing updates.
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

// Column `c` (0..7, one thread each) of the row-major 8x8 inverse of the lower-triangular block
// (jb, jb) of A: x_c = 1/l_cc, x_r = -(sum_{k<r} l_rk x_k) / l_rr for r > c, zeros above.
__device__ __forceinline__ void inv8_column(const double* A, int jb, const double* rinv, double* dinv, int c) {
    double x[8];
    const int o = 8 * jb;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < r; ++k) s = fma(A[tix(o + r, o + k)], x[k], s);
        const double ri = rinv[r];
        x[r] = (r < c) ? 0.0 : ((r == c) ? ri : -s * ri);
        dinv[r * 8 + c] = x[r];
    }
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

// ---------------------------------------------------------------------------------------------
// In-tile Cholesky of a diagonal tile (lower part valid, strictly-upper 8x8 blocks must be zero).
// On exit A = L (lower), dinv[8][64] = inverses of the eight 8x8 diagonal blocks of L.
// Per 8-column step:
//   (a) thread 0: potrf8 of block (jb, jb) (serial chain)   ||  warp 1: hook(jb-1)
//   (b) one thread per row below the block: x = a L8^-T by an 8-step substitution with the
//       reciprocal pivots (right-looking, so only mul -> fma is on each row's dependent chain)
//       ||  thread 224: inv8_block(jb) for later consumers (panel solves, hook(jb))
//   (c) DMMA rank-8 update of the trailing blocks inside the tile
// `hook(jb)` (all threads of warp 1) overlaps the forward substitution of the right-hand side with
// the serial chain; hook(7) runs at the end.  `shadow(jb, v)` (all threads of warp v + 2, v = 0..5)
// runs independent work -- the covariance build of this column's panel tiles -- in the same window.  *fail (shared int) receives the first failing
// 1-based column within the tile.  `rinv` is 8 doubles of shared scratch.
// All 256 threads must call.  Ends with a __syncthreads().
// ---------------------------------------------------------------------------------------------
template <typename Hook, typename Shadow>
__device__ __forceinline__ void tile_potrf(double* A, double* dinv, double* rinv, int* fail, Hook hook, Shadow shadow) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int jb = 0; jb < 8; ++jb) {
        if (threadIdx.x == 0) {
            PHB(5);
            const int f = potrf8(A, jb, rinv);
            if (f && *fail == 0) *fail = f;
            PHE(5);
        } else if (w == 1) {
            if (jb > 0) hook(jb - 1);
        } else if (w >= 2) {
            shadow(jb, w - 2);               // independent work of warps 2..7 in the shadow of the serial chain
        }
        __syncthreads();
        PHB(6);
        if (w == 7) {
            // 8x8 inverse for later consumers (panel solves, hook(jb)): off the critical path, it
            // spans phases (b) and (c) of the other seven warps
            if (lane < 8) inv8_column(A, jb, rinv, dinv + 64 * jb, lane);
        } else {
            // (b) rows below the diagonal block, one thread each
            const int row = 8 * (jb + 1) + threadIdx.x;
            if (row < TS) {
                double x[8], l[8][8], ri[8];
#pragma unroll
                for (int c = 0; c < 8; c += 2) {                  // 16-byte accesses: (c, c+1) stay adjacent under the swizzle
                    const double2 v = *reinterpret_cast<const double2*>(&A[tix(row, 8 * jb + c)]);
                    x[c] = v.x; x[c + 1] = v.y;
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    ri[c] = rinv[c];
#pragma unroll
                    for (int k = 0; k < 8; ++k) l[c][k] = (k < c) ? A[tix(8 * jb + c, 8 * jb + k)] : 0.0;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    x[k] *= ri[k];
#pragma unroll
                    for (int c = k + 1; c < 8; ++c) x[c] = fma(-x[k], l[c][k], x[c]);
                }
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    double2 v; v.x = x[c]; v.y = x[c + 1];
                    *reinterpret_cast<double2*>(&A[tix(row, 8 * jb + c)]) = v;
                }
            }
            if (jb < 7) {
                named_bar_sync(1, NTHREADS - 32);
                // (c) trailing update inside the tile: blocks (ib, kb), jb < kb <= ib, over 7 warps
                const int m = 7 - jb;                 // trailing block rows
                const int nblk = m * (m + 1) / 2;
                for (int q = w; q < nblk; q += NWARPS - 1) {
                    int ri_ = 0, acc_ = 0;            // unrank q -> (ri_, rk), rk <= ri_, row-major lower triangle
                    while (acc_ + ri_ + 1 <= q) { acc_ += ri_ + 1; ++ri_; }
                    const int rk = q - acc_;
                    const int ib = jb + 1 + ri_, kb = jb + 1 + rk;
                    double c0, c1;
                    blk_load(c0, c1, A, ib, kb);
#pragma unroll
                    for (int k0 = 0; k0 < 8; k0 += 4) {
                        const double a = -A[tix(8 * ib + g, 8 * jb + k0 + t)];
                        const double b = A[tix(8 * kb + g, 8 * jb + k0 + t)];
                        dmma(c0, c1, a, b);
                    }
                    blk_store(c0, c1, A, ib, kb);
                }
            }
        }
        __syncthreads();
        PHE(6);
    }
    if (w == 1) hook(7);
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Right solve  X * L^T = A  (L = factored diagonal tile with its 8x8 inverses) for up to NB panel
// tiles at once; X overwrites A.  Each warp owns the same 8-row strip of every tile and interleaves
// them, so the serial 8-step chain of one strip hides behind the others (no block barriers inside).
// Ends with __syncthreads().
// ---------------------------------------------------------------------------------------------
template <int NB>
__device__ __forceinline__ void tile_trsm_right_multi(double* const (&T)[NB], int n, const double* __restrict__ L,
                                                      const double* __restrict__ dinv) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int s = w;                                  // strip = rows 8s .. 8s+7
    for (int jb = 0; jb < 8; ++jb) {
        double c0[NB], c1[NB];
#pragma unroll
        for (int q = 0; q < NB; ++q) if (q < n) blk_load(c0[q], c1[q], T[q], s, jb);
        for (int kb = 0; kb < jb; ++kb) {
#pragma unroll
            for (int k0 = 0; k0 < 8; k0 += 4) {
                const double b = L[tix(8 * jb + g, 8 * kb + k0 + t)];        // (L_(jb,kb))^T
#pragma unroll
                for (int q = 0; q < NB; ++q)
                    if (q < n) dmma(c0[q], c1[q], -T[q][tix(8 * s + g, 8 * kb + k0 + t)], b);   // X_(s,kb)
            }
        }
        // Y = acc * inv8_jb^T : round-trip acc through the tile to turn it into an A operand
#pragma unroll
        for (int q = 0; q < NB; ++q) if (q < n) blk_store(c0[q], c1[q], T[q], s, jb);
        __syncwarp();
        double y0[NB], y1[NB];
#pragma unroll
        for (int q = 0; q < NB; ++q) y0[q] = y1[q] = 0.0;
#pragma unroll
        for (int k0 = 0; k0 < 8; k0 += 4) {
            const double b = dinv[64 * jb + g * 8 + k0 + t];
#pragma unroll
            for (int q = 0; q < NB; ++q)
                if (q < n) dmma(y0[q], y1[q], T[q][tix(8 * s + g, 8 * jb + k0 + t)], b);
        }
#pragma unroll
        for (int q = 0; q < NB; ++q) if (q < n) blk_store(y0[q], y1[q], T[q], s, jb);
        __syncwarp();
    }
    __syncthreads();
}

__device__ __forceinline__ void tile_trsm_right(double* A, const double* __restrict__ L,
                                                const double* __restrict__ dinv) {
    double* const T[1] = {A};
    tile_trsm_right_multi<1>(T, 1, L, dinv);
}

// ---------------------------------------------------------------------------------------------
// Left solve  L * X = R  (X overwrites R).  Each warp owns an 8-column strip.
// If UNIT_RHS, R is taken to be the identity (R need not be initialised) -> X = L^-1.
// If NEG_RHS, solves L * X = -R.   Ends with __syncthreads().
// ---------------------------------------------------------------------------------------------
template <bool UNIT_RHS, bool NEG_RHS>
__device__ __forceinline__ void tile_trsm_left(double* R, const double* __restrict__ L,
                                               const double* __restrict__ dinv) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int s = w;                                  // strip = cols 8s .. 8s+7
    for (int ib = 0; ib < 8; ++ib) {
        double c0, c1;
        if (UNIT_RHS) {
            if (ib < s) {                             // X is lower triangular: zero block above the diagonal
                blk_store(0.0, 0.0, R, ib, s);
                continue;                             // warp-uniform
            }
            c0 = (ib == s && g == 2 * t) ? 1.0 : 0.0;
            c1 = (ib == s && g == 2 * t + 1) ? 1.0 : 0.0;
        } else {
            blk_load(c0, c1, R, ib, s);
            if (NEG_RHS) { c0 = -c0; c1 = -c1; }
        }
        for (int kb = (UNIT_RHS ? s : 0); kb < ib; ++kb) {
#pragma unroll
            for (int k0 = 0; k0 < 8; k0 += 4) {
                const double a = -L[tix(8 * ib + g, 8 * kb + k0 + t)];       // L_(ib,kb)
                const double b = R[tix(8 * kb + k0 + t, 8 * s + g)];         // X_(kb,s)
                dmma(c0, c1, a, b);
            }
        }
        blk_store(c0, c1, R, ib, s);
        __syncwarp();
        double y0 = 0.0, y1 = 0.0;
#pragma unroll
        for (int k0 = 0; k0 < 8; k0 += 4) {
            const double a = dinv[64 * ib + g * 8 + k0 + t];                 // inv8_ib
            const double b = R[tix(8 * ib + k0 + t, 8 * s + g)];
            dmma(y0, y1, a, b);
        }
        blk_store(y0, y1, R, ib, s);
        __syncwarp();
    }
    __syncthreads();
}

}  // namespace fab
