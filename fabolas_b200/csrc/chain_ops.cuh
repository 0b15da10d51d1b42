// chain_ops.cuh -- the latency-bound serial work of one diagonal tile, written for the four "chain" warps
// of gp_dag_kernel (one per SM sub-partition) so that it runs concurrently with the eight throughput warps:
//   potrf64_chain : in-place Cholesky of a 64x64 tile (8x8 register Cholesky on one lane, row solves one
//                   lane per row, DMMA rank-8 trailing updates dealt to the warps), producing the 8x8 inverses
//   inv64_chain   : W = L^-1 of the factored tile into another tile, row block by row block; the column
//                   blocks of a row are independent DMMA chains dealt to the warps
//   matvec helpers for z = W rhs and alpha = W^T z.
// Synchronisation among the chain warps only: named barrier 2 (chain_bar) and __syncwarp.
#pragma once
#include "tiles.cuh"

namespace fab {

#ifndef CHT
#define CHT(jb, k) do {} while (0)        // (tools/microbench/potrf64_bench.cu: clock stamps of warp 0)
#endif
constexpr int NCW = 4;                 // chain warps (one per SM sub-partition)
__device__ __forceinline__ void chain_bar() { named_bar_sync(2, 32 * NCW); }

// (row, column) of the q-th block of a lower-triangular block matrix in row-major order, q < 28
__constant__ unsigned char kTriRow[28] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 6};
__constant__ unsigned char kTriCol[28] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5, 6};

// One row of the panel below an 8x8 diagonal block: x = a L8^-T by substitution (in place in tile A).
__device__ __forceinline__ void solve_row8(double* A, int row, int o, const double* __restrict__ rinv) {
    double x[8];
#pragma unroll
    for (int cc = 0; cc < 8; cc += 2) {
        const double2 v = *reinterpret_cast<const double2*>(&A[tix(row, o + cc)]);
        x[cc] = v.x; x[cc + 1] = v.y;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        x[k] *= rinv[k];
#pragma unroll
        for (int cc = k + 1; cc < 8; ++cc) x[cc] = fma(-x[k], A[tix(o + cc, o + k)], x[cc]);
    }
#pragma unroll
    for (int cc = 0; cc < 8; cc += 2) {
        double2 v; v.x = x[cc]; v.y = x[cc + 1];
        *reinterpret_cast<double2*>(&A[tix(row, o + cc)]) = v;
    }
}

// Row block IB of W = L^-1 (64x64, lower triangular) from the factored tile L and the 8x8 diagonal-block
// inverses, by the three worker warps (wi = 0..2) of the chain:
//   W(IB, IB) = inv8(IB);   W(IB, s) = -inv8(IB) * sum_{kb = s}^{IB - 1} L(IB, kb) W(kb, s),  s < IB
// The column blocks s of a row are independent; worker wi owns s = wi, wi + 3, wi + 6 in every row, so a
// worker only ever reads its own earlier results.  The blocks right of the diagonal are zero-filled (the tile
// products read whole tiles).  Needs rows < IB of W, row block IB of L and dinv[IB] final.
template <int IB>
__device__ __forceinline__ void inv64_row(double* W, const double* __restrict__ L, const double* __restrict__ dinv, int wi) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    if (wi == (IB % 3)) {          // zero blocks right of the diagonal, inverse block on it
#pragma unroll
        for (int s = IB + 1; s < 8; ++s) blk_store(0.0, 0.0, W, IB, s);
        const double2 v = *reinterpret_cast<const double2*>(&dinv[64 * IB + g * 8 + 2 * t]);
        blk_store(v.x, v.y, W, IB, IB);
    }
    if (IB == 0) return;
    constexpr int NU = 3;
    double c0[NU], c1[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) c0[u] = c1[u] = 0.0;
    for (int kb = 0; kb < IB; ++kb) {
#pragma unroll
        for (int k0 = 0; k0 < 8; k0 += 4) {
            const double a = L[tix(8 * IB + g, 8 * kb + k0 + t)];
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const int s = wi + 3 * u;
                if (s <= kb) dmma(c0[u], c1[u], a, W[tix(8 * kb + k0 + t, 8 * s + g)]);
            }
        }
    }
    // T_s through the tile (accumulator -> B operand), then Y_s = -inv8(IB) T_s; a warp only touches its own blocks
#pragma unroll
    for (int u = 0; u < NU; ++u) if (wi + 3 * u < IB) blk_store(c0[u], c1[u], W, IB, wi + 3 * u);
    __syncwarp();
    double y0[NU], y1[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) y0[u] = y1[u] = 0.0;
#pragma unroll
    for (int k0 = 0; k0 < 8; k0 += 4) {
        const double a = -dinv[64 * IB + g * 8 + k0 + t];
#pragma unroll
        for (int u = 0; u < NU; ++u) {
            const int s = wi + 3 * u;
            if (s < IB) dmma(y0[u], y1[u], a, W[tix(8 * IB + k0 + t, 8 * s + g)]);
        }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < NU; ++u) if (wi + 3 * u < IB) blk_store(y0[u], y1[u], W, IB, wi + 3 * u);
}

// Cholesky of tile A (lower part valid; strictly-upper 8x8 blocks are ignored and left untouched) by the
// NCW chain warps, as a two-stage pipeline with one step of look-ahead:
//   warp 0 (the critical path), step jb:  8x8 Cholesky of block (jb, jb) on one lane (the serial rsqrt chain)
//       -> the 8 rows of row block jb + 1 solved against it -> block (jb + 1, jb + 1) updated,
//       i.e. exactly what the next 8x8 Cholesky needs;
//   warps 1..3 (one step behind):  8x8 inverse of block (jb, jb), the remaining row solves (one lane per row),
//       the DMMA rank-8 update of every other trailing block, and row block jb of W = L^-1 (inv64_row).
// Named barriers 3 (warp 0 -> workers: "step jb factored") and 4 (workers -> warp 0: "step jb applied").
// dinv[8][64]: inverses of the eight 8x8 diagonal blocks of L (row-major).  rinv: 8 x 8 doubles scratch (one
// set per step).  *failp (shared, zero on entry) receives the 1-based tile-local index of the first
// non-positive pivot.  W receives L^-1.  Ends with chain_bar().
// nb (1..8): number of 8-row blocks of the tile that hold observations; the blocks beyond (padding of the last tile
// of a ragged problem: identity on the diagonal, zero elsewhere) are their own factor and inverse and are not worked on.
__device__ __forceinline__ void potrf64_chain(double* A, double* W, double* dinv, double* rinv, int* failp, int cw, int nb) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    constexpr int NT128 = 32 * NCW;
    if (cw == 0) {
        for (int jb = 0; jb < nb; ++jb) {
            const int o = 8 * jb;
            double* rv = rinv + 8 * jb;
            PHBC(60);
            CHT(jb, 0);
            if (lane == 0) {
                const int f = potrf8(A, jb, rv);
                if (f && *failp == 0) *failp = f;
            }
            __syncwarp();
            PHEC(60);
            PHBC(61);
            CHT(jb, 1);
            if (jb >= 1) named_bar_sync(4, NT128);          // the workers have applied step jb - 1 to the rest of the tile
            CHT(jb, 2);
            if (jb + 1 < nb) {
                if (lane < 8) solve_row8(A, o + 8 + lane, o, rv);
                __syncwarp();
                CHT(jb, 4);
                double c0, c1;
                blk_load(c0, c1, A, jb + 1, jb + 1);
#pragma unroll
                for (int k0 = 0; k0 < 8; k0 += 4) {
                    const double x = A[tix(o + 8 + g, o + k0 + t)];
                    dmma(c0, c1, -x, x);
                }
                blk_store(c0, c1, A, jb + 1, jb + 1);
                __syncwarp();
            }
            CHT(jb, 3);
            named_bar_arrive(3, NT128);
            PHEC(61);
        }
    } else {
        for (int jb = 0; jb < nb; ++jb) {
            const int o = 8 * jb;
            const double* rv = rinv + 8 * jb;
            named_bar_sync(3, NT128);                       // block (jb, jb) factored, row block jb + 1 solved
            if (cw == NCW - 1) {
                // 8x8 inverse: every lane computes column lane & 7 (identical values), lanes 0..7 store
                const int c = lane & 7;
                double xi[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < r; ++k) s = fma(A[tix(o + r, o + k)], xi[k], s);
                    const double ri = rv[r];
                    xi[r] = (r < c) ? 0.0 : ((r == c) ? ri : -s * ri);
                }
                if (lane < 8) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) dinv[64 * jb + r * 8 + c] = xi[r];
                }
            } else {
                const int row = o + 16 + lane + 32 * (cw - 1);       // rows of the row blocks >= jb + 2
                if (row < 8 * nb) solve_row8(A, row, o, rv);
            }
            named_bar_sync(5, NT128 - 32);                  // (workers only) column block jb solved, dinv[jb] written
            if (jb + 2 < nb) {
                // trailing update: block (ib, kb) -= L(ib, jb) L(kb, jb)^T, jb < kb <= ib < nb, except (jb+1, jb+1)
                const int m = nb - 1 - jb;
                const int nblk = m * (m + 1) / 2;
                for (int q0 = cw; q0 < nblk; q0 += 2 * (NCW - 1)) {      // q0 = 1, 2, 3 first: q = 0 is warp 0's block
                    int ib[2], kb[2];
                    bool on[2];
                    double c0[2], c1[2], a0[2], a1[2], b0[2], b1[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int q = q0 + (NCW - 1) * u;
                        on[u] = q < nblk;
                        ib[u] = jb + 1 + kTriRow[on[u] ? q : 0]; kb[u] = jb + 1 + kTriCol[on[u] ? q : 0];
                        if (on[u]) {
                            blk_load(c0[u], c1[u], A, ib[u], kb[u]);
                            a0[u] = -A[tix(8 * ib[u] + g, o + t)]; a1[u] = -A[tix(8 * ib[u] + g, o + 4 + t)];
                            b0[u] = A[tix(8 * kb[u] + g, o + t)]; b1[u] = A[tix(8 * kb[u] + g, o + 4 + t)];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) if (on[u]) dmma(c0[u], c1[u], a0[u], b0[u]);
#pragma unroll
                    for (int u = 0; u < 2; ++u) if (on[u]) dmma(c0[u], c1[u], a1[u], b1[u]);
#pragma unroll
                    for (int u = 0; u < 2; ++u) if (on[u]) blk_store(c0[u], c1[u], A, ib[u], kb[u]);
                }
            }
            if (jb + 1 < nb) named_bar_arrive(4, NT128);
            switch (jb) {                                   // row block jb of the inverse, off the critical path
            case 0: inv64_row<0>(W, A, dinv, cw - 1); break;
            case 1: inv64_row<1>(W, A, dinv, cw - 1); break;
            case 2: inv64_row<2>(W, A, dinv, cw - 1); break;
            case 3: inv64_row<3>(W, A, dinv, cw - 1); break;
            case 4: inv64_row<4>(W, A, dinv, cw - 1); break;
            case 5: inv64_row<5>(W, A, dinv, cw - 1); break;
            case 6: inv64_row<6>(W, A, dinv, cw - 1); break;
            default: inv64_row<7>(W, A, dinv, cw - 1); break;
            }
        }
        // padding rows of the inverse: identity blocks on the diagonal, zero elsewhere (the tile products read whole tiles)
        for (int q = (cw - 1); q < (8 - nb) * 8; q += NCW - 1) {
            const int ib = nb + (q >> 3), s = q & 7;
            const double one = (ib == s && g == 2 * t) ? 1.0 : 0.0, one1 = (ib == s && g == 2 * t + 1) ? 1.0 : 0.0;
            blk_store(one, one1, W, ib, s);
        }
    }
    chain_bar();
}

// Rows 16 cw .. 16 cw + 15 of z = W v (W lower triangular, v a 64-vector in shared memory): written to
// z_out[r] (two destinations) by lane r & 15 of the calling warp; returns this lane's share of sum z^2.
__device__ __forceinline__ double tile_matvec_rows(const double* __restrict__ W, const double* __restrict__ v, double* z_a,
                                                   double* z_b, int cw) {
    const int lane = threadIdx.x & 31;
    const double v0 = v[lane], v1 = v[lane + 32];
    double zz = 0.0;
#pragma unroll 4
    for (int rr = 0; rr < 16; ++rr) {
        const int r = 16 * cw + rr;
        double s = W[tix(r, lane)] * v0;
        if (r >= 32) s = fma(W[tix(r, lane + 32)], v1, s);       // columns 32.. of rows < 32 are zero
        s = warp_sum(s);
        if (lane == rr) { z_a[r] = s; z_b[r] = s; zz = s * s; }
    }
    return zz;
}
// Share of rows 16 cw .. 16 cw + 15 in a = W^T z: lane l returns the partial sums of columns l and l + 32.
__device__ __forceinline__ void tile_matvec_t_rows(const double* __restrict__ W, const double* __restrict__ z, double& a0,
                                                   double& a1, int cw) {
    const int lane = threadIdx.x & 31;
    a0 = 0.0; a1 = 0.0;
#pragma unroll 8
    for (int rr = 0; rr < 16; ++rr) {
        const int r = 16 * cw + rr;
        const double zr = z[r];
        a0 = fma(W[tix(r, lane)], zr, a0);
        a1 = fma(W[tix(r, lane + 32)], zr, a1);
    }
}

}  // namespace fab
