// gp_factor.cuh -- kernel A: fused covariance build + tiled Cholesky + forward solve + log-likelihood.
//
// One CTA (256 threads, one SM) per walker / hyper-parameter sample.  Replaces, for a whole batch of
// walkers at once, the george calls  GP.compute (fabolas.py:62, entropy_search.py:73)  and
// GP.log_likelihood (fabolas.py:66, entropy_search.py:77):
//     K = k(X, X) + (yerr^2 + 1.25e-12) I ;  K = L L^T ;  z = L^-1 (y - mean)
//     ll = -0.5 (N log 2pi + 2 sum log L_ii) - 0.5 z.z          (-inf when K is not positive definite)
// The covariance matrix is never materialised in HBM: tiles are generated from 64-row slices of the
// N x D observation matrix (staged in shared memory, pre-scaled by 1/sqrt(M_k)) straight into the
// shared-memory tile cache and consumed by the left-looking block-column factorisation.  Factor
// tiles are written through to the HBM workspace only when a later stage (gradient, prediction)
// needs them or when the factor does not fit the cache (N > 256) and must be streamed back by TMA.
// Shared memory use is independent of N: tile slots + a few coordinate slices + 64-entry vectors.
#pragma once
#include "gp_common.cuh"

namespace fab {

constexpr int FACTOR_GB = 3;       // panel tiles solved together (ILP across their strips)

struct FactorArgs {
    GpData gd;
    GpWork wk;
    const double* theta;   // B x Pk
    const double* yerr2;   // B
    double* ll;            // B
    int* info;             // B   0 ok, k>0: leading minor k not positive definite
    int nslot;
    int store;             // write L tiles / Dinv / z to the workspace
};

__host__ __device__ inline size_t factor_smem_bytes(int nslot, int d) {
    return (size_t)nslot * TILE_BYTES + 8 * (size_t)(512 + (1 + FACTOR_GB) * (d + 1) * TS + 2 * TS + 64) +
           8 * MAX_SLOTS + 64 + 8 * EXP_TAB_N;
}

__global__ void __launch_bounds__(NTHREADS, 1) gp_factor_kernel(FactorArgs a) {
    FAB_DYN_SMEM(smem_raw);
    const GpData& gd = a.gd;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int Np = gd.Np, NT = gd.NT, N = gd.N, d = gd.d;
    const int SL = slice_doubles(d);

    double* slots = reinterpret_cast<double*>(smem_raw);
    double* dinv = slots + (size_t)a.nslot * TILE_ELEMS;
    double* sl_col = dinv + 512;                 // slice of block column K (also the rows of the diagonal tile)
    double* sl_row = sl_col + SL;                // FACTOR_GB row slices of the panel group
    double* rhs = sl_row + FACTOR_GB * SL;       // 64: r_K - sum_{J<K} L_KJ z_J
    double* zl = rhs + TS;                       // 64: z_K
    double* red = zl + TS;                       // 64 scratch
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(red + 64);
    int* fail = reinterpret_cast<int*>(bar + MAX_SLOTS);
    double* etab = reinterpret_cast<double*>(bar + MAX_SLOTS + 8);   // exp table of the fast Matern evaluation
    double* rinv = red + 32;                     // 8 reciprocal pivots of the block being factored

    PHB(0);
    const Hyper h = decode_hyper(gd, a.theta + (size_t)b * gd.Pk, a.yerr2[b]);
    double* accw = a.wk.accw + (size_t)b * Np;
    for (int i = tid; i < Np; i += NTHREADS) accw[i] = 0.0;
    if (tid == 0) *fail = 0;
    exp_table_fill(etab);

    TileCache tc; tc.init(slots, a.nslot, bar);
    __syncthreads();
    double* Lw = a.wk.Lw ? a.wk.Lw + (size_t)b * (NT * (NT + 1) / 2) * TILE_ELEMS : nullptr;
    double* Dw = a.wk.Dinv ? a.wk.Dinv + (size_t)b * NT * 512 : nullptr;
    const WarpTile wt = warp_tile();
    int first_fail = 0;
    double part[2] = {0.0, 0.0};                 // per-thread partial sums: log diag(L), z^2

    for (int K = 0; K < NT; ++K) {
        auto dead = [K](int id) { return id_row(id) < K; };   // rows above the current block column are done
        // ================= diagonal tile (K, K) =================
        PHB(1);
        const int sd = tc.alloc(tile_id(0, K, K), dead, 0u);  // (block barrier inside: slices are free to overwrite)
        double* Td = tc.ptr(sd);
        stage_slice(sl_col, gd, h, K);
        if (tid < TS) {
            const int gi = TS * K + tid;
            rhs[tid] = (gi < N) ? (__ldg(&gd.y[gi]) - gd.mean) - accw[gi] : 0.0;
        }
        __syncthreads();
        Coords cd; cd.tab = etab; cd.zr = sl_col; cd.zc = sl_col; cd.N = N; cd.d = d; cd.family = gd.family;
        build_tile(Td, cd, h, K, K);
        __syncthreads();
        PHE(1);
        PHB(2);
        if (K > 0) {
            double acc[4][2][2];
            acc_load(acc, Td, wt);
            for (int J = 0; J < K; ++J) {
                const int sj = tc.acquire(Lw, tile_id(0, K, J), dead, pin(sd));
                const double* Lj = tc.ptr(sj);
                tile_mma<false, true, true>(acc, Lj, Lj, wt, 0, TS);
            }
            acc_store(acc, Td, wt);
            __syncthreads();
        }
        PHE(2);
        PHB(3);
        // First group of panel tiles of this column: slots reserved now, covariance entries generated
        // by warps 2..7 in the shadow of the serial 8x8 factorisations of the diagonal tile
        // (6 half-strips of 4 rows per 8-column step => 3 tiles per diagonal tile).
        const int gb = max(1, min(FACTOR_GB, a.nslot - 3));   // diag + group + two streamed operands must fit
        const int ng0 = min(gb, NT - K - 1);
        int st0[FACTOR_GB];
        double* Tt0[FACTOR_GB];
        unsigned pins0 = pin(sd);
#pragma unroll
        for (int q = 0; q < FACTOR_GB; ++q) {
            st0[q] = -1; Tt0[q] = nullptr;
            if (q < ng0) {
                st0[q] = tc.alloc(tile_id(0, K + 1 + q, K), dead, pins0);
                pins0 |= pin(st0[q]);
                Tt0[q] = tc.ptr(st0[q]);
                stage_slice(sl_row + q * SL, gd, h, K + 1 + q);
            }
        }
        __syncthreads();
        PHE(3);
        auto shadow = [&](int jb, int v) {
            const int hs = 6 * jb + v;                 // half-strip index: tile hs / 16, rows 4 * (hs % 16) ..
            const int q = hs >> 4;
            if (q < ng0) {
                double* T = (q == 0) ? Tt0[0] : ((q == 1) ? Tt0[1] : Tt0[2]);
                Coords cp; cp.tab = etab; cp.zr = sl_row + q * SL; cp.zc = sl_col; cp.N = N; cp.d = d; cp.family = gd.family;
                build_rows(T, cp, h, K + 1 + q, K, 4 * (hs & 15), 4);
            }
        };
        {
            // forward substitution of the right-hand side, 8 rows at a time, overlapped with the
            // serial 8x8 factorisations (runs on warp 1 while thread 0 factors the next block)
            auto hook = [&](int jb) {
                const int q = lane >> 2, p = lane & 3;
                const int row = 8 * jb + q;
                double s = 0.0;
                for (int c = p; c < 8 * jb; c += 4) s = fma(Td[tix(row, c)], zl[c], s);
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                const double v = rhs[row] - s;
                double zz = 0.0;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const double vc = __shfl_sync(0xffffffffu, v, 4 * c);
                    zz = fma(dinv[64 * jb + q * 8 + c], vc, zz);     // inv8 is lower triangular (zeros above)
                }
                if (p == 0) zl[row] = zz;
                __syncwarp();
            };
            PHB(4);
            tile_potrf(Td, dinv, rinv, fail, hook, shadow);
            PHE(4);
        }
        PHB(7);
        if (*fail && first_fail == 0) first_fail = TS * K + *fail;
        if (tid < TS && TS * K + tid < N) {
            part[0] += log(Td[tix(tid, tid)]);
            part[1] += zl[tid] * zl[tid];
        }
        if (a.store) {
            if (tid < TS) a.wk.zvec[(size_t)b * Np + TS * K + tid] = zl[tid];
            for (int e = tid; e < 512; e += NTHREADS) Dw[(size_t)K * 512 + e] = dinv[e];
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) bulk_s2g(Lw + (size_t)tri_index(K, K) * TILE_ELEMS, Td, TILE_BYTES);
        }
        PHE(7);
        // ================= panel tiles (I, K), I > K, in groups of up to gb =================
        for (int I0 = K + 1; I0 < NT; I0 += gb) {
            const int ng = min(gb, NT - I0);
            const bool prebuilt = I0 == K + 1;             // the first group was generated in the shadow
            int st[FACTOR_GB];
            double* Tt[FACTOR_GB];
            unsigned pins = prebuilt ? pins0 : pin(sd);
#pragma unroll
            for (int q = 0; q < FACTOR_GB; ++q) {
                st[q] = -1; Tt[q] = nullptr;
                if (q < ng) {
                    const int I = I0 + q;
                    if (prebuilt) {
                        st[q] = st0[q]; Tt[q] = Tt0[q];
                    } else {
                        PHB(8);
                        st[q] = tc.alloc(tile_id(0, I, K), dead, pins);     // (block barrier inside)
                        pins |= pin(st[q]);
                        Tt[q] = tc.ptr(st[q]);
                        stage_slice(sl_row, gd, h, I);
                        __syncthreads();
                        Coords cp; cp.tab = etab; cp.zr = sl_row; cp.zc = sl_col; cp.N = N; cp.d = d; cp.family = gd.family;
                        build_tile(Tt[q], cp, h, I, K);
                        __syncthreads();
                        PHE(8);
                    }
                    PHB(9);
                    if (K > 0) {
                        double acc[4][2][2];
                        acc_load(acc, Tt[q], wt);
                        for (int J = 0; J < K; ++J) {
                            const int sa = tc.acquire(Lw, tile_id(0, I, J), dead, pins);
                            const int sb = tc.acquire(Lw, tile_id(0, K, J), dead, pins | pin(sa));
                            tile_mma<false, true, true>(acc, tc.ptr(sa), tc.ptr(sb), wt, 0, TS);
                        }
                        acc_store(acc, Tt[q], wt);
                    }
                    PHE(9);
                }
            }
            PHB(10);
            __syncthreads();
            tile_trsm_right_multi<FACTOR_GB>(Tt, ng, Td, dinv);
            PHE(10);
            PHB(11);
            // accw_I += L_IK z_K   (warp w: rows 8w..8w+7, lanes across columns; lane 0 owns the row's entry)
            const double z0 = zl[lane], z1 = zl[lane + 32];
#pragma unroll
            for (int q = 0; q < FACTOR_GB; ++q) {
                if (q < ng) {
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const int r = 8 * w + rr;
                        double s = Tt[q][tix(r, lane)] * z0 + Tt[q][tix(r, lane + 32)] * z1;
                        s = warp_sum(s);
                        if (lane == 0) accw[TS * (I0 + q) + r] += s;
                    }
                }
            }
            PHE(11);
            PHB(12);
            if (a.store) {
                fence_proxy_async();
                __syncthreads();
                if (tid == 0)
                    for (int q = 0; q < ng; ++q)
                        bulk_s2g(Lw + (size_t)tri_index(I0 + q, K) * TILE_ELEMS, Tt[q], TILE_BYTES);
            }
            PHE(12);
        }
        __syncthreads();
    }
    PHB(13);

    // ---- log-likelihood ----
    block_sum<2>(part, red);
    if (tid == 0) {
        double ll = -0.5 * (N * kLog2Pi + 2.0 * part[0]) - 0.5 * part[1];
        int inf = first_fail;
        if (inf != 0 || !isfinite(ll)) ll = -INFINITY;      // george: non-finite -> -inf (fabolas.py:58-66)
        a.ll[b] = ll;
        a.info[b] = inf;
        bulk_s2g_wait_all();
    }
    PHE(13);
    PHE(0);
}

}  // namespace fab
