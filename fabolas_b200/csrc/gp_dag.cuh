// gp_dag.cuh -- the fused GP kernel: covariance build + tiled Cholesky + log-likelihood and, by mode,
// W = L^-1 + alpha (fit) and the K^-1 tile products + gradient contraction (grad), one CTA per walker.
//
// Replaces, for a whole batch of walkers at once, george's GP.compute / GP.log_likelihood
// (fabolas.py:62-66, entropy_search.py:73-77), the solver state GP.predict needs (fabolas.py:187) and
// grad_log_likelihood.  The CTA is warp-specialised: eight BULK warps execute the
// throughput stream (Matern tile generation straight into DMMA accumulator registers, 64^3 DMMA tile
// products, panel solves in GEMM form against the inverted diagonal tile, the dK contraction) and one
// CHAIN warp executes the latency-bound stream (in-tile Cholesky and triangular inverse of each diagonal
// tile, forward-substitution vectors, log-determinant).  Both interpret the static program built on the
// host (dag_program.h): slots, TMA prefetches, write-through stores and cross-stream waits are all
// precomputed; the streams meet only through two shared-memory completion counters (release/acquire).
// The covariance matrix and its inverse are never materialised.
#pragma once
#include "gp_common.cuh"
#include "gp_grad.cuh"
#include "chain_ops.cuh"
#include "dag_program.h"

namespace fab {

constexpr int DAG_BULK_THREADS = NTHREADS;         // warps 0..7
constexpr int DAG_THREADS = NTHREADS + 32;         // + the chain warp

struct DagArgs {
    GpData gd;
    GpWork wk;
    const double* theta;   // B x Pk
    const double* yerr2;   // B
    double* ll;            // B
    int* info;             // B   0 ok, k>0: leading minor k not positive definite
    double* grad;          // B x (Pk + 1), DAG_GRAD only
    const DagOp* bulk;     // device copies of the two instruction streams
    const DagOp* chain;
    int nbulk, nchain;
    int nslot, mode;
    int resident;          // 1: scaled coordinates, z, accw, alpha of all Np rows live in shared memory
    int ntiles;            // NT (NT + 1) / 2: workspace tile indices >= ntiles address the spill region wk_spill
    double* wk_spill;      // B x NT tiles (FACTOR mode only: inverted diagonal tiles that had to leave shared memory)
};

__host__ __device__ inline size_t dag_smem_bytes(int nslot, int d, int resident_np) {
    const size_t coords = resident_np ? (size_t)(d + 4) * resident_np : (size_t)(2 * (d + 1) * TS + 2 * TS);
    return (size_t)nslot * TILE_BYTES + 8 * (size_t)(512 + EXP_TAB_N + 512 + 128 + 64 + 64 + 16) + 8 * coords +
           8 * MAX_SLOTS + 64;
}

// ---- stream synchronisation: completion counters in shared memory ----
__device__ __forceinline__ void sync_arrive(int* p) {
#ifdef FAB_CUSIM
    *p += 1;
#else
    asm volatile("red.release.cta.shared.add.s32 [%0], 1;" ::"r"(smem_u32(p)) : "memory");
#endif
}
__device__ __forceinline__ void spin_until(int* p, int target) {
#ifdef FAB_CUSIM
    while (*reinterpret_cast<volatile int*>(p) < target) cusim::yield();
    CUSIM_RESTORE_TID();
#else
    int v;
    for (;;) {
        asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
        if (v >= target) break;
        __nanosleep(20);
    }
#endif
}

// Covariance tile (I, J) straight into the DMMA accumulator fragment layout (16 entries per thread,
// all independent: the Matern evaluations overlap).  Padding rows/columns get the identity; for a
// diagonal tile the strictly-upper 8x8 blocks are zero.
template <int DD>
__device__ __forceinline__ void build_acc(double (&acc)[4][2][2], const double* zr_s, const double* zc_s,
                                          const double* ur_s, const double* uc_s, int ldz, const Hyper& h, int N,
                                          int family, int I, int J, WarpTile wt, const double* __restrict__ etab) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const bool interior = TS * I + TS <= N && TS * J + TS <= N;
    const bool basis = family == 1;
    double zr[4][DD], ur[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int li = wt.m0 + 8 * i + g;
#pragma unroll
        for (int k = 0; k < DD; ++k) zr[i][k] = zr_s[k * ldz + li];
        ur[i] = ur_s[li] * h.inv_g2;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int lj0 = wt.n0 + 8 * j + 2 * t;
        const int bcol = (wt.n0 >> 3) + j;
        double2 zc[DD];
#pragma unroll
        for (int k = 0; k < DD; ++k) zc[k] = *reinterpret_cast<const double2*>(&zc_s[k * ldz + lj0]);
        const double2 uc = *reinterpret_cast<const double2*>(&uc_s[lj0]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int li = wt.m0 + 8 * i + g;
            const int gi = TS * I + li;
            const int brow = (wt.m0 >> 3) + i;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gj = TS * J + lj0 + e;
                double v;
                if (I == J && bcol > brow) {
                    v = 0.0;
                } else if (!interior && (gi >= N || gj >= N)) {
                    v = (gi == gj) ? 1.0 : 0.0;
                } else {
                    double r2 = 0.0;
#pragma unroll
                    for (int k = 0; k < DD; ++k) {
                        const double dz = zr[i][k] - (e ? zc[k].y : zc[k].x);
                        r2 = fma(dz, dz, r2);
                    }
                    v = h.amp * matern52_fast_value(r2, etab);
                    if (basis) v *= fma(ur[i], e ? uc.y : uc.x, h.cb);
                    if (gi == gj) v += h.diag_add;
                }
                acc[i][j][e] = v;
            }
        }
    }
}

// Sum of NV per-thread values over the 256 bulk threads (named barrier 1); result valid in every bulk thread.
template <int NV>
__device__ __forceinline__ void bulk_sum(double (&v)[NV], double* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const double s = warp_sum(v[i]);
        if (lane == 0) scratch[i * NWARPS + w] = s;
    }
    named_bar_sync(1, DAG_BULK_THREADS);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = 0.0;
        for (int k = 0; k < NWARPS; ++k) s += scratch[i * NWARPS + k];
        v[i] = s;
    }
}

// 288 threads: the register file is allocated in groups of four warps, so the budget is that of 384
// threads (168 registers each) -- what ptxas assumes under __launch_bounds__(288).
#define FAB_DAG_KERNEL_ATTR __global__ void __launch_bounds__(DAG_THREADS, 1)
FAB_DAG_KERNEL_ATTR gp_dag_kernel(DagArgs a) {
    FAB_DYN_SMEM(smem_raw);
    const GpData& gd = a.gd;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int Np = gd.Np, N = gd.N, d = gd.d, Pk = gd.Pk;
    const int SL = slice_doubles(d);

    double* slots = reinterpret_cast<double*>(smem_raw);
    double* dinv = slots + (size_t)a.nslot * TILE_ELEMS;      // 512: 8x8 inverses of the tile the chain works on
    double* etab = dinv + 512;                                // exp table
    double* part = etab + EXP_TAB_N;                          // 512: partial sums of the bulk matvec epilogues
    double* red = part + 512;                                 // 128: final reductions
    double* crhs = red + 128;                                 // 64: chain right-hand side
    double* cz = crhs + 64;                                   // 64: chain z_K
    double* rinv = cz + 64;                                   // 8 (+8)
    double* cbase = rinv + 16;
    // resident: zs[k * Np + i] (k < d), us[i], alpha[i], z[i], accw[i];  otherwise two staged slices + alpha slices
    double* zs_all = cbase;
    double* us_all = zs_all + (size_t)d * Np;
    double* sl_r = cbase;
    double* sl_c = sl_r + SL;
    double* al_r = sl_c + SL;
    double* al_c = al_r + TS;
    double* alv = a.resident ? us_all + Np : a.wk.alpha + (size_t)b * Np;
    double* zv = a.resident ? alv + Np : a.wk.zvec + (size_t)b * Np;
    double* accw = a.resident ? zv + Np : a.wk.accw + (size_t)b * Np;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(
        cbase + (a.resident ? (size_t)(d + 4) * Np : (size_t)(2 * SL + 2 * TS)));
    int* sync = reinterpret_cast<int*>(bars + MAX_SLOTS);     // [0] bulk warp arrivals, [1] chain ops done, [2] failure

    const Hyper h = decode_hyper(gd, a.theta + (size_t)b * Pk, a.yerr2[b]);
    double* Lw = a.wk.Lw ? a.wk.Lw + (size_t)b * a.ntiles * TILE_ELEMS : nullptr;
    double* Lx = a.wk_spill ? a.wk_spill + (size_t)b * gd.NT * TILE_ELEMS : nullptr;
    auto ws_tile = [&](int tile) { return tile < a.ntiles ? Lw + (size_t)tile * TILE_ELEMS : Lx + (size_t)(tile - a.ntiles) * TILE_ELEMS; };

    exp_table_fill(etab);
    for (int i = tid; i < Np; i += DAG_THREADS) accw[i] = 0.0;
    if (a.resident) {
        for (int e = tid; e < N * gd.D; e += DAG_THREADS) {
            const int i = e / gd.D, k = e - i * gd.D;
            const double x = __ldg(&gd.X[e]);
            if (k < d) zs_all[k * Np + i] = x * h.scale[k];
            else us_all[i] = x;
        }
        for (int i = tid; i < Np; i += DAG_THREADS) {
            if (i >= N) { for (int k = 0; k < d; ++k) zs_all[k * Np + i] = 0.0; }
            if (i >= N || gd.family != 1) us_all[i] = 0.0;
        }
    }
    if (tid == 0) {
        for (int s = 0; s < MAX_SLOTS; ++s) mbar_init(&bars[s], 1);
        sync[0] = 0; sync[1] = 0; sync[2] = 0;
    }
    __syncthreads();                      // the only block-wide barrier: the two streams part here

    if (w == NWARPS) {
        // ======================================= CHAIN stream =======================================
        double logdet = 0.0, zz = 0.0;
        int first_fail = 0;
        PHBC(55);
        for (int j = 0; j < a.nchain; ++j) {
            const DagOp* op = a.chain + j;
            const int K = __ldg(&op->I), sA = __ldg(&op->slotA), sC = __ldg(&op->slotC), wo = __ldg(&op->wait_other);
            PHBC(56);
            if (wo > 0) {
                if (lane == 0) spin_until(&sync[0], NWARPS * wo);
                __syncwarp();
            }
            PHEC(56);
            double* A = slots + (size_t)sA * TILE_ELEMS;
            double* W = slots + (size_t)sC * TILE_ELEMS;
            PHBC(57);
            const int f = potrf64_warp(A, dinv, rinv);
            PHEC(57);
            if (f && first_fail == 0) first_fail = TS * K + f;
            logdet += log(A[tix(lane, lane)]) + log(A[tix(lane + 32, lane + 32)]);
            PHBC(58);
            inv64_warp(W, A, dinv);
            PHEC(58);
            PHBC(59);
            // z_K = W_KK (r_K - sum_{J<K} L_KJ z_J)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int r = lane + 32 * half, gi = TS * K + r;
                crhs[r] = (gi < N) ? (__ldg(&gd.y[gi]) - gd.mean) - accw[gi] : 0.0;
            }
            __syncwarp();
            double z0, z1;
            tile_matvec_warp(W, crhs, z0, z1);
            zz = fma(z0, z0, fma(z1, z1, zz));
            zv[TS * K + lane] = z0; zv[TS * K + lane + 32] = z1;
            cz[lane] = z0; cz[lane + 32] = z1;
            __syncwarp();
            if (a.mode >= DAG_FIT) {      // alpha_K starts as W_KK^T z_K; the bulk stream adds W_IK^T z_I, I > K
                double a0, a1;
                tile_matvec_t_warp(W, cz, a0, a1);
                alv[TS * K + lane] = a0; alv[TS * K + lane + 32] = a1;
            }
            fence_proxy_async();          // W (and, in FACTOR mode, L) may be written through by a bulk store
            __syncwarp();
            if (lane == 0) sync_arrive(&sync[1]);
            PHEC(59);
        }
        PHEC(55);
        logdet = warp_sum(logdet);
        zz = warp_sum(zz);
        if (lane == 0) {
            double ll = -0.5 * (N * kLog2Pi + 2.0 * logdet) - 0.5 * zz;
            if (first_fail != 0 || !isfinite(ll)) ll = -INFINITY;      // george: non-finite -> -inf (fabolas.py:58-66)
            a.ll[b] = ll;
            a.info[b] = first_fail;
            sync[2] = first_fail;
            sync_arrive(&sync[1]);        // "everything of the chain, including the failure flag, is published"
        }
        return;
    }

    // ========================================= BULK stream =========================================
    const WarpTile wt = warp_tile();
    const int g = lane >> 2, t = lane & 3;
    unsigned phases = 0u;                 // mbarrier phase parity per slot (uniform)
    double acc[4][2][2];
    acc_zero(acc);
    double gacc[MAX_PK + 1];              // [0] log_c0, [1..d] log_M_k, [MAX_D+1] log_gamma2, [MAX_D+2] log_cb, [MAX_PK] yerr2
#pragma unroll
    for (int k = 0; k <= MAX_PK; ++k) gacc[k] = 0.0;
    int cur_r = -1, cur_c = -1;           // tile rows whose coordinate slices are staged (slice mode)

    PHB(40);
    for (int i = 0; i < a.nbulk; ++i) {
        PHB(41);
        const DagOp* op = a.bulk + i;
        const int type = __ldg(&op->type), I = __ldg(&op->I), J = __ldg(&op->J), k = __ldg(&op->k);
        const int sA = __ldg(&op->slotA), sB = __ldg(&op->slotB), sC = __ldg(&op->slotC);
        const int wo = __ldg(&op->wait_other), wait_mask = __ldg(&op->wait_mask), flags = __ldg(&op->flags);
        if (tid == 0 && (flags & DOPF_WAIT_READ)) bulk_s2g_wait_read();
        named_bar_sync(1, DAG_BULK_THREADS);          // every bulk warp is done with the previous operation
        if (wo > 0) {
            if (lane == 0) spin_until(&sync[1], wo);
            __syncwarp();
        }
        if (tid == 0) {
            const int nload = __ldg(&op->nload), st_slot = __ldg(&op->st_slot);
            if (flags & DOPF_DRAIN) bulk_s2g_wait_all();
            if (nload > 0) fence_proxy_async();
            for (int l = 0; l < nload; ++l) {
                const int ls = __ldg(&op->ld_slot[l]), lt = __ldg(&op->ld_tile[l]);
                bulk_g2s(slots + (size_t)ls * TILE_ELEMS, ws_tile(lt), TILE_BYTES, &bars[ls]);
            }
            if (st_slot >= 0)
                bulk_s2g(ws_tile(__ldg(&op->st_tile)), slots + (size_t)st_slot * TILE_ELEMS, TILE_BYTES);
        }
        PHE(41);
        PHB(42);
        for (int s = 0; s < MAX_SLOTS; ++s)
            if ((wait_mask >> s) & 1) {
                mbar_wait(&bars[s], (phases >> s) & 1u);
                phases ^= 1u << s;
            }
        PHE(42);
        PHB(42 + type);
        const double* A = slots + (size_t)(sA < 0 ? 0 : sA) * TILE_ELEMS;
        const double* Bt = slots + (size_t)(sB < 0 ? 0 : sB) * TILE_ELEMS;
        double* C = slots + (size_t)(sC < 0 ? 0 : sC) * TILE_ELEMS;

        if (type == DOP_BUILD || (type == DOP_P2_MMA && (flags & DOPF_LAST))) {
            if (!a.resident) {            // slice mode: stage the coordinates (and alpha) of the two tile rows
                bool staged = false;
                if (cur_r != I) { stage_slice(sl_r, gd, h, I); cur_r = I; staged = true; }
                if (cur_c != J) { stage_slice(sl_c, gd, h, J); cur_c = J; staged = true; }
                if (type == DOP_P2_MMA) {
                    if (tid < TS) al_r[tid] = alv[TS * I + tid];
                    else if (tid < 2 * TS) al_c[tid - TS] = alv[TS * J + tid - TS];
                    staged = true;
                }
                if (staged) named_bar_sync(1, DAG_BULK_THREADS);
            }
        }
        const double* zr_p = a.resident ? zs_all + TS * I : sl_r;
        const double* zc_p = a.resident ? zs_all + TS * J : sl_c;
        const double* ur_p = a.resident ? us_all + TS * I : sl_r + d * TS;
        const double* uc_p = a.resident ? us_all + TS * J : sl_c + d * TS;
        const int ldz = a.resident ? Np : TS;

        switch (type) {
        case DOP_BUILD: {
#define FAB_BUILD_ACC(DD) case DD: build_acc<DD>(acc, zr_p, zc_p, ur_p, uc_p, ldz, h, N, gd.family, I, J, wt, etab); break;
            switch (d) {
                FAB_BUILD_ACC(1) FAB_BUILD_ACC(2) FAB_BUILD_ACC(3) FAB_BUILD_ACC(4)
                FAB_BUILD_ACC(5) FAB_BUILD_ACC(6) FAB_BUILD_ACC(7) FAB_BUILD_ACC(8)
            }
#undef FAB_BUILD_ACC
            break;
        }
        case DOP_MMA_NT:
            tile_mma<false, true, true>(acc, A, Bt, wt, 0, TS);
            break;
        case DOP_PUT:
            acc_store(acc, C, wt);
            fence_proxy_async();
            break;
        case DOP_TRSM: {
            // L(I,J) = U(I,J) W(J,J)^T in place (W lower triangular: k <= n), then accw_I += L(I,J) z_J
            double r2[4][2][2];
            acc_zero(r2);
            tile_mma<false, true, false>(r2, C, Bt, wt, 0, wt.n0 + 16);
            const double* zJ = zv + TS * J;
            double zc0[2], zc1[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) { zc0[j] = zJ[wt.n0 + 8 * j + 2 * t]; zc1[j] = zJ[wt.n0 + 8 * j + 2 * t + 1]; }
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < 2; ++j) s = fma(r2[i4][j][0], zc0[j], fma(r2[i4][j][1], zc1[j], s));
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (t == 0) part[(wt.n0 >> 4) * TS + wt.m0 + 8 * i4 + g] = s;
            }
            named_bar_sync(1, DAG_BULK_THREADS);       // every warp has read its rows of U
            acc_store(r2, C, wt);
            fence_proxy_async();
            if (tid < TS) accw[TS * I + tid] += (part[tid] + part[TS + tid]) + (part[2 * TS + tid] + part[3 * TS + tid]);
            break;
        }
        case DOP_P1_MMA:
            if (flags & DOPF_FIRST) acc_zero(acc);
            // W(J,J) is lower triangular: rows k' < n contribute nothing
            tile_mma<false, false, false>(acc, A, Bt, wt, (k == J) ? wt.n0 : 0, TS);
            break;
        case DOP_P1_FIN: {
            // W(I,J) = -W(I,I) acc  (W(I,I) lower triangular: k <= m), then alpha_J += W(I,J)^T z_I
            acc_store(acc, C, wt);
            named_bar_sync(1, DAG_BULK_THREADS);
            double r2[4][2][2];
            acc_zero(r2);
            tile_mma<false, false, true>(r2, A, C, wt, 0, wt.m0 + 32);
            const double* zI = zv + TS * I;
            double zr4[4];
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) zr4[i4] = zI[wt.m0 + 8 * i4 + g];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double s = 0.0;
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) s = fma(r2[i4][j][e], zr4[i4], s);
                    s += __shfl_xor_sync(0xffffffffu, s, 4);
                    s += __shfl_xor_sync(0xffffffffu, s, 8);
                    s += __shfl_xor_sync(0xffffffffu, s, 16);
                    if (g == 0) part[(wt.m0 >> 5) * TS + wt.n0 + 8 * j + 2 * t + e] = s;
                }
            named_bar_sync(1, DAG_BULK_THREADS);       // every warp has read its columns of the product
            acc_store(r2, C, wt);
            fence_proxy_async();
            if (tid < TS) alv[TS * J + tid] += part[tid] + part[TS + tid];
            break;
        }
        case DOP_P2_MMA:
            if (flags & DOPF_FIRST) acc_zero(acc);
            // W(I,I) is lower triangular: (W_II)[k'][m] = 0 for k' < m
            tile_mma<true, false, false>(acc, A, Bt, wt, (k == I) ? wt.m0 : 0, TS);
            if (flags & DOPF_LAST) {
                PHB(50);
                const double* ar_p = a.resident ? alv + TS * I : al_r;
                const double* ac_p = a.resident ? alv + TS * J : al_c;
#define FAB_CONTRACT_CASE(DD) case DD: contract_tile<DD>(acc, gacc, zr_p, zc_p, ur_p, uc_p, ar_p, ac_p, ldz, h, N, gd.family, I, J, wt, etab); break;
                switch (d) {
                    FAB_CONTRACT_CASE(1) FAB_CONTRACT_CASE(2) FAB_CONTRACT_CASE(3) FAB_CONTRACT_CASE(4)
                    FAB_CONTRACT_CASE(5) FAB_CONTRACT_CASE(6) FAB_CONTRACT_CASE(7) FAB_CONTRACT_CASE(8)
                }
#undef FAB_CONTRACT_CASE
                PHE(50);
            }
            break;
        default: break;
        }
        PHE(42 + type);
        __syncwarp();
        if (lane == 0) sync_arrive(&sync[0]);
    }
    PHB(51);

    // the chain has published the log-likelihood, the failure flag and its share of alpha
    if (lane == 0) spin_until(&sync[1], a.nchain + 1);
    __syncwarp();
    named_bar_sync(1, DAG_BULK_THREADS);
    if (a.mode == DAG_FIT && a.resident)
        for (int i = tid; i < Np; i += DAG_BULK_THREADS) a.wk.alpha[(size_t)b * Np + i] = alv[i];
    if (a.mode == DAG_GRAD && a.grad) {
        bulk_sum<MAX_PK + 1>(gacc, red);
        if (tid == 0) {
            double* go = a.grad + (size_t)b * (Pk + 1);
            const bool bad = sync[2] != 0;             // not positive definite: no factor to differentiate
            go[0] = bad ? NAN : 0.5 * gacc[0];
            for (int q = 0; q < d; ++q) go[1 + q] = bad ? NAN : 0.5 * gacc[1 + q];
            if (gd.family == 1) {
                go[1 + d] = bad ? NAN : 0.5 * gacc[MAX_D + 1];
                go[2 + d] = bad ? NAN : 0.5 * gacc[MAX_D + 2];
            }
            go[Pk] = bad ? NAN : 0.5 * gacc[MAX_PK];
        }
    }
    if (tid == 0) bulk_s2g_wait_all();
    PHE(51);
    PHE(40);
}

}  // namespace fab
