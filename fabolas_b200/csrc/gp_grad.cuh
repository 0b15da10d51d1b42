// gp_grad.cuh -- kernel B: in-place triangular inverse W = L^-1, alpha = K^-1 r, and the gradient of
// the marginal log-likelihood in george's grad_log_likelihood convention:
//     g_k = 0.5 * sum_ij (alpha_i alpha_j - K^-1_ij) dK_ij / dtheta_k ,   K^-1 = W^T W
// One CTA per walker, run after gp_factor_kernel (which left L, the 8x8 inverses and z in the
// workspace).  K^-1 is never stored: each 64x64 tile of W^T W is accumulated in DMMA registers
// and contracted on the spot with dK regenerated from staged 64-row coordinate slices.
//   phase 1 (trtri): column by column, W_JJ = L_JJ^-1, W_IJ = -L_II^-1 sum_{J<=k<I} L_Ik W_kJ,
//                    written over L in the workspace; alpha_J = sum_I W_IJ^T z_I
//   phase 2 (lauum + contraction): (K^-1)_IJ = sum_{k>=I} W_kI^T W_kJ
// The kernel is an interpreter of the static tile program built on the host (tile_program.h):
// slot assignment, TMA prefetches (up to three operations ahead) and waits are all precomputed, so
// no thread spends instructions on cache bookkeeping and shared memory use is independent of N.
#pragma once
#include "gp_common.cuh"
#include "tile_program.h"

namespace fab {

// Coordinates / alpha for the contraction: 64-row slices staged per (I, J) pair (any N), or -- when
// (d + 2) * Np doubles fit next to the tile slots -- everything resident in shared memory from the start.
struct GradArgs {
    GpData gd;
    GpWork wk;
    const double* theta;   // B x Pk
    const double* yerr2;   // B
    double* grad;          // B x (Pk + 1) or null (then only W and alpha are produced)
    const int* info;       // B (from the factor kernel); walkers with info != 0 get NaN gradients
    const TileOp* prog;    // device copy of the tile program
    int nops;
    int nslot;
    int resident;          // 1: coordinates and alpha of all Np rows live in shared memory
};

__host__ __device__ inline size_t grad_smem_bytes(int nslot, int d, int resident_np) {
    const size_t coords = resident_np ? (size_t)(d + 2) * resident_np : (size_t)(2 * (d + 1) * TS + 2 * TS);
    return (size_t)nslot * TILE_BYTES + 8 * (size_t)(2 * 512 + 512 + 128) + 8 * coords + 8 * (MAX_SLOTS + 2) + 64 +
           8 * EXP_TAB_N;
}

// alpha_J += T^T z_I for one 64x64 tile T (rows = block I, cols = block J); alpha lives in the HBM
// workspace (thread t < 64 owns entry t of the block).  Two barriers.
__device__ __forceinline__ void tile_matvec_t(const double* T, const double* __restrict__ zI, double* alphaJ, double* part) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
        const int r = 8 * w + rr;
        const double zr = __ldg(&zI[r]);
        s0 = fma(T[tix(r, lane)], zr, s0);
        s1 = fma(T[tix(r, lane + 32)], zr, s1);
    }
    part[w * 64 + lane] = s0;
    part[w * 64 + lane + 32] = s1;
    __syncthreads();
    if (threadIdx.x < 64) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < NWARPS; ++q) s += part[q * 64 + threadIdx.x];
        alphaJ[threadIdx.x] += s;
    }
    __syncthreads();
}

// Contract one 64x64 tile of (alpha alpha^T - K^-1), held in this warp's DMMA accumulator fragment,
// with dK/dtheta regenerated from the staged slices.  DD = number of Matern axes (compile time).
// Diagonal tiles: only 8x8 blocks on or below the diagonal are visited (weight 2 below, 1 on it).
// gacc: [0] log_c0, [1..DD] log_M_k, [MAX_D+1] log_gamma2, [MAX_D+2] log_cb, [MAX_PK] yerr2.
template <int DD>
__device__ __forceinline__ void contract_tile(const double (&acc)[4][2][2], double (&gacc)[MAX_PK + 1],
                                              const double* zr_s, const double* zc_s, const double* ur_s,
                                              const double* uc_s, const double* al_r, const double* al_c, int ldz,
                                              const Hyper& h, int N, int family, int I, int J, WarpTile wt,
                                              const double* __restrict__ etab) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const bool interior = TS * I + TS <= N && TS * J + TS <= N;
    const bool basis = family == 1;
    double gm[DD];
#pragma unroll
    for (int k = 0; k < DD; ++k) gm[k] = 0.0;
    double g0 = 0.0, gg2 = 0.0, gcb = 0.0, gtr = 0.0;
#pragma unroll 1
    for (int j2 = 0; j2 < 2; ++j2) {
        const int lj0 = wt.n0 + 8 * j2 + 2 * t;
        const int bcol = (wt.n0 >> 3) + j2;
        double2 zc[DD];
#pragma unroll
        for (int k = 0; k < DD; ++k) zc[k] = *reinterpret_cast<const double2*>(&zc_s[k * ldz + lj0]);
        const double2 uc = *reinterpret_cast<const double2*>(&uc_s[lj0]);
        const double2 ac = *reinterpret_cast<const double2*>(&al_c[lj0]);
        double kinv[4][2];                                             // this column block's K^-1 entries (static indices)
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4)
#pragma unroll
            for (int e = 0; e < 2; ++e) kinv[i4][e] = (j2 == 0) ? acc[i4][0][e] : acc[i4][1][e];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
            const int li = wt.m0 + 8 * i4 + g;
            const int gi = TS * I + li;
            const int brow = (wt.m0 >> 3) + i4;
            if (I == J && bcol > brow) continue;                       // warp-uniform
            const double wgt = (I != J || bcol < brow) ? 2.0 : 1.0;
            double zr[DD];
#pragma unroll
            for (int k = 0; k < DD; ++k) zr[k] = zr_s[k * ldz + li];
            const double ur = ur_s[li] * h.inv_g2, ar = al_r[li];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gj = TS * J + lj0 + e;
                if (interior || (gi < N && gj < N)) {
                    const double wij = wgt * (ar * (e ? ac.y : ac.x) - kinv[i4][e]);
                    double tk[DD], r2 = 0.0;
#pragma unroll
                    for (int k = 0; k < DD; ++k) {
                        const double dz = zr[k] - (e ? zc[k].y : zc[k].x);
                        tk[k] = dz * dz;
                        r2 += tk[k];
                    }
                    double m, gf;
                    matern52_fast(r2, m, gf, etab);
                    const double uu = ur * (e ? uc.y : uc.x);
                    const double beta = basis ? (uu + h.cb) : 1.0;
                    const double wa = wij * h.amp;
                    const double wam = wa * m;
                    g0 = fma(wam, beta, g0);
                    const double wg = wa * beta * gf;
#pragma unroll
                    for (int k = 0; k < DD; ++k) gm[k] = fma(wg, tk[k], gm[k]);
                    if (basis) {
                        gg2 = fma(-wam, uu, gg2);
                        gcb = fma(wam, h.cb, gcb);
                    }
                    if (gi == gj) gtr += wij;
                }
            }
        }
    }
    gacc[0] += g0;
#pragma unroll
    for (int k = 0; k < DD; ++k) gacc[1 + k] += gm[k];
    gacc[MAX_D + 1] += gg2;
    gacc[MAX_D + 2] += gcb;
    gacc[MAX_PK] += gtr;
}

__global__ void __launch_bounds__(NTHREADS, 1) gp_grad_kernel(GradArgs a) {
    FAB_DYN_SMEM(smem_raw);
    const GpData& gd = a.gd;
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int Np = gd.Np, NT = gd.NT, N = gd.N, d = gd.d, Pk = gd.Pk;
    const int SL = slice_doubles(d);

    double* slots = reinterpret_cast<double*>(smem_raw);
    double* dinvb = slots + (size_t)a.nslot * TILE_ELEMS;     // 2 x 512
    double* part = dinvb + 2 * 512;                           // 512
    double* red = part + 512;                                 // 128
    double* cbase = red + 128;                                // coordinates: slices or resident arrays
    double* sl_r = cbase;                                     // slice mode: coordinate slice of the row tile I
    double* sl_c = sl_r + SL;                                 //             ... of the column tile J
    double* al_r = sl_c + SL;                                 //             64: alpha of row tile
    double* al_c = al_r + TS;                                 //             64
    double* zs_all = cbase;                                   // resident mode: zs[k * Np + i], us[i], alpha[i]
    double* us_all = zs_all + (size_t)d * Np;
    double* al_all = us_all + Np;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(
        cbase + (a.resident ? (size_t)(d + 2) * Np : (size_t)(2 * SL + 2 * TS)));   // MAX_SLOTS + 2 mbarriers
    double* etab = reinterpret_cast<double*>(bars + MAX_SLOTS + 2 + 6);             // exp table (fast Matern evaluation)

    PHB(20); PHB(21);
    const Hyper h = decode_hyper(gd, a.theta + (size_t)b * Pk, a.yerr2[b]);
    double* Lw = a.wk.Lw + (size_t)b * (NT * (NT + 1) / 2) * TILE_ELEMS;
    const double* Dw = a.wk.Dinv + (size_t)b * NT * 512;
    const double* zg = a.wk.zvec + (size_t)b * Np;
    double* alg = a.wk.alpha + (size_t)b * Np;

    if (a.info[b] != 0) {      // not positive definite: no factor to differentiate
        if (a.grad) for (int k = tid; k <= Pk; k += NTHREADS) a.grad[(size_t)b * (Pk + 1) + k] = NAN;
        return;
    }
    for (int i = tid; i < Np; i += NTHREADS) alg[i] = 0.0;
    exp_table_fill(etab);
    if (a.resident) {
        for (int e = tid; e < N * gd.D; e += NTHREADS) {
            const int i = e / gd.D, k = e - i * gd.D;
            const double x = __ldg(&gd.X[e]);
            if (k < d) zs_all[k * Np + i] = x * h.scale[k];
            else us_all[i] = x;
        }
        for (int i = tid; i < Np; i += NTHREADS) {
            if (i >= N) { for (int k = 0; k < d; ++k) zs_all[k * Np + i] = 0.0; }
            if (i >= N || gd.family != 1) us_all[i] = 0.0;
        }
    }
    if (tid == 0)
        for (int s = 0; s < MAX_SLOTS + 2; ++s) mbar_init(&bars[s], 1);
    __syncthreads();

    const WarpTile wt = warp_tile();
    unsigned phases = 0u;                       // mbarrier phase parity per slot / dinv buffer (uniform)
    double acc[4][2][2];
    acc_zero(acc);
    // accumulators: [0] log_c0, [1..d] log_M_k, [MAX_D+1] log_gamma2, [MAX_D+2] log_cb, [MAX_PK] yerr2
    double gacc[MAX_PK + 1];
#pragma unroll
    for (int k = 0; k <= MAX_PK; ++k) gacc[k] = 0.0;

    PHE(21);
    for (int i = 0; i < a.nops; ++i) {
        PHB(22);
        const TileOp* opp = a.prog + i;
        const int type = __ldg(&opp->type), I = __ldg(&opp->I), J = __ldg(&opp->J), k = __ldg(&opp->k);
        const int sA = __ldg(&opp->slotA), sB = __ldg(&opp->slotB), sC = __ldg(&opp->slotC);
        const int wait_mask = __ldg(&opp->wait_mask), dbuf = __ldg(&opp->dinv_buf), flags = __ldg(&opp->flags);
        if (tid == 0 && (flags & OPF_WAIT_READ)) bulk_s2g_wait_read();   // a stored tile's slot is overwritten from here on
        __syncthreads();                        // every warp is done with the previous operation
        if (tid == 0) {
            const int nload = __ldg(&opp->nload);
            if (nload > 0 && (flags & OPF_DRAIN_STORES)) bulk_s2g_wait_all();
            for (int l = 0; l < nload; ++l) {
                const int ls = __ldg(&opp->ld_slot[l]), lt = __ldg(&opp->ld_tile[l]);
                if (ls >= TP_DINV_SLOT0)
                    bulk_g2s(dinvb + (ls - TP_DINV_SLOT0) * 512, Dw + (size_t)lt * 512, 4096, &bars[MAX_SLOTS + ls - TP_DINV_SLOT0]);
                else
                    bulk_g2s(slots + (size_t)ls * TILE_ELEMS, Lw + (size_t)lt * TILE_ELEMS, TILE_BYTES, &bars[ls]);
            }
        }
        for (int s = 0; s < MAX_SLOTS + 2; ++s) {
            const int bit = (s < MAX_SLOTS) ? s : TP_DINV_SLOT0 + (s - MAX_SLOTS);
            if ((wait_mask >> bit) & 1) {
                mbar_wait(&bars[s], (phases >> s) & 1u);
                phases ^= 1u << s;
            }
        }
        PHE(22);
        const double* A = slots + (size_t)sA * TILE_ELEMS;
        if (type == OP_INV_DIAG) {
            double* C = slots + (size_t)sC * TILE_ELEMS;
            PHB(23);
            tile_trsm_left<true, false>(C, A, dinvb + dbuf * 512);
            PHE(23); PHB(24);
            tile_matvec_t(C, zg + TS * J, alg + TS * J, part);
            PHE(24); PHB(25);
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) bulk_s2g(Lw + (size_t)tri_index(J, J) * TILE_ELEMS, C, TILE_BYTES);
            PHE(25);
        } else if (type == OP_P1_MMA) {
            PHB(26);
            if (flags & OPF_FIRST) acc_zero(acc);
            // W_JJ is lower triangular: rows k' < n contribute nothing
            tile_mma<false, false, false>(acc, A, slots + (size_t)sB * TILE_ELEMS, wt, (k == J) ? wt.n0 : 0, TS);
            PHE(26);
        } else if (type == OP_P1_FIN) {
            double* C = slots + (size_t)sC * TILE_ELEMS;
            PHB(27);
            acc_store(acc, C, wt);
            __syncthreads();
            tile_trsm_left<false, true>(C, A, dinvb + dbuf * 512);
            PHE(27); PHB(24);
            tile_matvec_t(C, zg + TS * I, alg + TS * J, part);
            PHE(24); PHB(25);
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) bulk_s2g(Lw + (size_t)tri_index(I, J) * TILE_ELEMS, C, TILE_BYTES);
            PHE(25);
        } else {   // OP_P2_MMA
            PHB(28);
            if (a.resident) {
                if (i == 0 || (__ldg(&(opp - 1)->type) != OP_P2_MMA)) {       // first phase-2 operation: alpha is final
                    for (int q = tid; q < Np; q += NTHREADS) al_all[q] = alg[q];
                    __syncthreads();
                }
            } else if (flags & OPF_NEW_PAIR) {
                stage_slice(sl_r, gd, h, I);
                stage_slice(sl_c, gd, h, J);
                if (tid < TS) al_r[tid] = alg[TS * I + tid];
                else if (tid < 2 * TS) al_c[tid - TS] = alg[TS * J + tid - TS];
            }
            if (flags & OPF_FIRST) acc_zero(acc);
            // W_II is lower triangular: (W_II)[k'][m] = 0 for k' < m
            tile_mma<true, false, false>(acc, A, slots + (size_t)sB * TILE_ELEMS, wt, (k == I) ? wt.m0 : 0, TS);
            PHE(28);
            if (flags & OPF_LAST) {
                PHB(29);
                if (!a.resident && (flags & OPF_NEW_PAIR)) __syncthreads();      // slices staged in this very operation
                const double* zr_p = a.resident ? zs_all + TS * I : sl_r;
                const double* zc_p = a.resident ? zs_all + TS * J : sl_c;
                const double* ur_p = a.resident ? us_all + TS * I : sl_r + d * TS;
                const double* uc_p = a.resident ? us_all + TS * J : sl_c + d * TS;
                const double* ar_p = a.resident ? al_all + TS * I : al_r;
                const double* ac_p = a.resident ? al_all + TS * J : al_c;
                const int ldz = a.resident ? Np : TS;
#define FAB_CONTRACT_CASE(DD) case DD: contract_tile<DD>(acc, gacc, zr_p, zc_p, ur_p, uc_p, ar_p, ac_p, ldz, h, N, gd.family, I, J, wt, etab); break;
                switch (d) {
                    FAB_CONTRACT_CASE(1) FAB_CONTRACT_CASE(2) FAB_CONTRACT_CASE(3) FAB_CONTRACT_CASE(4)
                    FAB_CONTRACT_CASE(5) FAB_CONTRACT_CASE(6) FAB_CONTRACT_CASE(7) FAB_CONTRACT_CASE(8)
                }
#undef FAB_CONTRACT_CASE
                PHE(29);
            }
        }
    }
    PHB(30);
    __syncthreads();
    if (!a.grad) {
        if (tid == 0) bulk_s2g_wait_all();
        return;
    }
    block_sum<MAX_PK + 1>(gacc, red);
    if (tid == 0) {
        double* go = a.grad + (size_t)b * (Pk + 1);
        go[0] = 0.5 * gacc[0];
        for (int k = 0; k < d; ++k) go[1 + k] = 0.5 * gacc[1 + k];
        if (gd.family == 1) {
            go[1 + d] = 0.5 * gacc[MAX_D + 1];
            go[2 + d] = 0.5 * gacc[MAX_D + 2];
        }
        go[Pk] = 0.5 * gacc[MAX_PK];
        bulk_s2g_wait_all();
    }
    PHE(30); PHE(20);
}

}  // namespace fab
