// gp_grad.cuh -- kernel B: in-place triangular inverse W = L^-1, alpha = K^-1 r, and the gradient of
// the marginal log-likelihood in george's grad_log_likelihood convention:
//     g_k = 0.5 * sum_ij (alpha_i alpha_j - K^-1_ij) dK_ij / dtheta_k ,   K^-1 = W^T W
// One CTA per walker, run after gp_factor_kernel (which left L, the 8x8 inverses and z in the
// workspace).  K^-1 is never stored: each 64x64 tile of W^T W is accumulated in DMMA registers
// and contracted on the spot with dK regenerated from the staged observations.
//   phase 1 (trtri): column by column, W_JJ = L_JJ^-1, W_IJ = -L_II^-1 sum_{J<=k<I} L_Ik W_kJ,
//                    written over L in the workspace; alpha_J = sum_I W_IJ^T z_I
//   phase 2 (lauum + contraction): (K^-1)_IJ = sum_{k>=I} W_kI^T W_kJ
#pragma once
#include "gp_common.cuh"

namespace fab {

struct GradArgs {
    GpData gd;
    GpWork wk;
    const double* theta;   // B x Pk
    const double* yerr2;   // B
    double* grad;          // B x (Pk + 1) or null (then only W and alpha are produced)
    const int* info;       // B (from the factor kernel); walkers with info != 0 get NaN gradients
    int nslot;
};

__host__ __device__ inline size_t grad_smem_bytes(int nslot, int Np, int d) {
    return (size_t)nslot * TILE_BYTES + 8 * (size_t)(512 + (d + 3) * Np + 512 + 128) + 8 * MAX_SLOTS + 64;
}

// alpha_J += T^T z_I for one 64x64 tile T (rows = block I, cols = block J).  Two barriers.
__device__ __forceinline__ void tile_matvec_t(const double* T, const double* zI, double* alphaJ, double* part) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
        const int r = 8 * w + rr;
        const double zr = zI[r];
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
// with dK/dtheta regenerated from the staged observations.  DD = number of Matern axes (compile time).
// Diagonal tiles: only 8x8 blocks on or below the diagonal are visited (weight 2 below, 1 on it).
// gacc: [0] log_c0, [1..DD] log_M_k, [MAX_D+1] log_gamma2, [MAX_D+2] log_cb, [MAX_PK] yerr2.
template <int DD>
__device__ __forceinline__ void contract_tile(const double (&acc)[4][2][2], double (&gacc)[MAX_PK + 1],
                                              const double* zs, const double* us, const double* al, const Hyper& h,
                                              int Np, int N, int family, int I, int J, WarpTile wt) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const bool interior = TS * I + TS <= N && TS * J + TS <= N;
    const bool basis = family == 1;
    double gm[DD];
#pragma unroll
    for (int k = 0; k < DD; ++k) gm[k] = 0.0;
    double g0 = 0.0, gg2 = 0.0, gcb = 0.0, gtr = 0.0;
#pragma unroll 1
    for (int j2 = 0; j2 < 2; ++j2) {
        const int gj0 = TS * J + wt.n0 + 8 * j2 + 2 * t;
        const int bcol = (wt.n0 >> 3) + j2;
        double2 zc[DD];
#pragma unroll
        for (int k = 0; k < DD; ++k) zc[k] = *reinterpret_cast<const double2*>(&zs[k * Np + gj0]);
        const double2 uc = *reinterpret_cast<const double2*>(&us[gj0]);
        const double2 ac = *reinterpret_cast<const double2*>(&al[gj0]);
        double kinv[4][2];                                             // this column block's K^-1 entries (static indices)
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4)
#pragma unroll
            for (int e = 0; e < 2; ++e) kinv[i4][e] = (j2 == 0) ? acc[i4][0][e] : acc[i4][1][e];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
            const int gi = TS * I + wt.m0 + 8 * i4 + g;
            const int brow = (wt.m0 >> 3) + i4;
            if (I == J && bcol > brow) continue;                       // warp-uniform
            const double wgt = (I != J || bcol < brow) ? 2.0 : 1.0;
            double zr[DD];
#pragma unroll
            for (int k = 0; k < DD; ++k) zr[k] = zs[k * Np + gi];
            const double ur = us[gi] * h.inv_g2, ar = al[gi];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gj = gj0 + e;
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
                    matern52(r2, m, gf);
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

    double* slots = reinterpret_cast<double*>(smem_raw);
    double* dinv = slots + (size_t)a.nslot * TILE_ELEMS;
    double* zs = dinv + 512;
    double* us = zs + (size_t)d * Np;
    double* zv = us + Np;
    double* al = zv + Np;
    double* part = al + Np;            // 512
    double* red = part + 512;          // 128
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(red + 128);

    const Hyper h = decode_hyper(gd, a.theta + (size_t)b * Pk, a.yerr2[b]);
    double* Lw = a.wk.Lw + (size_t)b * (NT * (NT + 1) / 2) * TILE_ELEMS;
    const double* Dw = a.wk.Dinv + (size_t)b * NT * 512;

    if (a.info[b] != 0) {      // not positive definite: no factor to differentiate
        if (a.grad) for (int k = tid; k <= Pk; k += NTHREADS) a.grad[(size_t)b * (Pk + 1) + k] = NAN;
        return;
    }

    for (int e = tid; e < N * gd.D; e += NTHREADS) {
        const int i = e / gd.D, k = e - i * gd.D;
        const double x = __ldg(&gd.X[e]);
        if (k < d) zs[k * Np + i] = x * h.scale[k];
        if (gd.family == 1 && k == d) us[i] = x;
    }
    for (int i = tid; i < Np; i += NTHREADS) {
        if (i >= N) {
            for (int k = 0; k < d; ++k) zs[k * Np + i] = 0.0;
            us[i] = 0.0;
        }
        if (gd.family != 1 && i < N) us[i] = 0.0;
        zv[i] = a.wk.zvec[(size_t)b * Np + i];
        al[i] = 0.0;
    }
    TileCache tc; tc.init(slots, a.nslot, bar);
    __syncthreads();
    const WarpTile wt = warp_tile();

    // ======================= phase 1: W = L^-1 in place, alpha = W^T z =======================
    for (int J = 0; J < NT; ++J) {
        // W tiles of earlier columns are not needed again in this phase
        auto dead = [J](int id) { return id_kind(id) == 1 && id_col(id) < J; };
        {
            const int sl = tc.acquire(Lw, tile_id(0, J, J), dead, 0u);
            const int sw = tc.alloc(tile_id(1, J, J), dead, pin(sl));
            for (int e = tid; e < 512; e += NTHREADS) dinv[e] = Dw[(size_t)J * 512 + e];
            __syncthreads();
            tile_trsm_left<true, false>(tc.ptr(sw), tc.ptr(sl), dinv);
            tile_matvec_t(tc.ptr(sw), zv + TS * J, al + TS * J, part);
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) bulk_s2g(Lw + (size_t)tri_index(J, J) * TILE_ELEMS, tc.ptr(sw), TILE_BYTES);
        }
        for (int I = J + 1; I < NT; ++I) {
            const int ss = tc.alloc(tile_id(1, I, J), dead, 0u);
            double acc[4][2][2];
            acc_zero(acc);
            for (int k = J; k < I; ++k) {
                const int sa = tc.acquire(Lw, tile_id(0, I, k), dead, pin(ss));
                const int sb = tc.acquire(Lw, tile_id(1, k, J), dead, pin(ss) | pin(sa));
                // W_JJ is lower triangular: rows k' < n contribute nothing
                tile_mma<false, false, false>(acc, tc.ptr(sa), tc.ptr(sb), wt, (k == J) ? wt.n0 : 0, TS);
            }
            acc_store(acc, tc.ptr(ss), wt);
            const int sl = tc.acquire(Lw, tile_id(0, I, I), dead, pin(ss));
            for (int e = tid; e < 512; e += NTHREADS) dinv[e] = Dw[(size_t)I * 512 + e];
            __syncthreads();
            tile_trsm_left<false, true>(tc.ptr(ss), tc.ptr(sl), dinv);
            tile_matvec_t(tc.ptr(ss), zv + TS * I, al + TS * J, part);
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) bulk_s2g(Lw + (size_t)tri_index(I, J) * TILE_ELEMS, tc.ptr(ss), TILE_BYTES);
        }
    }
    __syncthreads();
    if (a.wk.alpha) for (int i = tid; i < Np; i += NTHREADS) a.wk.alpha[(size_t)b * Np + i] = al[i];
    if (!a.grad) {
        if (tid == 0) bulk_s2g_wait_all();
        return;
    }

    // ======================= phase 2: K^-1 tiles + gradient contraction =======================
    // accumulators: [0] log_c0, [1..d] log_M_k, [d+1] log_gamma2, [d+2] log_cb, [MAX_PK] yerr2
    double gacc[MAX_PK + 1];
#pragma unroll
    for (int k = 0; k <= MAX_PK; ++k) gacc[k] = 0.0;

    // Static schedule of tile products (J, I, k), k = I..NT-1 innermost; the operands of step s+1 are
    // requested (TMA bulk copies on per-slot mbarriers) before the DMMA work of step s starts.
    {
        int J = 0, I = 0, k = 0;
        auto dead0 = [](int id) { return id_kind(id) == 0; };
        int sa = tc.issue(Lw, tile_id(1, k, I), dead0, 0u);
        int sb = tc.issue(Lw, tile_id(1, k, J), dead0, pin(sa));
        double acc[4][2][2];
        acc_zero(acc);
        bool more = true;
        while (more) {
            // successor of (J, I, k)
            int nJ = J, nI = I, nk = k + 1;
            if (nk >= NT) { nI = I + 1; nk = nI; if (nI >= NT) { nJ = J + 1; nI = nJ; nk = nJ; } }
            const bool has_next = nJ < NT;
            __syncthreads();                   // every warp is done with the previous step's operands
            int na = -1, nb = -1;
            if (has_next) {
                auto dead = [nJ](int id) { return id_kind(id) == 0 || id_col(id) < nJ; };
                na = tc.issue(Lw, tile_id(1, nk, nI), dead, pin(sa) | pin(sb));
                nb = tc.issue(Lw, tile_id(1, nk, nJ), dead, pin(sa) | pin(sb) | pin(na));
            }
            tc.wait(sa);
            tc.wait(sb);
            // W_II is lower triangular: (W_II)[k'][m] = 0 for k' < m
            tile_mma<true, false, false>(acc, tc.ptr(sa), tc.ptr(sb), wt, (k == I) ? wt.m0 : 0, TS);
            if (k == NT - 1) {
                // contraction with dK over this warp's 32 x 16 fragment
#define FAB_CONTRACT_CASE(DD) case DD: contract_tile<DD>(acc, gacc, zs, us, al, h, Np, N, gd.family, I, J, wt); break;
                switch (d) {
                    FAB_CONTRACT_CASE(1) FAB_CONTRACT_CASE(2) FAB_CONTRACT_CASE(3) FAB_CONTRACT_CASE(4)
                    FAB_CONTRACT_CASE(5) FAB_CONTRACT_CASE(6) FAB_CONTRACT_CASE(7) FAB_CONTRACT_CASE(8)
                }
#undef FAB_CONTRACT_CASE
                acc_zero(acc);
            }
            J = nJ; I = nI; k = nk; sa = na; sb = nb;
            more = has_next;
        }
    }
    __syncthreads();
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
}

}  // namespace fab
