// gp_predict.cuh -- kernel C: batched GP prediction for M candidates x B resident models.
//
// Replaces george's GP.predict (fabolas.py:187; entropy_search.py:130; acquisitions.py:31,103,109):
//     mu* = K* alpha + mean ,   cov* = k(t,t) - K* K^-1 K*^T = k(t,t) - V^T V ,  V = W K*^T ,  W = L^-1
// One CTA per (candidate block, model).  The cross-covariance panel K*^T (N x CB) is generated in
// shared memory from the staged observations and candidates and never touches HBM; the W tiles of
// the model stream from L2/HBM through a TMA bulk copy and feed DMMA tile products; variance,
// the optional dot product with a per-model vector (used by the information-gain kernel:
// cov(x, last representer) = k - vrl . V_x) and the optional full covariance of the block are
// reduced straight from the accumulator fragments.
#pragma once
#include "gp_common.cuh"

namespace fab {

struct PredictArgs {
    GpData gd;
    GpWork wk;               // Lw holds W = L^-1 tiles, alpha
    const double* theta;     // B x Pk (context copy made by factorize)
    const double* yerr2;     // B
    const double* T;         // M x D candidates (last column = basis input as given by the caller)
    long long T_bstride;     // 0: all models share T ; M*D: model b reads T + b*T_bstride (per-model points)
    int M;
    int B;                   // resident models
    double* mu;              // B x M or null
    double* var;             // B x M or null
    double* cov;             // B x M x M or null (M <= 64 only)
    const double* dotvec;    // B x Np or null
    double* dot;             // B x M or null:  dot[b][c] = sum_obs V[obs][c] * dotvec[b][obs]
    double* vcol;            // B x Np or null: column `vcol_index` of V = W K*^T (first block only)
    int vcol_index;
    int nbuf;                // W tile buffers in shared memory: 2 (load of tile t+1 overlaps the product of tile t) or 1
    int jc;                  // tile rows of the cross-covariance panel resident at a time (>= NT: the whole panel)
    double* vpart;           // jc < NT: per-CTA scratch for the partial rows of V (grid x NT x 64 x CB doubles)
};

// Shared memory: the resident part of the panel (jc x 64 x CB), then a region used twice -- first by the scaled
// observation coordinates, the basis column and alpha of the rows being generated, then by the W tile buffers of the
// product loop --, the optional covariance scratch, the candidate coordinates, the exp table, partial sums.
__host__ __device__ inline size_t predict_union_doubles(int rows, int d, int nbuf) {
    const size_t coords = (size_t)(d + 2) * rows, wb = (size_t)nbuf * TILE_ELEMS;
    return coords > wb ? coords : wb;
}
__host__ __device__ inline size_t predict_smem_bytes(int CB, int jc, int d, bool cov, int nbuf) {
    size_t doubles = (size_t)jc * TS * CB + predict_union_doubles(jc * TS, d, nbuf) + (cov ? TS * CB : 0) +
                     (size_t)(d + 1) * CB + EXP_TAB_N + 3 * 8 * CB + 16;
    return doubles * 8 + 64;
}

// One 8-row strip of the cross-covariance panel per warp and pass: lane = candidate column, four rows in
// flight (independent Matern chains), candidate coordinates in registers, observation coordinates broadcast.
// Generates the `nt` tile rows whose observations are staged in zs / us / al (row stride `ld`); `row0` is the
// global index of the first staged observation.  pmu accumulates the mean partials.
template <int DD, int CB>
__device__ __forceinline__ void predict_panel(double* Ks, const double* zs, const double* us, const double* al,
                                              const double* tz, const double* tu, const Hyper& h, int family, int N,
                                              int ld, int nt, int row0, int ncand, const double* __restrict__ etab,
                                              double (&pmu)[CB / 32]) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool basis = family == 1;
#pragma unroll
    for (int q = 0; q < CB / 32; ++q) {
        const int c = lane + 32 * q;
        double tc[DD];
#pragma unroll
        for (int k = 0; k < DD; ++k) tc[k] = tz[k * CB + c];
        const double tuc = basis ? tu[c] * h.inv_g2 : 0.0;
        const double cbv = basis ? h.cb : 1.0;
        const bool cok = c < ncand;
        double pm = 0.0;
        for (int J = 0; J < nt; ++J) {
            double* P = Ks + (size_t)J * TS * CB;
#pragma unroll
            for (int r4 = 0; r4 < 2; ++r4) {
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int li = TS * J + 8 * w + 4 * r4 + u;
                    double r2 = 0.0;
#pragma unroll
                    for (int k = 0; k < DD; ++k) {
                        const double dz = zs[k * ld + li] - tc[k];
                        r2 = fma(dz, dz, r2);
                    }
                    const double kv = h.amp * matern52_fast_value(r2, etab) * fma(us[li], tuc, cbv);
                    v[u] = (row0 + li < N && cok) ? kv : 0.0;
                    pm = fma(v[u], al[li], pm);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) P[tixw<CB>(8 * w + 4 * r4 + u, c)] = v[u];
            }
        }
        pmu[q] += pm;
    }
}

// Grid: every CTA takes the work items (candidate block, model) item = linear CTA index, + number of CTAs, ...
// (the whole panel fits: one item per CTA, grid (blocks, B); panel in chunks: a persistent grid, one scratch region
// for the partial rows of V per CTA).
template <int CB>
__global__ void __launch_bounds__(NTHREADS, 1) gp_predict_kernel(PredictArgs a) {
    constexpr int MI = (CB == 64) ? 4 : 2;
    constexpr int NJ = 2;
    constexpr int NRG = (CB == 64) ? 2 : 4;        // warp row groups
    constexpr int NACC = MI * NJ * 2;
    FAB_DYN_SMEM(smem_raw);
    const GpData& gd = a.gd;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int Np = gd.Np, NT = gd.NT, N = gd.N, d = gd.d, D = gd.D;
    const bool want_cov = a.cov != nullptr;
    const int JC = a.jc < NT ? a.jc : NT;
    const int rows_c = JC * TS;

    double* Ks = reinterpret_cast<double*>(smem_raw);            // JC panels of 64 x CB
    double* uni = Ks + (size_t)JC * TS * CB;                      // coordinates, then W tile buffers
    double* zs = uni;
    double* us = zs + (size_t)d * rows_c;
    double* al = us + rows_c;
    double* Wbuf = uni;
    double* Vscr = uni + predict_union_doubles(rows_c, d, a.nbuf);
    double* tz = Vscr + (want_cov ? TS * CB : 0);                  // d x CB
    double* tu = tz + (size_t)d * CB;                              // CB
    double* etab = tu + CB;
    double* part = etab + EXP_TAB_N;                               // 3 x 8 x CB
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(part + 3 * 8 * CB);   // [2]

    exp_table_fill(etab);
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
    unsigned waits[2] = {0u, 0u};          // completed waits per W buffer: the parity of the next one (uniform)
    const WarpTile wt = (CB == 64) ? warp_tile() : warp_tile_32();
    const int nb = a.nbuf;
    const long long nblk = (a.M + CB - 1) / CB;
    const long long ncta = (long long)gridDim.x * gridDim.y, cta = (long long)blockIdx.y * gridDim.x + blockIdx.x;
    double* vpart = a.vpart ? a.vpart + (size_t)cta * NT * NACC * NTHREADS : nullptr;

    for (long long item = cta; item < nblk * a.B; item += ncta) {
        const int b = (int)(item / nblk), blk = (int)(item - (long long)b * nblk);
        const Hyper h = decode_hyper(gd, a.theta + (size_t)b * gd.Pk, a.yerr2[b]);
        const double* Lw = a.wk.Lw + (size_t)b * (NT * (NT + 1) / 2) * TILE_ELEMS;
        const int c_base = blk * CB;
        __syncthreads();                   // the previous item is done with tz / tu / part (and the barriers are set up)
        for (int e = tid; e < CB * D; e += NTHREADS) {
            const int c = e / D, k = e - c * D;
            const int gc = c_base + c;
            const double x = (gc < a.M) ? __ldg(&a.T[(size_t)b * a.T_bstride + (size_t)gc * D + k]) : 0.0;
            if (k < d) tz[k * CB + c] = x * h.scale[k];
            if (k == d) tu[c] = x;
        }
        if (gd.family != 1) for (int c = tid; c < CB; c += NTHREADS) tu[c] = 0.0;

        double pmu[CB / 32];
#pragma unroll
        for (int q = 0; q < CB / 32; ++q) pmu[q] = 0.0;
        double css[NJ][2], cdot[NJ][2];
#pragma unroll
        for (int j = 0; j < NJ; ++j) css[j][0] = css[j][1] = cdot[j][0] = cdot[j][1] = 0.0;
        double cacc[4][2][2];
        acc_zero(cacc);

        for (int c0 = 0; c0 < NT; c0 += JC) {
            const int c1 = (c0 + JC < NT) ? c0 + JC : NT;
            // ---- observations of the tile rows [c0, c1): scaled coordinates, basis column, alpha ----
            __syncthreads();               // the W tile buffers of the previous chunk are no longer read
            const int r0 = TS * c0, nrow = TS * (c1 - c0);
            for (int e = tid; e < nrow * D; e += NTHREADS) {
                const int li = e / D, k = e - li * D;
                const int gi = r0 + li;
                const double x = (gi < N) ? __ldg(&gd.X[(size_t)gi * D + k]) : 0.0;
                if (k < d) zs[k * rows_c + li] = x * h.scale[k];
                if (gd.family == 1 && k == d) us[li] = x;
            }
            for (int li = tid; li < nrow; li += NTHREADS) {
                if (gd.family != 1) us[li] = 0.0;
                al[li] = a.wk.alpha[(size_t)b * Np + r0 + li];
            }
            __syncthreads();

            // ---- cross-covariance panel rows K*^T (obs x cand) + mean partials ----
#define FAB_PANEL_CASE(DD) case DD: predict_panel<DD, CB>(Ks, zs, us, al, tz, tu, h, gd.family, N, rows_c, c1 - c0, r0, a.M - c_base, etab, pmu); break;
            switch (d) {
                FAB_PANEL_CASE(1) FAB_PANEL_CASE(2) FAB_PANEL_CASE(3) FAB_PANEL_CASE(4)
                FAB_PANEL_CASE(5) FAB_PANEL_CASE(6) FAB_PANEL_CASE(7) FAB_PANEL_CASE(8)
            }
#undef FAB_PANEL_CASE
            fence_proxy_async();           // the coordinate region is about to be overwritten by bulk copies
            __syncthreads();

            // ---- V_I (+)= sum_{J in chunk, J <= I} W_IJ K*_J, reduced as soon as a row block is complete; W tiles
            // stream through `nbuf` buffers in the order (I, J) = (c0, c0) (c0+1, c0) (c0+1, c0+1) ... ----
            int tcount = 0;
            if (tid == 0) bulk_g2s(Wbuf, Lw + (size_t)tri_index(c0, c0) * TILE_ELEMS, TILE_BYTES, &bar[0]);
            for (int I = c0; I < NT; ++I) {
                double acc[MI][NJ][2];
                if (c0 > 0) {              // rows of V started by the earlier chunks
#pragma unroll
                    for (int i = 0; i < MI; ++i)
#pragma unroll
                        for (int j = 0; j < NJ; ++j)
#pragma unroll
                            for (int e = 0; e < 2; ++e)
                                acc[i][j][e] = vpart[((size_t)I * NACC + (i * NJ + j) * 2 + e) * NTHREADS + tid];
                } else {
                    acc_zero(acc);
                }
                const int jhi = (I < c1 - 1) ? I : c1 - 1;
                for (int J = c0; J <= jhi; ++J, ++tcount) {
                    const int cur = (nb == 2) ? (tcount & 1) : 0;
                    // the tile after (I, J) in this chunk's order
                    int In = I, Jn = J + 1;
                    if (Jn > jhi) { In = I + 1; Jn = c0; }
                    const bool more = In < NT;
                    if (nb == 2 && tid == 0 && more)       // the other buffer was released by the barrier below
                        bulk_g2s(Wbuf + (size_t)(cur ^ 1) * TILE_ELEMS, Lw + (size_t)tri_index(In, Jn) * TILE_ELEMS, TILE_BYTES,
                                 &bar[cur ^ 1]);
                    mbar_wait(&bar[cur], waits[cur] & 1u);
                    ++waits[cur];
                    // W_II is lower triangular: columns k > m contribute nothing
                    tile_mma<false, false, false, MI, NJ, TS, CB>(acc, Wbuf + (size_t)cur * TILE_ELEMS, Ks + (size_t)(J - c0) * TS * CB,
                                                                  wt, 0, (J == I) ? wt.m0 + 8 * MI : TS);
                    __syncthreads();       // everyone done with this buffer before the next bulk copy lands in it
                    if (nb == 1 && tid == 0 && more)
                        bulk_g2s(Wbuf, Lw + (size_t)tri_index(In, Jn) * TILE_ELEMS, TILE_BYTES, &bar[0]);
                }
                if (I >= c1) {             // the later chunks continue this row block
#pragma unroll
                    for (int i = 0; i < MI; ++i)
#pragma unroll
                        for (int j = 0; j < NJ; ++j)
#pragma unroll
                            for (int e = 0; e < 2; ++e)
                                vpart[((size_t)I * NACC + (i * NJ + j) * 2 + e) * NTHREADS + tid] = acc[i][j][e];
                    continue;
                }
#pragma unroll
                for (int i = 0; i < MI; ++i) {
                    const double dvr = a.dotvec ? __ldg(&a.dotvec[(size_t)b * Np + TS * I + wt.m0 + 8 * i + g]) : 0.0;
#pragma unroll
                    for (int j = 0; j < NJ; ++j)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            css[j][e] = fma(acc[i][j][e], acc[i][j][e], css[j][e]);
                            cdot[j][e] = fma(acc[i][j][e], dvr, cdot[j][e]);
                        }
                }
                if (a.vcol && blk == 0) {
#pragma unroll
                    for (int i = 0; i < MI; ++i)
#pragma unroll
                        for (int j = 0; j < NJ; ++j)
#pragma unroll
                            for (int e = 0; e < 2; ++e)
                                if (wt.n0 + 8 * j + 2 * t + e == a.vcol_index)
                                    a.vcol[(size_t)b * Np + TS * I + wt.m0 + 8 * i + g] = acc[i][j][e];
                }
                if (CB == 64 && want_cov) {
                    acc_store<MI, NJ, CB>(acc, Vscr, wt);
                    __syncthreads();
                    tile_mma<true, false, false, 4, 2, CB, CB>(cacc, Vscr, Vscr, warp_tile(), 0, TS);
                    __syncthreads();
                }
            }
        }
#pragma unroll
        for (int q = 0; q < CB / 32; ++q) part[w * CB + lane + 32 * q] = pmu[q];
        // reduce the per-thread column partials over the 8 row lanes (g) of the warp
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    css[j][e] += __shfl_xor_sync(0xffffffffu, css[j][e], o);
                    cdot[j][e] += __shfl_xor_sync(0xffffffffu, cdot[j][e], o);
                }
            }
        const int rg = (CB == 64) ? (w >> 2) : (w >> 1);
        if (g == 0) {
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = wt.n0 + 8 * j + 2 * t + e;
                    part[(8 + rg) * CB + c] = css[j][e];
                    part[(16 + rg) * CB + c] = cdot[j][e];
                }
        }
        __syncthreads();
        if (tid < CB && c_base + tid < a.M) {
            const int c = tid;
            double m = 0.0, ss = 0.0, dd = 0.0;
#pragma unroll
            for (int q = 0; q < NWARPS; ++q) m += part[q * CB + c];
#pragma unroll
            for (int q = 0; q < NRG; ++q) { ss += part[(8 + q) * CB + c]; dd += part[(16 + q) * CB + c]; }
            const size_t o = (size_t)b * a.M + c_base + c;
            if (a.mu) a.mu[o] = m + gd.mean;
            if (a.var) {
                const double kss = h.amp * ((gd.family == 1) ? fma(tu[c] * tu[c], h.inv_g2, h.cb) : 1.0);
                a.var[o] = kss - ss;
            }
            if (a.dot) a.dot[o] = dd;
        }
        if (CB == 64 && want_cov) {
            const WarpTile wc = warp_tile();
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int c1 = wc.m0 + 8 * i + g, c2 = wc.n0 + 8 * j + 2 * t + e;
                        if (c1 < a.M && c2 < a.M) {
                            double r2 = 0.0;
                            for (int k = 0; k < d; ++k) {
                                const double dz = tz[k * CB + c1] - tz[k * CB + c2];
                                r2 = fma(dz, dz, r2);
                            }
                            double m, gf;
                            matern52(r2, m, gf);
                            double v = h.amp * m;
                            if (gd.family == 1) v *= fma(tu[c1] * tu[c2], h.inv_g2, h.cb);
                            a.cov[((size_t)b * a.M + c1) * a.M + c2] = v - cacc[i][j][e];
                        }
                    }
        }
    }
}

// EI (acquisitions.py:40-75) over B x M predictions + per-model argmax.  One CTA per model.
//   sd = sqrt(var) ; u = (mu - y_max - xi) / sd ; ei = sd (u Phi(u) + phi(u)) ; 0 where sd <= 0
struct EiArgs {
    const double* mu; const double* var; int M;
    const double* y_max;      // B
    double xi;
    double* ei;               // B x M or null
    double* best_val;         // B or null
    long long* best_idx;      // B or null
};

__device__ __forceinline__ double ei_value(double mu, double var, double y_max, double xi) {
    const double sd = sqrt(var);
    if (!(sd > 0.0)) return (sd <= 0.0) ? 0.0 : NAN;      // acquisitions.py:73 ; NaN variance stays NaN
    const double u = (mu - y_max - xi) / sd;
    const double cdf = 0.5 * erfc(-u * 0.70710678118654752440);
    const double pdf = exp(-0.5 * u * u) * 0.39894228040143267794;
    return sd * (u * cdf + pdf);
}

// np.argmax ordering of (value, index) pairs: the first occurrence of the maximum wins, and a NaN counts as the
// maximum (np.argmax returns the index of the first NaN -- expected_improvement.py:92-93 inherits that when a
// predictive variance is NaN); index < 0 = "no candidate yet".
__device__ __forceinline__ bool argmax_takes(double ov, long long oi, double best, long long bi) {
    if (oi < 0) return false;
    if (bi < 0) return true;
    const bool on = isnan(ov), bn = isnan(best);
    if (on != bn) return on;
    if (on) return oi < bi;
    return ov > best || (ov == best && oi < bi);
}

// Grid (G, B): CTA (x, b) scans candidates [x * chunk, (x + 1) * chunk) of model b.  G == 1 writes the result,
// otherwise the (value, index) pair of the chunk goes to part_val / part_idx [b * G + x] and ei_argmax_final_kernel
// finishes (np.argmax: the first occurrence of the maximum wins at every level).
__global__ void __launch_bounds__(NTHREADS) ei_argmax_kernel(EiArgs a, long long chunk, double* part_val,
                                                             long long* part_idx) {
    __shared__ double sval[NWARPS];
    __shared__ long long sidx[NWARPS];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const double ym = a.y_max[b];
    double best = -INFINITY;
    long long bi = -1;
    const long long c0 = (long long)blockIdx.x * chunk;
    const long long c1 = (c0 + chunk < (long long)a.M) ? c0 + chunk : (long long)a.M;
    for (long long c = c0 + tid; c < c1; c += NTHREADS) {
        const double v = ei_value(a.mu[(size_t)b * a.M + c], a.var[(size_t)b * a.M + c], ym, a.xi);
        if (a.ei) a.ei[(size_t)b * a.M + c] = v;
        if (argmax_takes(v, c, best, bi)) { best = v; bi = c; }      // first maximum wins within a thread (ascending c)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (argmax_takes(ov, oi, best, bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { sval[w] = best; sidx[w] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int q = 1; q < NWARPS; ++q)
            if (argmax_takes(sval[q], sidx[q], best, bi)) { best = sval[q]; bi = sidx[q]; }
        if (gridDim.x == 1) {
            if (a.best_val) a.best_val[b] = best;
            if (a.best_idx) a.best_idx[b] = bi;          // np.argmax: first occurrence of the maximum
        } else {
            part_val[(size_t)b * gridDim.x + blockIdx.x] = best;
            part_idx[(size_t)b * gridDim.x + blockIdx.x] = bi;
        }
    }
}

__global__ void ei_argmax_final_kernel(EiArgs a, int G, const double* part_val, const long long* part_idx) {
    const int b = blockIdx.x, lane = threadIdx.x;       // one warp per model
    double best = -INFINITY;
    long long bi = -1;
    for (int x = lane; x < G; x += 32) {                // chunks ascend with x: the first maximum wins
        const double v = part_val[(size_t)b * G + x];
        const long long i = part_idx[(size_t)b * G + x];
        if (argmax_takes(v, i, best, bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (argmax_takes(ov, oi, best, bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) {
        if (a.best_val) a.best_val[b] = best;
        if (a.best_idx) a.best_idx[b] = bi;
    }
}

}  // namespace fab
