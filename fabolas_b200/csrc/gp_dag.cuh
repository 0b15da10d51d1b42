// gp_dag.cuh -- the fused GP kernel: covariance build + tiled Cholesky + log-likelihood and, by mode,
// W = L^-1 + alpha (fit) and the K^-1 tile products + gradient contraction (grad), one CTA per walker.
//
// Replaces, for a whole batch of walkers at once, george's GP.compute / GP.log_likelihood
// (fabolas.py:62-66, entropy_search.py:73-77), the solver state GP.predict needs (fabolas.py:187) and
// grad_log_likelihood.  The CTA is warp-specialised: eight BULK warps execute the
// throughput stream (Matern tile generation straight into DMMA accumulator registers, 64^3 DMMA tile
// products, panel solves in GEMM form against the inverted diagonal tile, the dK contraction) and four
// CHAIN warps execute the latency-bound stream (in-tile Cholesky and triangular inverse of each diagonal
// tile, forward-substitution vectors, log-determinant).  Both interpret the static program built on the
// host (dag_program.h): slots, TMA prefetches, write-through stores and cross-stream waits are all
// precomputed; the streams meet only through two shared-memory completion counters (release/acquire).
// The covariance matrix and its inverse are never materialised.
#pragma once
#include "gp_common.cuh"
#include "chain_ops.cuh"
#include "dag_program.h"

namespace fab {

// Accumulator layout of the bulk warps (FAB_DAG_LAYOUT, prims.cuh).  A warp issues one DMMA every ~32 cycles whatever
// else runs on its SM sub-partition (two warps together fill the pipe, one alone gets half of it), so an op lasts as
// long as its BUSIEST WARP -- and most ops are not full products: panel solves and inverse products have a triangular
// operand (k-range proportional to the column or to the row of the output block), diagonal tiles need their lower
// triangle only.  With one 32 x 16 block per warp (round 1) the busiest warp of every such op did the work of a full
// product (profiles/r02m_fine_trace_layout0.txt: bodies of 0.5 .. 2.1 us side by side, 64 us of arrival spread per
// evaluation).  Layout 2 gives every warp TWO 16 x 16 units (r, c) and (3 - r, 3 - c) of the 4 x 4 unit grid: the
// k-ranges of a warp's two units add up to 5/8 of a full product for every triangular operand, whether the range goes
// with the row or with the column, and a diagonal tile costs its busiest warp 6 of 8 blocks.
constexpr int DAG_LAYOUT = FAB_DAG_LAYOUT;
static_assert(DAG_LAYOUT >= 0 && DAG_LAYOUT <= 2, "gp_dag_kernel: accumulator layout 0, 1 or 2");
constexpr int DAG_NBW = FAB_DAG_NBW;               // bulk warps
constexpr int DAG_NU = (DAG_LAYOUT == 2) ? 2 : 1;  // accumulator units per warp
constexpr int DAG_MI = (DAG_LAYOUT == 0) ? 4 : 2;  // 8-row blocks per unit: 32 x 16 or 16 x 16
constexpr int DAG_RB = TS / (8 * DAG_MI);          // row blocks of units over a tile (2 or 4)
constexpr int DAG_BULK_THREADS = 32 * DAG_NBW;     // warps 0..DAG_NBW-1
constexpr int DAG_THREADS = DAG_BULK_THREADS + 32 * NCW;   // + the chain warps
typedef double DagUnit[DAG_MI][2][2];              // one unit in DMMA accumulator-fragment layout
struct DagAcc { DagUnit u[DAG_NU]; };
struct DagTiles { WarpTile u[DAG_NU]; };           // origins of this warp's units inside the 64 x 64 tile

__device__ __forceinline__ DagTiles dag_tiles() {
    const int w = threadIdx.x >> 5;
    DagTiles wt;
    if constexpr (DAG_LAYOUT == 0) {
        wt.u[0] = warp_tile();
    } else if constexpr (DAG_LAYOUT == 1) {
        // 16 warps as 4 x 4 units.  Warps w, w + 4, w + 8, w + 12 share an SM sub-partition s = w & 3: row r = w >> 2
        // gets column (s - r) & 3, one unit of every row and of every column per sub-partition
        wt.u[0].m0 = (w >> 2) * 16;
        wt.u[0].n0 = (((w & 3) - (w >> 2)) & 3) * 16;
    } else {
        // warp w = (h, cp, rp): rows {ra, rb} = {0, 3} or {1, 2}, columns {ca, cb} likewise; h = 0: units (ra, ca) and
        // (rb, cb), h = 1: (ra, cb) and (rb, ca).  Warps w and w + 4 share an SM sub-partition and the 2 x 2 unit set.
        const int rp = w & 1, cp = (w >> 1) & 1, hh = (w >> 2) & 1;
        const int ra = rp ? 1 : 0, rb = 3 - ra, ca = cp ? 1 : 0, cb = 3 - ca;
        wt.u[0].m0 = 16 * ra; wt.u[0].n0 = 16 * (hh ? cb : ca);
        wt.u[DAG_NU - 1].m0 = 16 * rb; wt.u[DAG_NU - 1].n0 = 16 * (hh ? ca : cb);
    }
    return wt;
}
__device__ __forceinline__ void dag_acc_zero(DagAcc& acc) {
#pragma unroll
    for (int u = 0; u < DAG_NU; ++u) acc_zero(acc.u[u]);
}
__device__ __forceinline__ void dag_acc_load(DagAcc& acc, const double* C, const DagTiles& wt) {
#pragma unroll
    for (int u = 0; u < DAG_NU; ++u) acc_load(acc.u[u], C, wt.u[u]);
}
__device__ __forceinline__ void dag_acc_store(const DagAcc& acc, double* C, const DagTiles& wt) {
#pragma unroll
    for (int u = 0; u < DAG_NU; ++u) acc_store(acc.u[u], C, wt.u[u]);
}
// acc += sign * opA(A) opB(B) over k in [k_lo, k_hi) per unit; the bounds follow the unit for a triangular operand:
enum { KR_ZERO = 0, KR_N0 = 1, KR_M0 = 2 };        // k_lo = 0 | n0 | m0
enum { KR_FULL = 0, KR_N0_END = 1, KR_M0_END = 2 };// k_hi = 64 | n0 + 16 | m0 + rows of the unit
// lower: only the blocks on and below the block diagonal of the result are needed (diagonal tile)
template <bool TA, bool TB, bool NEG, int L = DAG_LAYOUT>
__device__ __forceinline__ void dag_mma(DagAcc& acc, const double* __restrict__ A, const double* __restrict__ B,
                                        const DagTiles& wt, int lo_sel, int hi_sel, bool lower) {
    if constexpr (L == 2) {
        // full product: both units in ONE loop, eight independent DMMA chains in flight (N = 2048: 64.1 -> 62.7 ms; the
        // same for the shared part of two triangular k-ranges measured no gain)
        if (lo_sel == KR_ZERO && hi_sel == KR_FULL && !lower) {
            tile_mma_pair<TA, TB, NEG>(acc.u[0], acc.u[DAG_NU - 1], A, B, wt.u[0], wt.u[DAG_NU - 1], 0, TS);
            return;
        }
    }
#pragma unroll
    for (int u = 0; u < DAG_NU; ++u) {
        const WarpTile t = wt.u[u];
        const int k_lo = lo_sel == KR_ZERO ? 0 : (lo_sel == KR_N0 ? t.n0 : t.m0);
        const int k_hi = hi_sel == KR_FULL ? TS : (hi_sel == KR_N0_END ? t.n0 + 16 : t.m0 + 8 * DAG_MI);
        if (!lower) {
            tile_mma<TA, TB, NEG, DAG_MI, 2>(acc.u[u], A, B, t, k_lo, k_hi);
        } else if constexpr (L == 0) {
            tile_mma_lower<TA, TB, NEG>(acc.u[u], A, B, t, k_lo);       // (k_hi == 64 for every lower product)
        } else {
            if (t.n0 < t.m0) tile_mma<TA, TB, NEG, DAG_MI, 2>(acc.u[u], A, B, t, k_lo, k_hi);
            else if (t.n0 == t.m0) tile_mma<TA, TB, NEG, DAG_MI, 2, TS, TS, true>(acc.u[u], A, B, t, k_lo, k_hi);
        }
    }
}

// Peer exchange of the per-walker results (multi-GPU likelihood step): instead of an
// all-gather after the kernel, every CTA stores its walker's result row [grad (Pk+1) | ll | info] straight into the
// gather buffer of EVERY rank (plain NVLink / NVSwitch peer stores by a few lanes; the local rank is just another
// entry) as soon as the walker is done -- the transfers of the early CTAs overlap the evaluations still running.
// The LAST CTA of the launch to finish (local completion counter, device-scope release/acquire chain from every
// CTA's stores) issues one system-scope fence, publishes this rank's step number in ITS OWN arrival slot on every
// rank, and then waits until every rank's slot on THIS rank has reached the step: when the kernel completes, the
// gather buffer of the step is complete -- one launch per step, no collective, no separate wait kernel.  One slot
// per source rank (a generation number, not a summed counter): a fast peer's arrival for step s + 1 can never stand
// in for a missing step-s row.  world == 0: no exchange.
constexpr int FAB_MAX_PEERS = 8;
struct PeerTable {
    double* buf[FAB_MAX_PEERS];        // this step's gather buffer of rank r (world x rows_per_rank x row_doubles)
    unsigned* slots[FAB_MAX_PEERS];    // arrival slots of rank r: slots[r][src] = last step rank `src` has delivered
    unsigned* done;                    // local: CTAs of this rank that have finished (monotonic over steps)
    int* err;                          // local: set when a wait gave up (a peer never delivered)
    unsigned done_target;              // value of *done once every CTA of THIS step has finished
    unsigned step;                     // generation number of this step (1, 2, ...)
    int world, rank, rows_per_rank, row_doubles;
};

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
    int host_io;           // 1: theta / yerr2 (and ll / info / grad) are mapped HOST memory: one warp fetches the walker's
                           //    record over PCIe and the CTA decodes it from shared memory (fab_gp_loglik_batch_host)
    int resident;          // 1: scaled coordinates, z, accw, alpha of all Np rows live in shared memory
    int ntiles;            // NT (NT + 1) / 2: workspace tile indices >= ntiles address the spill region wk_spill
    double* wk_spill;      // B x NT tiles (FACTOR mode only: inverted diagonal tiles that had to leave shared memory)
    PeerTable peer;
};

__host__ __device__ inline size_t dag_smem_bytes(int nslot, int d, int resident_np) {
    const size_t coords = resident_np ? (size_t)(d + 4) * resident_np : (size_t)(2 * (d + 1) * TS + 2 * TS);
    return (size_t)nslot * TILE_BYTES + 8 * (size_t)(512 + EXP_TAB_N + 512 + 256 + 64 + 64 + 64 + 4 * TS) + 8 * coords +
           8 * MAX_SLOTS + 64 + 512;       // mbarriers, stream counters, op-record ring (4 x 112 B)
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

// Covariance tile (I, J) straight into the DMMA accumulator fragment layout (16 entries per thread).
// Four passes of four entries (two rows x two adjacent columns) keep the register footprint small enough
// for the four Matern chains of a pass to overlap; evaluation is branch-free (padding and the strictly-upper
// blocks of a diagonal tile are selected afterwards: identity resp. zero).  The pass index is a template
// parameter: with a run-time pass loop the accumulator registers could only be addressed through a switch, and 40 %
// of the 439 instructions of a pass were register moves and selects (the ops are issue-bound, not FP64-bound).
template <int DD, int Q>
__device__ __forceinline__ void build_pass(DagUnit& acc, const double* zr_s, const double* zc_s,
                                           const double* ur_s, const double* uc_s, int ldz, const Hyper& h, int N,
                                           bool basis, double cb, int I, int J, WarpTile wt, const double* __restrict__ etab) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    constexpr int ip = Q >> 1, j = Q & 1;
    const int lj0 = wt.n0 + 8 * j + 2 * t;
    const int bcol = (wt.n0 >> 3) + j;
    // diagonal tile: both 8x8 blocks of this pass lie above the block diagonal -> zero, nothing to evaluate
    if (I == J && bcol > (wt.m0 >> 3) + 2 * ip + 1) {
        acc[2 * ip][j][0] = acc[2 * ip][j][1] = acc[2 * ip + 1][j][0] = acc[2 * ip + 1][j][1] = 0.0;
        return;
    }
    double2 zc[DD];
#pragma unroll
    for (int k = 0; k < DD; ++k) zc[k] = *reinterpret_cast<const double2*>(&zc_s[k * ldz + lj0]);
    const double2 uc = *reinterpret_cast<const double2*>(&uc_s[lj0]);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int li = wt.m0 + 8 * (2 * ip + r) + g;
        const int gi = TS * I + li;
        const bool upper = I == J && bcol > (li >> 3);
        double zr[DD];
#pragma unroll
        for (int k = 0; k < DD; ++k) zr[k] = zr_s[k * ldz + li];
        const double ur = basis ? ur_s[li] * h.inv_g2 : 0.0;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int gj = TS * J + lj0 + e;
            double r2 = 0.0;
#pragma unroll
            for (int k = 0; k < DD; ++k) {
                const double dz = zr[k] - (e ? zc[k].y : zc[k].x);
                r2 = fma(dz, dz, r2);
            }
            double v = h.amp * matern52_fast_value(r2, etab) * fma(ur, e ? uc.y : uc.x, cb);
            v = (gi == gj) ? v + h.diag_add : v;
            v = (gi >= N || gj >= N) ? ((gi == gj) ? 1.0 : 0.0) : v;
            acc[2 * ip + r][j][e] = upper ? 0.0 : v;
        }
    }
}
template <int DD>
__device__ __forceinline__ void build_acc(DagUnit& acc, const double* zr_s, const double* zc_s,
                                          const double* ur_s, const double* uc_s, int ldz, const Hyper& h, int N,
                                          int family, int I, int J, WarpTile wt, const double* __restrict__ etab) {
    const bool basis = family == 1;
    const double cb = basis ? h.cb : 1.0;
    // (the warp barriers keep the compiler from overlapping the passes, which costs registers and gained nothing)
    build_pass<DD, 0>(acc, zr_s, zc_s, ur_s, uc_s, ldz, h, N, basis, cb, I, J, wt, etab); __syncwarp();
    build_pass<DD, 1>(acc, zr_s, zc_s, ur_s, uc_s, ldz, h, N, basis, cb, I, J, wt, etab);
    if constexpr (DAG_MI == 4) {
        __syncwarp();
        build_pass<DD, 2>(acc, zr_s, zc_s, ur_s, uc_s, ldz, h, N, basis, cb, I, J, wt, etab); __syncwarp();
        build_pass<DD, 3>(acc, zr_s, zc_s, ur_s, uc_s, ldz, h, N, basis, cb, I, J, wt, etab);
    }
}
template <int DD>
__device__ __forceinline__ void dag_build(DagAcc& acc, const double* zr_s, const double* zc_s, const double* ur_s,
                                          const double* uc_s, int ldz, const Hyper& h, int N, int family, int I, int J,
                                          const DagTiles& wt, const double* __restrict__ etab) {
#pragma unroll
    for (int u = 0; u < DAG_NU; ++u) {
        if (u) __syncwarp();
        build_acc<DD>(acc.u[u], zr_s, zc_s, ur_s, uc_s, ldz, h, N, family, I, J, wt.u[u], etab);
    }
}

// Contract one 64x64 tile of (alpha alpha^T - K^-1), held in this warp's DMMA accumulator fragment, with
// dK/dtheta regenerated from the coordinates: branch-free (weights select what counts), four entries per
// pass as in build_acc.  Diagonal tiles: weight 2 below the diagonal 8x8 blocks, 1 on them, 0 above.
// The nine partial sums of the warp go to gsum[warp][.] (fixed order => deterministic):
//   [0] log_c0, [1..DD] log_M_k, [MAX_D+1] log_gamma2, [MAX_D+2] log_cb, [MAX_PK] yerr2.
struct ContractSums {            // per-thread partial sums of one contraction
    double g0, gg2, gcb, gtr;
};
template <int DD, int Q>
__device__ __forceinline__ void contract_pass(const DagUnit& acc, ContractSums& cs, double (&gm)[DD],
                                              const double* zr_s, const double* zc_s, const double* ur_s,
                                              const double* uc_s, const double* al_r, const double* al_c, int ldz,
                                              const Hyper& h, int N, bool basis, double cb, int I, int J, WarpTile wt,
                                              const double* __restrict__ etab) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    constexpr int ip = Q >> 1, j = Q & 1;
    const int lj0 = wt.n0 + 8 * j + 2 * t;
    const int bcol = (wt.n0 >> 3) + j;
    if (I == J && bcol > (wt.m0 >> 3) + 2 * ip + 1) return;     // weight 0 on both blocks of this pass
    double2 zc[DD];
#pragma unroll
    for (int k = 0; k < DD; ++k) zc[k] = *reinterpret_cast<const double2*>(&zc_s[k * ldz + lj0]);
    const double2 uc = *reinterpret_cast<const double2*>(&uc_s[lj0]);
    const double2 ac = *reinterpret_cast<const double2*>(&al_c[lj0]);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int li = wt.m0 + 8 * (2 * ip + r) + g;
        const int gi = TS * I + li;
        const int brow = li >> 3;
        double zr[DD];
#pragma unroll
        for (int k = 0; k < DD; ++k) zr[k] = zr_s[k * ldz + li];
        const double ur = basis ? ur_s[li] * h.inv_g2 : 0.0;
        const double ar = al_r[li];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int gj = TS * J + lj0 + e;
            double wgt = (I != J || bcol < brow) ? 2.0 : ((bcol == brow) ? 1.0 : 0.0);
            wgt = (gi >= N || gj >= N) ? 0.0 : wgt;
            const double wij = wgt * (ar * (e ? ac.y : ac.x) - acc[2 * ip + r][j][e]);
            double r2 = 0.0;
#pragma unroll
            for (int k = 0; k < DD; ++k) {
                const double dz = zr[k] - (e ? zc[k].y : zc[k].x);
                r2 = fma(dz, dz, r2);
            }
            double m, gf;
            matern52_fast(r2, m, gf, etab);
            const double uu = ur * (e ? uc.y : uc.x);
            const double beta = uu + cb;
            const double wa = wij * h.amp;
            const double wam = wa * m;
            cs.g0 = fma(wam, beta, cs.g0);
            const double wg = wa * beta * gf;
#pragma unroll
            for (int k = 0; k < DD; ++k) {          // dz^2 recomputed: cheaper than keeping it across the Matern chain
                const double dz = zr[k] - (e ? zc[k].y : zc[k].x);
                gm[k] = fma(wg * dz, dz, gm[k]);
            }
            cs.gg2 = fma(-wam, uu, cs.gg2);
            cs.gcb = fma(wam, cb, cs.gcb);
            cs.gtr += (gi == gj) ? wij : 0.0;
        }
    }
}
template <int DD>
__device__ __forceinline__ void contract_acc(const DagAcc& acc, double* gsum_w,
                                             const double* zr_s, const double* zc_s, const double* ur_s,
                                             const double* uc_s, const double* al_r, const double* al_c, int ldz,
                                             const Hyper& h, int N, int family, int I, int J, const DagTiles& wts,
                                             const double* __restrict__ etab) {
    const int lane = threadIdx.x & 31;
    const bool basis = family == 1;
    const double cb = basis ? h.cb : 1.0;
    double gm[DD];
#pragma unroll
    for (int k = 0; k < DD; ++k) gm[k] = 0.0;
    ContractSums cs;
    cs.g0 = cs.gg2 = cs.gcb = cs.gtr = 0.0;
    // the pass index is a template parameter (static accumulator registers, see build_pass); two passes (eight Matern
    // chains) in flight: 267 -> 263 us; four do not fit the register budget
#pragma unroll
    for (int u = 0; u < DAG_NU; ++u) {
        const WarpTile wt = wts.u[u];
        if (u) __syncwarp();
        contract_pass<DD, 0>(acc.u[u], cs, gm, zr_s, zc_s, ur_s, uc_s, al_r, al_c, ldz, h, N, basis, cb, I, J, wt, etab);
        contract_pass<DD, 1>(acc.u[u], cs, gm, zr_s, zc_s, ur_s, uc_s, al_r, al_c, ldz, h, N, basis, cb, I, J, wt, etab);
        if constexpr (DAG_MI == 4) {
            __syncwarp();
            contract_pass<DD, 2>(acc.u[u], cs, gm, zr_s, zc_s, ur_s, uc_s, al_r, al_c, ldz, h, N, basis, cb, I, J, wt, etab);
            contract_pass<DD, 3>(acc.u[u], cs, gm, zr_s, zc_s, ur_s, uc_s, al_r, al_c, ldz, h, N, basis, cb, I, J, wt, etab);
        }
    }
    double g0 = warp_sum(cs.g0), gg2 = warp_sum(cs.gg2), gcb = warp_sum(cs.gcb), gtr = warp_sum(cs.gtr);
#pragma unroll
    for (int k = 0; k < DD; ++k) gm[k] = warp_sum(gm[k]);
    if (lane == 0) {
        gsum_w[0] += g0;
#pragma unroll
        for (int k = 0; k < DD; ++k) gsum_w[1 + k] += gm[k];
        if (basis) { gsum_w[MAX_D + 1] += gg2; gsum_w[MAX_D + 2] += gcb; }
        gsum_w[MAX_PK] += gtr;
    }
}

// One evaluation by the whole CTA (DAG_THREADS threads): walker record `b` of a.theta / a.yerr2 / a.ll /
// a.info / a.grad and of the per-walker workspaces.  A CTA may run any number of evaluations one after the
// other (the device-resident ensemble sampler, mcmc.cuh): `first` marks the CTA's first one (exp table and
// mbarriers are set up once), `phases` carries the mbarrier parities of the bulk warps across evaluations, and
// the caller separates two evaluations by a __syncthreads().
// `side`: scalar work of the caller that does not depend on the result (the sampler's prior density and acceptance
// draw); thread 0 runs it when the bulk stream has finished its program and would otherwise only wait for the chain.
struct DagNoSide { __device__ __forceinline__ void operator()() const {} };
template <int DD, class Side = DagNoSide>
__device__ __forceinline__ void dag_eval(const DagArgs& a, const int b, unsigned char* smem_raw, unsigned& phases,
                                         const bool first, Side side = Side()) {
    const GpData& gd = a.gd;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int Np = gd.Np, N = gd.N, Pk = gd.Pk;
    constexpr int d = DD;                 // == gd.d (the host dispatches on it)
    const int SL = slice_doubles(d);

    double* slots = reinterpret_cast<double*>(smem_raw);
    double* dinv = slots + (size_t)a.nslot * TILE_ELEMS;      // 512: 8x8 inverses of the tile the chain works on
    double* etab = dinv + 512;                                // exp table
    double* part = etab + EXP_TAB_N;                          // 512: partial sums of the bulk matvec epilogues
    double* red = part + 512;                                 // 256: final reductions (16 warps x 12)
    double* crhs = red + 256;                                 // 64: chain right-hand side
    double* cz = crhs + 64;                                   // 64: chain z_K
    double* rinv = cz + 64;                                   // 64: reciprocal pivots, 8 per 8-column step
    double* cpart = rinv + 64;                                // 256: partial sums across the chain warps
    double* cbase = cpart + 4 * TS;
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
    int* sync = reinterpret_cast<int*>(bars + MAX_SLOTS);     // [0] bulk warp arrivals, [1] chain ops done, [2] failure, [3] per-tile failure
    // ring of the next bulk op records (28 ints each), filled by cp.async three ops ahead by lanes 0..6 of warp 0:
    // no thread ever waits for an op record or a load list in global memory
    int* ring = sync + 16;
    const bool ring_lane = w == 0 && lane < 7;
    auto ring_fetch = [&](int opi) {          // every call commits exactly one group (empty past the end)
        if (ring_lane) {
            if (opi < a.nbulk) cp_async16(ring + (opi & 3) * 28 + 4 * lane, reinterpret_cast<const int4*>(a.bulk + opi) + lane);
            cp_async_commit();
        }
    };
    ring_fetch(0); ring_fetch(1); ring_fetch(2);

    const double* thp = a.theta + (size_t)b * Pk;
    double ye;
    if (a.host_io) {                      // zero-copy input: Pk + 1 loads per CTA cross the bus, not Pk + 1 per warp
        if (tid <= Pk) part[tid] = (tid < Pk) ? thp[tid] : a.yerr2[b];
        __syncthreads();
        thp = part;                       // (free until the first panel-solve epilogue)
        ye = part[Pk];
    } else {
        ye = a.yerr2[b];
    }
    const Hyper hfull = decode_hyper(gd, thp, ye);
    Hyper h;                              // scalars only: hfull.scale[] is indexed dynamically and lives in local memory
    h.amp = hfull.amp; h.inv_g2 = hfull.inv_g2; h.cb = hfull.cb; h.diag_add = hfull.diag_add;
    double* Lw = a.wk.Lw ? a.wk.Lw + (size_t)b * a.ntiles * TILE_ELEMS : nullptr;
    double* Lx = a.wk_spill ? a.wk_spill + (size_t)b * gd.NT * TILE_ELEMS : nullptr;
    auto ws_tile = [&](int tile) { return tile < a.ntiles ? Lw + (size_t)tile * TILE_ELEMS : Lx + (size_t)(tile - a.ntiles) * TILE_ELEMS; };

    if (first) exp_table_fill(etab);
    for (int i = tid; i < Np; i += DAG_THREADS) accw[i] = 0.0;
    if (a.resident) {
        for (int e = tid; e < N * gd.D; e += DAG_THREADS) {
            const int i = e / gd.D, k = e - i * gd.D;
            const double x = __ldg(&gd.X[e]);
            if (k < d) zs_all[k * Np + i] = x * hfull.scale[k];
            else us_all[i] = x;
        }
        for (int i = tid; i < Np; i += DAG_THREADS) {
            if (i >= N) { for (int k = 0; k < d; ++k) zs_all[k * Np + i] = 0.0; }
            if (i >= N || gd.family != 1) us_all[i] = 0.0;
        }
    }
    if (tid == 0) {
        if (first)
            for (int s = 0; s < MAX_SLOTS; ++s) mbar_init(&bars[s], 1);
        sync[0] = 0; sync[1] = 0; sync[2] = 0; sync[3] = 0;
    }
    if (ring_lane) cp_async_wait<2>();    // op record 0 has landed; the barrier publishes it
    __syncthreads();                      // the only block-wide barrier of an evaluation: the two streams part here

    if (w >= DAG_NBW) {
        // ======================================= CHAIN stream =======================================
        const int cw = w - DAG_NBW;
        double logdet = 0.0, zz = 0.0;
        int first_fail = 0;                   // (warp 0, lane 0)
        PHBC(55);
        for (int j = 0; j < a.nchain; ++j) {
            const DagOp* op = a.chain + j;
            const int K = __ldg(&op->I), sA = __ldg(&op->slotA), sC = __ldg(&op->slotC), wo = __ldg(&op->wait_other);
            PHBC(56);
            TRC(j, 0);
            if (wo > 0) {
                if (lane == 0) spin_until(&sync[0], DAG_NBW * wo);
                __syncwarp();
            }
            PHEC(56);
            TRC(j, 1);
            double* A = slots + (size_t)sA * TILE_ELEMS;
            double* W = slots + (size_t)sC * TILE_ELEMS;
            // the observations of this tile row are fetched now (L2 latency) and used after the factorisation
            double yv[2] = {0.0, 0.0};
            if (cw == 0) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int gi = TS * K + lane + 32 * half;
                    if (gi < N) yv[half] = __ldg(&gd.y[gi]) - gd.mean;
                }
            }
            PHBC(57);
            const int rows_here = min(TS, N - TS * K);
            potrf64_chain(A, W, dinv, rinv, &sync[3], cw, (rows_here + 7) >> 3);
            PHEC(57);
            if (cw == 0 && lane == 0 && sync[3] != 0) {
                if (first_fail == 0) first_fail = TS * K + sync[3];
                sync[3] = 0;
            }
            if (lane < 16) logdet += log(A[tix(16 * cw + lane, 16 * cw + lane)]);
            fence_proxy_async();          // W (and, in FACTOR mode, L) may be written through by a bulk store
            chain_bar();
            if (cw == 0 && lane == 0) sync_arrive(&sync[1]);      // event 2j + 1: L(K,K) and W(K,K) are final
            TRC(j, 2);
            PHBC(59);
            // z_K = W_KK (r_K - sum_{J<K} L_KJ z_J)
            if (cw == 0) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int r = lane + 32 * half, gi = TS * K + r;
                    crhs[r] = (gi < N) ? yv[half] - accw[gi] : 0.0;
                }
            }
            chain_bar();
            zz += tile_matvec_rows(W, crhs, zv + TS * K, cz, cw);
            if (a.mode >= DAG_FIT) {      // alpha_K starts as W_KK^T z_K; the bulk stream adds W_IK^T z_I, I > K
                chain_bar();
                double a0, a1;
                tile_matvec_t_rows(W, cz, a0, a1, cw);
                cpart[cw * TS + lane] = a0; cpart[cw * TS + lane + 32] = a1;
                chain_bar();
                if (cw == 0) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int c = lane + 32 * half;
                        alv[TS * K + c] = (cpart[c] + cpart[TS + c]) + (cpart[2 * TS + c] + cpart[3 * TS + c]);
                    }
                }
            }
            chain_bar();
            if (cw == 0 && lane == 0) sync_arrive(&sync[1]);      // event 2j + 2: z_K (and alpha_K) published
            TRC(j, 3);
            PHEC(59);
        }
        PHEC(55);
        logdet = warp_sum(logdet);
        zz = warp_sum(zz);
        if (lane == 0) { cpart[2 * cw] = logdet; cpart[2 * cw + 1] = zz; }
        chain_bar();
        if (cw == 0 && lane == 0) {
            logdet = (cpart[0] + cpart[2]) + (cpart[4] + cpart[6]);
            zz = (cpart[1] + cpart[3]) + (cpart[5] + cpart[7]);
            double ll = -0.5 * (N * kLog2Pi + 2.0 * logdet) - 0.5 * zz;
            if (first_fail != 0 || !isfinite(ll)) ll = -INFINITY;      // george: non-finite -> -inf (fabolas.py:58-66)
            a.ll[b] = ll;
            a.info[b] = first_fail;
            sync[2] = first_fail;
            sync_arrive(&sync[1]);        // "everything of the chain, including the failure flag, is published"
        }
        return;                           // (inlined: leaves dag_eval, not the kernel)
    }

    // ========================================= BULK stream =========================================
    const DagTiles wt = dag_tiles();
    const int g = lane >> 2, t = lane & 3;
    DagAcc acc;
    dag_acc_zero(acc);
    double* gsum_w = red + w * (MAX_PK + 1);   // this warp's gradient partial sums (contract_acc)
    if (lane <= MAX_PK) gsum_w[lane] = 0.0;
    int cur_r = -1, cur_c = -1;           // tile rows whose coordinate slices are staged (slice mode)

    PHB(40);
    for (int i = 0; i < a.nbulk; ++i) {
        PHB(41);
        TRB(i, 0);
        TRW(i, 0);
        const int* rec = ring + (i & 3) * 28;         // published by the previous barrier
        const int type = rec[0], I = rec[1], J = rec[2], k = rec[3];
        const int sA = rec[4], sB = rec[5], sC = rec[6];
        const int wo = rec[7], wait_mask = rec[8], flags = rec[9], st_slot = rec[10], st_tile = rec[11];
        if (tid == 0 && (flags & DOPF_WAIT_READ)) bulk_s2g_wait_read();
        if (ring_lane) cp_async_wait<1>();            // record i + 1 has landed (i + 2 may still be in flight)
        named_bar_sync(1, DAG_BULK_THREADS);          // every bulk warp is done with the previous operation
        TRW(i, 1);
        ring_fetch(i + 3);                            // into the slot of record i - 1
        if (wo > 0) {
            // chain op c publishes two events: 2c + 1 (tiles final) and 2c + 2 (vectors); a panel solve that only
            // waits for its own diagonal tile starts on the first and takes z before its epilogue
#ifdef FAB_TRSM_EARLY
            const int need = (type == DOP_TRSM && wo == J + 1) ? 2 * wo - 1 : 2 * wo;
#else
            const int need = 2 * wo;
#endif
            if (lane == 0) spin_until(&sync[1], need);
            __syncwarp();
        }
        TRW(i, 2);
        if (tid == 0) {
            const int nload = rec[12];
            if (flags & DOPF_DRAIN) bulk_s2g_wait_all();
            if (nload > 0) fence_proxy_async();
            for (int l = 0; l < nload; ++l) {
                const int ls = rec[13 + l], lt = rec[19 + l];
                bulk_g2s(slots + (size_t)ls * TILE_ELEMS, ws_tile(lt), TILE_BYTES, &bars[ls]);
            }
            if (st_slot >= 0)
                bulk_s2g(ws_tile(st_tile), slots + (size_t)st_slot * TILE_ELEMS, TILE_BYTES);
        }
        TRW(i, 3);
        PHE(41);
        const double* A = slots + (size_t)(sA < 0 ? 0 : sA) * TILE_ELEMS;
        const double* Bt = slots + (size_t)(sB < 0 ? 0 : sB) * TILE_ELEMS;
        double* C = slots + (size_t)(sC < 0 ? 0 : sC) * TILE_ELEMS;

        if (type == DOP_BUILD || (flags & DOPF_BUILD_FIRST) || (type == DOP_P2_MMA && (flags & DOPF_LAST))) {
            if (!a.resident) {            // slice mode: stage the coordinates (and alpha) of the two tile rows
                bool staged = false;
                if (cur_r != I) { stage_slice(sl_r, gd, hfull, I, DAG_BULK_THREADS); cur_r = I; staged = true; }
                if (cur_c != J) { stage_slice(sl_c, gd, hfull, J, DAG_BULK_THREADS); cur_c = J; staged = true; }
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

        // a covariance tile fused with its first update is generated while the operand tiles are still in flight
        // (single call site for both kinds of op; a plain BUILD op has no tile to wait for)
        TRB(i, 1);
        if (type == DOP_BUILD || (flags & DOPF_BUILD_FIRST)) {
            PHB(43);
            dag_build<DD>(acc, zr_p, zc_p, ur_p, uc_p, ldz, h, N, gd.family, I, J, wt, etab);
            PHE(43);
        }
        PHB(42);
        for (int s = 0; s < MAX_SLOTS; ++s)
            if ((wait_mask >> s) & 1) {
                mbar_wait(&bars[s], (phases >> s) & 1u);
                phases ^= 1u << s;
            }
        PHE(42);
        if (type != DOP_BUILD) PHB(42 + type);
        TRW(i, 4);
        if (flags & DOPF_GET_BEFORE) dag_acc_load(acc, C, wt);  // partial tile P(I,J) continues in the accumulators
#ifdef FAB_EXP_TWICE          // experiment: every op body twice (same instructions: second run = warm instruction caches)
#pragma unroll 1
        for (int rep = 0; rep < 2; ++rep) {
#endif
        switch (type) {
        case DOP_BUILD: break;            // (generated above)
        case DOP_MMA_NT:
            dag_mma<false, true, true>(acc, A, Bt, wt, KR_ZERO, KR_FULL, I == J);
            break;
        case DOP_PUT:
            dag_acc_store(acc, C, wt);
            fence_proxy_async();
            break;
        case DOP_GET:
            dag_acc_load(acc, C, wt);
            break;
        case DOP_TRSM: {
            // L(I,J) = U(I,J) W(J,J)^T in place (W lower triangular: k <= n), then accw_I += L(I,J) z_J
            DagAcc r2;
            dag_acc_zero(r2);
            dag_mma<false, true, false>(r2, C, Bt, wt, KR_ZERO, KR_N0_END, false);
            if (lane == 0) spin_until(&sync[1], 2 * J + 2);
            __syncwarp();
            const double* zJ = zv + TS * J;
#pragma unroll
            for (int u = 0; u < DAG_NU; ++u) {
                const int m0 = wt.u[u].m0, n0 = wt.u[u].n0;
                double zc0[2], zc1[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) { zc0[j] = zJ[n0 + 8 * j + 2 * t]; zc1[j] = zJ[n0 + 8 * j + 2 * t + 1]; }
#pragma unroll
                for (int i4 = 0; i4 < DAG_MI; ++i4) {
                    double s = 0.0;
#pragma unroll
                    for (int j = 0; j < 2; ++j) s = fma(r2.u[u][i4][j][0], zc0[j], fma(r2.u[u][i4][j][1], zc1[j], s));
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    if (t == 0) part[(n0 >> 4) * TS + m0 + 8 * i4 + g] = s;
                }
            }
            named_bar_sync(1, DAG_BULK_THREADS);       // every warp has read its rows of U
            dag_acc_store(r2, C, wt);
            fence_proxy_async();
            if (tid < TS) accw[TS * I + tid] += (part[tid] + part[TS + tid]) + (part[2 * TS + tid] + part[3 * TS + tid]);
            break;
        }
        case DOP_P1_MMA:
            if (flags & DOPF_FIRST) dag_acc_zero(acc);
            // W(J,J) is lower triangular: rows k' < n contribute nothing
            dag_mma<false, false, false>(acc, A, Bt, wt, (k == J) ? KR_N0 : KR_ZERO, KR_FULL, false);
            break;
        case DOP_P1_FIN: {
            // W(I,J) = -W(I,I) T(I,J) in place  (W(I,I) lower triangular: k <= m), then alpha_J += W(I,J)^T z_I
            DagAcc r2;
            dag_acc_zero(r2);
            dag_mma<false, false, true>(r2, A, C, wt, KR_ZERO, KR_M0_END, false);
            const double* zI = zv + TS * I;
#pragma unroll
            for (int u = 0; u < DAG_NU; ++u) {
                const int m0 = wt.u[u].m0, n0 = wt.u[u].n0;
                double zr4[DAG_MI];
#pragma unroll
                for (int i4 = 0; i4 < DAG_MI; ++i4) zr4[i4] = zI[m0 + 8 * i4 + g];
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        double s = 0.0;
#pragma unroll
                        for (int i4 = 0; i4 < DAG_MI; ++i4) s = fma(r2.u[u][i4][j][e], zr4[i4], s);
                        s += __shfl_xor_sync(0xffffffffu, s, 4);
                        s += __shfl_xor_sync(0xffffffffu, s, 8);
                        s += __shfl_xor_sync(0xffffffffu, s, 16);
                        if (g == 0) part[(m0 / (8 * DAG_MI)) * TS + n0 + 8 * j + 2 * t + e] = s;
                    }
            }
            named_bar_sync(1, DAG_BULK_THREADS);       // every warp has read its columns of the product
            dag_acc_store(r2, C, wt);
            fence_proxy_async();
            if (tid < TS) {
                double s = part[tid] + part[TS + tid];
                if (DAG_RB == 4) s += part[2 * TS + tid] + part[3 * TS + tid];
                alv[TS * J + tid] += s;
            }
            break;
        }
        case DOP_P2_MMA:
            if (flags & DOPF_FIRST) dag_acc_zero(acc);
            // W(I,I) is lower triangular: (W_II)[k'][m] = 0 for k' < m
            dag_mma<true, false, false>(acc, A, Bt, wt, (k == I) ? KR_M0 : KR_ZERO, KR_FULL, I == J);
            if (flags & DOPF_LAST) {
                PHB(50);
                const double* ar_p = a.resident ? alv + TS * I : al_r;
                const double* ac_p = a.resident ? alv + TS * J : al_c;
                contract_acc<DD>(acc, gsum_w, zr_p, zc_p, ur_p, uc_p, ar_p, ac_p, ldz, h, N, gd.family, I, J, wt, etab);
                PHE(50);
            }
            break;
        default: break;
        }
#ifdef FAB_EXP_TWICE
        if (rep == 0) TRB(i, 3);
        }
#endif
        if (flags & DOPF_PUT_AFTER) {         // the tile leaves the accumulators: slot C is no operand of this op ...
            if (flags & DOPF_PUT_BARRIER) named_bar_sync(1, DAG_BULK_THREADS);     // ... or everybody is done reading it
            dag_acc_store(acc, C, wt);
            fence_proxy_async();
        }
        if (type != DOP_BUILD) PHE(42 + type);
        TRB(i, 2);
        TRW(i, 5);
        __syncwarp();
        if (lane == 0) sync_arrive(&sync[0]);
    }
    PHB(51);

    if (tid == 0) side();
    // the chain has published the log-likelihood, the failure flag and its share of alpha
    if (lane == 0) spin_until(&sync[1], 2 * a.nchain + 1);
    __syncwarp();
    named_bar_sync(1, DAG_BULK_THREADS);
    if (a.mode == DAG_FIT && a.resident)
        for (int i = tid; i < Np; i += DAG_BULK_THREADS) a.wk.alpha[(size_t)b * Np + i] = alv[i];
    if (a.mode == DAG_GRAD && a.grad) {
        if (tid == 0) {
            double gacc[MAX_PK + 1];
            for (int q = 0; q <= MAX_PK; ++q) {
                double sgm = 0.0;
                for (int ww = 0; ww < DAG_NBW; ++ww) sgm += red[ww * (MAX_PK + 1) + q];
                gacc[q] = sgm;
            }
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
    if (a.peer.world > 0 && w == 0) {
        // this walker's ll / info (chain, published through sync[1]) and gradient row (lane 0, above) are final:
        // the first row_doubles lanes of warp 0 store the row into every rank's gather buffer
        __syncwarp();
        const int rd = a.peer.row_doubles;
        if (lane < rd) {
            double v;
            if (lane == rd - 2) v = a.ll[b];
            else if (lane == rd - 1) v = (double)a.info[b];
            else v = (a.mode == DAG_GRAD && a.grad) ? a.grad[(size_t)b * (Pk + 1) + lane] : 0.0;
            const size_t off = ((size_t)a.peer.rank * a.peer.rows_per_rank + b) * rd + lane;
            for (int r = 0; r < a.peer.world; ++r) a.peer.buf[r][off] = v;
        }
        __syncwarp();
        if (lane == 0) {
#ifdef FAB_CUSIM
            const unsigned prev = *a.peer.done; *a.peer.done = prev + 1u;      // (the emulator has no peers: world == 1)
            if (prev + 1u == a.peer.done_target) a.peer.slots[a.peer.rank][a.peer.rank] = a.peer.step;
#else
            __threadfence();                          // the row stores of this CTA before its count (release, device scope)
            const unsigned prev = atomicAdd(a.peer.done, 1u);
            if (prev + 1u == a.peer.done_target) {    // last CTA of this rank: every row of the block has been stored
                __threadfence_system();               // acquire the other CTAs' counts, release everything to the box
                for (int r = 0; r < a.peer.world; ++r)
                    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(a.peer.slots[r] + a.peer.rank), "r"(a.peer.step) : "memory");
                const long long t0 = clock64();
                for (int r = 0; r < a.peer.world; ++r) {
                    const unsigned* sl = a.peer.slots[a.peer.rank] + r;
                    for (;;) {
                        unsigned v;
                        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(sl) : "memory");
                        if ((int)(v - a.peer.step) >= 0) break;
                        if (clock64() - t0 > 4000000000LL) { *a.peer.err = 1; break; }      // ~2 s: a peer died
                        __nanosleep(40);
                    }
                }
            }
#endif
        }
    }
    if (tid == 0) bulk_s2g_wait_all();
    PHE(51);
    PHE(40);
}

// 12 warps: 65536 / 384 = 170 -> 168 registers per thread; 20 warps: 65536 / 640 = 102 -> 96.
#define FAB_DAG_KERNEL_ATTR __global__ void __launch_bounds__(DAG_THREADS, 1)
// One instantiation per number of Matern axes d = gd.d (1..MAX_D): the covariance generation and the contraction are
// unrolled over d, and a kernel that carried all eight versions was 406 KB of code (the sampler 921 KB) -- far more
// than the instruction caches hold, for straight-line tile code that runs once per op (profiles/r02g_*).
template <int DD>
FAB_DAG_KERNEL_ATTR gp_dag_kernel(DagArgs a) {
    FAB_DYN_SMEM(smem_raw);
    unsigned phases = 0u;                 // mbarrier phase parity per slot (uniform over the bulk warps)
    dag_eval<DD>(a, blockIdx.x, smem_raw, phases, true);
}

// host side: run `...` with `constexpr int DD = d`  (development builds: -DFAB_DEV_D=5 instantiates one d only)
#ifdef FAB_DEV_D
#define FAB_DISPATCH_D(d, ...) { constexpr int DD = FAB_DEV_D; __VA_ARGS__; }
#else
#define FAB_DISPATCH_D(d, ...)                                   \
    switch (d) {                                                 \
    case 1: { constexpr int DD = 1; __VA_ARGS__; } break;        \
    case 2: { constexpr int DD = 2; __VA_ARGS__; } break;        \
    case 3: { constexpr int DD = 3; __VA_ARGS__; } break;        \
    case 4: { constexpr int DD = 4; __VA_ARGS__; } break;        \
    case 5: { constexpr int DD = 5; __VA_ARGS__; } break;        \
    case 6: { constexpr int DD = 6; __VA_ARGS__; } break;        \
    case 7: { constexpr int DD = 7; __VA_ARGS__; } break;        \
    default: { constexpr int DD = 8; __VA_ARGS__; } break;       \
    }
#endif

}  // namespace fab
