// gp_common.cuh -- problem description shared by the GP kernels, the covariance function
// (george Matern-5/2 x dataset-size basis) and the shared-memory tile cache.
#pragma once
#include "tiles.cuh"

namespace fab {

constexpr int MAX_D = 8;        // Matern axes
constexpr int MAX_PK = MAX_D + 3;
constexpr int MAX_SLOTS = 6;    // 6 x 32 KB tile slots + vectors fit the 227 KB of one SM

// family 0: amp * M52(r2)                        theta = [log_c0, log_M_0..log_M_{d-1}]
// family 1: amp * M52(r2) * (u u'/gamma2 + cb)   theta = [..., log_gamma2, log_cb], u = column d
// amp = amp_mult * exp(log_c0): george sums a ConstantKernel over its axes (fabolas.py:57 writes the
// raw parameter into that slot; scalar*kernel builds log(c/ndim) -- oracle/george_oracle.py).
struct GpData {
    const double* X;   // N x D row-major (device)
    const double* y;   // N
    double mean;
    double amp_mult;
    int N, D, d, family, Pk;
    int NT, Np;        // tiles per edge, padded size 64*NT
};

struct GpWork {        // per-walker workspaces in HBM (L2-resident at the benchmark sizes)
    double* Lw;        // B x ntiles x 4096   swizzled lower tiles of L (later W = L^-1 in place)
    double* Dinv;      // B x NT x 512        inverses of the 8x8 diagonal blocks of L
    double* zvec;      // B x Np              z = L^-1 (y - mean)
    double* alpha;     // B x Np              alpha = K^-1 (y - mean)
    double* accw;      // B x Np              running sum_J L_IJ z_J of the forward substitution
};

struct Hyper {         // decoded hyper-parameters of one walker
    double amp, inv_g2, cb, diag_add;
    double scale[MAX_D];      // 1/sqrt(M_k)
};

__device__ __forceinline__ Hyper decode_hyper(const GpData& gd, const double* __restrict__ th, double yerr2) {
    Hyper h;
    h.amp = gd.amp_mult * exp(th[0]);
#pragma unroll
    for (int k = 0; k < MAX_D; ++k) h.scale[k] = (k < gd.d) ? exp(-0.5 * th[1 + k]) : 0.0;
    if (gd.family == 1) {
        h.inv_g2 = exp(-th[1 + gd.d]);
        h.cb = exp(th[2 + gd.d]);
    } else {
        h.inv_g2 = 0.0;
        h.cb = 1.0;
    }
    h.diag_add = yerr2 + kTiny;
    return h;
}

// exp(x) for x <= 0 (the only range the Matern kernels need), ~1 ulp: round-to-nearest range
// reduction with a two-part ln2, degree-13 Taylor polynomial on |r| <= ln2/2 (remainder 5e-18),
// exponent patched in with an integer add.  21 FP64-pipe instructions instead of the ~46 of the
// general-purpose library exp() that profiles/r01a showed dominating the build and gradient kernels.
// Results below 2^-1000 are flushed to zero (they are < 1e-301 of the prior variance).
__device__ __forceinline__ double exp_nonpos(double x) {
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52
    const double t = fma(x, 1.4426950408889634, MAGIC);
    const double kd = t - MAGIC;
    const long long k = (long long)(int)(__double_as_longlong(t) & 0xffffffffLL);
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;                       // 1/13!
    p = fma(p, r, 2.08767569878681e-09);                     // 1/12!
    p = fma(p, r, 2.505210838544172e-08);                    // 1/11!
    p = fma(p, r, 2.755731922398589e-07);                    // 1/10!
    p = fma(p, r, 2.7557319223985893e-06);                   // 1/9!
    p = fma(p, r, 2.48015873015873e-05);                     // 1/8!
    p = fma(p, r, 1.984126984126984e-04);                    // 1/7!
    p = fma(p, r, 1.388888888888889e-03);                    // 1/6!
    p = fma(p, r, 8.333333333333333e-03);                    // 1/5!
    p = fma(p, r, 4.1666666666666664e-02);                   // 1/4!
    p = fma(p, r, 1.6666666666666666e-01);                   // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double v = __longlong_as_double(__double_as_longlong(p) + (k << 52));
    return (x > -693.0) ? v : 0.0;                            // NaN input -> 0 never reached: callers pass finite r
}

// Matern-5/2 pieces from the squared scaled distance.
__device__ __forceinline__ void matern52(double r2, double& m, double& gfac) {
    const double r = sqrt(5.0 * r2);
    const double e = exp_nonpos(-r);
    m = (1.0 + r + (5.0 / 3.0) * r2) * e;
    gfac = (5.0 / 6.0) * (1.0 + r) * e;      // -dm/dr2
}
__device__ __forceinline__ double matern52_value(double r2) {
    const double r = sqrt(5.0 * r2);
    return (1.0 + r + (5.0 / 3.0) * r2) * exp_nonpos(-r);
}

// ---------------------------------------------------------------------------------------------
// Fast Matern-5/2 evaluation (the inner loop of the covariance build and of the dK contraction):
// 33 FP64-pipe instructions per entry instead of ~49, short dependent chains.
//   r = sqrt(5 r2): hardware rsqrt seed (~2^-20) + two coupled Newton steps on the root itself
//       (residual x - r^2 by FMA, correction e * y/2) => ~2^-60 relative; the +1e-300 (folded into
//       the FMA that forms 5 r2) makes r2 == 0 a regular input (r = 1e-150, m = 1 exactly).
//   exp(-r) = 2^-q * tab[j] * p(rr),  n = round(r * 128 / ln2) = 128 q + j,  |rr| <= ln2/256,
//       tab[j] = 2^(-j/128) from a 1 KB shared-memory table (exp_table_fill), degree-5 Taylor
//       polynomial (remainder 6e-19), exponent patched with an integer subtract.  ~1.5 ulp.
// r >= 690 (exp < 1e-299 of the prior variance) is flushed to zero.
// ---------------------------------------------------------------------------------------------
constexpr int EXP_TAB_N = 128;
__device__ __forceinline__ void exp_table_fill(double* tab) {   // all threads; caller syncs
    for (int j = threadIdx.x; j < EXP_TAB_N; j += blockDim.x) tab[j] = exp2(-(double)j / EXP_TAB_N);
}
__device__ __forceinline__ double sqrt5_fast(double r2) {       // sqrt(5 r2 + 1e-300), r2 >= 0 finite
    const double x = fma(5.0, r2, 1e-300);
#ifdef FAB_CUSIM
    return std::sqrt(x);
#else
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hy = 0.5 * y;
    double r = x * y;
    r = fma(fma(-r, r, x), hy, r);
    r = fma(fma(-r, r, x), hy, r);
    return r;
#endif
}
__device__ __forceinline__ double exp_neg_fast(double r, const double* __restrict__ tab) {   // exp(-r), r >= 0
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52
    const double t = fma(r, 184.66496523378731, MAGIC);      // 128 / ln2
    const double nd = t - MAGIC;
    const int n = (int)(__double_as_longlong(t) & 0xffffffffLL);
    double rr = fma(nd, 5.41521234663377981633e-03, -r);     // ln2_hi / 128 (32 significant bits: nd * hi exact)
    rr = fma(nd, 1.49079291349264664064e-12, rr);            // ln2_lo / 128
    double p = 8.333333333333333e-03;                        // 1/5!
    p = fma(p, rr, 4.1666666666666664e-02);
    p = fma(p, rr, 1.6666666666666666e-01);
    p = fma(p, rr, 0.5);
    p = fma(p, rr, 1.0);
    p = fma(p, rr, 1.0);
    const double v = tab[n & (EXP_TAB_N - 1)] * p;
    const double w = __longlong_as_double(__double_as_longlong(v) - ((long long)(n >> 7) << 52));
    return (r < 690.0) ? w : 0.0;
}
__device__ __forceinline__ double matern52_fast_value(double r2, const double* __restrict__ tab) {
    const double r = sqrt5_fast(r2);
    return fma(r2, 5.0 / 3.0, 1.0 + r) * exp_neg_fast(r, tab);
}
__device__ __forceinline__ void matern52_fast(double r2, double& m, double& gfac, const double* __restrict__ tab) {
    const double r = sqrt5_fast(r2);
    const double e = exp_neg_fast(r, tab);
    const double opr = 1.0 + r;
    m = fma(r2, 5.0 / 3.0, opr) * e;
    gfac = (5.0 / 6.0) * opr * e;            // -dm/dr2
}

// Scaled coordinates of the 64 rows and the 64 columns of the covariance tile being generated,
// staged in shared memory as structure-of-arrays slices: z[k * 64 + r] = x_rk / sqrt(M_k) (k < d)
// followed by u[r] (basis column, family 1; zero otherwise).  Slices are O(64 (d+1)) regardless of N.
__host__ __device__ constexpr int slice_doubles(int d) { return (d + 1) * TS; }

struct Coords {
    const double* zr;   // row slice
    const double* zc;   // column slice
    const double* tab;  // exp table (exp_table_fill) in shared memory
    int N, d, family;
    __device__ __forceinline__ const double* ur() const { return zr + d * TS; }
    __device__ __forceinline__ const double* uc() const { return zc + d * TS; }
};

// Stage the slice of tile-row `I` (observations 64 I .. 64 I + 63) into `buf`; all threads, no barrier.
__device__ __forceinline__ void stage_slice(double* buf, const GpData& gd, const Hyper& h, int I) {
    const int D = gd.D, d = gd.d;
    for (int e = threadIdx.x; e < TS * D; e += NTHREADS) {
        const int r = e / D, k = e - r * D;
        const int gi = TS * I + r;
        const double x = (gi < gd.N) ? __ldg(&gd.X[(size_t)gi * D + k]) : 0.0;
        if (k < d) buf[k * TS + r] = x * h.scale[k];
        else buf[d * TS + r] = x;                       // k == d only happens for family 1 (D = d + 1)
    }
    if (gd.family != 1)
        for (int r = threadIdx.x; r < TS; r += NTHREADS) buf[d * TS + r] = 0.0;
}

__device__ __forceinline__ double kern_value(const Coords& c, const Hyper& h, int r, int cc) {
    double r2 = 0.0;
    for (int k = 0; k < c.d; ++k) {
        const double dz = c.zr[k * TS + r] - c.zc[k * TS + cc];
        r2 = fma(dz, dz, r2);
    }
    double v = h.amp * matern52_fast_value(r2, c.tab);
    if (c.family == 1) v *= fma(c.ur()[r] * c.uc()[cc], h.inv_g2, h.cb);
    return v;
}

// Rows [row0, row0 + nrows) of tile (I, J) by the CALLING WARP, lanes across the 64 columns (two per
// lane, conflict-free stores).  Generic version: any padding, diagonal tiles included.
__device__ __forceinline__ void build_rows_generic(double* tile, const Coords& c, const Hyper& h, int I, int J,
                                                   int row0, int nrows) {
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int rr = 0; rr < nrows; ++rr) {
        const int r = row0 + rr;
        const int gi = TS * I + r;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int cc = lane + 32 * half;
            const int gj = TS * J + cc;
            double v;
            if (I == J && (cc >> 3) > (r >> 3)) {
                v = 0.0;
            } else if (gi >= c.N || gj >= c.N) {
                v = (gi == gj) ? 1.0 : 0.0;
            } else {
                v = kern_value(c, h, r, cc);
                if (gi == gj) v += h.diag_add;
            }
            tile[tix(r, cc)] = v;
        }
    }
}

// Interior off-diagonal tile, DD Matern axes known at compile time: each lane keeps the scaled
// coordinates of its two columns in registers, the row coordinates arrive by broadcast loads.
template <int DD>
__device__ __forceinline__ void build_rows_offdiag(double* tile, const Coords& c, const Hyper& h, int row0, int nrows) {
    const int lane = threadIdx.x & 31;
    const bool basis = c.family == 1;
    double zc0[DD], zc1[DD];
#pragma unroll
    for (int k = 0; k < DD; ++k) {
        zc0[k] = c.zc[k * TS + lane];
        zc1[k] = c.zc[k * TS + lane + 32];
    }
    const double uc0 = c.uc()[lane] * h.inv_g2, uc1 = c.uc()[lane + 32] * h.inv_g2;
#pragma unroll 2
    for (int rr = 0; rr < nrows; ++rr) {
        const int r = row0 + rr;
        double r20 = 0.0, r21 = 0.0;
#pragma unroll
        for (int k = 0; k < DD; ++k) {
            const double zr = c.zr[k * TS + r];
            const double d0 = zr - zc0[k], d1 = zr - zc1[k];
            r20 = fma(d0, d0, r20);
            r21 = fma(d1, d1, r21);
        }
        double v0 = h.amp * matern52_fast_value(r20, c.tab), v1 = h.amp * matern52_fast_value(r21, c.tab);
        if (basis) {
            const double ur = c.ur()[r];
            v0 *= fma(ur, uc0, h.cb);
            v1 *= fma(ur, uc1, h.cb);
        }
        tile[tix(r, lane)] = v0;
        tile[tix(r, lane + 32)] = v1;
    }
}

// Rows of a tile by the calling warp (dispatch on padding and on the number of axes).
__device__ __forceinline__ void build_rows(double* tile, const Coords& c, const Hyper& h, int I, int J, int row0,
                                           int nrows) {
    const bool interior = TS * I + TS <= c.N && TS * J + TS <= c.N;
    if (!interior || I == J) { build_rows_generic(tile, c, h, I, J, row0, nrows); return; }
#define FAB_ROWS_CASE(DD) case DD: build_rows_offdiag<DD>(tile, c, h, row0, nrows); break;
    switch (c.d) {
        FAB_ROWS_CASE(1) FAB_ROWS_CASE(2) FAB_ROWS_CASE(3) FAB_ROWS_CASE(4)
        FAB_ROWS_CASE(5) FAB_ROWS_CASE(6) FAB_ROWS_CASE(7) FAB_ROWS_CASE(8)
        default: build_rows_generic(tile, c, h, I, J, row0, nrows);
    }
#undef FAB_ROWS_CASE
}

// Interior diagonal tile: only the 36 lower 8x8 blocks are evaluated (one block per warp pass,
// two adjacent columns per lane, 16-byte stores); the 28 strictly-upper blocks are zero-filled
// (tile_potrf contract).  Row and column slices coincide.
template <int DD>
__device__ __forceinline__ void build_tile_diag(double* tile, const Coords& c, const Hyper& h) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = lane >> 2, cp = (lane & 3) * 2;
    const bool basis = c.family == 1;
    // upper blocks: zero fill (28 blocks), lower blocks: 36 evaluated, both spread evenly over the warps
    for (int q = w; q < 28; q += NWARPS) {
        int bi = 0, acc_ = 0;
        while (acc_ + (7 - bi) <= q) { acc_ += 7 - bi; ++bi; }
        const int bj = bi + 1 + (q - acc_);
        double2 z; z.x = 0.0; z.y = 0.0;
        *reinterpret_cast<double2*>(&tile[tix(8 * bi + r, 8 * bj + cp)]) = z;
    }
    for (int q = w; q < 36; q += NWARPS) {
        int bi = 0, acc_ = 0;
        while (acc_ + bi + 1 <= q) { acc_ += bi + 1; ++bi; }
        const int bj = q - acc_;
        double2 v; v.x = 0.0; v.y = 0.0;
        {
            const int li = 8 * bi + r, lj = 8 * bj + cp;
            double r20 = 0.0, r21 = 0.0;
#pragma unroll
            for (int k = 0; k < DD; ++k) {
                const double zr = c.zr[k * TS + li];
                const double2 zc = *reinterpret_cast<const double2*>(&c.zc[k * TS + lj]);
                const double d0 = zr - zc.x, d1 = zr - zc.y;
                r20 = fma(d0, d0, r20);
                r21 = fma(d1, d1, r21);
            }
            v.x = h.amp * matern52_fast_value(r20, c.tab);
            v.y = h.amp * matern52_fast_value(r21, c.tab);
            if (basis) {
                const double ur = c.ur()[li] * h.inv_g2;
                const double2 uc = *reinterpret_cast<const double2*>(&c.uc()[lj]);
                v.x *= fma(ur, uc.x, h.cb);
                v.y *= fma(ur, uc.y, h.cb);
            }
            if (li == lj) v.x += h.diag_add;
            if (li == lj + 1) v.y += h.diag_add;
        }
        *reinterpret_cast<double2*>(&tile[tix(8 * bi + r, 8 * bj + cp)]) = v;
    }
}

// Fill tile (I, J) of the covariance matrix into `tile` (swizzled), all 8 warps.  Padding rows/cols
// get the identity.  For diagonal tiles the strictly-upper 8x8 blocks are zeroed.  No barrier inside.
__device__ __forceinline__ void build_tile(double* tile, const Coords& c, const Hyper& h, int I, int J) {
    const int w = threadIdx.x >> 5;
    const bool interior = TS * I + TS <= c.N && TS * J + TS <= c.N;
    if (I != J || !interior) { build_rows(tile, c, h, I, J, 8 * w, 8); return; }
#define FAB_BUILD_CASE(DD) case DD: build_tile_diag<DD>(tile, c, h); break;
    switch (c.d) {
        FAB_BUILD_CASE(1) FAB_BUILD_CASE(2) FAB_BUILD_CASE(3) FAB_BUILD_CASE(4)
        FAB_BUILD_CASE(5) FAB_BUILD_CASE(6) FAB_BUILD_CASE(7) FAB_BUILD_CASE(8)
        default: build_rows_generic(tile, c, h, I, J, 8 * w, 8);
    }
#undef FAB_BUILD_CASE
}

// ---------------------------------------------------------------------------------------------
// Shared-memory tile cache.  Every thread keeps an identical copy of the (tiny) directory in
// registers and runs the same deterministic policy, so no communication is needed to agree on
// hits, victims and slots.  Tile ids: (kind << 30) | (I << 16) | J, kind 0 = L tile, 1 = W = L^-1.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int tile_id(int kind, int I, int J) { return (kind << 30) | (I << 16) | J; }
__device__ __forceinline__ int id_row(int id) { return (id >> 16) & 0x3fff; }
__device__ __forceinline__ int id_col(int id) { return id & 0xffff; }
__device__ __forceinline__ int id_kind(int id) { return (id >> 30) & 1; }

__device__ __forceinline__ unsigned pin(int s) { return s >= 0 ? (1u << s) : 0u; }

struct TileCache {
    int id[MAX_SLOTS];
    int stamp[MAX_SLOTS];
    int nslot, clock;
    unsigned inflight;              // slots with a bulk load in flight (bit mask, uniform)
    unsigned phases;                // per-slot mbarrier phase parity (bit mask, uniform)
    double* base;                   // nslot * TILE_ELEMS doubles of shared memory
    unsigned long long* bars;       // MAX_SLOTS mbarriers, one per slot

    // `mb` must point to MAX_SLOTS mbarriers; thread 0 initialises them (caller syncs afterwards).
    __device__ __forceinline__ void init(double* b, int n, unsigned long long* mb) {
        base = b; nslot = n; bars = mb; clock = 0; inflight = 0u; phases = 0u;
#pragma unroll
        for (int s = 0; s < MAX_SLOTS; ++s) { id[s] = -1; stamp[s] = 0; }
        if (threadIdx.x == 0)
            for (int s = 0; s < MAX_SLOTS; ++s) mbar_init(&bars[s], 1);
    }
    __device__ __forceinline__ double* ptr(int s) const { return base + (size_t)s * TILE_ELEMS; }
    __device__ __forceinline__ void touch(int s) {
        ++clock;
#pragma unroll
        for (int q = 0; q < MAX_SLOTS; ++q) if (q == s) stamp[q] = clock;
    }
    __device__ __forceinline__ int find(int tid_) {
        int f = -1;
#pragma unroll
        for (int s = 0; s < MAX_SLOTS; ++s) if (s < nslot && id[s] == tid_) f = s;
        if (f >= 0) touch(f);
        return f;
    }
    // Pick a slot to (re)use: free first, then dead (per the caller's predicate), then least
    // recently used; slots whose bit is set in `pins` are never chosen.
    template <typename Dead>
    __device__ __forceinline__ int victim(Dead dead, unsigned pins) {
        int best = -1, best_score = 0x7fffffff;
#pragma unroll
        for (int s = 0; s < MAX_SLOTS; ++s) {
            if (s >= nslot || ((pins >> s) & 1u)) continue;
            int score;
            if (id[s] < 0) score = -2000000000;
            else if (dead(id[s])) score = -1000000000 + stamp[s];
            else score = stamp[s];
            if (score < best_score) { best_score = score; best = s; }
        }
        return best;
    }
    __device__ __forceinline__ void set(int s, int tid_) {
#pragma unroll
        for (int q = 0; q < MAX_SLOTS; ++q) if (q == s) id[q] = tid_;
        touch(s);
    }
    // Reserve a slot for a tile that is about to be produced in shared memory.  Contains a block
    // barrier (orders earlier readers of the victim and drains its pending bulk store).
    template <typename Dead>
    __device__ __forceinline__ int alloc(int tid_, Dead dead, unsigned pins) {
        const int s = victim(dead, pins | inflight);
        if (threadIdx.x == 0) bulk_s2g_wait_read();
        __syncthreads();
        set(s, tid_);
        return s;
    }
    // Start making a tile resident: on a miss a TMA bulk copy from the HBM/L2 workspace is issued
    // into a victim slot and completes on that slot's mbarrier.  The caller guarantees (block
    // barrier) that no thread still reads any unpinned slot.  No barrier inside.
    template <typename Dead>
    __device__ __forceinline__ int issue(const double* Lw_walker, int tid_, Dead dead, unsigned pins) {
        int s = find(tid_);
        if (s >= 0) return s;
        s = victim(dead, pins | inflight);
        if (threadIdx.x == 0) {
            bulk_s2g_wait_all();               // write-through stores must have landed
            bulk_g2s(ptr(s), Lw_walker + (size_t)tri_index(id_row(tid_), id_col(tid_)) * TILE_ELEMS, TILE_BYTES,
                     &bars[s]);
        }
        inflight |= 1u << s;
        set(s, tid_);
        return s;
    }
    // Block until the bulk load into slot s (if any) has landed.
    __device__ __forceinline__ void wait(int s) {
        if ((inflight >> s) & 1u) {
            mbar_wait(&bars[s], (phases >> s) & 1u);
            phases ^= 1u << s;
            inflight &= ~(1u << s);
        }
    }
    // Synchronous convenience: make a tile resident now (a block barrier on a miss).
    template <typename Dead>
    __device__ __forceinline__ int acquire(const double* Lw_walker, int tid_, Dead dead, unsigned pins) {
        int s = find(tid_);
        if (s < 0) {
            __syncthreads();                   // nobody may still be reading the victim
            s = issue(Lw_walker, tid_, dead, pins);
        }
        wait(s);
        return s;
    }
};

}  // namespace fab
