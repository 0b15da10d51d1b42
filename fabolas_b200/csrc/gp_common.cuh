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
    double* hyp;       // B x 4 + MAX_D       decoded hyper-parameters (amp, inv_g2, cb, yerr2, scale_k..)
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

// Matern-5/2 pieces from the squared scaled distance.
__device__ __forceinline__ void matern52(double r2, double& m, double& gfac) {
    const double r = sqrt(5.0 * r2);
    const double e = exp(-r);
    m = (1.0 + r + (5.0 / 3.0) * r2) * e;
    gfac = (5.0 / 6.0) * (1.0 + r) * e;      // -dm/dr2
}

// Scaled coordinates in shared memory, structure-of-arrays: zs[k * Np + i] = x_ik / sqrt(M_k),
// k < d ; us[i] = basis column (x_id for family 1, unused otherwise).  Rows i >= N are padding.
struct Coords {
    const double* zs;
    const double* us;
    int Np, N, d, family;
};

__device__ __forceinline__ double kern_value(const Coords& c, const Hyper& h, int i, int j) {
    double r2 = 0.0;
    for (int k = 0; k < c.d; ++k) {
        const double dz = c.zs[k * c.Np + i] - c.zs[k * c.Np + j];
        r2 = fma(dz, dz, r2);
    }
    double m, gf;
    matern52(r2, m, gf);
    double v = h.amp * m;
    if (c.family == 1) v *= fma(c.us[i] * c.us[j], h.inv_g2, h.cb);
    return v;
}

// Fill tile (I, J) of the covariance matrix into `tile` (swizzled).  Padding rows/cols get the
// identity.  For diagonal tiles the strictly-upper 8x8 blocks are zeroed (tile_potrf contract).
// Warp w writes rows 8w..8w+7, lanes across columns (conflict-free).  No barrier inside.
__device__ __forceinline__ void build_tile(double* tile, const Coords& c, const Hyper& h, int I, int J) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 1
    for (int rr = 0; rr < 8; ++rr) {
        const int r = 8 * w + rr;
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
                v = kern_value(c, h, gi, gj);
                if (gi == gj) v += h.diag_add;
            }
            tile[tix(r, cc)] = v;
        }
    }
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

struct TileCache {
    int id[MAX_SLOTS];
    int stamp[MAX_SLOTS];
    int nslot, clock;
    unsigned phase;                 // mbarrier phase counter (uniform)
    double* base;                   // nslot * TILE_ELEMS doubles of shared memory
    unsigned long long* bar;

    __device__ __forceinline__ void init(double* b, int n, unsigned long long* mb) {
        base = b; nslot = n; bar = mb; clock = 0; phase = 0;
#pragma unroll
        for (int s = 0; s < MAX_SLOTS; ++s) { id[s] = -1; stamp[s] = 0; }
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
    // recently used; slots p0/p1/p2 are pinned.
    template <typename Dead>
    __device__ __forceinline__ int victim(Dead dead, int p0, int p1, int p2) {
        int best = -1, best_score = 0x7fffffff;
#pragma unroll
        for (int s = 0; s < MAX_SLOTS; ++s) {
            if (s >= nslot || s == p0 || s == p1 || s == p2) continue;
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
    __device__ __forceinline__ int alloc(int tid_, Dead dead, int p0, int p1, int p2) {
        const int s = victim(dead, p0, p1, p2);
        if (threadIdx.x == 0) bulk_s2g_wait_read();
        __syncthreads();
        set(s, tid_);
        return s;
    }
    // Make a tile resident, loading it from the HBM/L2 workspace on a miss (TMA bulk copy).
    template <typename Dead>
    __device__ __forceinline__ int acquire(const double* Lw_walker, int tid_, Dead dead, int p0, int p1, int p2) {
        int s = find(tid_);
        if (s >= 0) return s;
        s = victim(dead, p0, p1, p2);
        __syncthreads();                       // nobody may still be reading the victim
        if (threadIdx.x == 0) {
            bulk_s2g_wait_all();               // write-through stores must have landed
            bulk_g2s(ptr(s), Lw_walker + (size_t)tri_index(id_row(tid_), id_col(tid_)) * TILE_ELEMS, TILE_BYTES, bar);
        }
        mbar_wait(bar, phase & 1u);
        ++phase;
        set(s, tid_);
        return s;
    }
};

}  // namespace fab
