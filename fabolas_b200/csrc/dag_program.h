// dag_program.h -- host-side static scheduler of the fused GP kernel (gp_dag.cuh).
//
// One walker's whole evaluation -- covariance build, tiled Cholesky, forward solve, log-likelihood and
// (optionally) W = L^-1, alpha, the K^-1 tile products and the gradient contraction -- is a DAG of
// 64x64 tile tasks.  The kernel runs it on two concurrent instruction streams inside one CTA:
//   * the BULK stream (8 warps): throughput work -- Matern tile generation, DMMA tile products,
//     panel solves in GEMM form, the dK contraction;
//   * the CHAIN stream (4 warps): the latency-bound serial work -- the in-tile Cholesky of each
//     diagonal tile, its triangular inverse, the forward-substitution vectors.
// This file decides everything ahead of time, once per (NT, slots, mode): the order of the tasks on
// each stream (list scheduling on estimated costs, critical path first), the shared-memory slot of
// every tile (farthest-next-use replacement over the merged order, write-through to the L2-resident
// workspace so that clean tiles can be dropped and streamed back by TMA), the prefetch point of every
// reload, and the cross-stream waits (every RAW/WAR/WAW hazard between the streams becomes an explicit
// "other stream has completed op #n" wait, so correctness never depends on the cost estimates).
// dag_verify() re-checks the emitted program (contents of every slot at every access, and that every
// cross-stream conflict is ordered by a wait).  Plain C++ so the CPU emulator build shares it.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

namespace fab {

enum DagMode { DAG_LL = 0, DAG_FACTOR = 1, DAG_FIT = 2, DAG_GRAD = 3 };

enum DagOpType {
    DOP_NOP = 0,
    DOP_BUILD = 1,     // acc = K(I,J) entries (registers)
    DOP_MMA_NT = 2,    // acc -= A * B^T                 A = L(I,k), B = L(J,k)
    DOP_PUT = 3,       // slot C = acc                   k = kind of the tile put: unfinished U(I,J), partial P / T
    DOP_TRSM = 4,      // C = C * B^T                    B = W(J,J);  accw_I += C z_J     -> L(I,J)
    DOP_P1_MMA = 5,    // acc (+)= A * B                 A = L(I,k), B = W(k,J)
    DOP_P1_FIN = 6,    // C = -A * C                     A = W(I,I), C = T(I,J);  alpha_J += C^T z_I   -> W(I,J)
    DOP_P2_MMA = 7,    // acc (+)= A^T * B               A = W(k,I), B = W(k,J); LAST: contraction with dK
    DOP_CH_POTRF = 8,  // chain: slot A: U(K,K) -> L(K,K) in place; slot C = W(K,K); z_K, log det, alpha_K
    DOP_GET = 9        // acc = slot C                   partial tile P(I,J) picked up again
};
enum DagOpFlags {
    DOPF_FIRST = 1,       // first product of an accumulation: acc = 0
    DOPF_LAST = 2,        // last product of a K^-1 tile: contraction
    DOPF_WAIT_READ = 4,   // a slot that is the source of an earlier bulk store is overwritten from here on
    DOPF_DRAIN = 8,       // a tile loaded here was stored since the last drain: wait for all bulk stores first
    DOPF_INPLACE = 16,    // PUT over the partial tile P(I,J) that the preceding GET picked up (same slot)
    // PUT and GET do not run as ops of their own: they ride on the op that produces / continues the accumulators
    DOPF_GET_BEFORE = 32, // acc = slot C (partial tile P(I,J)) before the op's own work
    DOPF_PUT_AFTER = 64,  // slot C = acc after the op's own work; put_kind = kind of the tile (U, P or T)
    DOPF_PUT_BARRIER = 128,// slot C is an operand slot of this very op (no other slot was free): the bulk warps meet first
    // A covariance tile followed by its first update is ONE op: the tile is generated into the accumulators while the
    // operand tiles of the update are still on their way (one op start less per tile, the TMA latency hidden)
    DOPF_BUILD_FIRST = 256
};
constexpr int DAG_MAX_LOADS = 6;
// A reload is issued at most this many ops ahead of its consumer.  Measured (N = 256, ll+grad): 1 -> 267 us, 2 -> 270,
// 3 -> 272, 4 -> 275, 6 -> 280: one op (~2 us) covers the L2 / HBM latency of a tile, and a reload hoisted further
// drags the cross-stream waits of its slot to an earlier op.  N = 2048 does not care (slot availability decides).
#ifndef FAB_DAG_PREFETCH_DEPTH
#define FAB_DAG_PREFETCH_DEPTH 1
#endif
constexpr int DAG_PREFETCH_DEPTH = FAB_DAG_PREFETCH_DEPTH;
// List scheduling of the bulk stream: tasks that become ready within this many (estimated) microseconds compete with
// the ready ones (see the scheduling loop).
constexpr double DAG_LOOKAHEAD_US = 3.0;   // measured at N = 256: 0 -> 256.4 / 132.4 us (ll+grad / ll), 2..4 -> 253.6 / 129.0, 8 -> 255.6 / 131.2

struct DagOp {             // 28 ints, read by every thread of its stream
    int type, I, J, k;
    int slotA, slotB, slotC;
    int wait_other;        // number of completed ops of the other stream required before this op starts
    int wait_mask;         // bit s: wait for the TMA load into slot s
    int flags;
    int st_slot, st_tile;  // bulk: bulk-store slot -> workspace tile at the start of this op (-1: none)
    int nload;
    int ld_slot[DAG_MAX_LOADS];
    int ld_tile[DAG_MAX_LOADS];
    int put_kind;          // DOPF_PUT_AFTER: TK_U / TK_P / TK_T
    int pad[2];
};

static_assert(sizeof(DagOp) == 112, "DagOp is read as int4 words by the kernel");

struct DagProgram {
    std::vector<DagOp> bulk, chain;
    int NT = 0, nslot = 0, mode = 0;
    bool ok = false;
    double est_us = 0.0;   // estimated makespan of the schedule (cost model, for diagnostics)
    std::string err;
};

namespace dagdetail {

// L, W: final tiles (have a workspace location); U: updated covariance tile awaiting its chain op / panel
// solve; P: partially updated diagonal tile; T: sum_k L(I,k) W(k,J) awaiting W(I,I).  U, P, T never leave
// shared memory (pinned).
enum { TK_L = 0, TK_W = 1, TK_U = 2, TK_P = 3, TK_T = 4 };
inline int tid_of(int kind, int I, int J) { return (kind << 28) | (I << 14) | J; }
inline int tid_kind(int id) { return id >> 28; }
inline int tid_I(int id) { return (id >> 14) & 0x3fff; }
inline int tid_J(int id) { return id & 0x3fff; }
inline int tri(int I, int J) { return I * (I + 1) / 2 + J; }
// Workspace location of a tile: L(I,J) and W(I,J) share tri(I,J) (W replaces L), except that in FACTOR
// mode, where the L(K,K) are outputs, a W(K,K) that has to be spilled goes to the NT extra tiles.
inline int ws_index(int id, int mode, int NT) {
    if (mode == 1 /* DAG_FACTOR */ && tid_kind(id) == 1 /* TK_W */) return NT * (NT + 1) / 2 + tid_I(id);
    return tri(tid_I(id), tid_J(id));
}

enum { T_FA = 0, T_CH = 1, T_FB = 2, T_P1 = 3, T_P2 = 4, T_FA1 = 5, T_FA2 = 6, T_P1A = 7, T_P1B = 8 };
struct Task {
    int kind, I, J;
    std::vector<int> deps, succ;
    double cost = 0, prio = 0, start = -1, end = -1;
    bool scheduled = false;
};
// estimated costs in microseconds (B200, profiles/r01f_phase_cycles.txt)
struct DagCosts { double build, mma, put, trsm, p1fin, contract, chain; };
inline const DagCosts& dag_costs() {
    static DagCosts c = [] {
        DagCosts d{2.7, 2.3, 0.3, 3.2, 2.7, 3.8, 21.0};
        if (const char* e = getenv("FAB_DAG_COSTS"))          // (development: schedule experiments)
            sscanf(e, "%lf,%lf,%lf,%lf,%lf,%lf,%lf", &d.build, &d.mma, &d.put, &d.trsm, &d.p1fin, &d.contract, &d.chain);
        return d;
    }();
    return c;
}
#define C_BUILD (dag_costs().build)
#define C_MMA (dag_costs().mma)
#define C_PUT (dag_costs().put)
#define C_TRSM (dag_costs().trsm)
#define C_P1FIN (dag_costs().p1fin)
#define C_CONTRACT (dag_costs().contract)
#define C_CHAIN (dag_costs().chain)

struct SymOp {            // one op before slot assignment
    int stream;           // 0 bulk, 1 chain
    int type, I, J, k, flags;
    int in[2];            // input tile ids (-1 none)
    int inplace;          // tile id transformed in place (-1 none) ...
    int inplace_to;       // ... into this id
    int out;              // new tile id produced into a fresh slot (-1 none)
    double t;             // estimated start
    int idx;              // index within its stream
    int get_in;           // DOPF_GET_BEFORE: partial tile picked up first (-1 none)
    int put_out;          // DOPF_PUT_AFTER: tile id put into a fresh slot (-1: none or in place)
    int put_inplace, put_inplace_to;   // DOPF_PUT_AFTER | DOPF_INPLACE: over the slot of the partial tile
    int put_kind;
};

}  // namespace dagdetail

// split = true: tasks that would hold the accumulators across a wait for the chain are cut in two at a
// PUT / GET boundary -- the diagonal tile's build and early updates run before the panel solve it still needs
// (FA1 | FA2), the products of an inverse tile run before its diagonal inverse exists (P1A | P1B).
inline DagProgram build_dag_program_impl(int NT, int nslot, int mode, bool all_stores, bool split);
inline DagProgram build_dag_program(int NT, int nslot, int mode, bool all_stores = false) {
    DagProgram P = build_dag_program_impl(NT, nslot, mode, all_stores, true);
    if (!P.ok) P = build_dag_program_impl(NT, nslot, mode, all_stores, false);   // fewer pinned tiles
    return P;
}

// -------------------------------------------------------------------------------------------------
inline DagProgram build_dag_program_impl(int NT, int nslot, int mode, bool all_stores, bool split) {
    using namespace dagdetail;
    DagProgram P;
    P.NT = NT; P.nslot = nslot; P.mode = mode;
    if (NT < 1 || nslot < 4 || nslot > 30) { P.err = "need NT >= 1 and 4 <= slots <= 30"; return P; }
    const bool with_p1 = mode >= DAG_FIT, with_p2 = mode == DAG_GRAD;

    // ---- tasks and dependencies ----
    std::vector<Task> T;
    std::map<long long, int> key;
    auto K3 = [](int kind, int I, int J) { return ((long long)kind << 40) | ((long long)I << 20) | J; };
    auto add = [&](int kind, int I, int J, double cost) {
        Task t; t.kind = kind; t.I = I; t.J = J; t.cost = cost;
        key[K3(kind, I, J)] = (int)T.size();
        T.push_back(t);
    };
    auto id = [&](int kind, int I, int J) { return key.at(K3(kind, I, J)); };
    for (int K = 0; K < NT; ++K) {
        for (int I = K; I < NT; ++I) {
            if (split && I == K && K >= 1) {
                add(T_FA1, K, K, C_BUILD + C_MMA * (K - 1) + C_PUT);
                add(T_FA2, K, K, C_PUT + C_MMA + C_PUT);
            } else {
                add(T_FA, I, K, C_BUILD + C_MMA * K + C_PUT);
            }
        }
        add(T_CH, K, K, C_CHAIN);
        for (int I = K + 1; I < NT; ++I) add(T_FB, I, K, C_TRSM);
    }
    if (with_p1)
        for (int I = 1; I < NT; ++I)
            for (int J = 0; J < I; ++J) {
                if (split) { add(T_P1A, I, J, C_MMA * (I - J) + C_PUT); add(T_P1B, I, J, C_P1FIN); }
                else add(T_P1, I, J, C_MMA * (I - J) + C_P1FIN);
            }
    if (with_p2)
        for (int J = 0; J < NT; ++J)
            for (int I = J; I < NT; ++I) add(T_P2, I, J, C_MMA * (NT - I) + C_CONTRACT);
    auto dep = [&](int t, int d) { T[t].deps.push_back(d); };
    const bool dsplit = split;
    auto fa_last = [&](int I, int K) { return (dsplit && I == K && K >= 1) ? id(T_FA2, K, K) : id(T_FA, I, K); };   // task that completes U(I,K)
    auto p1_done = [&](int I, int J) { return split ? id(T_P1B, I, J) : id(T_P1, I, J); };                          // task that completes W(I,J)
    for (int K = 0; K < NT; ++K) {
        for (int I = K; I < NT; ++I) {
            if (dsplit && I == K && K >= 1) {
                const int t1 = id(T_FA1, K, K), t2 = id(T_FA2, K, K);
                for (int J = 0; J + 1 < K; ++J) dep(t1, id(T_FB, K, J));
                dep(t2, t1); dep(t2, id(T_FB, K, K - 1));
            } else {
                const int t = id(T_FA, I, K);
                for (int J = 0; J < K; ++J) { dep(t, id(T_FB, I, J)); if (I != K) dep(t, id(T_FB, K, J)); }
            }
        }
        dep(id(T_CH, K, K), fa_last(K, K));
        for (int I = K + 1; I < NT; ++I) { dep(id(T_FB, I, K), id(T_FA, I, K)); dep(id(T_FB, I, K), id(T_CH, K, K)); }
    }
    if (with_p1)
        for (int I = 1; I < NT; ++I)
            for (int J = 0; J < I; ++J) {
                // ta: the products sum_k L(I,k) W(k,J); tb: the multiplication by -W(I,I) (and the store of W(I,J))
                const int ta = split ? id(T_P1A, I, J) : id(T_P1, I, J), tb = p1_done(I, J);
                dep(tb, id(T_CH, I, I)); dep(ta, id(T_CH, J, J));
                if (ta != tb) dep(tb, ta);
                for (int k = J; k < I; ++k) dep(ta, id(T_FB, I, k));
                for (int k = J + 1; k < I; ++k) dep(ta, p1_done(k, J));
                // W(I,J) takes the place of L(I,J) in the workspace: every reader of L(I,J) comes first
                for (int I2 = I + 1; I2 < NT; ++I2) dep(tb, id(T_FA, I2, I));
                for (int J2 = 0; J2 < J; ++J2) dep(tb, split ? id(T_P1A, I, J2) : id(T_P1, I, J2));
                dep(tb, id(T_CH, I, I));
            }
    if (with_p2)
        for (int J = 0; J < NT; ++J)
            for (int I = J; I < NT; ++I) {
                const int t = id(T_P2, I, J);
                dep(t, id(T_CH, NT - 1, NT - 1));
                dep(t, id(T_CH, I, I)); dep(t, id(T_CH, J, J));
                for (int k = I + 1; k < NT; ++k) dep(t, p1_done(k, I));      // W(k,I), and alpha_I final
                for (int k = J + 1; k < NT; ++k) dep(t, p1_done(k, J));      // W(k,J), and alpha_J final
            }
    const int nt = (int)T.size();
    for (int t = 0; t < nt; ++t) {
        std::sort(T[t].deps.begin(), T[t].deps.end());
        T[t].deps.erase(std::unique(T[t].deps.begin(), T[t].deps.end()), T[t].deps.end());
        for (int d : T[t].deps) T[d].succ.push_back(t);
    }
    // priority = longest path to a sink (tasks were created in a topological order)
    for (int t = nt - 1; t >= 0; --t) {
        double m = 0;
        for (int s : T[t].succ) m = std::max(m, T[s].prio);
        T[t].prio = T[t].cost + m;
    }

    // ---- phase 1: list scheduling of the tasks on the two streams ----
    const int ucap = std::max(1, nslot - 3);       // unfinished (pinned, never stored) tiles alive at once
    // tasks that leave a pinned tile behind, and the task that consumes it
    auto makes_pinned = [&](int t) { return T[t].kind == T_FA || T[t].kind == T_FA1 || T[t].kind == T_P1A; };
    auto consumer_of = [&](int t) {
        switch (T[t].kind) {
        case T_FA1: return id(T_FA2, T[t].I, T[t].J);
        case T_P1A: return id(T_P1B, T[t].I, T[t].J);
        default: return T[t].I == T[t].J ? id(T_CH, T[t].I, T[t].I) : id(T_FB, T[t].I, T[t].J);
        }
    };
    double tfree[2] = {0.0, 0.0};
    std::vector<int> order[2];
    double look = DAG_LOOKAHEAD_US;
    if (const char* e = getenv("FAB_DAG_LOOK")) look = atof(e);          // (development: schedule experiments)
    int remaining = nt;
    while (remaining > 0) {
        int best[2] = {-1, -1};
        double best_est[2] = {1e300, 1e300};
        for (int s = 0; s < 2; ++s) {
            int ready_best = -1, early_best = -1;
            double early_est = 1e300, ready_est = 0.0;
            for (int t = 0; t < nt; ++t) {
                if (T[t].scheduled || (T[t].kind == T_CH) != (s == 1)) continue;
                double est = tfree[s];
                bool okd = true;
                for (int d : T[t].deps) { if (!T[d].scheduled) { okd = false; break; } est = std::max(est, T[d].end); }
                if (!okd) continue;
                if (makes_pinned(t)) {                 // cap on live pinned tiles at the end of this task
                    int unsched = 0;
                    std::vector<double> ends;
                    for (int f = 0; f < nt; ++f)
                        if (makes_pinned(f) && T[f].scheduled) {
                            const Task& c = T[consumer_of(f)];
                            if (!c.scheduled) ++unsched; else if (c.end > est + T[t].cost) ends.push_back(c.end);
                        }
                    if (unsched >= ucap) continue;     // blocked until a consumer gets scheduled
                    std::sort(ends.begin(), ends.end());
                    const int excess = unsched + (int)ends.size() - (ucap - 1);   // consumers that must finish first
                    if (excess > 0) est = std::max(est, ends[excess - 1] - T[t].cost);
                }
                // a task that becomes ready within `look` us competes with the ready ones: the bulk stream rather idles
                // for a moment than starts a long low-priority task right before the panel solve the chain waits for
                if (est <= tfree[s] + (s == 0 ? look : 0.0) + 1e-9) {
                    if (ready_best < 0 || T[t].prio > T[ready_best].prio) { ready_best = t; ready_est = std::max(est, tfree[s]); }
                }
                else if (est < early_est - 1e-9 || (est < early_est + 1e-9 && T[t].prio > T[early_best].prio)) { early_est = est; early_best = t; }
            }
            if (ready_best >= 0) { best[s] = ready_best; best_est[s] = ready_est; }
            else if (early_best >= 0) { best[s] = early_best; best_est[s] = early_est; }
        }
        int s;
        if (best[0] < 0 && best[1] < 0) { P.err = "scheduler stalled (slot cap)"; return P; }
        if (best[0] < 0) s = 1; else if (best[1] < 0) s = 0; else s = best_est[1] < best_est[0] ? 1 : 0;
        Task& t = T[best[s]];
        t.start = best_est[s]; t.end = t.start + t.cost; t.scheduled = true;
        tfree[s] = t.end;
        order[s].push_back(best[s]);
        --remaining;
    }

    P.est_us = std::max(tfree[0], tfree[1]);

    // ---- expand into symbolic ops with estimated start times ----
    std::vector<SymOp> ops;
    const bool fuse_build = getenv("FAB_DAG_NOFUSE") == nullptr;      // (development: A/B of the fused BUILD + MMA op)
    int pending_get = -1;                      // partial tile a GET wants: rides on the next op pushed
    double pending_get_t = 0.0;
    auto push = [&](int stream, int type, int I, int J, int k, int flags, int in0, int in1, int inpl, int inpl_to, int out,
                    double t) {
        SymOp o; o.stream = stream; o.type = type; o.I = I; o.J = J; o.k = k; o.flags = flags; o.in[0] = in0; o.in[1] = in1;
        o.inplace = inpl; o.inplace_to = inpl_to; o.out = out; o.t = t; o.idx = -1;
        o.get_in = -1; o.put_out = -1; o.put_inplace = -1; o.put_inplace_to = -1; o.put_kind = -1;
        if (type == DOP_GET) { pending_get = in0; pending_get_t = t; return; }
        if (type == DOP_PUT) {                 // rides on the op that has just produced the accumulators
            SymOp& p = ops.back();
            p.flags |= DOPF_PUT_AFTER | (flags & DOPF_INPLACE);
            p.put_kind = k; p.put_out = out; p.put_inplace = inpl; p.put_inplace_to = inpl_to;
            return;
        }
        if (pending_get >= 0 && stream == 0) {
            o.flags |= DOPF_GET_BEFORE; o.get_in = pending_get; o.t = pending_get_t;
            pending_get = -1;
        }
        if (fuse_build && type == DOP_MMA_NT && !ops.empty() && !(o.flags & DOPF_GET_BEFORE)) {
            SymOp& p = ops.back();
            if (p.stream == 0 && p.type == DOP_BUILD && p.I == I && p.J == J && p.flags == 0) {
                o.flags |= DOPF_BUILD_FIRST; o.t = p.t;
                p = o;
                return;
            }
        }
        ops.push_back(o);
    };
    for (int s = 0; s < 2; ++s)
        for (int ti : order[s]) {
            const Task& t = T[ti];
            double tm = t.start;
            const int I = t.I, J = t.J;
            switch (t.kind) {
            case T_FA:
                push(0, DOP_BUILD, I, J, 0, 0, -1, -1, -1, -1, -1, tm); tm += C_BUILD;
                for (int k = 0; k < J; ++k) { push(0, DOP_MMA_NT, I, J, k, 0, tid_of(TK_L, I, k), tid_of(TK_L, J, k), -1, -1, -1, tm); tm += C_MMA; }
                push(0, DOP_PUT, I, J, TK_U, 0, -1, -1, -1, -1, tid_of(TK_U, I, J), tm);
                break;
            case T_FA1:
                push(0, DOP_BUILD, I, J, 0, 0, -1, -1, -1, -1, -1, tm); tm += C_BUILD;
                for (int k = 0; k + 1 < J; ++k) { push(0, DOP_MMA_NT, I, J, k, 0, tid_of(TK_L, I, k), tid_of(TK_L, J, k), -1, -1, -1, tm); tm += C_MMA; }
                push(0, DOP_PUT, I, J, TK_P, 0, -1, -1, -1, -1, tid_of(TK_P, I, J), tm);
                break;
            case T_FA2:
                push(0, DOP_GET, I, J, TK_P, 0, tid_of(TK_P, I, J), -1, -1, -1, -1, tm); tm += C_PUT;
                push(0, DOP_MMA_NT, I, J, J - 1, 0, tid_of(TK_L, I, J - 1), tid_of(TK_L, J, J - 1), -1, -1, -1, tm); tm += C_MMA;
                push(0, DOP_PUT, I, J, TK_U, DOPF_INPLACE, -1, -1, tid_of(TK_P, I, J), tid_of(TK_U, I, J), -1, tm);
                break;
            case T_CH:
                push(1, DOP_CH_POTRF, I, I, I, 0, -1, -1, tid_of(TK_U, I, I), tid_of(TK_L, I, I), tid_of(TK_W, I, I), tm);
                break;
            case T_FB:
                push(0, DOP_TRSM, I, J, J, 0, tid_of(TK_W, J, J), -1, tid_of(TK_U, I, J), tid_of(TK_L, I, J), -1, tm);
                break;
            case T_P1: case T_P1A:
                for (int k = J; k < I; ++k) {
                    push(0, DOP_P1_MMA, I, J, k, k == J ? DOPF_FIRST : 0, tid_of(TK_L, I, k), tid_of(TK_W, k, J), -1, -1, -1, tm);
                    tm += C_MMA;
                }
                push(0, DOP_PUT, I, J, TK_T, 0, -1, -1, -1, -1, tid_of(TK_T, I, J), tm); tm += C_PUT;
                if (t.kind == T_P1) push(0, DOP_P1_FIN, I, J, I, 0, tid_of(TK_W, I, I), -1, tid_of(TK_T, I, J), tid_of(TK_W, I, J), -1, tm);
                break;
            case T_P1B:
                push(0, DOP_P1_FIN, I, J, I, 0, tid_of(TK_W, I, I), -1, tid_of(TK_T, I, J), tid_of(TK_W, I, J), -1, tm);
                break;
            case T_P2:
                for (int k = I; k < NT; ++k) {
                    push(0, DOP_P2_MMA, I, J, k, (k == I ? DOPF_FIRST : 0) | (k == NT - 1 ? DOPF_LAST : 0), tid_of(TK_W, k, I),
                         tid_of(TK_W, k, J), -1, -1, -1, tm);
                    tm += C_MMA;
                }
                break;
            }
        }
    {   // per-stream indices, then the merged (estimated-time) order used for slot assignment
        int n0 = 0, n1 = 0;
        for (SymOp& o : ops) o.idx = o.stream == 0 ? n0++ : n1++;
    }
    std::vector<int> seq(ops.size());
    for (size_t i = 0; i < ops.size(); ++i) seq[i] = (int)i;
    std::stable_sort(seq.begin(), seq.end(), [&](int a, int b) {
        if (ops[a].t != ops[b].t) return ops[a].t < ops[b].t;
        return ops[a].stream > ops[b].stream;
    });
    // positions must respect per-stream program order (estimated times are monotone per stream)
    std::vector<int> pos(ops.size());
    for (size_t g = 0; g < seq.size(); ++g) pos[seq[g]] = (int)g;

    // future uses of every tile id (global positions), for farthest-next-use replacement
    std::map<int, std::vector<int>> uses;
    for (size_t g = 0; g < seq.size(); ++g) {
        const SymOp& o = ops[seq[g]];
        for (int q = 0; q < 2; ++q) if (o.in[q] >= 0) uses[o.in[q]].push_back((int)g);
        if (o.inplace >= 0) uses[o.inplace].push_back((int)g);
        if (o.get_in >= 0) uses[o.get_in].push_back((int)g);
        if (o.put_inplace >= 0) uses[o.put_inplace].push_back((int)g);
    }
    auto next_use = [&](int tile, int after) {
        auto it = uses.find(tile);
        if (it == uses.end()) return 1 << 30;
        auto p = std::upper_bound(it->second.begin(), it->second.end(), after);
        return p == it->second.end() ? (1 << 30) : *p;
    };
    auto is_output = [&](int tile) {           // must end up in the workspace whatever the schedule
        if (mode == DAG_FACTOR) return tid_kind(tile) == TK_L;
        if (mode == DAG_FIT) return tid_kind(tile) == TK_W;
        return false;
    };

    // ---- phase 2: slots, loads, stores, waits (possibly twice: the second pass drops useless stores) ----
    std::map<int, char> reloaded_prev;
    for (int pass = 0; pass < 2; ++pass) {
        const bool store_all = all_stores || pass == 0;
        std::vector<DagOp> prog[2];
        int cnt[2] = {0, 0};
        for (const SymOp& o : ops) ++cnt[o.stream];
        prog[0].assign(cnt[0], DagOp{}); prog[1].assign(cnt[1], DagOp{});
        for (const SymOp& o : ops) {
            DagOp& d = prog[o.stream][o.idx];
            d.type = o.type; d.I = o.I; d.J = o.J; d.k = o.k; d.flags = o.flags;
            d.slotA = d.slotB = d.slotC = -1; d.st_slot = d.st_tile = -1; d.put_kind = o.put_kind;
        }
        struct Slot { int id = -1; bool clean = false; int lastB = -1, lastC = -1, storeB = -1; };
        std::vector<Slot> S(nslot);
        struct Load { int slot, tile, needed_at, free_from, min_chain, stored_at; bool after_store; };
        std::vector<Load> loads;
        std::map<int, int> stored_at;              // tile id -> bulk op that issues its store
        std::map<int, char> reloaded;
        struct Pend { int slot, tile_id; int chain_op; double ready_t; };
        std::vector<Pend> pending;                 // stores waiting for a bulk op to carry them
        bool failed = false;
        std::vector<int> bulk_pos_of_idx(cnt[0]);
        for (size_t i = 0; i < ops.size(); ++i) if (ops[i].stream == 0) bulk_pos_of_idx[ops[i].idx] = pos[i];
        std::vector<int> chain_pos_of_idx(cnt[1]);
        std::vector<double> chain_end_t(cnt[1]);
        for (size_t i = 0; i < ops.size(); ++i)
            if (ops[i].stream == 1) { chain_pos_of_idx[ops[i].idx] = pos[i]; chain_end_t[ops[i].idx] = ops[i].t + C_CHAIN; }
        auto need_store = [&](int tile) {
            if (tid_kind(tile) >= TK_U) return false;           // U, P, T never leave shared memory
            if (tid_kind(tile) == TK_L && tid_I(tile) == tid_J(tile)) return mode == DAG_FACTOR;   // L(K,K) is dead after its chain op
            if (is_output(tile) || store_all) return true;
            auto it = reloaded_prev.find(tile);
            return it != reloaded_prev.end();
        };

        for (size_t g = 0; g < seq.size() && !failed; ++g) {
            const SymOp& o = ops[seq[g]];
            DagOp& d = prog[o.stream][o.idx];
            unsigned busy = 0;
            auto wait_for_chain = [&](DagOp& bd, int c) { if (c >= 0) bd.wait_other = std::max(bd.wait_other, c + 1); };
            auto touch = [&](int q) {
                busy |= 1u << q;
                if (o.stream == 0) { wait_for_chain(d, S[q].lastC); S[q].lastB = o.idx; }
                else { if (S[q].lastB >= 0) d.wait_other = std::max(d.wait_other, S[q].lastB + 1); S[q].lastC = o.idx; }
            };
            auto find = [&](int tile) { for (int q = 0; q < nslot; ++q) if (S[q].id == tile) return q; return -1; };
            auto victim = [&](bool for_chain) {
                int best = -1; long long best_score = -1;
                for (int q = 0; q < nslot; ++q) {
                    if ((busy >> q) & 1u) continue;
                    long long score;
                    const int nu = S[q].id < 0 ? (1 << 30) : next_use(S[q].id, (int)g - 1);
                    const bool awaiting_store = S[q].id >= 0 && need_store(S[q].id) && !S[q].clean;
                    // free or dead slots first, the one idle for longest first (its reload can be hoisted furthest)
                    if (S[q].id < 0) score = (1LL << 40) - S[q].lastB;
                    else if (nu == (1 << 30) && !awaiting_store) score = (1LL << 36) - S[q].lastB;   // dead
                    else if (S[q].clean) {
                        // farthest next use -- except that a slot the bulk stream has only just finished with cannot be
                        // refilled ahead of time (its reload would be issued and awaited in the same op): last resort.
                        // Costs a few more reloads, removes most of the waits for them (N = 2048: -15 %).
                        score = nu;
                        if (o.stream == 0 && S[q].lastB >= o.idx - 1) score = 2 + (score >> 6);
                    }
                    else continue;                                   // live and not in the workspace: pinned
                    if (for_chain && S[q].storeB >= 0) {
                        // the chain can only take a slot whose bulk store is known to have been read out
                        const int r = S[q].storeB + 1;
                        if (r >= cnt[0] || bulk_pos_of_idx[r] >= (int)g) continue;
                    }
                    // prefer not to take a slot the other stream is (estimated to be) still working on
                    if (o.stream == 0 && S[q].lastC >= 0 && chain_end_t[S[q].lastC] > o.t + 1e-9) score = 1;
                    if (score > best_score) { best_score = score; best = q; }
                }
                return best;
            };
            auto overwrite = [&](int q, int new_id, bool clean) {        // q becomes new_id (produced in shared memory)
                if (S[q].storeB >= 0) {
                    if (o.stream == 0) { d.flags |= DOPF_WAIT_READ; }
                    else {
                        DagOp& r = prog[0][S[q].storeB + 1];
                        r.flags |= DOPF_WAIT_READ;
                        d.wait_other = std::max(d.wait_other, S[q].storeB + 2);
                    }
                    S[q].storeB = -1;
                }
                S[q].id = new_id; S[q].clean = clean;
                touch(q);
            };
            auto resident = [&](int tile) {                              // bulk input tile -> slot
                int q = find(tile);
                if (q >= 0) { touch(q); return q; }
                if (o.stream != 0 || !stored_at.count(tile)) { failed = true; P.err = "tile needed but not in the workspace"; return 0; }
                q = victim(false);
                if (q < 0) { failed = true; P.err = "no evictable slot"; return 0; }
                Load l;
                l.slot = q; l.tile = ws_index(tile, mode, NT); l.needed_at = o.idx;
                l.free_from = std::max(S[q].lastB + 1, stored_at[tile] + 1);     // an op issues its loads before its store
                l.min_chain = S[q].lastC; l.stored_at = stored_at[tile]; l.after_store = S[q].storeB >= 0;
                if (S[q].storeB >= 0) l.free_from = std::max(l.free_from, S[q].storeB + 1);
                if (S[q].lastC >= 0) {     // do not make the bulk stream wait for the chain earlier than the schedule expects
                    int e = 0;
                    while (e < cnt[0] && bulk_pos_of_idx[e] < chain_pos_of_idx[S[q].lastC]) ++e;
                    l.free_from = std::max(l.free_from, std::min(e, o.idx));
                }
                loads.push_back(l);
                S[q].storeB = -1;
                S[q].id = tile; S[q].clean = true; S[q].lastC = -1;
                S[q].lastB = o.idx; busy |= 1u << q;
                d.wait_mask |= 1 << q;
                reloaded[tile] = 1;
                return q;
            };
            // a bulk op carries at most one pending store (issued right after its start barrier)
            int pend_i = -1;
            if (o.stream == 0)
                for (size_t q = 0; q < pending.size(); ++q) if (pending[q].ready_t <= o.t + 1e-9) { pend_i = (int)q; break; }
            if (pend_i >= 0) {
                const Pend p = pending[pend_i];
                pending.erase(pending.begin() + pend_i);
                d.st_slot = p.slot; d.st_tile = ws_index(p.tile_id, mode, NT);
                if (p.chain_op >= 0) wait_for_chain(d, p.chain_op);
                S[p.slot].clean = true; S[p.slot].storeB = o.idx; S[p.slot].lastB = std::max(S[p.slot].lastB, o.idx);
                S[p.slot].lastC = -1;                 // ordered through this op's wait from here on
                stored_at[p.tile_id] = o.idx;
                busy |= 1u << p.slot;                 // the store reads the slot during this very op
            }
            if (o.flags & DOPF_GET_BEFORE) {
                const int q = find(o.get_in);
                if (q < 0) { failed = true; P.err = "partial tile lost"; break; }
                touch(q);
                d.slotC = q;
            }
            switch (o.type) {
            case DOP_BUILD: break;
            case DOP_MMA_NT: case DOP_P1_MMA: case DOP_P2_MMA:
                d.slotA = resident(o.in[0]);
                if (failed) break;
                d.slotB = o.in[1] == o.in[0] ? d.slotA : resident(o.in[1]);
                break;
            case DOP_PUT: {
                if (o.inplace >= 0) {                       // over the partial tile the preceding GET picked up
                    const int q = find(o.inplace);
                    if (q < 0) { failed = true; P.err = "partial tile lost"; break; }
                    touch(q);
                    S[q].id = o.inplace_to; S[q].clean = false;
                    d.slotC = q;
                    break;
                }
                const int q = victim(false);
                if (q < 0) { failed = true; P.err = "no slot for an unfinished tile"; break; }
                overwrite(q, o.out, false);
                d.slotC = q;
                break;
            }
            case DOP_GET: {
                const int q = find(o.in[0]);
                if (q < 0) { failed = true; P.err = "partial tile lost"; break; }
                touch(q);
                d.slotC = q;
                break;
            }
            case DOP_TRSM: {
                const int q = find(o.inplace);
                if (q < 0) { failed = true; P.err = "unfinished tile lost"; break; }
                touch(q);
                d.slotC = q;
                d.slotB = resident(o.in[0]);
                if (failed) break;
                S[q].id = o.inplace_to; S[q].clean = false;
                if (need_store(o.inplace_to)) pending.push_back({q, o.inplace_to, -1, -1.0});
                break;
            }
            case DOP_P1_FIN: {
                const int q = find(o.inplace);
                if (q < 0) { failed = true; P.err = "product tile lost"; break; }
                touch(q);
                d.slotC = q;
                d.slotA = resident(o.in[0]);
                if (failed) break;
                S[q].id = o.inplace_to; S[q].clean = false;
                if (need_store(o.inplace_to)) pending.push_back({q, o.inplace_to, -1, -1.0});
                break;
            }
            case DOP_CH_POTRF: {
                const int q = find(o.inplace);
                if (q < 0) { failed = true; P.err = "diagonal tile lost"; break; }
                touch(q);
                d.slotA = q;
                const int w = victim(true);
                if (w < 0) { failed = true; P.err = "no slot for the inverse of a diagonal tile"; break; }
                overwrite(w, o.out, false);
                d.slotC = w;
                S[q].id = o.inplace_to; S[q].clean = false;
                if (need_store(o.inplace_to)) pending.push_back({q, o.inplace_to, o.idx, chain_end_t[o.idx]});
                if (need_store(o.out)) pending.push_back({w, o.out, o.idx, chain_end_t[o.idx]});
                break;
            }
            default: break;
            }
            if (!failed && (o.flags & DOPF_PUT_AFTER)) {
                if (o.put_inplace >= 0) {                   // over the partial tile this op (or an earlier one) picked up
                    const int q = find(o.put_inplace);
                    if (q < 0) { failed = true; P.err = "partial tile lost"; break; }
                    touch(q);
                    S[q].id = o.put_inplace_to; S[q].clean = false;
                    d.slotC = q;
                } else {
                    int q = victim(false);                  // not one of this op's operand slots (busy) ...
                    if (q < 0) {                            // ... unless nothing else is free: then after a barrier
                        const unsigned keep = busy;
                        // (not the slot this op's own bulk store is reading from)
                        if (d.slotA >= 0 && d.slotA != d.st_slot) busy &= ~(1u << d.slotA);
                        if (d.slotB >= 0 && d.slotB != d.st_slot) busy &= ~(1u << d.slotB);
                        q = victim(false);
                        busy = keep;
                        if (q >= 0) d.flags |= DOPF_PUT_BARRIER;
                    }
                    if (q < 0) { failed = true; P.err = "no slot for an unfinished tile"; break; }
                    overwrite(q, o.put_out, false);
                    d.slotC = q;
                }
            }
        }
        if (failed) return P;
        // trailing stores ride on NOP ops appended to the bulk stream
        while (!pending.empty()) {
            const Pend p = pending.front();
            pending.erase(pending.begin());
            DagOp n{};
            n.type = DOP_NOP; n.slotA = n.slotB = n.slotC = -1;
            n.st_slot = p.slot; n.st_tile = ws_index(p.tile_id, mode, NT);
            if (p.chain_op >= 0) n.wait_other = p.chain_op + 1;
            stored_at[p.tile_id] = (int)prog[0].size();
            prog[0].push_back(n);
        }
        // hoist the loads; decide the drains
        std::sort(loads.begin(), loads.end(), [](const Load& a, const Load& b) { return a.needed_at < b.needed_at; });
        for (const Load& l : loads) {
            int e = std::max(l.free_from, l.needed_at - DAG_PREFETCH_DEPTH);
            e = std::max(e, 0);
            e = std::min(e, l.needed_at);
            while (e < l.needed_at && prog[0][e].nload >= DAG_MAX_LOADS) ++e;
            DagOp& d = prog[0][e];
            if (d.nload >= DAG_MAX_LOADS) { P.err = "too many loads on one op"; return P; }
            d.ld_slot[d.nload] = l.slot; d.ld_tile[d.nload] = l.tile; ++d.nload;
            if (l.min_chain >= 0) d.wait_other = std::max(d.wait_other, l.min_chain + 1);
            if (l.after_store) d.flags |= DOPF_WAIT_READ;
        }
        {   // drains: a load must see the completed store of its tile
            std::map<int, int> store_op_of_tile;    // workspace tile index -> last bulk op storing it so far
            int last_drain = -1;
            for (int e = 0; e < (int)prog[0].size(); ++e) {
                DagOp& d = prog[0][e];
                for (int q = 0; q < d.nload; ++q) {
                    auto it = store_op_of_tile.find(d.ld_tile[q]);
                    if (it != store_op_of_tile.end() && it->second >= last_drain) { d.flags |= DOPF_DRAIN; last_drain = e; }
                }
                // (the store of this op is issued after its loads)
                if (d.st_slot >= 0) store_op_of_tile[d.st_tile] = e;
            }
        }
        if (pass == 0) {
            reloaded_prev = reloaded;
            P.bulk = prog[0]; P.chain = prog[1];
            if (all_stores) break;
        } else {
            bool consistent = true;                 // pass 2 must not reload a tile whose store it dropped
            for (auto& kv : reloaded) if (!reloaded_prev.count(kv.first) && !is_output(kv.first)) consistent = false;
            if (consistent) { P.bulk = prog[0]; P.chain = prog[1]; }
        }
    }
    P.ok = true;
    return P;
}

// -------------------------------------------------------------------------------------------------
// Independent check of an emitted program: (1) replayed in an order consistent with the waits, every
// op finds the tile it expects in the slot it names, every load fetches a tile version that has been
// stored, no slot with an unread store or an unawaited load is overwritten; (2) every pair of
// conflicting slot accesses from different streams is ordered by an explicit wait.
// Returns an empty string when the program is valid.
// -------------------------------------------------------------------------------------------------
inline std::string dag_verify(const DagProgram& P) {
    using namespace dagdetail;
    if (!P.ok) return "program not built: " + P.err;
    const int nb = (int)P.bulk.size(), nc = (int)P.chain.size(), ns = P.nslot;
    char buf[256];
    // cumulative knowledge of the other stream
    std::vector<int> KC(nb, 0), KB(nc, 0);
    for (int i = 0, m = 0; i < nb; ++i) { m = std::max(m, P.bulk[i].wait_other); KC[i] = m; if (m > nc) return "bulk op waits for a chain op that does not exist"; }
    for (int j = 0, m = 0; j < nc; ++j) { m = std::max(m, P.chain[j].wait_other); KB[j] = m; if (m > nb) return "chain op waits for a bulk op that does not exist"; }
    // (2) accesses per slot
    struct Acc { int stream, idx; bool write; };
    std::vector<std::vector<Acc>> acc(ns);
    auto rec = [&](int slot, int stream, int idx, bool w) { if (slot >= 0) acc[slot].push_back({stream, idx, w}); };
    for (int i = 0; i < nb; ++i) {
        const DagOp& d = P.bulk[i];
        for (int q = 0; q < d.nload; ++q) rec(d.ld_slot[q], 0, i, true);
        if (d.st_slot >= 0) rec(d.st_slot, 0, i, false);
        if (d.flags & DOPF_GET_BEFORE) rec(d.slotC, 0, i, false);
        if (d.flags & DOPF_PUT_AFTER) {
            rec(d.slotC, 0, i, true);
            if ((d.slotC == d.slotA || d.slotC == d.slotB) && !(d.flags & DOPF_PUT_BARRIER))
                return "a fused PUT targets an operand slot of its own op without a barrier";
        }
        switch (d.type) {
        case DOP_MMA_NT: case DOP_P1_MMA: case DOP_P2_MMA: rec(d.slotA, 0, i, false); rec(d.slotB, 0, i, false); break;
        case DOP_PUT: rec(d.slotC, 0, i, true); break;
        case DOP_GET: rec(d.slotC, 0, i, false); break;
        case DOP_TRSM: rec(d.slotC, 0, i, true); rec(d.slotB, 0, i, false); break;
        case DOP_P1_FIN: rec(d.slotA, 0, i, false); rec(d.slotC, 0, i, true); break;
        default: break;
        }
    }
    for (int j = 0; j < nc; ++j) { rec(P.chain[j].slotA, 1, j, true); rec(P.chain[j].slotC, 1, j, true); }
    // a chain op that overwrites a slot after a bulk store of it must know that the store has been read out
    for (int i = 0; i < nb; ++i) {
        const int s = P.bulk[i].st_slot;
        if (s < 0) continue;
        int r = i + 1;
        while (r < nb && !(P.bulk[r].flags & DOPF_WAIT_READ)) ++r;
        for (int j = 0; j < nc; ++j) {
            if (P.chain[j].slotA != s && P.chain[j].slotC != s) continue;
            if (KC[i] >= j + 1) continue;                       // the chain op came first
            if (r >= nb || KB[j] < r + 1) {
                snprintf(buf, sizeof buf, "chain op %d overwrites slot %d without waiting for the read-out of the store of bulk op %d", j, s, i);
                return buf;
            }
        }
    }
    for (int s = 0; s < ns; ++s)
        for (const Acc& a : acc[s]) for (const Acc& b : acc[s]) {
            if (a.stream != 0 || b.stream != 1 || !(a.write || b.write)) continue;
            const bool a_first = KB[b.idx] >= a.idx + 1, b_first = KC[a.idx] >= b.idx + 1;
            if (!a_first && !b_first) {
                snprintf(buf, sizeof buf, "slot %d: bulk op %d and chain op %d conflict without an ordering wait", s, a.idx, b.idx);
                return buf;
            }
        }
    // (1) replay: advance whichever stream's next op has its wait satisfied (bulk first)
    std::vector<int> content(ns, -1);                 // tile id in the slot
    std::vector<int> load_pending(ns, 0), store_unread(ns, 0);
    std::map<int, int> ws;                            // workspace tile index -> tile id stored
    int ib = 0, ic = 0;
    auto expect = [&](int slot, int tile, const char* what, int stream, int idx) -> std::string {
        if (slot < 0 || slot >= ns) { snprintf(buf, sizeof buf, "%s op %d (%s): bad slot", stream ? "chain" : "bulk", idx, what); return buf; }
        if (content[slot] != tile) {
            snprintf(buf, sizeof buf, "%s op %d (%s): slot %d holds %x, expected %x", stream ? "chain" : "bulk", idx, what, slot, content[slot], tile);
            return buf;
        }
        if (load_pending[slot]) { snprintf(buf, sizeof buf, "%s op %d (%s): slot %d read before its load was awaited", stream ? "chain" : "bulk", idx, what, slot); return buf; }
        return "";
    };
    auto clobber = [&](int slot, const char* what, int stream, int idx) -> std::string {
        if (store_unread[slot]) { snprintf(buf, sizeof buf, "%s op %d (%s): slot %d overwritten while a bulk store may still read it", stream ? "chain" : "bulk", idx, what, slot); return buf; }
        if (load_pending[slot]) { snprintf(buf, sizeof buf, "%s op %d (%s): slot %d overwritten with a load in flight", stream ? "chain" : "bulk", idx, what, slot); return buf; }
        return "";
    };
    std::string e;
    int acc_owner = -1;                               // tile whose partial result the bulk accumulators hold
    const int NT = P.NT;
    std::vector<char> haveL(NT * NT, 0), haveW(NT * NT, 0);
    while (ib < nb || ic < nc) {
        bool progressed = false;
        if (ib < nb && P.bulk[ib].wait_other <= ic) {
            const DagOp& d = P.bulk[ib];
            if (d.flags & DOPF_WAIT_READ) for (int s = 0; s < ns; ++s) store_unread[s] = 0;
            for (int q = 0; q < d.nload; ++q) {
                const int s = d.ld_slot[q];
                if (!(e = clobber(s, "load", 0, ib)).empty()) return e;
                auto it = ws.find(d.ld_tile[q]);
                if (it == ws.end()) { snprintf(buf, sizeof buf, "bulk op %d loads workspace tile %d that was never stored", ib, d.ld_tile[q]); return buf; }
                content[s] = it->second; load_pending[s] = 1;
            }
            for (int s = 0; s < ns; ++s) if ((d.wait_mask >> s) & 1) {
                if (!load_pending[s]) { snprintf(buf, sizeof buf, "bulk op %d waits for slot %d with no load in flight", ib, s); return buf; }
                load_pending[s] = 0;
            }
            if (d.st_slot >= 0) {
                if (content[d.st_slot] < 0 || load_pending[d.st_slot]) { snprintf(buf, sizeof buf, "bulk op %d stores an invalid slot", ib); return buf; }
                if (ws_index(content[d.st_slot], P.mode, P.NT) != d.st_tile) { snprintf(buf, sizeof buf, "bulk op %d stores slot %d to the wrong workspace tile", ib, d.st_slot); return buf; }
                ws[d.st_tile] = content[d.st_slot]; store_unread[d.st_slot] = 1;
            }
            const int I = d.I, J = d.J, k = d.k;
            if (d.flags & DOPF_GET_BEFORE) {
                if (!(e = expect(d.slotC, tid_of(TK_P, I, J), "get", 0, ib)).empty()) return e;
                acc_owner = tid_of(TK_U, I, J);
            }
            switch (d.type) {
            case DOP_MMA_NT:
                if (d.flags & DOPF_BUILD_FIRST) acc_owner = tid_of(TK_U, I, J);
                if (acc_owner != tid_of(TK_U, I, J)) { snprintf(buf, sizeof buf, "bulk op %d updates accumulators that belong to another tile", ib); return buf; }
                if (!(e = expect(d.slotA, tid_of(TK_L, I, k), "A", 0, ib)).empty()) return e;
                if (!(e = expect(d.slotB, tid_of(TK_L, J, k), "B", 0, ib)).empty()) return e;
                break;
            case DOP_PUT:
                if (acc_owner != tid_of(k == TK_T ? TK_T : TK_U, I, J)) { snprintf(buf, sizeof buf, "bulk op %d puts accumulators that belong to another tile", ib); return buf; }
                if (d.flags & DOPF_INPLACE) { if (!(e = expect(d.slotC, tid_of(TK_P, I, J), "put over", 0, ib)).empty()) return e; }
                else if (!(e = clobber(d.slotC, "put", 0, ib)).empty()) return e;
                content[d.slotC] = tid_of(k, I, J);
                acc_owner = -1;
                break;
            case DOP_GET:
                if (!(e = expect(d.slotC, tid_of(TK_P, I, J), "get", 0, ib)).empty()) return e;
                acc_owner = tid_of(TK_U, I, J);
                break;
            case DOP_BUILD:
                acc_owner = tid_of(TK_U, I, J);
                break;
            case DOP_TRSM:
                if (!(e = expect(d.slotC, tid_of(TK_U, I, J), "C", 0, ib)).empty()) return e;
                if (!(e = expect(d.slotB, tid_of(TK_W, J, J), "B", 0, ib)).empty()) return e;
                content[d.slotC] = tid_of(TK_L, I, J); haveL[I * NT + J] = 1;
                break;
            case DOP_P1_MMA:
                if (d.flags & DOPF_FIRST) acc_owner = tid_of(TK_T, I, J);
                if (acc_owner != tid_of(TK_T, I, J)) { snprintf(buf, sizeof buf, "bulk op %d accumulates into accumulators that belong to another tile", ib); return buf; }
                if (!(e = expect(d.slotA, tid_of(TK_L, I, k), "A", 0, ib)).empty()) return e;
                if (!(e = expect(d.slotB, tid_of(TK_W, k, J), "B", 0, ib)).empty()) return e;
                break;
            case DOP_P1_FIN:
                if (!(e = expect(d.slotA, tid_of(TK_W, I, I), "A", 0, ib)).empty()) return e;
                if (!(e = expect(d.slotC, tid_of(TK_T, I, J), "C", 0, ib)).empty()) return e;
                if (d.slotC == d.slotA) return "P1_FIN output slot equals its operand";
                content[d.slotC] = tid_of(TK_W, I, J); haveW[I * NT + J] = 1;
                break;
            case DOP_P2_MMA:
                if (d.flags & DOPF_FIRST) acc_owner = tid_of(TK_W + 8, I, J);
                if (acc_owner != tid_of(TK_W + 8, I, J)) { snprintf(buf, sizeof buf, "bulk op %d accumulates into accumulators that belong to another tile", ib); return buf; }
                if (!(e = expect(d.slotA, tid_of(TK_W, k, I), "A", 0, ib)).empty()) return e;
                if (!(e = expect(d.slotB, tid_of(TK_W, k, J), "B", 0, ib)).empty()) return e;
                break;
            default: break;
            }
            if (d.flags & DOPF_PUT_AFTER) {
                const int pk = d.put_kind;
                if (acc_owner != tid_of(pk == TK_T ? TK_T : TK_U, I, J)) { snprintf(buf, sizeof buf, "bulk op %d puts accumulators that belong to another tile", ib); return buf; }
                if (d.flags & DOPF_INPLACE) { if (!(e = expect(d.slotC, tid_of(TK_P, I, J), "put over", 0, ib)).empty()) return e; }
                else if (!(e = clobber(d.slotC, "put", 0, ib)).empty()) return e;
                content[d.slotC] = tid_of(pk, I, J);
                acc_owner = -1;
            }
            ++ib; progressed = true;
        }
        if (ic < nc && P.chain[ic].wait_other <= ib) {
            const DagOp& d = P.chain[ic];
            if (d.type != DOP_CH_POTRF) return "unknown chain op";
            if (!(e = expect(d.slotA, tid_of(TK_U, d.I, d.I), "diag", 1, ic)).empty()) return e;
            if (!(e = clobber(d.slotC, "inverse", 1, ic)).empty()) return e;
            if (d.slotA == d.slotC) return "chain op output slot equals its input";
            content[d.slotA] = tid_of(TK_L, d.I, d.I); content[d.slotC] = tid_of(TK_W, d.I, d.I);
            haveL[d.I * NT + d.I] = 1; haveW[d.I * NT + d.I] = 1;
            ++ic; progressed = true;
        }
        if (!progressed) { snprintf(buf, sizeof buf, "deadlock at bulk op %d / chain op %d", ib, ic); return buf; }
    }
    for (int I = 0; I < NT; ++I) for (int J = 0; J <= I; ++J) {
        if (!haveL[I * NT + J]) return "a factor tile is never produced";
        if (P.mode >= DAG_FIT && !haveW[I * NT + J]) return "an inverse tile is never produced";
        if (P.mode == DAG_FACTOR && (!ws.count(tri(I, J)) || tid_kind(ws[tri(I, J)]) != TK_L)) return "FACTOR mode: a factor tile is not in the workspace at the end";
        if (P.mode == DAG_FIT && (!ws.count(tri(I, J)) || tid_kind(ws[tri(I, J)]) != TK_W)) return "FIT mode: an inverse tile is not in the workspace at the end";
    }
    return "";
}

}  // namespace fab
