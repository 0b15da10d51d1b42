// tile_program.h -- host-side static scheduler for the tile algebra of the gradient / fit kernel.
//
// The sequence of 64x64 tile operations of one walker (in-place triangular inverse, then the
// K^-1 tile products) depends only on (NT, number of shared-memory slots).  Instead of letting
// every thread of the CTA run cache bookkeeping at run time, the host simulates the slot cache once
// with Belady (farthest-next-use) replacement, hoists every HBM/L2 -> shared-memory tile load as
// early as its destination slot is free (up to PREFETCH_DEPTH operations ahead) and emits a flat
// "tile program".  The kernel is then a tiny interpreter: per operation one block barrier, the
// TMA bulk loads listed for it (issued by one thread, completing on per-slot mbarriers), the waits
// listed for it, and the DMMA work.  Plain C++ (no CUDA) so the emulator build shares it.
#pragma once
#include <algorithm>
#include <map>
#include <vector>

namespace fab {

enum TileOpType {
    OP_INV_DIAG = 0,   // C = W(J,J) = L(J,J)^-1                         A = L(J,J)
    OP_P1_MMA = 1,     // acc (+)= L(I,k) * W(k,J)                       A = L(I,k), B = W(k,J)   flags: first k
    OP_P1_FIN = 2,     // C = W(I,J) = -L(I,I)^-1 acc                    A = L(I,I)
    OP_P2_MMA = 3      // acc (+)= W(k,I)^T W(k,J); last k: contraction  A = W(k,I), B = W(k,J)
};
enum TileOpFlags {
    OPF_DRAIN_STORES = 1,   // wait for all bulk stores before issuing this op's loads (RAW through HBM)
    OPF_FIRST = 2,          // first product of an accumulation: acc = 0
    OPF_LAST = 4,           // last product: contraction (phase 2)
    OPF_NEW_PAIR = 8,       // first op of a (I, J) pair in phase 2: stage coordinate / alpha slices
    OPF_WAIT_READ = 16      // a slot whose tile is the source of an earlier bulk store gets overwritten here
};
constexpr int TP_DINV_SLOT0 = 8;      // ld_slot values >= 8 address the two 4 KB dinv buffers
constexpr int TP_MAX_LOADS = 12;     // <= 3 loads per op, hoisted from at most 4 ops (depth 3) into one
constexpr int TP_PREFETCH_DEPTH = 3;

struct TileOp {            // 36 ints, read by every thread of the CTA
    int type, I, J, k;
    int slotA, slotB, slotC;
    int wait_mask;         // bit s (< 8): tile slot s, bit 8 + b: dinv buffer b
    int dinv_buf;          // which dinv buffer this op reads (-1: none)
    int flags;
    int nload;
    int ld_slot[TP_MAX_LOADS];
    int ld_tile[TP_MAX_LOADS];   // tri_index(row, col) of the tile, or the block row for a dinv buffer
    int pad;
};

inline int tp_tri(int I, int J) { return I * (I + 1) / 2 + J; }
inline int tp_id(int kind, int I, int J) { return (kind << 30) | (I << 16) | J; }   // kind 0 = L, 1 = W

// Build the program.  with_grad = false stops after phase 1 (fit: W and alpha only).
inline std::vector<TileOp> build_grad_program(int NT, int nslot, bool with_grad) {
    struct Sym { int type, I, J, k, idA, idB, idC, dinv, flags; };
    std::vector<Sym> sym;
    for (int J = 0; J < NT; ++J) {
        sym.push_back({OP_INV_DIAG, J, J, J, tp_id(0, J, J), -1, tp_id(1, J, J), J, 0});
        for (int I = J + 1; I < NT; ++I) {
            for (int k = J; k < I; ++k)
                sym.push_back({OP_P1_MMA, I, J, k, tp_id(0, I, k), tp_id(1, k, J), -1, -1, k == J ? OPF_FIRST : 0});
            sym.push_back({OP_P1_FIN, I, J, I, tp_id(0, I, I), -1, tp_id(1, I, J), I, 0});
        }
    }
    const int n_phase1 = (int)sym.size();
    if (with_grad)
        for (int J = 0; J < NT; ++J)
            for (int I = J; I < NT; ++I)
                for (int k = I; k < NT; ++k)
                    sym.push_back({OP_P2_MMA, I, J, k, tp_id(1, k, I), tp_id(1, k, J), -1, -1,
                                   (k == I ? OPF_FIRST | OPF_NEW_PAIR : 0) | (k == NT - 1 ? OPF_LAST : 0)});
    const int n = (int)sym.size();

    // uses of every tile id (as an input), for Belady
    std::map<int, std::vector<int>> uses;
    for (int i = 0; i < n; ++i) {
        uses[sym[i].idA].push_back(i);
        if (sym[i].idB >= 0 && sym[i].idB != sym[i].idA) uses[sym[i].idB].push_back(i);
    }
    auto next_use = [&](int id, int after) {          // first use index > after, or a large number
        if (id < 0) return 1 << 30;
        auto it = uses.find(id);
        if (it == uses.end()) return 1 << 30;
        auto p = std::upper_bound(it->second.begin(), it->second.end(), after);
        return p == it->second.end() ? (1 << 30) : *p;
    };

    std::vector<TileOp> prog(n);
    std::vector<int> slot_id(nslot, -1), slot_last(nslot, -1);
    struct Load { int slot, tile, needed_at, free_from; bool is_w; bool after_store; };
    std::vector<char> slot_stored(nslot, 0);          // current content was written through to HBM by a bulk store
    std::vector<Load> loads;
    int dinv_row[2] = {-1, -1}, dinv_last[2] = {-1, -1};

    for (int i = 0; i < n; ++i) {
        const Sym& s = sym[i];
        TileOp& op = prog[i];
        op = TileOp{};
        op.type = s.type; op.I = s.I; op.J = s.J; op.k = s.k; op.flags = s.flags;
        op.slotA = op.slotB = op.slotC = -1; op.dinv_buf = -1;
        unsigned busy = 0;                             // slots this op touches
        auto locate = [&](int id) {                    // make `id` resident, return its slot
            for (int q = 0; q < nslot; ++q)
                if (slot_id[q] == id) {
                    busy |= 1u << q; slot_last[q] = i;
                    return q;
                }
            int best = -1, best_next = -1;
            for (int q = 0; q < nslot; ++q) {
                if ((busy >> q) & 1u) continue;
                const int nu = slot_id[q] < 0 ? (1 << 30) + 1 : next_use(slot_id[q], i - 1);
                if (nu > best_next) { best_next = nu; best = q; }
            }
            loads.push_back({best, tp_tri((id >> 16) & 0x3fff, id & 0xffff), i, slot_last[best] + 1, ((id >> 30) & 1) != 0,
                             slot_stored[best] != 0});
            slot_id[best] = id; slot_last[best] = i; slot_stored[best] = 0;
            op.wait_mask |= 1 << best;               // the demanding op waits; later hits find the tile landed
            busy |= 1u << best;
            return best;
        };
        op.slotA = locate(s.idA);
        if (s.idB >= 0) op.slotB = (s.idB == s.idA) ? op.slotA : locate(s.idB);
        if (s.idC >= 0) {                              // output tile: any slot not used by this op
            int best = -1, best_next = -1;
            for (int q = 0; q < nslot; ++q) {
                if ((busy >> q) & 1u) continue;
                const int nu = slot_id[q] < 0 ? (1 << 30) + 1 : next_use(slot_id[q], i);
                if (nu > best_next) { best_next = nu; best = q; }
            }
            if (slot_stored[best]) op.flags |= OPF_WAIT_READ;
            slot_id[best] = s.idC; slot_last[best] = i; slot_stored[best] = 1;   // every output is stored at the end of its op
            op.slotC = best;
        }
        if (s.dinv >= 0) {                             // 4 KB block of 8x8 inverses of L(dinv, dinv)
            int bsel = (dinv_row[0] == s.dinv) ? 0 : ((dinv_row[1] == s.dinv) ? 1 : -1);
            if (bsel < 0) {
                bsel = dinv_last[0] <= dinv_last[1] ? 0 : 1;
                loads.push_back({TP_DINV_SLOT0 + bsel, s.dinv, i, dinv_last[bsel] + 1, false, false});
                dinv_row[bsel] = s.dinv;
                op.wait_mask |= 1 << (TP_DINV_SLOT0 + bsel);
            }
            dinv_last[bsel] = i;
            op.dinv_buf = bsel;
        }
    }
    // hoist every load to the earliest operation at which its destination is free
    for (const Load& l : loads) {
        int e = std::max(l.free_from, l.needed_at - TP_PREFETCH_DEPTH);
        e = std::max(e, 0);
        while (e < l.needed_at && prog[e].nload >= TP_MAX_LOADS) ++e;   // capacity argument above: op needed_at has room
        TileOp& op = prog[e];
        op.ld_slot[op.nload] = l.slot;
        op.ld_tile[op.nload] = l.tile;
        ++op.nload;
        if (l.is_w && e < n_phase1) op.flags |= OPF_DRAIN_STORES;   // W tiles are re-read after an in-place store
        if (l.after_store) op.flags |= OPF_WAIT_READ;
    }
    if (n > n_phase1) prog[n_phase1].flags |= OPF_DRAIN_STORES;     // phase 2 starts with every W tile landed in HBM
    return prog;
}

}  // namespace fab
