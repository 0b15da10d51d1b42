// api.cu -- C ABI (include/fabolas_b200.h): context, workspaces, launch logic.
#include "../../include/fabolas_b200.h"
#include "gp_dag.cuh"
#include "gp_predict.cuh"
#include "epmgp.cuh"
#include "ig.cuh"
#include "kmat.cuh"
#include "mcmc.cuh"

#include <cstdio>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

using namespace fab;

constexpr int FAB_NEV = 8;      // events of the optional per-kernel timing: up to 7 intervals per call

struct fab_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    // problem
    bool has_data = false;
    GpData gd{};
    double* dX = nullptr; size_t capX = 0;
    double* dy = nullptr; size_t capy = 0;
    // workspaces
    GpWork wk{};
    size_t capL = 0, capZ = 0, capA = 0, capW = 0;
    double* dTheta = nullptr; size_t capT = 0;   // hyper-parameters of the resident models
    double* dYerr2 = nullptr; size_t capY = 0;
    double* hTheta = nullptr; size_t capHT = 0;  // staging of fab_gp_loglik_batch_host for pageable host memory
    double* hYerr2 = nullptr; size_t capHY = 0;
    double* hOut = nullptr; size_t capHO = 0;
    int* dInfo = nullptr; size_t capI = 0;
    bool resident = false;  // factorize_batch completed: W, alpha, theta resident
    // Entropy-Search state (fab_es_setup) and scratch for information-gain sweeps
    EsState es{};
    bool es_ready = false;
    size_t cap_es[7] = {0, 0, 0, 0, 0, 0, 0};
    double* sc_mu = nullptr; double* sc_var = nullptr; double* sc_dot = nullptr; double* sc_c0 = nullptr;
    double* sc_part = nullptr; double* sc_vmix = nullptr; double* sc_mmix = nullptr;
    size_t cap_sc[7] = {0, 0, 0, 0, 0, 0, 0};
    int wsB = 0;            // walkers the factor workspace currently describes
    bool factored = false;
    bool inverted = false;  // workspace holds W = L^-1 (and alpha) instead of L
    // fused warp-specialised kernel (gp_dag.cuh): cached static programs per (NT, slots, mode)
    struct DagDev { DagOp* bulk = nullptr; DagOp* chain = nullptr; int nbulk = 0, nchain = 0; DagProgram host; };
    std::map<long long, DagDev> dag;
    double* wkSpill = nullptr; size_t capSpill = 0;
    int smem_optin = 232448;
    int max_slots = MAX_SLOTS;   // test knob: fewer shared-memory tile slots => exercises the streaming path
    int force_slices = 0;        // test knob: never keep all coordinates resident in the gradient kernel
    int mcmc_no_cluster = getenv("FAB_MCMC_NO_CLUSTER") != nullptr;   // development: A/B of the cluster launch of small ensembles
    // peer exchange of the likelihood step (gp_dag.cuh PeerTable): one IPC-exported allocation per rank holding two
    // gather buffers (step parity) + arrival counter + error flag; the peers' mappings
    struct Peer {
        int world = 0, rank = 0, rows_per_rank = 0, row_doubles = 0;
        void* base[FAB_MAX_PEERS] = {nullptr};      // base[rank] is the local allocation, the others are IPC mappings
        size_t buf_bytes = 0;
        bool connected = false;
        unsigned long long steps = 0;
    } peer;
    // sharded device sampler (mcmc.cuh): one IPC-exported allocation per rank holding the replica of coords (K x P)
    // and lp (K), the barrier flag block and the error flag; the peers' mappings
    struct McPeer {
        int world = 0, rank = 0, K = 0, P = 0;
        void* base[FAB_MAX_PEERS] = {nullptr};
        bool connected = false;
        unsigned xgen = 0;
        size_t off_lp = 0, off_flag = 0, off_err = 0, bytes = 0;
    } mpeer;
    // device-resident ensemble sampler (mcmc.cuh): per-CTA scratch records
    double* mcTheta = nullptr; double* mcYerr2 = nullptr; double* mcLl = nullptr; double* mcQ = nullptr;
    int* mcInfo = nullptr; unsigned long long* mcKeys = nullptr; int* mcOrder = nullptr; unsigned* mcBar = nullptr;
    size_t cap_mc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int num_sms = 148;
    double* eiVal = nullptr; long long* eiIdx = nullptr; size_t cap_ei[2] = {0, 0};   // per-chunk argmax partials
    double* vpart = nullptr; size_t capVpart = 0;     // prediction with a chunked panel: partial rows of V per CTA
    int predict_jc = 0;          // test knob: cap on the resident tile rows of the cross-covariance panel (0 = whatever fits)
    int mcmc_max_ctas = 0;       // test knob: cap the sampler's grid (0 = one CTA per proposal, up to the SM count)
    // optional per-kernel timing (CUDA events on the launching stream)
    bool timing = false;
    cudaEvent_t ev[FAB_NEV] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int ev_used = 0;        // number of intervals recorded by the last call
};

namespace {

int fail(fab_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess) {
    if (c) {
        c->err = what;
        if (e != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(e); }
    }
    return code;
}

#define FAB_CUDA(c, call)                                                      \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) return fail((c), FAB_ECUDA, #call, e_);         \
    } while (0)

// optional per-kernel timing: mark(c, i) records event i on the stream (i = 0 before the first launch of a call,
// i = k after the k-th); the call ends with marks_done(c, k)
#define FAB_MARK(c, i)                                                         \
    do {                                                                       \
        if ((c)->timing) FAB_CUDA((c), cudaEventRecord((c)->ev[(i)], (c)->stream)); \
    } while (0)

template <typename T>
int ensure(fab_ctx* c, T** p, size_t* cap, size_t need) {
    if (*cap >= need && *p) return FAB_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    if (cudaMalloc(reinterpret_cast<void**>(p), need * sizeof(T)) != cudaSuccess) {
        cudaGetLastError();
        return fail(c, FAB_ENOMEM, "workspace allocation failed");
    }
    *cap = need;
    return FAB_OK;
}

int pk_of(const GpData& gd) { return gd.family == FAB_FAMILY_M52_BASIS ? gd.d + 3 : gd.d + 1; }

__global__ void untile_kernel(const double* __restrict__ Lw, double* __restrict__ out, int N, int NT) {
    const int i = blockIdx.x;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        double v = 0.0;
        if (j <= i) {
            const int I = i / TS, J = j / TS;
            v = Lw[(size_t)tri_index(I, J) * TILE_ELEMS + tix(i % TS, j % TS)];
        }
        out[(size_t)i * N + j] = v;
    }
}

void peer_release(fab_ctx* c) {
#ifndef FAB_CUSIM
    for (int r = 0; r < c->peer.world; ++r)
        if (r != c->peer.rank && c->peer.base[r]) cudaIpcCloseMemHandle(c->peer.base[r]);
#endif
    if (c->peer.world > 0 && c->peer.base[c->peer.rank]) cudaFree(c->peer.base[c->peer.rank]);
    c->peer = fab_ctx::Peer();
}
void mpeer_release(fab_ctx* c) {
#ifndef FAB_CUSIM
    for (int r = 0; r < c->mpeer.world; ++r)
        if (r != c->mpeer.rank && c->mpeer.base[r]) cudaIpcCloseMemHandle(c->mpeer.base[r]);
#endif
    if (c->mpeer.world > 0 && c->mpeer.base[c->mpeer.rank]) cudaFree(c->mpeer.base[c->mpeer.rank]);
    c->mpeer = fab_ctx::McPeer();
}
constexpr size_t ES_ENTROPY_SMEM_OPTIN = 100 * 1024;
inline size_t peer_ctrl_offset(const fab_ctx::Peer& p) { return 2 * p.buf_bytes; }     // +0 arrival slots (one u32 per source rank), +64 local CTAs done, +128 error flag

}  // namespace

extern "C" {

int fab_version(void) { return 100; }

int fab_ctx_create(fab_ctx** out, int device, void* cuda_stream) {
    if (!out) return FAB_EINVAL;
    fab_ctx* c = new (std::nothrow) fab_ctx();
    if (!c) return FAB_ENOMEM;
    c->device = device;
    c->stream = static_cast<cudaStream_t>(cuda_stream);
    if (cudaSetDevice(device) != cudaSuccess) { delete c; return FAB_ECUDA; }
#ifndef FAB_CUSIM
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) == cudaSuccess && v > 0)
        c->smem_optin = v;
    for (int d = 1; d <= MAX_D; ++d)
        FAB_DISPATCH_D(d, cudaFuncSetAttribute(gp_dag_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin);
                          cudaFuncSetAttribute(mcmc_stretch_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               c->smem_optin - MCMC_STATIC_SMEM));
    cudaFuncSetAttribute(gp_predict_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin);
    cudaFuncSetAttribute(epmgp_ep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(epmgp_normalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(es_entropy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ES_ENTROPY_SMEM_OPTIN);
    cudaFuncSetAttribute(gp_predict_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin);
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) c->num_sms = v;
#endif
    *out = c;
    return FAB_OK;
}

int fab_ctx_destroy(fab_ctx* c) {
    if (!c) return FAB_EINVAL;
    cudaSetDevice(c->device);
    cudaFree(c->dX); cudaFree(c->dy); cudaFree(c->dTheta); cudaFree(c->dYerr2); cudaFree(c->dInfo);
    cudaFree(c->hTheta); cudaFree(c->hYerr2); cudaFree(c->hOut);
    cudaFree(c->es.rl); cudaFree(c->es.vrl); cudaFree(c->es.coef); cudaFree(c->es.logP); cudaFree(c->es.logU);
    cudaFree(c->es.base); cudaFree(c->es.omega);
    cudaFree(c->sc_mu); cudaFree(c->sc_var); cudaFree(c->sc_dot); cudaFree(c->sc_c0); cudaFree(c->sc_part);
    cudaFree(c->sc_vmix); cudaFree(c->sc_mmix);
#ifndef FAB_CUSIM
    for (int i = 0; i < FAB_NEV; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
#endif
    for (auto& kv : c->dag) { cudaFree(kv.second.bulk); cudaFree(kv.second.chain); }
    cudaFree(c->wkSpill);
    peer_release(c);
    mpeer_release(c);
    cudaFree(c->eiVal); cudaFree(c->eiIdx); cudaFree(c->vpart);
    cudaFree(c->mcTheta); cudaFree(c->mcYerr2); cudaFree(c->mcLl); cudaFree(c->mcQ); cudaFree(c->mcInfo);
    cudaFree(c->mcKeys); cudaFree(c->mcOrder); cudaFree(c->mcBar);
    cudaFree(c->wk.Lw); cudaFree(c->wk.zvec); cudaFree(c->wk.alpha); cudaFree(c->wk.accw);
    delete c;
    return FAB_OK;
}

int fab_ctx_debug_set_max_slots(fab_ctx* c, int n) {
    if (!c || n < 4 || n > MAX_SLOTS) return FAB_EINVAL;
    c->max_slots = n;
    return FAB_OK;
}

int fab_debug_dag_verify(int NT, int nslot, int mode, char* msg, int msg_len, int* stats) {
    if (mode < DAG_LL || mode > DAG_GRAD) return FAB_EINVAL;
    const DagProgram P = build_dag_program(NT, nslot, mode);
    const std::string e = dag_verify(P);
    if (msg && msg_len > 0) { strncpy(msg, e.c_str(), msg_len - 1); msg[msg_len - 1] = 0; }
    if (stats) {
        int nl = 0, ns = 0;
        for (const DagOp& d : P.bulk) { nl += d.nload; ns += d.st_slot >= 0; }
        stats[0] = (int)P.bulk.size(); stats[1] = (int)P.chain.size(); stats[2] = nl; stats[3] = ns;
    }
    return e.empty() ? 0 : -1;
}

int fab_ctx_debug_set_predict_chunk(fab_ctx* c, int jc) {
    if (!c || jc < 0) return FAB_EINVAL;
    c->predict_jc = jc;
    return FAB_OK;
}

int fab_ctx_debug_force_slices(fab_ctx* c, int on) {
    if (!c) return FAB_EINVAL;
    c->force_slices = on != 0;
    return FAB_OK;
}

int fab_ctx_set_timing(fab_ctx* c, int enable) {
    if (!c) return FAB_EINVAL;
#ifndef FAB_CUSIM
    if (enable && !c->ev[0])
        for (int i = 0; i < FAB_NEV; ++i) FAB_CUDA(c, cudaEventCreate(&c->ev[i]));
#endif
    c->timing = enable != 0;
    c->ev_used = 0;
    return FAB_OK;
}

int fab_ctx_get_timing(fab_ctx* c, double* ms_out, int n) {
    if (!c || !ms_out) return FAB_EINVAL;
    for (int i = 0; i < n; ++i) ms_out[i] = 0.0;
#ifndef FAB_CUSIM
    if (!c->timing || c->ev_used == 0) return fail(c, FAB_ESTATE, "no timed launch recorded");
    FAB_CUDA(c, cudaEventSynchronize(c->ev[c->ev_used]));
    for (int i = 0; i < c->ev_used && i < n; ++i) {
        float ms = 0.f;
        FAB_CUDA(c, cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]));
        ms_out[i] = ms;
    }
#endif
    return c->ev_used;
}

const char* fab_last_error(const fab_ctx* c) { return c ? c->err.c_str() : "null context"; }

int fab_gp_set_data(fab_ctx* c, int family, const double* X, int N, int D, const double* y, double mean,
                    double amp_mult) {
    if (!c || !X || !y) return fail(c, FAB_EINVAL, "null pointer");
    if (family != FAB_FAMILY_M52 && family != FAB_FAMILY_M52_BASIS) return fail(c, FAB_EINVAL, "unknown family");
    const int d = family == FAB_FAMILY_M52_BASIS ? D - 1 : D;
    if (N < 1 || d < 1 || d > MAX_D) return fail(c, FAB_EINVAL, "need N >= 1 and 1 <= d <= 8 Matern axes");
    FAB_CUDA(c, cudaSetDevice(c->device));
    int rc;
    if ((rc = ensure(c, &c->dX, &c->capX, (size_t)N * D)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->dy, &c->capy, (size_t)N)) != FAB_OK) return rc;
    FAB_CUDA(c, cudaMemcpyAsync(c->dX, X, sizeof(double) * N * D, cudaMemcpyDeviceToDevice, c->stream));
    FAB_CUDA(c, cudaMemcpyAsync(c->dy, y, sizeof(double) * N, cudaMemcpyDeviceToDevice, c->stream));
    GpData& gd = c->gd;
    gd.X = c->dX; gd.y = c->dy; gd.mean = mean; gd.amp_mult = amp_mult;
    gd.N = N; gd.D = D; gd.d = d; gd.family = family;
    gd.NT = (N + TS - 1) / TS; gd.Np = gd.NT * TS;
    gd.Pk = pk_of(gd);
    c->has_data = true;
    c->factored = false;
    c->resident = false;
    return FAB_OK;
}

// Everything one launch of the fused evaluation needs for B walker records: the static tile program for
// (NT, slots, mode) -- built and uploaded once --, the per-walker workspaces and the shared-memory size.
// mode: DAG_LL (ll only), DAG_FACTOR (+ L tiles in the workspace), DAG_FIT (+ W = L^-1 tiles and alpha in the
// workspace), DAG_GRAD (+ gradient).  `smem_limit`: dynamic shared memory the launching kernel may use.
static int dag_setup(fab_ctx* c, int B, int mode, size_t smem_limit, DagArgs* out, size_t* smem_out) {
    const GpData& gd = c->gd;
    const int ntiles = gd.NT * (gd.NT + 1) / 2;
    int nslot = c->max_slots;
    int resident = (!c->force_slices && dag_smem_bytes(nslot, gd.d, gd.Np) <= smem_limit) ? 1 : 0;
    while (nslot > 0 && dag_smem_bytes(nslot, gd.d, resident ? gd.Np : 0) > smem_limit) --nslot;
    if (nslot < 4) return fail(c, FAB_EUNSUPPORTED, "not enough shared memory for the fused kernel");
    const long long key = ((long long)gd.NT << 16) | (nslot << 4) | mode;
    auto it = c->dag.find(key);
    if (it == c->dag.end()) {
        fab_ctx::DagDev dv;
        dv.host = build_dag_program(gd.NT, nslot, mode);
        if (!dv.host.ok) return fail(c, FAB_EUNSUPPORTED, dv.host.err.c_str());
        dv.nbulk = (int)dv.host.bulk.size(); dv.nchain = (int)dv.host.chain.size();
        if (cudaMalloc(reinterpret_cast<void**>(&dv.bulk), (dv.nbulk + 1) * sizeof(DagOp)) != cudaSuccess ||
            cudaMalloc(reinterpret_cast<void**>(&dv.chain), (dv.nchain + 1) * sizeof(DagOp)) != cudaSuccess) {
            cudaGetLastError();
            cudaFree(dv.bulk);                        // null when the first allocation is the one that failed
            return fail(c, FAB_ENOMEM, "tile program allocation failed");
        }
        it = c->dag.emplace(key, std::move(dv)).first;          // host vectors stay alive: async copy sources
        FAB_CUDA(c, cudaMemcpyAsync(it->second.bulk, it->second.host.bulk.data(), it->second.nbulk * sizeof(DagOp),
                                    cudaMemcpyHostToDevice, c->stream));
        FAB_CUDA(c, cudaMemcpyAsync(it->second.chain, it->second.host.chain.data(), it->second.nchain * sizeof(DagOp),
                                    cudaMemcpyHostToDevice, c->stream));
    }
    const fab_ctx::DagDev& dv = it->second;
    int rc;
    if ((rc = ensure(c, &c->wk.Lw, &c->capL, (size_t)B * ntiles * TILE_ELEMS)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->wk.accw, &c->capW, (size_t)B * gd.Np)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->wk.zvec, &c->capZ, (size_t)B * gd.Np)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->wk.alpha, &c->capA, (size_t)B * gd.Np)) != FAB_OK) return rc;
    if (mode == DAG_FACTOR && (rc = ensure(c, &c->wkSpill, &c->capSpill, (size_t)B * gd.NT * TILE_ELEMS)) != FAB_OK) return rc;
    DagArgs a;
    a.gd = gd; a.wk = c->wk; a.theta = nullptr; a.yerr2 = nullptr; a.ll = nullptr; a.info = nullptr; a.grad = nullptr;
    a.bulk = dv.bulk; a.chain = dv.chain; a.nbulk = dv.nbulk; a.nchain = dv.nchain;
    a.nslot = nslot; a.mode = mode; a.resident = resident; a.ntiles = ntiles; a.wk_spill = c->wkSpill;
    a.peer.world = 0; a.host_io = 0;
    *out = a;
    *smem_out = dag_smem_bytes(nslot, gd.d, resident ? gd.Np : 0);
    c->resident = false;
    c->es_ready = false;
    c->wsB = B;
    c->factored = mode == DAG_FACTOR;
    c->inverted = mode == DAG_FIT;
    return FAB_OK;
}

// One launch of the fused kernel (one CTA per walker).
static int launch_dag(fab_ctx* c, const double* theta, const double* yerr2, int B, double* ll, int* info, double* grad,
                      int mode, bool exchange = false, bool host_io = false) {
    DagArgs a;
    size_t smem = 0;
    const int rc = dag_setup(c, B, mode, (size_t)c->smem_optin, &a, &smem);
    if (rc != FAB_OK) return rc;
    a.theta = theta; a.yerr2 = yerr2; a.ll = ll; a.info = info; a.grad = grad; a.host_io = host_io ? 1 : 0;
    if (exchange) {
        const fab_ctx::Peer& p = c->peer;
        const size_t parity = (size_t)(p.steps & 1ull) * p.buf_bytes;
        for (int r = 0; r < p.world; ++r) {
            a.peer.buf[r] = reinterpret_cast<double*>(static_cast<char*>(p.base[r]) + parity);
            a.peer.slots[r] = reinterpret_cast<unsigned*>(static_cast<char*>(p.base[r]) + peer_ctrl_offset(p));
        }
        a.peer.done = reinterpret_cast<unsigned*>(static_cast<char*>(p.base[p.rank]) + peer_ctrl_offset(p) + 64);
        a.peer.err = reinterpret_cast<int*>(static_cast<char*>(p.base[p.rank]) + peer_ctrl_offset(p) + 128);
        a.peer.done_target = (unsigned)((p.steps + 1) * (unsigned long long)B);
        a.peer.step = (unsigned)(p.steps + 1);
        a.peer.world = p.world; a.peer.rank = p.rank; a.peer.rows_per_rank = p.rows_per_rank; a.peer.row_doubles = p.row_doubles;
    }
    if (c->timing) { FAB_CUDA(c, cudaEventRecord(c->ev[0], c->stream)); }
    FAB_DISPATCH_D(c->gd.d, FAB_LAUNCH(gp_dag_kernel<DD>, B, DAG_THREADS, smem, c->stream, a));
    FAB_CUDA(c, cudaGetLastError());
    if (c->timing) { FAB_CUDA(c, cudaEventRecord(c->ev[1], c->stream)); c->ev_used = 1; }
    return FAB_OK;
}

int fab_gp_loglik_batch(fab_ctx* c, const double* theta, const double* yerr2, int B, double* ll, double* grad,
                        int* info) {
    if (!c) return FAB_EINVAL;
    if (!c->has_data) return fail(c, FAB_ESTATE, "fab_gp_set_data has not been called");
    if (!theta || !yerr2 || !ll || !info || B < 1) return fail(c, FAB_EINVAL, "bad argument");
    FAB_CUDA(c, cudaSetDevice(c->device));
    return launch_dag(c, theta, yerr2, B, ll, info, grad, grad ? DAG_GRAD : DAG_LL);
}

// Device alias of a host pointer the CUDA driver knows (cudaHostAlloc / cudaHostRegister / torch pin_memory): the fused
// kernel reads and writes such memory directly.  nullptr for pageable memory.
static void* mapped_alias(const void* p) {
#ifdef FAB_CUSIM
    return const_cast<void*>(p);                       // (the emulator's "device" memory is host memory)
#else
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
    return at.devicePointer;
#endif
}

int fab_gp_loglik_batch_host(fab_ctx* c, const double* theta, const double* yerr2, int B, double* ll, double* grad,
                             int* info) {
    if (!c) return FAB_EINVAL;
    if (!c->has_data) return fail(c, FAB_ESTATE, "fab_gp_set_data has not been called");
    if (!theta || !yerr2 || !ll || !info || B < 1) return fail(c, FAB_EINVAL, "bad argument");
    FAB_CUDA(c, cudaSetDevice(c->device));
    const int Pk = c->gd.Pk;
    const double* d_theta = static_cast<const double*>(mapped_alias(theta));
    const double* d_yerr2 = static_cast<const double*>(mapped_alias(yerr2));
    double* d_ll = static_cast<double*>(mapped_alias(ll));
    int* d_info = static_cast<int*>(mapped_alias(info));
    double* d_grad = grad ? static_cast<double*>(mapped_alias(grad)) : nullptr;
    if (d_theta && d_yerr2 && d_ll && d_info && (!grad || d_grad))       // zero-copy: the kernel does the transfers itself
        return launch_dag(c, d_theta, d_yerr2, B, d_ll, d_info, d_grad, grad ? DAG_GRAD : DAG_LL, false, true);
    // pageable host memory: staged through buffers of the context (the copies are stream-ordered)
    int rc;
    if ((rc = ensure(c, &c->hTheta, &c->capHT, (size_t)B * Pk)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->hYerr2, &c->capHY, (size_t)B)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->hOut, &c->capHO, (size_t)B * (Pk + 3))) != FAB_OK) return rc;
    double* o_ll = c->hOut; double* o_grad = c->hOut + B; int* o_info = reinterpret_cast<int*>(c->hOut + (size_t)B * (Pk + 2));
    FAB_CUDA(c, cudaMemcpyAsync(c->hTheta, theta, sizeof(double) * B * Pk, cudaMemcpyHostToDevice, c->stream));
    FAB_CUDA(c, cudaMemcpyAsync(c->hYerr2, yerr2, sizeof(double) * B, cudaMemcpyHostToDevice, c->stream));
    rc = launch_dag(c, c->hTheta, c->hYerr2, B, o_ll, o_info, grad ? o_grad : nullptr, grad ? DAG_GRAD : DAG_LL);
    if (rc != FAB_OK) return rc;
    FAB_CUDA(c, cudaMemcpyAsync(ll, o_ll, sizeof(double) * B, cudaMemcpyDeviceToHost, c->stream));
    if (grad) FAB_CUDA(c, cudaMemcpyAsync(grad, o_grad, sizeof(double) * B * (Pk + 1), cudaMemcpyDeviceToHost, c->stream));
    FAB_CUDA(c, cudaMemcpyAsync(info, o_info, sizeof(int) * B, cudaMemcpyDeviceToHost, c->stream));
    return FAB_OK;
}

int fab_ctx_synchronize(fab_ctx* c) {
    if (!c) return FAB_EINVAL;
    FAB_CUDA(c, cudaSetDevice(c->device));
    FAB_CUDA(c, cudaStreamSynchronize(c->stream));
    return FAB_OK;
}

int fab_gp_factorize_batch(fab_ctx* c, const double* theta, const double* yerr2, int B, double* ll, int* info) {
    if (!c) return FAB_EINVAL;
    if (!c->has_data) return fail(c, FAB_ESTATE, "fab_gp_set_data has not been called");
    if (!theta || !yerr2 || !ll || !info || B < 1) return fail(c, FAB_EINVAL, "bad argument");
    FAB_CUDA(c, cudaSetDevice(c->device));
    return launch_dag(c, theta, yerr2, B, ll, info, nullptr, DAG_FACTOR);
}

// factorize + W = L^-1 + alpha, hyper-parameters copied into the context: models become resident.
int fab_gp_fit_batch(fab_ctx* c, const double* theta, const double* yerr2, int B, double* ll, int* info) {
    if (!c) return FAB_EINVAL;
    if (!c->has_data) return fail(c, FAB_ESTATE, "fab_gp_set_data has not been called");
    if (!theta || !yerr2 || !ll || !info || B < 1) return fail(c, FAB_EINVAL, "bad argument");
    FAB_CUDA(c, cudaSetDevice(c->device));
    int rc = launch_dag(c, theta, yerr2, B, ll, info, nullptr, DAG_FIT);
    if (rc != FAB_OK) return rc;
    const GpData& gd = c->gd;
    if ((rc = ensure(c, &c->dTheta, &c->capT, (size_t)B * gd.Pk)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->dYerr2, &c->capY, (size_t)B)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->dInfo, &c->capI, (size_t)B)) != FAB_OK) return rc;
    FAB_CUDA(c, cudaMemcpyAsync(c->dTheta, theta, sizeof(double) * B * gd.Pk, cudaMemcpyDeviceToDevice, c->stream));
    FAB_CUDA(c, cudaMemcpyAsync(c->dYerr2, yerr2, sizeof(double) * B, cudaMemcpyDeviceToDevice, c->stream));
    FAB_CUDA(c, cudaMemcpyAsync(c->dInfo, info, sizeof(int) * B, cudaMemcpyDeviceToDevice, c->stream));
    c->resident = true;
    return FAB_OK;
}

static int launch_predict(fab_ctx* c, const double* T, long long T_bstride, int M, double* mu, double* var,
                          double* cov, const double* dotvec, double* dot, double* vcol = nullptr, int vcol_index = 0) {
    const GpData& gd = c->gd;
    PredictArgs a;
    a.gd = gd; a.wk = c->wk; a.theta = c->dTheta; a.yerr2 = c->dYerr2; a.T = T; a.T_bstride = T_bstride; a.M = M;
    a.mu = mu; a.var = var; a.cov = cov; a.dotvec = dotvec; a.dot = dot; a.vcol = vcol; a.vcol_index = vcol_index;
    const bool want_cov = cov != nullptr;
    if (want_cov && M > 64) return fail(c, FAB_EUNSUPPORTED, "full predictive covariance is limited to M <= 64 points");
    // widest candidate block whose whole cross-covariance panel (N x CB) fits, with two W tile buffers when they fit as
    // well; otherwise (N > 512) the panel is generated jc tile rows at a time and the rows of V that later chunks
    // continue wait in a per-CTA scratch region -- any N
    int CB = 0, nbuf = 0, jc = gd.NT;
    const int jc_cap = c->predict_jc > 0 ? c->predict_jc : gd.NT;          // test knob
    if (jc_cap >= gd.NT)
        for (int cb : {64, 32}) {
            if (cb == 32 && want_cov) break;
            for (int nb : {2, 1})
                if (!CB && predict_smem_bytes(cb, gd.NT, gd.d, want_cov, nb) <= (size_t)c->smem_optin) { CB = cb; nbuf = nb; }
            if (CB) break;
        }
    if (!CB) {
        CB = want_cov ? 64 : 32; nbuf = 2;
        jc = jc_cap < gd.NT ? jc_cap : gd.NT;
        while (jc > 1 && predict_smem_bytes(CB, jc, gd.d, want_cov, nbuf) > (size_t)c->smem_optin) --jc;
        if (predict_smem_bytes(CB, jc, gd.d, want_cov, nbuf) > (size_t)c->smem_optin)
            return fail(c, FAB_EUNSUPPORTED, "not enough shared memory for the prediction kernel");
    }
    a.nbuf = nbuf; a.jc = jc; a.B = c->wsB; a.vpart = nullptr;
    const long long nblk = ((long long)M + CB - 1) / CB;
    dim3 grid((unsigned)nblk, (unsigned)c->wsB);
    if (jc < gd.NT) {                       // persistent grid: one scratch region per CTA
        const long long items = nblk * c->wsB;
        const long long cap = c->predict_jc > 0 ? 2 : c->num_sms;      // (test knob: two CTAs walk all the items)
        const unsigned G = (unsigned)(items < cap ? items : cap);
        grid = dim3(G, 1);
        int rc = ensure(c, &c->vpart, &c->capVpart, (size_t)G * gd.NT * TS * CB);
        if (rc != FAB_OK) return rc;
        a.vpart = c->vpart;
    }
    const size_t smem = predict_smem_bytes(CB, jc, gd.d, want_cov, nbuf);
    if (CB == 64) {
        FAB_LAUNCH(gp_predict_kernel<64>, grid, NTHREADS, smem, c->stream, a);
    } else {
        FAB_LAUNCH(gp_predict_kernel<32>, grid, NTHREADS, smem, c->stream, a);
    }
    FAB_CUDA(c, cudaGetLastError());
    return FAB_OK;
}

int fab_gp_predict_batch(fab_ctx* c, const double* T, int per_model, int M, double* mu, double* var, double* cov) {
    if (!c) return FAB_EINVAL;
    if (!c->resident) return fail(c, FAB_ESTATE, "fab_gp_fit_batch has not been called");
    if (!T || M < 1 || (!mu && !var && !cov)) return fail(c, FAB_EINVAL, "bad argument");
    FAB_CUDA(c, cudaSetDevice(c->device));
    FAB_MARK(c, 0);
    const int rc = launch_predict(c, T, per_model ? (long long)M * c->gd.D : 0, M, mu, var, cov, nullptr, nullptr);
    if (rc != FAB_OK) return rc;
    FAB_MARK(c, 1);
    if (c->timing) c->ev_used = 1;
    return FAB_OK;
}

int fab_epmgp_joint_min_batch(fab_ctx* c, const double* mu, const double* Sigma, int B, int Z, double* logP,
                              double* dMu, double* dSigma, double* dMuMu, int* info, int* sweeps) {
    if (!c) return FAB_EINVAL;
    if (!mu || !Sigma || !logP || !info || B < 1) return fail(c, FAB_EINVAL, "bad argument");
    if (Z < 1 || Z > EP_MAXZ) return fail(c, FAB_EUNSUPPORTED, "EPMGP supports 1 <= Z <= 64 representer points");
    const int wd = (dMu && dSigma && dMuMu) ? 1 : 0;
    if (!wd && (dMu || dSigma || dMuMu)) return fail(c, FAB_EINVAL, "pass all three derivative buffers or none");
    FAB_CUDA(c, cudaSetDevice(c->device));
    FAB_CUDA(c, cudaMemsetAsync(info, 0, sizeof(int) * B, c->stream));
    EpArgs a;
    a.mu = mu; a.Sigma = Sigma; a.Z = Z; a.logP = logP; a.dMu = dMu; a.dSigma = dSigma; a.dMuMu = dMuMu;
    a.info = info; a.sweeps = sweeps; a.with_derivatives = wd;
    FAB_MARK(c, 0);
    FAB_LAUNCH(epmgp_ep_kernel, dim3(Z, B), EP_THREADS, ep_smem_bytes(Z), c->stream, a);
    FAB_CUDA(c, cudaGetLastError());
    FAB_MARK(c, 1);
    FAB_LAUNCH(epmgp_normalize_kernel, B, 256, ep_norm_smem_bytes(Z), c->stream, a);
    FAB_CUDA(c, cudaGetLastError());
    FAB_MARK(c, 2);
    if (c->timing) c->ev_used = 2;          // [epmgp_ep_kernel, epmgp_normalize_kernel]
    return FAB_OK;
}

namespace {
__global__ void gather_last_rep_kernel(const double* reps, int Z, int D, double* rl) {
    const int b = blockIdx.x;
    for (int k = threadIdx.x; k < D; k += blockDim.x) rl[(size_t)b * D + k] = reps[((size_t)b * Z + (Z - 1)) * D + k];
}
}  // namespace

int fab_es_setup(fab_ctx* c, const double* reps, int Z, const double* cov_reps, const double* logP, const double* dMu,
                 const double* dSigma, const double* dMuMu, const double* U, const double* Omega, int I) {
    if (!c) return FAB_EINVAL;
    if (!c->resident) return fail(c, FAB_ESTATE, "fab_gp_fit_batch has not been called");
    if (!reps || !cov_reps || !logP || !dMu || !dSigma || !dMuMu || !U || !Omega) return fail(c, FAB_EINVAL, "null pointer");
    if (Z < 1 || Z > EP_MAXZ || I < 1) return fail(c, FAB_EUNSUPPORTED, "need 1 <= Z <= 64 representers and I >= 1 innovations");
    if (es_entropy_smem_bytes(Z, I) > ES_ENTROPY_SMEM_OPTIN)
        return fail(c, FAB_EUNSUPPORTED, "too many innovations: the innovation vectors of one model (I x Z) must fit the 100 KB of shared memory of the entropy kernel");
    FAB_CUDA(c, cudaSetDevice(c->device));
    const GpData& gd = c->gd;
    const int B = c->wsB;
    EsState& st = c->es;
    int rc;
    if ((rc = ensure(c, &st.rl, &c->cap_es[0], (size_t)B * gd.D)) != FAB_OK) return rc;
    if ((rc = ensure(c, &st.vrl, &c->cap_es[1], (size_t)B * gd.Np)) != FAB_OK) return rc;
    if ((rc = ensure(c, &st.coef, &c->cap_es[2], (size_t)B * 8 * Z)) != FAB_OK) return rc;
    if ((rc = ensure(c, &st.logP, &c->cap_es[3], (size_t)B * Z)) != FAB_OK) return rc;
    if ((rc = ensure(c, &st.logU, &c->cap_es[4], (size_t)B * Z)) != FAB_OK) return rc;
    if ((rc = ensure(c, &st.base, &c->cap_es[5], (size_t)B)) != FAB_OK) return rc;
    if ((rc = ensure(c, &st.omega, &c->cap_es[6], (size_t)B * I * Z)) != FAB_OK) return rc;
    st.Z = Z; st.I = I;
    FAB_CUDA(c, cudaMemcpyAsync(st.omega, Omega, sizeof(double) * B * I * Z, cudaMemcpyDeviceToDevice, c->stream));
    FAB_LAUNCH(gather_last_rep_kernel, B, 32, 0, c->stream, reps, Z, gd.D, st.rl);
    FAB_CUDA(c, cudaGetLastError());
    // vrl_b = W_b k_b(X, rep_last_b): one-point per-model prediction, keeping column 0 of V
    FAB_CUDA(c, cudaMemsetAsync(st.vrl, 0, sizeof(double) * B * gd.Np, c->stream));
    rc = launch_predict(c, st.rl, gd.D, 1, nullptr, nullptr, nullptr, nullptr, nullptr, st.vrl, 0);
    if (rc != FAB_OK) return rc;
    EsSetupArgs a;
    a.st = st; a.cov_reps = cov_reps; a.logP = logP; a.dMu = dMu; a.dSigma = dSigma; a.dMuMu = dMuMu; a.U = U;
    FAB_LAUNCH(es_setup_kernel, B, ES_THREADS, 0, c->stream, a);
    FAB_CUDA(c, cudaGetLastError());
    c->es_ready = true;
    return FAB_OK;
}

int fab_es_information_gain_batch(fab_ctx* c, const double* Xc, int M, double* ig, int* nan_flag, double* mix_mean,
                                  double* mix_var) {
    if (!c) return FAB_EINVAL;
    if (!c->es_ready) return fail(c, FAB_ESTATE, "fab_es_setup has not been called");
    if (!Xc || !ig || M < 1) return fail(c, FAB_EINVAL, "bad argument");
    FAB_CUDA(c, cudaSetDevice(c->device));
    const GpData& gd = c->gd;
    const int B = c->wsB;
    int rc;
    if ((rc = ensure(c, &c->sc_mu, &c->cap_sc[0], (size_t)B * M)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->sc_var, &c->cap_sc[1], (size_t)B * M)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->sc_dot, &c->cap_sc[2], (size_t)B * M)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->sc_c0, &c->cap_sc[3], (size_t)B * M)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->sc_part, &c->cap_sc[4], (size_t)B * M)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->sc_vmix, &c->cap_sc[5], (size_t)M)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->sc_mmix, &c->cap_sc[6], (size_t)M)) != FAB_OK) return rc;
    FAB_MARK(c, 0);
    rc = launch_predict(c, Xc, 0, M, c->sc_mu, c->sc_var, nullptr, c->es.vrl, c->sc_dot);
    if (rc != FAB_OK) return rc;
    FAB_MARK(c, 1);
    EsMixArgs m;
    m.gd = gd; m.theta = c->dTheta; m.yerr2 = c->dYerr2; m.st = c->es; m.Xc = Xc; m.M = M; m.B = B;
    m.mu = c->sc_mu; m.var = c->sc_var; m.dot = c->sc_dot; m.vmix = c->sc_vmix; m.mmix = c->sc_mmix; m.c0 = c->sc_c0;
    FAB_LAUNCH(es_mix_kernel, M, ES_THREADS, 0, c->stream, m);
    FAB_CUDA(c, cudaGetLastError());
    FAB_MARK(c, 2);
    EsEntropyArgs e;
    e.st = c->es; e.M = M; e.B = B; e.vmix = c->sc_vmix; e.c0 = c->sc_c0; e.partial = c->sc_part;
    FAB_LAUNCH(es_entropy_kernel, dim3((M + ES_XPB - 1) / ES_XPB, B), ES_THREADS,
               es_entropy_smem_bytes(c->es.Z, c->es.I), c->stream, e);
    FAB_CUDA(c, cudaGetLastError());
    FAB_MARK(c, 3);
    FAB_LAUNCH(es_reduce_kernel, (M + 127) / 128, 128, 0, c->stream, c->sc_part, M, B, ig, nan_flag);
    FAB_CUDA(c, cudaGetLastError());
    FAB_MARK(c, 4);
    if (c->timing) c->ev_used = 4;          // [gp_predict_kernel, es_mix_kernel, es_entropy_kernel, es_reduce_kernel]
    if (mix_mean) FAB_CUDA(c, cudaMemcpyAsync(mix_mean, c->sc_mmix, sizeof(double) * M, cudaMemcpyDeviceToDevice, c->stream));
    if (mix_var) FAB_CUDA(c, cudaMemcpyAsync(mix_var, c->sc_vmix, sizeof(double) * M, cudaMemcpyDeviceToDevice, c->stream));
    return FAB_OK;
}

int fab_kernel_matrix(fab_ctx* c, int family, int nu_code, double nu, const double* theta, int B, const double* Xa, int Na,
                      const double* Xb, int Nb, int D, double amp_mult, double* K, double* dK) {
    if (!c) return FAB_EINVAL;
    if (!theta || !Xa || !Xb || !K || B < 1 || Na < 1 || Nb < 1) return fail(c, FAB_EINVAL, "bad argument");
    if (family != FAB_FAMILY_M52 && family != FAB_FAMILY_M52_BASIS) return fail(c, FAB_EINVAL, "unknown family");
    if (nu_code < 0 || nu_code > 4) return fail(c, FAB_EINVAL, "nu_code must be 0..4");
    if (nu_code == 4 && (!(nu > 0.0) || dK)) return fail(c, FAB_EINVAL, "general nu: nu > 0 and no analytic gradient");
    const int d = family == FAB_FAMILY_M52_BASIS ? D - 1 : D;
    if (d < 1 || d > MAX_D) return fail(c, FAB_EINVAL, "need 1 <= d <= 8 Matern axes");
    FAB_CUDA(c, cudaSetDevice(c->device));
    KmatArgs a;
    a.family = family; a.D = D; a.d = d; a.Pk = family == FAB_FAMILY_M52_BASIS ? d + 3 : d + 1; a.nu_code = nu_code; a.nu = nu;
    a.amp_mult = amp_mult; a.theta = theta; a.Xa = Xa; a.Na = Na; a.Xb = Xb; a.Nb = Nb; a.K = K; a.dK = dK;
    const long long n = (long long)Na * Nb;
    FAB_LAUNCH(kmat_kernel, dim3((unsigned)((n + 255) / 256), B), 256, 0, c->stream, a);
    FAB_CUDA(c, cudaGetLastError());
    return FAB_OK;
}

int fab_gp_prior_draw_batch(fab_ctx* c, const double* reps, int Z, unsigned long long seed, double* out_max, double* sample,
                            int* info) {
    if (!c) return FAB_EINVAL;
    if (!c->resident) return fail(c, FAB_ESTATE, "fab_gp_fit_batch has not been called");
    if (!reps || !out_max) return fail(c, FAB_EINVAL, "null pointer");
    if (Z < 1 || Z > 64) return fail(c, FAB_EUNSUPPORTED, "prior draws are limited to 1 <= Z <= 64 points per model");
    FAB_CUDA(c, cudaSetDevice(c->device));
    PriorDrawArgs a;
    a.gd = c->gd; a.theta = c->dTheta; a.reps = reps; a.Z = Z; a.seed = seed; a.out_max = out_max; a.sample = sample; a.info = info;
    FAB_LAUNCH(prior_draw_kernel, c->wsB, 128, 0, c->stream, a);
    FAB_CUDA(c, cudaGetLastError());
    return FAB_OK;
}

int fab_ei_batch(fab_ctx* c, const double* mu, const double* var, int B, int M, const double* y_max, double xi,
                 double* ei, double* best_val, long long* best_idx) {
    if (!c) return FAB_EINVAL;
    if (!mu || !var || !y_max || B < 1 || M < 1) return fail(c, FAB_EINVAL, "bad argument");
    FAB_CUDA(c, cudaSetDevice(c->device));
    EiArgs a;
    a.mu = mu; a.var = var; a.M = M; a.y_max = y_max; a.xi = xi; a.ei = ei; a.best_val = best_val; a.best_idx = best_idx;
    // enough CTAs to stream the moments at HBM speed whatever B is (one model x 1M candidates: config 4)
    int G = (int)(((long long)M + 4 * NTHREADS - 1) / (4 * NTHREADS));
    const int gmax = (4 * c->num_sms + B - 1) / B;
    if (G > gmax) G = gmax;
    if (G < 1) G = 1;
    const long long chunk = ((long long)M + G - 1) / G;
    G = (int)(((long long)M + chunk - 1) / chunk);
    if (G > 1) {
        int rc;
        if ((rc = ensure(c, &c->eiVal, &c->cap_ei[0], (size_t)B * G)) != FAB_OK) return rc;
        if ((rc = ensure(c, &c->eiIdx, &c->cap_ei[1], (size_t)B * G)) != FAB_OK) return rc;
    }
    FAB_LAUNCH(ei_argmax_kernel, dim3(G, B), NTHREADS, 0, c->stream, a, chunk, c->eiVal, c->eiIdx);
    FAB_CUDA(c, cudaGetLastError());
    if (G > 1 && (best_val || best_idx)) {
        FAB_LAUNCH(ei_argmax_final_kernel, B, 32, 0, c->stream, a, G, c->eiVal, c->eiIdx);
        FAB_CUDA(c, cudaGetLastError());
    }
    return FAB_OK;
}

// ---- multi-GPU likelihood step with the exchange fused into the kernel (peer memory over NVLink / NVSwitch) ----
int fab_peer_alloc(fab_ctx* c, int rank, int world, int rows_per_rank, void* handle_out) {
    if (!c) return FAB_EINVAL;
    if (!c->has_data) return fail(c, FAB_ESTATE, "fab_gp_set_data has not been called");
    if (world < 1 || world > FAB_MAX_PEERS || rank < 0 || rank >= world || rows_per_rank < 1 || !handle_out)
        return fail(c, FAB_EINVAL, "need 1 <= world <= 8, 0 <= rank < world, rows_per_rank >= 1");
    FAB_CUDA(c, cudaSetDevice(c->device));
    peer_release(c);
    fab_ctx::Peer& p = c->peer;
    p.world = world; p.rank = rank; p.rows_per_rank = rows_per_rank; p.row_doubles = c->gd.Pk + 3;
    p.buf_bytes = (((size_t)world * rows_per_rank * p.row_doubles * sizeof(double)) + 255) / 256 * 256;
    void* base = nullptr;
    if (cudaMalloc(&base, 2 * p.buf_bytes + 256) != cudaSuccess) { cudaGetLastError(); p = fab_ctx::Peer(); return fail(c, FAB_ENOMEM, "peer buffer allocation failed"); }
    p.base[rank] = base;
    FAB_CUDA(c, cudaMemsetAsync(base, 0, 2 * p.buf_bytes + 256, c->stream));
    FAB_CUDA(c, cudaStreamSynchronize(c->stream));         // peers may write as soon as they have the handle
    memset(handle_out, 0, 64);
#ifndef FAB_CUSIM
    cudaIpcMemHandle_t h;
    FAB_CUDA(c, cudaIpcGetMemHandle(&h, base));
    static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
    memcpy(handle_out, &h, 64);
#endif
    return FAB_OK;
}

int fab_peer_connect(fab_ctx* c, const void* handles) {
    if (!c) return FAB_EINVAL;
    fab_ctx::Peer& p = c->peer;
    if (p.world < 1 || !p.base[p.rank]) return fail(c, FAB_ESTATE, "fab_peer_alloc has not been called");
    if (!handles && p.world > 1) return fail(c, FAB_EINVAL, "null pointer");
    FAB_CUDA(c, cudaSetDevice(c->device));
#ifndef FAB_CUSIM
    for (int r = 0; r < p.world; ++r) {
        if (r == p.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + 64 * r, 64);
        FAB_CUDA(c, cudaIpcOpenMemHandle(&p.base[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
#else
    if (p.world > 1) return fail(c, FAB_EUNSUPPORTED, "the emulator has no peers");
#endif
    p.connected = true;
    p.steps = 0;
    return FAB_OK;
}

int fab_gp_loglik_exchange(fab_ctx* c, const double* theta, const double* yerr2, int B, double* ll, double* grad,
                           int* info) {
    if (!c) return FAB_EINVAL;
    if (!c->has_data) return fail(c, FAB_ESTATE, "fab_gp_set_data has not been called");
    if (!c->peer.connected) return fail(c, FAB_ESTATE, "fab_peer_connect has not been called");
    if (!theta || !yerr2 || !ll || !info || B < 1) return fail(c, FAB_EINVAL, "bad argument");
    if (B != c->peer.rows_per_rank) return fail(c, FAB_EINVAL, "B must equal the rows_per_rank given to fab_peer_alloc");
    if (c->peer.row_doubles != c->gd.Pk + 3) return fail(c, FAB_ESTATE, "the data changed shape since fab_peer_alloc");
    FAB_CUDA(c, cudaSetDevice(c->device));
    // theta / yerr2 may be pinned HOST memory (both or neither): the kernel then fetches them over the bus itself
    bool host_in = false;
#ifndef FAB_CUSIM
    {
        cudaPointerAttributes at;
        const bool th_host = cudaPointerGetAttributes(&at, theta) == cudaSuccess && at.type == cudaMemoryTypeHost;
        const bool ye_host = cudaPointerGetAttributes(&at, yerr2) == cudaSuccess && at.type == cudaMemoryTypeHost;
        cudaGetLastError();
        if (th_host != ye_host) return fail(c, FAB_EINVAL, "theta and yerr2 must both be device or both be pinned host memory");
        if (th_host) {
            theta = static_cast<const double*>(mapped_alias(theta)); yerr2 = static_cast<const double*>(mapped_alias(yerr2));
            if (!theta || !yerr2) return fail(c, FAB_EINVAL, "host inputs must be pinned (CUDA-registered) memory");
            host_in = true;
        }
    }
#endif
    const int rc = launch_dag(c, theta, yerr2, B, ll, info, grad, grad ? DAG_GRAD : DAG_LL, true, host_in);
    if (rc != FAB_OK) return rc;
    c->peer.steps += 1;
    return FAB_OK;
}

int fab_peer_wait(fab_ctx* c, const double** gathered, int* rows, int* row_doubles) {
    if (!c) return FAB_EINVAL;
    fab_ctx::Peer& p = c->peer;
    if (!p.connected || p.steps == 0) return fail(c, FAB_ESTATE, "no exchanged likelihood step is in flight");
    FAB_CUDA(c, cudaSetDevice(c->device));
    // nothing to enqueue: the kernel of the step itself returns only when every rank's rows have arrived
    char* base = static_cast<char*>(p.base[p.rank]);
    if (gathered) *gathered = reinterpret_cast<const double*>(base + (size_t)((p.steps - 1) & 1ull) * p.buf_bytes);
    if (rows) *rows = p.world * p.rows_per_rank;
    if (row_doubles) *row_doubles = p.row_doubles;
    return FAB_OK;
}

int fab_peer_error(fab_ctx* c) {          // synchronises: 1 if a wait timed out (a peer never delivered), else 0
    if (!c) return FAB_EINVAL;
    fab_ctx::Peer& p = c->peer;
    if (!p.connected) return 0;
    int e = 0;
    FAB_CUDA(c, cudaMemcpyAsync(&e, static_cast<char*>(p.base[p.rank]) + peer_ctrl_offset(p) + 128, sizeof(int),
                                cudaMemcpyDeviceToHost, c->stream));
    FAB_CUDA(c, cudaStreamSynchronize(c->stream));
    return e;
}

int fab_peer_free(fab_ctx* c) {
    if (!c) return FAB_EINVAL;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    peer_release(c);
    return FAB_OK;
}

// The whole ensemble chain on the device: one persistent cooperative launch (mcmc.cuh).  `sharded`: coords / lp
// are this rank's replica and the proposals are dealt to the ranks of the box.
static int mcmc_launch(fab_ctx* c, int prior, double* coords, double* log_prob, int K, int P, int nsteps, int init_lp,
                       unsigned long long seed, long long iter0, double a_scale, double* chain, double* chain_lp,
                       long long* naccepted, bool sharded) {
    if (!c->has_data) return fail(c, FAB_ESTATE, "fab_gp_set_data has not been called");
    if (!coords || !log_prob || nsteps < 0 || !(a_scale > 1.0)) return fail(c, FAB_EINVAL, "bad argument");
    if (K < 2 || (K & 1)) return fail(c, FAB_EINVAL, "an even number of walkers >= 2 is required for the red/blue stretch move");
    const GpData& gd = c->gd;
    if (prior != MCMC_PRIOR_FABOLAS && prior != MCMC_PRIOR_ES) return fail(c, FAB_EINVAL, "unknown prior");
    if ((prior == MCMC_PRIOR_FABOLAS) != (gd.family == FAB_FAMILY_M52_BASIS))
        return fail(c, FAB_EINVAL, "prior 0 (fabolas.py) goes with the basis-kernel family, prior 1 (entropy_search.py) with the Matern family");
    if (P != gd.Pk + 1) return fail(c, FAB_EINVAL, "walker dimension must be len(kernel) + 1 (the noise)");
    FAB_CUDA(c, cudaSetDevice(c->device));
    const int world = sharded ? c->mpeer.world : 1;
    int grid = (K / 2 + world - 1) / world;     // proposals of a half-step that this rank owns
#ifdef FAB_CUSIM
    grid = 1;                                   // the emulator runs CTAs one after the other: no grid barrier
#else
    if (grid > c->num_sms) grid = c->num_sms;   // one CTA per SM (shared memory), all co-resident
#endif
    if (c->mcmc_max_ctas > 0 && grid > c->mcmc_max_ctas) grid = c->mcmc_max_ctas;
    if (grid < 1) grid = 1;
    McmcArgs m;
    size_t smem = 0;
    int rc = dag_setup(c, grid, DAG_LL, (size_t)c->smem_optin - MCMC_STATIC_SMEM, &m.dag, &smem);
    if (rc != FAB_OK) return rc;
    if ((rc = ensure(c, &c->mcTheta, &c->cap_mc[0], (size_t)grid * gd.Pk)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->mcYerr2, &c->cap_mc[1], (size_t)grid)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->mcLl, &c->cap_mc[2], (size_t)grid)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->mcQ, &c->cap_mc[3], (size_t)grid * P)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->mcInfo, &c->cap_mc[4], (size_t)grid)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->mcKeys, &c->cap_mc[5], (size_t)grid * K)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->mcOrder, &c->cap_mc[6], (size_t)grid * K)) != FAB_OK) return rc;
    if ((rc = ensure(c, &c->mcBar, &c->cap_mc[7], (size_t)1)) != FAB_OK) return rc;
    FAB_CUDA(c, cudaMemsetAsync(c->mcBar, 0, sizeof(unsigned), c->stream));
    m.dag.theta = c->mcTheta; m.dag.yerr2 = c->mcYerr2; m.dag.ll = c->mcLl; m.dag.info = c->mcInfo;
    m.coords = coords; m.lp = log_prob; m.q = c->mcQ; m.keys = c->mcKeys; m.order = c->mcOrder;
    m.K = K; m.P = P; m.prior = prior; m.nsteps = nsteps; m.init_lp = init_lp != 0;
    m.seed = seed; m.iter0 = iter0; m.a = a_scale;
    m.chain = chain; m.chain_lp = chain_lp; m.nacc = naccepted; m.gbar = c->mcBar;
    m.world = 1; m.rank = 0; m.xgen0 = 0; m.peer_err = nullptr; m.cluster = 0;
    for (int r = 0; r < FAB_MAX_PEERS; ++r) { m.peer_coords[r] = nullptr; m.peer_lp[r] = nullptr; m.peer_flag[r] = nullptr; }
    if (sharded) {
        fab_ctx::McPeer& p = c->mpeer;
        m.world = p.world; m.rank = p.rank; m.xgen0 = p.xgen;
        for (int r = 0; r < p.world; ++r) {
            char* b = static_cast<char*>(p.base[r]);
            m.peer_coords[r] = reinterpret_cast<double*>(b);
            m.peer_lp[r] = reinterpret_cast<double*>(b + p.off_lp);
            m.peer_flag[r] = reinterpret_cast<unsigned*>(b + p.off_flag);
        }
        m.peer_err = reinterpret_cast<int*>(static_cast<char*>(p.base[p.rank]) + p.off_err);
        if (p.world > 1) p.xgen += 1u + (init_lp ? 1u : 0u) + 2u * (unsigned)nsteps;     // barriers of this launch
    }
    if (c->timing) { FAB_CUDA(c, cudaEventRecord(c->ev[0], c->stream)); }
#ifdef FAB_CUSIM
    FAB_DISPATCH_D(gd.d, FAB_LAUNCH(mcmc_stretch_kernel<DD>, grid, DAG_THREADS, smem, c->stream, m));
#else
    void* params[] = {&m};
    void* kern = nullptr;
    FAB_DISPATCH_D(gd.d, kern = reinterpret_cast<void*>(mcmc_stretch_kernel<DD>));
    // A small ensemble on one GPU (the reference's: 20 walkers, 10 CTAs) runs as ONE thread-block cluster, whose hardware
    // barrier is the per-half-step barrier; larger grids and sharded chains use the cooperative launch and the counter.
    bool launched = false;
    if (!sharded && grid >= 2 && grid <= 16 && !c->mcmc_no_cluster) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(DAG_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)grid; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
            cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) == cudaSuccess && nclusters >= 1) {
            m.cluster = 1;
            if (cudaLaunchKernelExC(&cfg, kern, params) == cudaSuccess) launched = true;
            else m.cluster = 0;
        }
        cudaGetLastError();                      // (a refused cluster configuration is not an error: fall back)
    }
    if (!launched)
        FAB_CUDA(c, cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(DAG_THREADS), params, smem, c->stream));
#endif
    FAB_CUDA(c, cudaGetLastError());
    if (c->timing) { FAB_CUDA(c, cudaEventRecord(c->ev[1], c->stream)); c->ev_used = 1; }
    return FAB_OK;
}

int fab_mcmc_run(fab_ctx* c, int prior, double* coords, double* log_prob, int K, int P, int nsteps, int init_lp,
                 unsigned long long seed, long long iter0, double a_scale, double* chain, double* chain_lp,
                 long long* naccepted) {
    if (!c) return FAB_EINVAL;
    return mcmc_launch(c, prior, coords, log_prob, K, P, nsteps, init_lp, seed, iter0, a_scale, chain, chain_lp, naccepted,
                       false);
}

// ---- the same chain with the walkers sharded over the GPUs of one box ----
int fab_mcmc_peer_alloc(fab_ctx* c, int rank, int world, int K, int P, void* handle_out) {
    if (!c) return FAB_EINVAL;
    if (world < 1 || world > FAB_MAX_PEERS || rank < 0 || rank >= world || K < 2 || P < 1 || !handle_out)
        return fail(c, FAB_EINVAL, "need 1 <= world <= 8, 0 <= rank < world, K >= 2");
    FAB_CUDA(c, cudaSetDevice(c->device));
    mpeer_release(c);
    fab_ctx::McPeer& p = c->mpeer;
    p.world = world; p.rank = rank; p.K = K; p.P = P;
    p.off_lp = (((size_t)K * P * sizeof(double)) + 255) / 256 * 256;
    p.off_flag = p.off_lp + (((size_t)K * sizeof(double)) + 255) / 256 * 256;
    p.off_err = p.off_flag + (size_t)FAB_MAX_PEERS * 64;
    p.bytes = p.off_err + 256;
    void* base = nullptr;
    if (cudaMalloc(&base, p.bytes) != cudaSuccess) { cudaGetLastError(); p = fab_ctx::McPeer(); return fail(c, FAB_ENOMEM, "replica allocation failed"); }
    p.base[rank] = base;
    FAB_CUDA(c, cudaMemsetAsync(base, 0, p.bytes, c->stream));
    FAB_CUDA(c, cudaStreamSynchronize(c->stream));
    memset(handle_out, 0, 64);
#ifndef FAB_CUSIM
    cudaIpcMemHandle_t h;
    FAB_CUDA(c, cudaIpcGetMemHandle(&h, base));
    memcpy(handle_out, &h, 64);
#endif
    return FAB_OK;
}

int fab_mcmc_peer_connect(fab_ctx* c, const void* handles) {
    if (!c) return FAB_EINVAL;
    fab_ctx::McPeer& p = c->mpeer;
    if (p.world < 1 || !p.base[p.rank]) return fail(c, FAB_ESTATE, "fab_mcmc_peer_alloc has not been called");
    if (!handles && p.world > 1) return fail(c, FAB_EINVAL, "null pointer");
    FAB_CUDA(c, cudaSetDevice(c->device));
#ifndef FAB_CUSIM
    for (int r = 0; r < p.world; ++r) {
        if (r == p.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + 64 * r, 64);
        FAB_CUDA(c, cudaIpcOpenMemHandle(&p.base[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
#else
    if (p.world > 1) return fail(c, FAB_EUNSUPPORTED, "the emulator has no peers");
#endif
    p.connected = true;
    p.xgen = 0;
    return FAB_OK;
}

int fab_mcmc_peer_state(fab_ctx* c, double* coords, double* log_prob, int to_replica) {
    if (!c) return FAB_EINVAL;
    fab_ctx::McPeer& p = c->mpeer;
    if (!p.connected) return fail(c, FAB_ESTATE, "fab_mcmc_peer_connect has not been called");
    FAB_CUDA(c, cudaSetDevice(c->device));
    char* b = static_cast<char*>(p.base[p.rank]);
    const size_t nc = sizeof(double) * p.K * p.P, nl = sizeof(double) * p.K;
    if (coords) FAB_CUDA(c, to_replica ? cudaMemcpyAsync(b, coords, nc, cudaMemcpyDeviceToDevice, c->stream)
                                       : cudaMemcpyAsync(coords, b, nc, cudaMemcpyDeviceToDevice, c->stream));
    if (log_prob) FAB_CUDA(c, to_replica ? cudaMemcpyAsync(b + p.off_lp, log_prob, nl, cudaMemcpyDeviceToDevice, c->stream)
                                         : cudaMemcpyAsync(log_prob, b + p.off_lp, nl, cudaMemcpyDeviceToDevice, c->stream));
    return FAB_OK;
}

int fab_mcmc_run_sharded(fab_ctx* c, int prior, int nsteps, int init_lp, unsigned long long seed, long long iter0,
                         double a_scale, long long* naccepted) {
    if (!c) return FAB_EINVAL;
    fab_ctx::McPeer& p = c->mpeer;
    if (!p.connected) return fail(c, FAB_ESTATE, "fab_mcmc_peer_connect has not been called");
    char* b = static_cast<char*>(p.base[p.rank]);
    return mcmc_launch(c, prior, reinterpret_cast<double*>(b), reinterpret_cast<double*>(b + p.off_lp), p.K, p.P, nsteps,
                       init_lp, seed, iter0, a_scale, nullptr, nullptr, naccepted, true);
}

int fab_mcmc_peer_error(fab_ctx* c) {
    if (!c) return FAB_EINVAL;
    fab_ctx::McPeer& p = c->mpeer;
    if (!p.connected) return 0;
    int e = 0;
    FAB_CUDA(c, cudaMemcpyAsync(&e, static_cast<char*>(p.base[p.rank]) + p.off_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    FAB_CUDA(c, cudaStreamSynchronize(c->stream));
    return e;
}

int fab_mcmc_peer_free(fab_ctx* c) {
    if (!c) return FAB_EINVAL;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    mpeer_release(c);
    return FAB_OK;
}

int fab_mcmc_log_prior(fab_ctx* c, int prior, const double* p, int K, int P, double* out, double* theta, double* yerr2) {
    if (!c) return FAB_EINVAL;
    if (!p || !out || K < 1) return fail(c, FAB_EINVAL, "bad argument");
    if (prior != MCMC_PRIOR_FABOLAS && prior != MCMC_PRIOR_ES) return fail(c, FAB_EINVAL, "unknown prior");
    const int Pk = P - 1, d = prior == MCMC_PRIOR_FABOLAS ? Pk - 3 : Pk - 1;
    if (d < 1 || d > MAX_D) return fail(c, FAB_EINVAL, "need 1 <= d <= 8 Matern axes");
    FAB_CUDA(c, cudaSetDevice(c->device));
    FAB_LAUNCH(mcmc_prior_kernel, (K + 127) / 128, 128, 0, c->stream, p, K, P, prior, d, Pk, out, theta, yerr2);
    FAB_CUDA(c, cudaGetLastError());
    return FAB_OK;
}

int fab_ctx_debug_set_mcmc_ctas(fab_ctx* c, int n) {
    if (!c || n < 0) return FAB_EINVAL;
    c->mcmc_max_ctas = n;
    return FAB_OK;
}

int fab_gp_debug_get_factor(fab_ctx* c, int b, double* L_out) {
    if (!c || !L_out) return FAB_EINVAL;
    if (!c->factored || c->inverted || b < 0 || b >= c->wsB)
        return fail(c, FAB_ESTATE, "no Cholesky factor in the workspace for this walker");
    const GpData& gd = c->gd;
    const int ntiles = gd.NT * (gd.NT + 1) / 2;
    FAB_LAUNCH(untile_kernel, gd.N, 128, 0, c->stream, c->wk.Lw + (size_t)b * ntiles * TILE_ELEMS, L_out, gd.N,
               gd.NT);
    FAB_CUDA(c, cudaGetLastError());
    return FAB_OK;
}

}  // extern "C"

#if defined(FAB_PHASE_PROF) && !defined(FAB_CUSIM)
// Development build only (tools/phase_prof.py): per-phase cycle counters of CTA 0.
extern "C" int fab_debug_phase_reset(void) {
    long long z[64] = {0};
    return cudaMemcpyToSymbol(fab::g_phase, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
extern "C" int fab_debug_trace_read(long long* out) {      // [2][128][4] clock64 stamps of CTA 0 (bulk ops, chain ops)
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, fab::g_trace, sizeof(long long) * 2 * 128 * 4) == cudaSuccess ? 0 : -1;
}
extern "C" int fab_debug_trace_w_read(long long* out) {    // [16][128][6] stamps of lane 0 of every bulk warp of CTA 0
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, fab::g_trace_w, sizeof(long long) * 16 * 128 * 6) == cudaSuccess ? 0 : -1;
}
extern "C" int fab_debug_program_read(int NT, int nslot, int mode, int* out, int max_ops) {   // (type, I, J, k) per bulk op
    const DagProgram P = build_dag_program(NT, nslot, mode);
    int n = 0;
    for (const DagOp& d : P.bulk) { if (n >= max_ops) break; out[4 * n] = d.type; out[4 * n + 1] = d.I; out[4 * n + 2] = d.J; out[4 * n + 3] = d.k; ++n; }
    return n;
}
extern "C" int fab_debug_phase_read(long long* out) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, fab::g_phase, sizeof(long long) * 64) == cudaSuccess ? 0 : -1;
}
#endif
