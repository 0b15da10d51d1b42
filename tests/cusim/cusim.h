// cusim.h -- a minimal fiber-based CUDA execution-model emulator for the CPU.
//
// TEST INFRASTRUCTURE ONLY.  The CUDA sources under fabolas_b200/csrc are compiled a second time
// with g++ (-DFAB_CUSIM -include cusim.h) into tests/cusim/_build/libfabolas_sim.so so that the
// kernel LOGIC (tile schedules, indexing, barriers, warp exchanges) can be checked against the
// oracle in the no-GPU test tier.  The product package never loads that library: fabolas_b200
// loads only the nvcc-built libfabolas_b200.so and fails loudly when it (or a GPU) is missing.
//
// Model: one OS thread; every CUDA thread of a block is a ucontext fiber; blocks run one after the
// other.  __syncthreads / __syncwarp / shuffles / mma.sync are rendezvous points implemented by
// yielding round-robin until all participants have arrived.
#pragma once
#include <ucontext.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <algorithm>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 { double x, y; };
struct double4 { double x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

namespace cusim {
struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
    unsigned long long bar_gen = 0;   // generation of the block barrier this fiber waits on
    bool at_block_bar = false;
    bool at_warp_bar = false;
};
struct State {
    ucontext_t sched;
    std::vector<Fiber> fibers;
    int cur = -1;
    int nthreads = 0;
    std::function<void()> body;
    // block barrier
    int block_arrived = 0;
    unsigned long long block_gen = 0;
    // per-warp barrier + exchange slots
    std::vector<int> warp_arrived;
    std::vector<int> warp_live;
    int block_live = 0;
    std::vector<unsigned long long> warp_gen;
    std::vector<double> xch_d;      // 2 doubles per lane
    std::vector<long long> xch_i;
    char* dyn_smem = nullptr;
};
inline State& S() { static State s; return s; }
}  // namespace cusim

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;
#ifdef CUSIM_DEFINE_GLOBALS
uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;
#endif
static const int warpSize = 32;

namespace cusim {
inline void yield() { State& s = S(); swapcontext(&s.fibers[s.cur].ctx, &s.sched); }

inline int live_in_warp(int w) { return S().warp_live[w]; }
inline int live_in_block() { return S().block_live; }
inline void block_barrier() {
    State& s = S();
    unsigned long long g = s.block_gen;
    ++s.block_arrived;
    while (s.block_gen == g) {
        if (s.block_arrived >= live_in_block()) { s.block_arrived = 0; ++s.block_gen; break; }
        yield();
    }
}
inline void warp_barrier() {
    State& s = S();
    int w = s.cur / 32;
    unsigned long long g = s.warp_gen[w];
    ++s.warp_arrived[w];
    while (s.warp_gen[w] == g) {
        if (s.warp_arrived[w] >= live_in_warp(w)) { s.warp_arrived[w] = 0; ++s.warp_gen[w]; break; }
        yield();
    }
}
inline void trampoline() {
    State& s = S();
    s.body();
    s.fibers[s.cur].done = true;
    --s.block_live; --s.warp_live[s.cur / 32];
    swapcontext(&s.fibers[s.cur].ctx, &s.sched);
}
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, std::function<void()> body) {
    State& s = S();
    const int nt = block.x * block.y * block.z;
    const size_t stack_sz = 256 * 1024;
    s.nthreads = nt;
    s.body = body;
    gridDim = grid; blockDim = block;
    std::vector<char> smem(smem_bytes + 1024);
    s.dyn_smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem.data()) + 1023) & ~uintptr_t(1023));
    std::vector<char*> stacks(nt);
    for (int t = 0; t < nt; ++t) stacks[t] = static_cast<char*>(malloc(stack_sz));
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        memset(s.dyn_smem, 0xFF, smem_bytes);   // NaN-poison dynamic shared memory
        s.fibers.assign(nt, Fiber());
        s.block_arrived = 0; s.block_gen = 0;
        int nw = (nt + 31) / 32;
        s.warp_arrived.assign(nw, 0); s.warp_gen.assign(nw, 0);
        s.block_live = nt; s.warp_live.assign(nw, 0);
        for (int t = 0; t < nt; ++t) ++s.warp_live[t / 32];
        s.xch_d.assign(nt * 2, 0.0); s.xch_i.assign(nt, 0);
        for (int t = 0; t < nt; ++t) {
            Fiber& f = s.fibers[t];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = stacks[t];
            f.ctx.uc_stack.ss_size = stack_sz;
            f.ctx.uc_link = &s.sched;
            makecontext(&f.ctx, (void (*)())trampoline, 0);
        }
        int remaining = nt;
        while (remaining > 0) {
            remaining = 0;
            for (int t = 0; t < nt; ++t) {
                if (s.fibers[t].done) continue;
                s.cur = t;
                blockIdx = uint3{bx, by, bz};
                threadIdx = uint3{unsigned(t % block.x), unsigned((t / block.x) % block.y), unsigned(t / (block.x * block.y))};
                swapcontext(&s.sched, &s.fibers[t].ctx);
                if (!s.fibers[t].done) ++remaining;
            }
        }
    }
    for (int t = 0; t < nt; ++t) free(stacks[t]);
    s.cur = -1;
}
}  // namespace cusim

// After a context switch the global threadIdx belongs to someone else; every rendezvous primitive
// restores it for the resuming fiber.
#define CUSIM_RESTORE_TID() do { int t_ = cusim::S().cur; \
    threadIdx = uint3{unsigned(t_ % blockDim.x), unsigned((t_ / blockDim.x) % blockDim.y), unsigned(t_ / (blockDim.x * blockDim.y))}; } while (0)

static inline void __syncthreads() { cusim::block_barrier(); CUSIM_RESTORE_TID(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cusim::warp_barrier(); CUSIM_RESTORE_TID(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <typename T>
static inline T cusim_shfl_idx(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    cusim::State& s = cusim::S();
    int me = s.cur, w = me / 32;
    long long raw = 0; memcpy(&raw, &v, sizeof(T));
    s.xch_i[me] = raw;
    cusim::warp_barrier();
    long long got = s.xch_i[w * 32 + (src & 31)];
    cusim::warp_barrier();
    CUSIM_RESTORE_TID();
    T out; memcpy(&out, &got, sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return cusim_shfl_idx(v, src); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return cusim_shfl_idx(v, (cusim::S().cur & 31) ^ m); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d, int = 32) {
    int l = cusim::S().cur & 31; return cusim_shfl_idx(v, l + d < 32 ? l + d : l);
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d, int = 32) {
    int l = cusim::S().cur & 31; return cusim_shfl_idx(v, l - d >= 0 ? l - d : l);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) { int p = cusim_shfl_idx(pred, l); if (p) r |= 1u << l; }
    return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) {
    cusim::State& s = cusim::S(); int w = s.cur / 32; unsigned full = 0;
    for (int l = 0; l < 32; ++l) if (w * 32 + l < s.nthreads) full |= 1u << l;
    return __ballot_sync(m, p) == full;
}

// mma.sync.aligned.m8n8k4.row.col.f64: warp-collective; lane = 4*g + t holds A[g][t], B[t][g],
// C[g][2t], C[g][2t+1].
static inline void cusim_dmma884(double& c0, double& c1, double a, double b) {
    cusim::State& s = cusim::S();
    int me = s.cur, w = me / 32, lane = me & 31, g = lane >> 2, t = lane & 3;
    s.xch_d[2 * me] = a; s.xch_d[2 * me + 1] = b;
    cusim::warp_barrier();
    double r0 = c0, r1 = c1;
    for (int k = 0; k < 4; ++k) {
        double av = s.xch_d[2 * (w * 32 + 4 * g + k)];
        double b0 = s.xch_d[2 * (w * 32 + 4 * (2 * t) + k) + 1];
        double b1 = s.xch_d[2 * (w * 32 + 4 * (2 * t + 1) + k) + 1];
        r0 = std::fma(av, b0, r0); r1 = std::fma(av, b1, r1);
    }
    cusim::warp_barrier();
    CUSIM_RESTORE_TID();
    c0 = r0; c1 = r1;
}

// math
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline int __double2loint(double d) { return (int)(unsigned)(__double_as_longlong(d) & 0xffffffffLL); }
static inline int __double2hiint(double d) { return (int)(__double_as_longlong(d) >> 32); }
static inline double __hiloint2double(int hi, int lo) { return __longlong_as_double(((long long)hi << 32) | (unsigned)lo); }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
static inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline int atomicMin(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
static inline int atomicCAS(int* p, int cmp, int v) { int o = *p; if (o == cmp) *p = v; return o; }
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
using std::min; using std::max; using std::isfinite; using std::isnan; using std::isinf;

// ------------------------------------------------------------------------------------------------
// host runtime stubs (device memory == host memory)
// ------------------------------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); if (*p) memset(*p, 0xFF, n); return *p ? 0 : 2; }
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "cusim"; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
// block-wide OR (CUDA's __syncthreads_or): two rendezvous points around a shared accumulator
static inline int __syncthreads_or(int pred) {
    static int acc_[2] = {0, 0};
    static unsigned long long gen_ = 0;
    const unsigned long long my = gen_;
    if (pred) acc_[my & 1] = 1;
    cusim::block_barrier();
    if (gen_ == my) { gen_ = my + 1; acc_[(my + 1) & 1] = 0; }
    const int r = acc_[my & 1];
    cusim::block_barrier();
    CUSIM_RESTORE_TID();
    return r;
}
// named barrier (PTX bar.sync / bar.arrive id, nthreads): completes when `count` threads of the block have
// arrived (bar.arrive counts a thread in without waiting)
namespace cusim { struct NamedBar { int arrived[16]; unsigned long long gen[16]; }; inline NamedBar& NB() { static NamedBar b{}; return b; } }
static inline void cusim_named_barrier(int id, int count) {
    cusim::NamedBar& b = cusim::NB();
    const unsigned long long g = b.gen[id];
    ++b.arrived[id];
    while (b.gen[id] == g) {
        if (b.arrived[id] >= count) { b.arrived[id] = 0; ++b.gen[id]; break; }
        cusim::yield();
    }
    CUSIM_RESTORE_TID();
}
static inline void cusim_named_barrier_arrive(int id, int count) {
    cusim::NamedBar& b = cusim::NB();
    ++b.arrived[id];
    if (b.arrived[id] >= count) { b.arrived[id] = 0; ++b.gen[id]; }
}
