// prims.cuh -- device primitives for the sm_100a kernels: FP64 tensor-core tile MMA (DMMA),
// TMA bulk copies completed on mbarriers, warp/block reductions, launch macro.
//
// Every primitive has a second body under FAB_CUSIM so that the kernel logic can be replayed by the
// fiber-based emulator in tests/cusim (test infrastructure; never part of the shipped library).
#pragma once
#ifndef FAB_CUSIM
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>
#endif

namespace fab {

#ifdef FAB_CUSIM
#define FAB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    cusim::launch(dim3(grid), dim3(block), (smem), [=]() { kernel(__VA_ARGS__); })
#define FAB_DYN_SMEM(name) unsigned char* name = reinterpret_cast<unsigned char*>(cusim::S().dyn_smem)
#else
#define FAB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define FAB_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#endif

// Optional per-phase cycle counters (development builds only: -DFAB_PHASE_PROF, tools/phase_prof.py).
// Thread 0 of CTA 0 brackets a region with PHB(id) / PHE(id); both are fire-and-forget reductions.
// Accumulator layout of the bulk warps of gp_dag_kernel (gp_dag.cuh): 2 = eight warps, two paired 16 x 16 units each
// (shipped); 0 = eight warps, one 32 x 16 block each (round 1); 1 = sixteen warps, one 16 x 16 unit each (experiment)
#ifndef FAB_DAG_LAYOUT
#define FAB_DAG_LAYOUT 2
#endif
#define FAB_DAG_NBW (FAB_DAG_LAYOUT == 1 ? 16 : 8)
#if defined(FAB_PHASE_PROF) && !defined(FAB_CUSIM)
__device__ long long g_phase[64];
#define PHB(id) do { if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd((unsigned long long*)&g_phase[id], (unsigned long long)(-clock64())); } while (0)
#define PHE(id) do { if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd((unsigned long long*)&g_phase[id], (unsigned long long)clock64()); } while (0)
__device__ long long g_trace[2][128][4];
#define TRB(i, k) do { if (threadIdx.x == 0 && blockIdx.x == 0 && (i) < 128) g_trace[0][i][k] = clock64(); } while (0)
__device__ long long g_trace_w[16][128][6];     // per bulk warp (lane 0): arrive, barrier, chain wait, TMA issued, loads landed, end
#define TRW(i, k) do { if ((threadIdx.x & 31) == 0 && blockIdx.x == 0 && (i) < 128) g_trace_w[threadIdx.x >> 5][i][k] = clock64(); } while (0)
#define TRC(i, k) do { if (threadIdx.x == 32 * FAB_DAG_NBW && blockIdx.x == 0 && (i) < 128) g_trace[1][i][k] = clock64(); } while (0)
#define PHBC(id) do { if (threadIdx.x == 32 * FAB_DAG_NBW && blockIdx.x == 0) atomicAdd((unsigned long long*)&g_phase[id], (unsigned long long)(-clock64())); } while (0)
#define PHEC(id) do { if (threadIdx.x == 32 * FAB_DAG_NBW && blockIdx.x == 0) atomicAdd((unsigned long long*)&g_phase[id], (unsigned long long)clock64()); } while (0)
#else
#define PHB(id) do {} while (0)
#define PHE(id) do {} while (0)
#define PHBC(id) do {} while (0)
#define PHEC(id) do {} while (0)
#define TRB(i, k) do {} while (0)
#define TRC(i, k) do {} while (0)
#define TRW(i, k) do {} while (0)
#endif

constexpr double kTiny = 1.25e-12;        // george's default white noise (oracle/george_oracle.py TINY)
constexpr double kLog2Pi = 1.8378770664093453;

// ---------------------------------------------------------------------------------------------
// DMMA: D(8x8) += A(8x4, row) * B(4x8, col), FP64 tensor-core path (mma.sync m8n8k4 f64).
// lane = 4*g + t :  a = A[g][t],  b = B[t][g],  c0 = C[g][2t],  c1 = C[g][2t+1].
// Measured on B200: 36.96 TFLOP/s (profiles/r01_fp64_peaks.json) == DFMA pipe peak, but with 8x
// fewer operand fetches per FMA than a register-tiled DFMA loop.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
#ifdef FAB_CUSIM
    cusim_dmma884(c0, c1, a, b);
#else
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
#endif
}

// ---------------------------------------------------------------------------------------------
// mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP).  One thread issues, everybody waits.
// ---------------------------------------------------------------------------------------------
#ifndef FAB_CUSIM
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
#endif

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
#ifdef FAB_CUSIM
    *bar = 0;
#else
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}

// Issue a global->shared bulk copy of `bytes` (multiple of 16, both sides 16B-aligned) and arm the
// barrier with the expected transaction count.  Call from ONE thread.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes,
                                         unsigned long long* bar) {
#ifdef FAB_CUSIM
    memcpy(dst_smem, src_gmem, bytes);
    *bar += 1;
#else
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
#endif
}

// Wait for the barrier phase with parity `parity` to complete (all threads may call).
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
#ifdef FAB_CUSIM
    while (((*bar) & 1ull) == parity) cusim::yield();   // *bar counts completed phases
    CUSIM_RESTORE_TID();
#else
    unsigned ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
#endif
}

// 16-byte asynchronous global -> shared copy (LDGSTS): no register staging, completion by commit / wait groups of
// the issuing thread.  Used for the op-record ring of the fused GP kernel.
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_gmem) {
#ifdef FAB_CUSIM
    memcpy(dst_smem, src_gmem, 16);
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef FAB_CUSIM
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
#ifndef FAB_CUSIM
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// Every thread that wrote (with ordinary stores) shared memory that a bulk store is about to read
// calls this BEFORE the block barrier that precedes bulk_s2g (generic -> async proxy ordering).
__device__ __forceinline__ void fence_proxy_async() {
#ifndef FAB_CUSIM
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

// shared->global bulk store; call from ONE thread after fence_proxy_async() + __syncthreads().
// The same thread must call bulk_s2g_wait_read() before the shared memory is overwritten and
// bulk_s2g_wait_all() before the kernel exits / before the data is re-loaded.
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, unsigned bytes) {
#ifdef FAB_CUSIM
    memcpy(dst_gmem, src_smem, bytes);
#else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void bulk_s2g_wait_read() {
#ifndef FAB_CUSIM
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}
__device__ __forceinline__ void bulk_s2g_wait_all() {
#ifndef FAB_CUSIM
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
}

// 1/sqrt(x) for normal positive x on the latency-critical Cholesky chain: hardware seed (upper 32
// bits, ~2^-20) + one third-order (Halley) step => relative error ~2^-59, 6 dependent FP64-pipe
// instructions instead of the library rsqrt()'s 13.  x <= 0 or NaN must be screened by the caller.
__device__ __forceinline__ double rsqrt_fast(double x) {
#ifdef FAB_CUSIM
    return 1.0 / std::sqrt(x);
#else
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);             // 1 - x y^2
    double q = fma(0.375, e, 0.5);
    q = q * e;                                    // e/2 + 3 e^2/8
    return fma(y, q, y);
#endif
}

// Named barrier among `nthreads` (multiple of 32) threads of the CTA; id 1..15 (0 is __syncthreads).
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
#ifdef FAB_CUSIM
    cusim_named_barrier(id, nthreads);
#else
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
#endif
}

// Producer side of a named barrier: counts this thread in without waiting (pairs with named_bar_sync on
// the consumer side; `nthreads` = arriving + waiting threads).
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
#ifdef FAB_CUSIM
    cusim_named_barrier_arrive(id, nthreads);
#else
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
#endif
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_max_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of NV values per thread; result valid in every thread.  `scratch` must hold
// NV * (blockDim.x / 32) doubles.  Contains two __syncthreads().
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = warp_sum(v[i]);
        if (lane == 0) scratch[i * nw + w] = s;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = 0.0;
        for (int k = 0; k < nw; ++k) s += scratch[i * nw + k];
        v[i] = s;
    }
    __syncthreads();
}

}  // namespace fab
