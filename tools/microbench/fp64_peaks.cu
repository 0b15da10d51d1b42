// FP64 pipe microbenchmark for the roofline denominator ("FP64 peak must be measured").
// Measures on one B200: (1) DFMA register-only peak, (2) DMMA mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16
// peak, (3) DFMA+DMMA issued together, (4) shared-memory-fed DFMA 4x4 and 8x8 register tiles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double s) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
    double b = s, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* d, const double* a, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double s) {
    double d[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = 0.0;
    double a = s + threadIdx.x * 1e-6, b = s * 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(d[2 * i], d[2 * i + 1], a, b);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void __launch_bounds__(256) k_dmma1684(double* out, int iters, double s) {
    double d[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) d[i] = 0.0;
    double a[2] = {s + threadIdx.x * 1e-6, s}; double b = s * 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma1684(d + 4 * i, a, b);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) r += d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double s) {
    double d[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) d[i] = 0.0;
    double a[4] = {s + threadIdx.x * 1e-6, s, s * 2, s * 3}; double b[2] = {s * 0.5, s * 0.25};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma1688(d + 4 * i, a, b);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) r += d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double s) {
    double d[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) d[i] = 0.0;
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = s + i + threadIdx.x * 1e-6;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = s * 0.5 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma16816(d + 4 * i, a, b);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) r += d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// DFMA and DMMA interleaved: are they the same pipe?
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double s) {
    double d[16], f[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = i;
    double a = s + threadIdx.x * 1e-6, b = s * 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { dmma884(d[2 * i], d[2 * i + 1], a, b); f[i] = fma(f[i], b, a); }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += d[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) r += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// smem-fed register-tiled DFMA: TM x TN tile per thread, operands from shared (rank-1 updates), 64-deep k.
template <int TM, int TN>
__global__ void __launch_bounds__(256) k_smem_tile(double* out, int iters) {
    __shared__ double As[16][16 * TM];   // k-major
    __shared__ double Bs[16][16 * TN];
    for (int i = threadIdx.x; i < 16 * 16 * TM; i += 256) (&As[0][0])[i] = 1e-3 * (i % 97);
    for (int i = threadIdx.x; i < 16 * 16 * TN; i += 256) (&Bs[0][0])[i] = 1e-3 * (i % 89);
    __syncthreads();
    const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
    double acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
            double a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 2) { double2 v = *reinterpret_cast<const double2*>(&As[k][ty * TM + i]); a[i] = v.x; a[i + 1] = v.y; }
#pragma unroll
            for (int j = 0; j < TN; j += 2) { double2 v = *reinterpret_cast<const double2*>(&Bs[k][tx * TN + j]); b[j] = v.x; b[j + 1] = v.y; }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) r += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F>
double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * 256 * sms * 8));
    const int iters = 20000;
    for (int bps = 1; bps <= 4; bps *= 2) {
        int grid = sms * bps;
        double ms;
        ms = time_ms([&] { k_dfma<<<grid, 256>>>(out, iters, 1.0000001); }, 5);
        printf(" \"dfma_tflops_bps%d\": %.3f,\n", bps, 2.0 * 16 * iters * 256.0 * grid / ms * 1e-9);
        ms = time_ms([&] { k_dmma884<<<grid, 256>>>(out, iters, 1.0000001); }, 5);
        printf(" \"dmma_m8n8k4_tflops_bps%d\": %.3f,\n", bps, 2.0 * 8 * 256 * iters * 8.0 * grid / ms * 1e-9);
        ms = time_ms([&] { k_dmma1684<<<grid, 256>>>(out, iters, 1.0000001); }, 5);
        printf(" \"dmma_m16n8k4_tflops_bps%d\": %.3f,\n", bps, 2.0 * 8 * 512 * iters * 8.0 * grid / ms * 1e-9);
        ms = time_ms([&] { k_dmma1688<<<grid, 256>>>(out, iters, 1.0000001); }, 5);
        printf(" \"dmma_m16n8k8_tflops_bps%d\": %.3f,\n", bps, 2.0 * 8 * 1024 * iters * 8.0 * grid / ms * 1e-9);
        ms = time_ms([&] { k_dmma16816<<<grid, 256>>>(out, iters, 1.0000001); }, 5);
        printf(" \"dmma_m16n8k16_tflops_bps%d\": %.3f,\n", bps, 2.0 * 8 * 2048 * iters * 8.0 * grid / ms * 1e-9);
        ms = time_ms([&] { k_mixed<<<grid, 256>>>(out, iters, 1.0000001); }, 5);
        printf(" \"mixed_dfma_plus_dmma884_tflops_bps%d\": %.3f,\n", bps, (2.0 * 8 * 256 * 8.0 + 2.0 * 8 * 256) * iters * grid / ms * 1e-9);
    }
    {
        int grid = sms; const int it2 = 1600; double ms;
        ms = time_ms([&] { k_smem_tile<4, 4><<<grid, 256>>>(out, it2); }, 5);
        printf(" \"smem_tile_4x4_tflops\": %.3f,\n", 2.0 * 16 * 16 * it2 * 256.0 * grid / ms * 1e-9);
        ms = time_ms([&] { k_smem_tile<8, 4><<<grid, 256>>>(out, it2); }, 5);
        printf(" \"smem_tile_8x4_tflops\": %.3f,\n", 2.0 * 32 * 16 * it2 * 256.0 * grid / ms * 1e-9);
        ms = time_ms([&] { k_smem_tile<8, 8><<<grid, 256>>>(out, it2 / 2); }, 5);
        printf(" \"smem_tile_8x8_tflops\": %.3f,\n", 2.0 * 64 * 16 * (it2 / 2) * 256.0 * grid / ms * 1e-9);
    }
    printf(" \"note\": \"best of 5, CUDA events, register-resident operands unless smem_tile\"}\n");
    return 0;
}
