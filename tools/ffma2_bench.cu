// Micro-benchmark: scalar FFMA vs packed FFMA2 issue rate on sm_100a, alone and mixed with ALU-pipe work.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_bench tools/ffma2_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ int iadd(int a, int b) { int r; asm volatile("add.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

// MODE 0: 16 scalar FFMA per iteration; 1: 8 FFMA2 (same flops); 2: 16 FFMA + 4 IADD; 3: 8 FFMA2 + 4 IADD; 4: 8 FFMA2 + 8 IADD
template <int MODE>
__global__ void bench(float* out, long long* cyc, int iters, float seed) {
    float x[16];
    u64 y[8];
    int z[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = seed + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) { y[i] = pk(x[2 * i], x[2 * i + 1]); z[i] = threadIdx.x + i; }
    const float a = seed * 0.5f, b = seed * 0.25f;
    const u64 a2 = pk(a, a), b2 = pk(b, b);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fma1(x[i], a, b);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = fma2(y[i], a2, b2);
        }
        if (MODE == 2 || MODE == 3) {
#pragma unroll
            for (int i = 0; i < 4; ++i) z[i] = iadd(z[i], it);
        }
        if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) z[i] = iadd(z[i], it);
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float p, q; upk(y[i], p, q); s += p + q + z[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, float* out, long long* cyc, int threads) {
    const int iters = 4096;
    bench<MODE><<<148, threads>>>(out, cyc, iters, 1.0f);
    bench<MODE><<<148, threads>>>(out, cyc, iters, 1.0f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0;
    for (int i = 0; i < 148; ++i) m += h[i];
    m /= 148;
    printf("%-28s threads %4d: %.2f cycles/iteration/SMSP-warp-set (%.3f per warp-iteration)\n", name, threads, m / iters, m / iters / (threads / 128.0));
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int threads : {128, 512}) {
        run<0>("16 FFMA", out, cyc, threads);
        run<1>("8 FFMA2", out, cyc, threads);
        run<2>("16 FFMA + 4 IADD", out, cyc, threads);
        run<3>("8 FFMA2 + 4 IADD", out, cyc, threads);
        run<4>("8 FFMA2 + 8 IADD", out, cyc, threads);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
