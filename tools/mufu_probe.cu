// Measures bias / rms / max error of the MUFU approximations the log-likelihood epilogue relies on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mufu_probe tools/mufu_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// stats[0]=sum err, [1]=sum err^2, [2]=max|err|, [3]=count
__global__ void probe(int which, float lo, float hi, int n, double* stats) {
    double s = 0, s2 = 0, mx = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // low-discrepancy-ish scan with an irrational stride so mantissa bits vary
        double t = fmod((double)i * 0.6180339887498949 + 0.137, 1.0);
        float x = (float)(lo + (hi - lo) * t);
        double err;
        if (which == 0) {  // ex2: relative error
            err = (double)ex2a(x) / exp2((double)x) - 1.0;
        } else if (which == 1) {  // lg2: absolute error
            err = (double)lg2a(x) - log2((double)x);
        } else {  // composed term n*log2(1+u) - k*u with u = 2^x, n = 1, k = 0.72: absolute error
            float u = ex2a(x);
            float lg = lg2a(1.0f + u);
            float v = fmaf(1.0f, lg, -0.72f * u);
            double ud = exp2((double)x);
            err = (double)v - (log2(1.0 + ud) - (double)0.72f * ud);
        }
        s += err; s2 += err * err; mx = fmax(mx, fabs(err));
    }
    atomicAdd(&stats[0], s); atomicAdd(&stats[1], s2);
    // max via CAS on bits (values are non-negative)
    unsigned long long* p = (unsigned long long*)&stats[2];
    unsigned long long old = *p, v = __double_as_longlong(mx);
    while (v > old) { unsigned long long prev = atomicCAS(p, old, v); if (prev == old) break; old = prev; }
}

int main() {
    double* d; cudaMalloc(&d, 4 * sizeof(double));
    struct { int which; float lo, hi; const char* name; } cases[] = {
        {0, -20.f, 20.f, "ex2 rel, x in [-20,20]"}, {0, -3.f, 3.f, "ex2 rel, x in [-3,3]"}, {0, 0.f, 1.f, "ex2 rel, x in [0,1]"},
        {1, 1.f, 2.f, "lg2 abs, w in [1,2]"}, {1, 1.f, 1.01f, "lg2 abs, w in [1,1.01]"}, {1, 1.f, 64.f, "lg2 abs, w in [1,64]"},
        {1, 2.f, 4.f, "lg2 abs, w in [2,4]"}, {1, 1.0f, 1.5f, "lg2 abs, w in [1,1.5]"},
        {2, -6.f, 6.f, "term abs, x in [-6,6]"}, {2, -1.f, 3.f, "term abs, x in [-1,3]"}, {2, -12.f, -6.f, "term abs, x in [-12,-6]"},
    };
    const int n = 1 << 26;
    for (auto& c : cases) {
        cudaMemset(d, 0, 4 * sizeof(double));
        probe<<<1184, 256>>>(c.which, c.lo, c.hi, n, d);
        double h[4]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-28s mean % .3e  rms %.3e  max %.3e\n", c.name, h[0] / n, sqrt(h[1] / n), h[2]);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
