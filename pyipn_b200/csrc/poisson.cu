// Posterior-predictive Poisson counts (drop-in for ppc_generator, pyipn/fit.py:694-705) and the two
// device micro-benchmarks that give bench.py its measured FP32-FMA / MUFU denominators.
//
// The reference draws from numba's global MT19937 stream, serially; that stream cannot be reproduced in
// parallel, so parity here is statistical: every (draw s, bin i) owns an independent
// Philox4x32-10 counter stream keyed by (seed, s*T+i).  lam < 10 uses Knuth's product method, lam >= 10 the
// PTRS transformed-rejection sampler (Hoermann 1993) -- the same two algorithms numpy/numba use.
#include <math.h>

#include "ipn_common.cuh"

namespace ipn {

struct Philox {
    uint32_t c[4];
    uint32_t k[2];
    uint32_t out[4];
    int have;
    __device__ Philox(uint64_t seed, uint64_t stream) {
        k[0] = (uint32_t)seed;
        k[1] = (uint32_t)(seed >> 32);
        c[0] = 0;
        c[1] = 0;
        c[2] = (uint32_t)stream;
        c[3] = (uint32_t)(stream >> 32);
        have = 0;
    }
    __device__ void round10() {
        uint32_t x0 = c[0], x1 = c[1], x2 = c[2], x3 = c[3];
        uint32_t k0 = k[0], k1 = k[1];
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            const uint32_t y0 = hi1 ^ x1 ^ k0, y1 = lo1, y2 = hi0 ^ x3 ^ k1, y3 = lo0;
            x0 = y0; x1 = y1; x2 = y2; x3 = y3;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
        if (++c[0] == 0) ++c[1];
    }
    // uniform double in (0, 1) from 53 random bits
    __device__ double uniform() {
        if (have == 0) {
            round10();
            have = 2;
        }
        const int o = (2 - have) * 2;
        --have;
        const uint64_t bits = ((uint64_t)out[o] << 21) ^ ((uint64_t)out[o + 1] >> 11);
        return ((double)(bits & ((1ull << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
    }
};

__device__ int64_t poisson_draw(Philox& rng, double lam) {
    if (!(lam > 0.0)) return 0;  // lam == 0 (and NaN/negative, which the reference would reject) -> 0
    if (lam < 10.0) {
        const double enlam = exp(-lam);
        int64_t x = 0;
        double prod = 1.0;
        while (true) {
            prod *= rng.uniform();
            if (prod > enlam)
                ++x;
            else
                return x;
        }
    }
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    while (true) {
        const double U = rng.uniform() - 0.5;
        const double V = rng.uniform();
        const double us = 0.5 - fabs(U);
        const double kf = floor((2.0 * a / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr) return (int64_t)kf;
        if (kf < 0.0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + kf * loglam - lgamma(kf + 1.0)) return (int64_t)kf;
    }
}

__global__ void poisson_counts_kernel(int64_t T, const double* __restrict__ exposure, int64_t S,
                                      const double* __restrict__ rates, const double* __restrict__ bkg, uint64_t seed,
                                      int64_t* __restrict__ out) {
    const int64_t n = S * T;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = idx / T, i = idx - s * T;
        Philox rng(seed, (uint64_t)idx);
        out[idx] = poisson_draw(rng, (rates[idx] + bkg[s]) * exposure[i]);
    }
}

int launch_poisson_counts(ipn_ctx* ctx, int64_t T, const double* exposure, int64_t S, const double* rates,
                          const double* bkg, uint64_t seed, int64_t* out) {
    const int64_t n = S * T;
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)ctx->sm_count * 16) blocks = (int64_t)ctx->sm_count * 16;
    poisson_counts_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(T, exposure, S, rates, bkg, seed, out);
    ctx->launches += 1;
    IPN_CUDA_CHECK(cudaGetLastError());
    return IPN_OK;
}

// ---- micro-benchmarks ------------------------------------------------------------------------------
constexpr int MB_ITERS = 4096;

__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, float a, float b) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < MB_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(256) mufu_peak_kernel(float* out, float a) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < MB_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(v[i]) : "f"(v[i] * a));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;
}

int run_umma_peak(ipn_ctx* ctx, double* result);  // umma_peak.cu

int run_microbench(ipn_ctx* ctx, int which, double* result) {
    IPN_REQUIRE(which >= 0 && which <= 2, "ipn_microbench: which must be 0 (FFMA), 1 (MUFU.EX2) or 2 (int8 UTCIMMA)");
    if (which == 2) return run_umma_peak(ctx, result);
    int rc = ctx->misc.ensure(256);
    if (rc) return rc;
    const int blocks = ctx->sm_count * 8;
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        IPN_CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
        for (int k = 0; k < 4; ++k) {
            if (which == 0)
                fma_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->misc.as<float>(), 0.999f, 0.001f);
            else
                mufu_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->misc.as<float>(), -0.5f);
            ctx->launches += 1;
        }
        IPN_CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
        IPN_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        IPN_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    const double threads = (double)blocks * 256.0;
    if (which == 0)
        *result = 4.0 * threads * MB_ITERS * 16.0 * 2.0 / (best * 1e-3) / 1e12;  // TFLOP/s
    else
        *result = 4.0 * threads * MB_ITERS * 8.0 / (best * 1e-3) / 1e12;  // Tops/s
    return IPN_OK;
}

}  // namespace ipn
