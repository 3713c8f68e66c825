// Photon arrival times of a (multi-)Norris pulse over a linear background by thinning, on the device
// (pyipn/possion_gen.py:9-62 pulse shape, :159-242 generators).
//
// The reference walks a homogeneous process of rate fmax sequentially (exponential steps) and keeps a candidate
// with probability rate(t)/fmax.  The same point process is obtained in parallel: N ~ Poisson(fmax (tstop-tstart))
// candidates placed i.i.d. uniformly, each thinned independently -- every candidate owns a Philox4x32-10 counter
// stream keyed by (seed, index), so the result is deterministic for a seed and independent of the launch shape.
// Parity with the reference's MT19937 stream is statistical (same law), not draw-for-draw; the reference's quirk of
// emitting `tstart` itself as the first "arrival" is reproduced.  Accepted times are compacted and sorted (cub).
#include <cub/device/device_radix_sort.cuh>
#include <math.h>

#include "ipn_common.cuh"

namespace ipn {

struct Pulses {
    int n;
    double K[8], t_start[8], t_rise[8], t_decay[8];
    double slope, intercept;  // linear background rate intercept + slope t (possion_gen.py:209-242)
};

__device__ __forceinline__ double rate_at(const Pulses& p, double x) {
    double out = p.intercept + p.slope * x;
    for (int i = 0; i < p.n; ++i) {
        if (x > p.t_start[i]) {
            const double d = x - p.t_start[i];
            out += p.K[i] * exp(2.0 * sqrt(p.t_rise[i] / p.t_decay[i])) * exp((-p.t_rise[i] / d) - d / p.t_decay[i]);
        }
    }
    return out;
}

__device__ __forceinline__ void philox2(uint64_t seed, uint64_t ctr, double& u0, double& u1) {
    uint32_t x0 = (uint32_t)ctr, x1 = (uint32_t)(ctr >> 32), x2 = 0x1f83d9abu, x3 = 0x5be0cd19u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
        const uint32_t y0 = hi1 ^ x1 ^ k0, y1 = lo1, y2 = hi0 ^ x3 ^ k1, y3 = lo0;
        x0 = y0; x1 = y1; x2 = y2; x3 = y3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    const uint64_t a = ((uint64_t)x0 << 21) ^ ((uint64_t)x1 >> 11), b = ((uint64_t)x2 << 21) ^ ((uint64_t)x3 >> 11);
    u0 = ((double)(a & ((1ull << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
    u1 = ((double)(b & ((1ull << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
}

// fmax over the reference's 1000-point grid linspace(tstart, tstop + 1, 1000)  (possion_gen.py:171-183)
__global__ void thin_fmax_kernel(Pulses p, double tstart, double tstop, double* out /* [0] = fmax */) {
    __shared__ double red[32];
    double m = 0.0;
    for (int i = threadIdx.x; i < 1000; i += blockDim.x) m = fmax(m, rate_at(p, tstart + (tstop + 1.0 - tstart) * (double)i / 999.0));
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
        out[0] = m;
    }
}

__global__ void thin_candidates_kernel(Pulses p, double tstart, double tstop, const double* __restrict__ fmax_p, int64_t n_cand,
                                       uint64_t seed, double* __restrict__ times, unsigned long long* __restrict__ n_out,
                                       int64_t capacity) {
    const double fm = fmax_p[0];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_cand; i += (int64_t)gridDim.x * blockDim.x) {
        double u0, u1;
        philox2(seed, (uint64_t)i, u0, u1);
        const double t = tstart + (tstop - tstart) * u0;
        if (u1 * fm <= rate_at(p, t)) {
            const unsigned long long slot = atomicAdd(n_out, 1ull);
            if ((int64_t)slot < capacity) times[slot] = t;
        }
    }
}

int run_thinning(ipn_ctx* ctx, const Pulses& p, double tstart, double tstop, uint64_t seed, int64_t n_cand, double* d_fmax,
                 double* d_times, double* d_sorted, unsigned long long* d_count, int64_t capacity, bool sort_only,
                 int64_t n_sort) {
    if (!sort_only) {
        IPN_CUDA_CHECK(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), ctx->stream));
        int64_t blocks = (n_cand + 255) / 256;
        if (blocks > (int64_t)ctx->sm_count * 16) blocks = (int64_t)ctx->sm_count * 16;
        if (blocks < 1) blocks = 1;
        thin_candidates_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(p, tstart, tstop, d_fmax, n_cand, seed, d_times, d_count, capacity);
        ctx->launches += 1;
        IPN_CUDA_CHECK(cudaGetLastError());
        return IPN_OK;
    }
    if (n_sort <= 0) return IPN_OK;
    size_t tmp_bytes = 0;
    IPN_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_times, d_sorted, (int)n_sort, 0, 64, ctx->stream));
    int rc = ctx->perm.ensure(tmp_bytes + 64);
    if (rc) return rc;
    IPN_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(ctx->perm.p, tmp_bytes, d_times, d_sorted, (int)n_sort, 0, 64, ctx->stream));
    ctx->launches += 3;
    return IPN_OK;
}

int launch_thin_fmax(ipn_ctx* ctx, const Pulses& p, double tstart, double tstop, double* d_fmax) {
    thin_fmax_kernel<<<1, 256, 0, ctx->stream>>>(p, tstart, tstop, d_fmax);
    ctx->launches += 1;
    IPN_CUDA_CHECK(cudaGetLastError());
    return IPN_OK;
}

}  // namespace ipn

using namespace ipn;

extern "C" int ipn_thinning_events(ipn_ctx* ctx, double tstart, double tstop, int64_t n_pulses, const double* K,
                                   const double* t_start, const double* t_rise, const double* t_decay, double bkg_slope,
                                   double bkg_intercept, uint64_t seed, int64_t capacity, double* out_times,
                                   int64_t* out_count) {
    if (!ctx) {
        set_error("ipn_ctx is NULL");
        return IPN_ERR_INVALID_ARGUMENT;
    }
    IPN_CUDA_CHECK(cudaSetDevice(ctx->device));
    IPN_REQUIRE(tstop > tstart && n_pulses >= 0 && n_pulses <= 8 && capacity >= 1 && out_times && out_count,
                "need tstop > tstart, 0 <= n_pulses <= 8, capacity >= 1 and non-NULL outputs");
    IPN_REQUIRE(n_pulses == 0 || (K && t_start && t_rise && t_decay), "NULL pulse parameter array");
    Pulses p;
    p.n = (int)n_pulses;
    for (int i = 0; i < p.n; ++i) {
        IPN_REQUIRE(K[i] >= 0.0 && t_rise[i] > 0.0 && t_decay[i] > 0.0, "pulse %d: need K >= 0, t_rise > 0, t_decay > 0", i);
        p.K[i] = K[i];
        p.t_start[i] = t_start[i];
        p.t_rise[i] = t_rise[i];
        p.t_decay[i] = t_decay[i];
    }
    p.slope = bkg_slope;
    p.intercept = bkg_intercept;
    int rc = ctx->io[0].ensure(16);
    if (rc) return rc;
    double* d_fmax = ctx->io[0].as<double>();
    if ((rc = launch_thin_fmax(ctx, p, tstart, tstop, d_fmax))) return rc;
    double fmax = 0.0;
    IPN_CUDA_CHECK(cudaMemcpyAsync(&fmax, d_fmax, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    IPN_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    out_times[0] = tstart;  // the reference appends tstart itself (possion_gen.py:192, 230)
    *out_count = 1;
    if (!(fmax > 0.0)) {
        *out_count = 0;  // possion_gen.py:188-190: no expected counts -> empty array
        return IPN_OK;
    }
    IPN_REQUIRE(isfinite(fmax), "rate is not finite on [tstart, tstop]");
    // number of candidates of the homogeneous process: Poisson(fmax * span), drawn on the host from the seed
    const double mean = fmax * (tstop - tstart);
    IPN_REQUIRE(mean < 4.0e9, "more than 4e9 candidate events: split the time range");
    uint64_t z = seed + 0x9E3779B97F4A7C15ull;  // splitmix64 -> normal approximation is exact enough? no: use PTRS below
    auto next_u = [&z]() {
        z += 0x9E3779B97F4A7C15ull;
        uint64_t x = z;
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        x ^= x >> 31;
        return ((double)(x >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    };
    int64_t n_cand = 0;
    if (mean < 10.0) {
        const double enlam = exp(-mean);
        double prod = 1.0;
        while (true) {
            prod *= next_u();
            if (prod > enlam) ++n_cand; else break;
        }
    } else {  // PTRS (Hoermann 1993)
        const double slam = sqrt(mean), loglam = log(mean), b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
        const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
        while (true) {
            const double U = next_u() - 0.5, V = next_u(), us = 0.5 - fabs(U);
            const double kf = floor((2.0 * a / us + b) * U + mean + 0.43);
            if (us >= 0.07 && V <= vr) { n_cand = (int64_t)kf; break; }
            if (kf < 0.0 || (us < 0.013 && V > us)) continue;
            if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -mean + kf * loglam - lgamma(kf + 1.0)) { n_cand = (int64_t)kf; break; }
        }
    }
    if (n_cand == 0) return IPN_OK;
    const int64_t cap_dev = n_cand;  // accepted <= candidates
    if ((rc = ctx->io[1].ensure((size_t)cap_dev * 8))) return rc;
    if ((rc = ctx->io[2].ensure((size_t)cap_dev * 8))) return rc;
    if ((rc = ctx->io[3].ensure(16))) return rc;
    double* d_times = ctx->io[1].as<double>();
    double* d_sorted = ctx->io[2].as<double>();
    unsigned long long* d_count = ctx->io[3].as<unsigned long long>();
    if ((rc = run_thinning(ctx, p, tstart, tstop, seed, n_cand, d_fmax, d_times, d_sorted, d_count, cap_dev, false, 0))) return rc;
    unsigned long long n_acc = 0;
    IPN_CUDA_CHECK(cudaMemcpyAsync(&n_acc, d_count, sizeof(n_acc), cudaMemcpyDeviceToHost, ctx->stream));
    IPN_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if ((int64_t)n_acc + 1 > capacity) {
        set_error("output capacity %lld too small for %llu events", (long long)capacity, n_acc + 1);
        *out_count = (int64_t)n_acc + 1;
        return IPN_ERR_INVALID_ARGUMENT;
    }
    if (n_acc > 0) {
        if ((rc = run_thinning(ctx, p, tstart, tstop, seed, n_cand, d_fmax, d_times, d_sorted, d_count, cap_dev, true, (int64_t)n_acc)))
            return rc;
        IPN_CUDA_CHECK(cudaMemcpyAsync(out_times + 1, d_sorted, (size_t)n_acc * 8, cudaMemcpyDeviceToHost, ctx->stream));
        IPN_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    *out_count = (int64_t)n_acc + 1;
    return IPN_OK;
}
