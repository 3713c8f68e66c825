// Sky post-processing: per-pixel log-likelihoods -> localisation products.
//
// The reference turns posterior (phi, theta) samples into a clustered KDE, rasterises it on a HEALPix multi-order map
// and cuts credible regions with MOC.from_valued_healpix_cells(uniq, prob, cumul_to=level)
// (pyipn/fit.py:193-206 _build_moc_map, :328-347 _contour_more_detectors).  With the likelihood already evaluated on
// the pixel grid (ipn_loglik_sky_grid) none of the KDE machinery is needed:
//   L[p]     = logsumexp_s ll[p, s] - log S + log_weight[p]       draw-marginal
//   prob[p]  = exp(L[p] - logZ),  logZ = logsumexp_p L[p]
//   order    = pixels by decreasing prob (stable: ties keep pixel order)
//   level[p] = cumulative prob up to and including p in that order; the c-credible region is {p : level[p] <= c}
//              plus the first pixel that crosses c (cumul_to semantics of mocpy: cells are added until the sum reaches c).
// All float64.  One warp per pixel for the draw reduction; cub for the pixel reduction, sort and scan.
#include <cub/cub.cuh>

#include "ipn_common.cuh"

namespace ipn {

__global__ void __launch_bounds__(256) sky_marginal_kernel(int64_t P, int64_t S, const double* __restrict__ ll,
                                                           const double* __restrict__ log_weight, double* __restrict__ L) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double log_s = log((double)S);
    for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < P; p += warps) {
        const double* row = ll + p * S;
        double m = -INFINITY;
        for (int64_t s = lane; s < S; s += 32) m = fmax(m, row[s]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        double acc = 0.0;
        if (m > -INFINITY && m < INFINITY) {
            for (int64_t s = lane; s < S; s += 32) acc += exp(row[s] - m);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        }
        if (lane == 0) {
            double v = (m > -INFINITY && m < INFINITY) ? m + log(acc) - log_s : m;  // all -inf stays -inf, +inf/nan propagate
            if (log_weight) v += log_weight[p];
            L[p] = v;
        }
    }
}

struct MaxOp {
    __device__ __forceinline__ double operator()(double a, double b) const { return fmax(a, b); }
};

// prob[p] = exp(L[p] - max) (unnormalised); the sum is taken by cub afterwards
__global__ void sky_exp_kernel(int64_t P, const double* __restrict__ L, const double* __restrict__ maxv, double* __restrict__ e,
                               int64_t* __restrict__ idx) {
    const double m = maxv[0];
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        e[p] = exp(L[p] - m);
        idx[p] = p;
    }
}

__global__ void sky_scale_kernel(int64_t P, double* __restrict__ e, const double* __restrict__ sum, const double* __restrict__ maxv,
                                 double* __restrict__ log_norm) {
    const double inv = 1.0 / sum[0];
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) e[p] *= inv;
    if (blockIdx.x == 0 && threadIdx.x == 0) log_norm[0] = maxv[0] + log(sum[0]);
}

__global__ void sky_scatter_kernel(int64_t P, const int64_t* __restrict__ order, const double* __restrict__ cum,
                                   double* __restrict__ level) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x)
        level[order[i]] = cum[i];
}

// All pointers are device pointers.  prob / level: P doubles, order: P int64 (pixels by decreasing prob), log_norm: 1 double.
int run_sky_credible(ipn_ctx* ctx, int64_t P, int64_t S, const double* ll, const double* log_weight, double* prob,
                     double* level, int64_t* order, double* log_norm) {
    if (P > 0x7fffffffLL) {
        set_error("at most 2^31 - 1 pixels (got %lld)", (long long)P);
        return IPN_ERR_UNSUPPORTED;
    }
    const int n = (int)P;
    int rc;
    // scratch: L[P], e_sorted[P], idx[P], cum[P], 2 scalars, cub temp
    size_t t_max = 0, t_sum = 0, t_sort = 0, t_scan = 0;
    double* dn = nullptr;
    int64_t* in64 = nullptr;
    IPN_CUDA_CHECK(cub::DeviceReduce::Reduce(nullptr, t_max, dn, dn, n, MaxOp(), -INFINITY, ctx->stream));
    IPN_CUDA_CHECK(cub::DeviceReduce::Sum(nullptr, t_sum, dn, dn, n, ctx->stream));
    IPN_CUDA_CHECK(cub::DeviceRadixSort::SortPairsDescending(nullptr, t_sort, dn, dn, in64, in64, n, 0, 64, ctx->stream));
    IPN_CUDA_CHECK(cub::DeviceScan::InclusiveSum(nullptr, t_scan, dn, dn, n, ctx->stream));
    size_t t_cub = t_max;
    if (t_sum > t_cub) t_cub = t_sum;
    if (t_sort > t_cub) t_cub = t_sort;
    if (t_scan > t_cub) t_cub = t_scan;
    const size_t nP = (size_t)P * 8;
    if ((rc = ctx->misc.ensure(4 * nP + 64 + t_cub + 256))) return rc;
    uint8_t* base = ctx->misc.as<uint8_t>();
    double* d_L = reinterpret_cast<double*>(base);
    double* d_sorted = reinterpret_cast<double*>(base + nP);
    int64_t* d_idx = reinterpret_cast<int64_t*>(base + 2 * nP);
    double* d_cum = reinterpret_cast<double*>(base + 3 * nP);
    double* d_scal = reinterpret_cast<double*>(base + 4 * nP);  // [0] max, [1] sum
    void* d_tmp = base + 4 * nP + 64;

    int blocks = (int)((P * 32 + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    if (blocks < 1) blocks = 1;
    sky_marginal_kernel<<<blocks, 256, 0, ctx->stream>>>(P, S, ll, log_weight, d_L);
    IPN_CUDA_CHECK(cub::DeviceReduce::Reduce(d_tmp, t_cub, d_L, d_scal, n, MaxOp(), -INFINITY, ctx->stream));
    int eb = (int)((P + 255) / 256);
    if (eb > ctx->sm_count * 8) eb = ctx->sm_count * 8;
    if (eb < 1) eb = 1;
    sky_exp_kernel<<<eb, 256, 0, ctx->stream>>>(P, d_L, d_scal, prob, d_idx);
    IPN_CUDA_CHECK(cub::DeviceReduce::Sum(d_tmp, t_cub, prob, d_scal + 1, n, ctx->stream));
    sky_scale_kernel<<<eb, 256, 0, ctx->stream>>>(P, prob, d_scal + 1, d_scal, log_norm);
    // stable LSD radix sort: equal probabilities keep increasing pixel index, like numpy argsort(-prob, kind="stable")
    IPN_CUDA_CHECK(cub::DeviceRadixSort::SortPairsDescending(d_tmp, t_cub, prob, d_sorted, d_idx, order, n, 0, 64, ctx->stream));
    IPN_CUDA_CHECK(cub::DeviceScan::InclusiveSum(d_tmp, t_cub, d_sorted, d_cum, n, ctx->stream));
    sky_scatter_kernel<<<eb, 256, 0, ctx->stream>>>(P, order, d_cum, level);
    ctx->launches += 10;
    IPN_CUDA_CHECK(cudaGetLastError());
    return IPN_OK;
}

}  // namespace ipn

using namespace ipn;

extern "C" int ipn_sky_credible_dev(ipn_ctx* ctx, int64_t P, int64_t S, const double* d_ll, const double* d_log_weight,
                                    double* d_prob, double* d_level, int64_t* d_order, double* d_log_norm) {
    if (!ctx) {
        set_error("ipn_ctx is NULL");
        return IPN_ERR_INVALID_ARGUMENT;
    }
    IPN_CUDA_CHECK(cudaSetDevice(ctx->device));
    IPN_REQUIRE(P >= 1 && S >= 1, "need P >= 1 and S >= 1 (got %lld, %lld)", (long long)P, (long long)S);
    IPN_REQUIRE(d_ll && d_prob && d_level && d_order && d_log_norm, "NULL device pointer");
    return run_sky_credible(ctx, P, S, d_ll, d_log_weight, d_prob, d_level, d_order, d_log_norm);
}

extern "C" int ipn_sky_credible(ipn_ctx* ctx, int64_t P, int64_t S, const double* ll, const double* log_weight, double* prob,
                                double* level, int64_t* order, double* log_norm) {
    if (!ctx) {
        set_error("ipn_ctx is NULL");
        return IPN_ERR_INVALID_ARGUMENT;
    }
    IPN_CUDA_CHECK(cudaSetDevice(ctx->device));
    IPN_REQUIRE(P >= 1 && S >= 1, "need P >= 1 and S >= 1 (got %lld, %lld)", (long long)P, (long long)S);
    IPN_REQUIRE(ll && prob && level && order && log_norm, "NULL pointer");
    bool any_finite = false;
    for (int64_t i = 0; i < P * S; ++i) {
        IPN_REQUIRE(!(ll[i] != ll[i]) && ll[i] < INFINITY, "ll[%lld] is NaN or +inf", (long long)i);
        any_finite = any_finite || ll[i] > -INFINITY;
    }
    IPN_REQUIRE(any_finite, "every log-likelihood is -inf: nothing to normalise");
    if (log_weight)
        for (int64_t p = 0; p < P; ++p) IPN_REQUIRE(!(log_weight[p] != log_weight[p]) && log_weight[p] < INFINITY, "log_weight[%lld] is NaN or +inf", (long long)p);
    int rc;
    const size_t nll = (size_t)P * S * 8, nP = (size_t)P * 8;
    if ((rc = ctx->io[0].ensure(nll))) return rc;
    if ((rc = ctx->io[1].ensure(nP))) return rc;
    if ((rc = ctx->io[2].ensure(nP))) return rc;
    if ((rc = ctx->io[3].ensure(nP))) return rc;
    if ((rc = ctx->io[4].ensure(nP))) return rc;
    if ((rc = ctx->io[5].ensure(16))) return rc;
    IPN_CUDA_CHECK(cudaMemcpyAsync(ctx->io[0].p, ll, nll, cudaMemcpyHostToDevice, ctx->stream));
    if (log_weight) IPN_CUDA_CHECK(cudaMemcpyAsync(ctx->io[1].p, log_weight, nP, cudaMemcpyHostToDevice, ctx->stream));
    rc = run_sky_credible(ctx, P, S, ctx->io[0].as<double>(), log_weight ? ctx->io[1].as<double>() : nullptr,
                          ctx->io[2].as<double>(), ctx->io[3].as<double>(), ctx->io[4].as<int64_t>(), ctx->io[5].as<double>());
    if (rc) return rc;
    IPN_CUDA_CHECK(cudaMemcpyAsync(prob, ctx->io[2].p, nP, cudaMemcpyDeviceToHost, ctx->stream));
    IPN_CUDA_CHECK(cudaMemcpyAsync(level, ctx->io[3].p, nP, cudaMemcpyDeviceToHost, ctx->stream));
    IPN_CUDA_CHECK(cudaMemcpyAsync(order, ctx->io[4].p, nP, cudaMemcpyDeviceToHost, ctx->stream));
    IPN_CUDA_CHECK(cudaMemcpyAsync(log_norm, ctx->io[5].p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    IPN_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return IPN_OK;
}
