// Event binning on the device: arrival times -> integer counts with exactly the
// semantics of LightCurve.get_binned_light_curve (pyipn/lightcurve.py:24-47): edges = np.arange(tstart, tstop, dt)
// (numpy fills float aranges as edge_i = start + i*delta with delta = (start + dt) - start, two roundings, no FMA:
// the caller passes that delta), np.histogram(times, bins=edges): half-open bins [e_i, e_{i+1}) except the
// last one, which is closed on the right; events outside [e_0, e_last] are dropped.
#include "ipn_common.cuh"

namespace ipn {

__global__ void bin_events_kernel(int64_t n_events, const double* __restrict__ times, double tstart, double dt,
                                  int64_t n_edges, unsigned long long* __restrict__ counts) {
    const int64_t n_bins = n_edges - 1;
    auto edge = [&](int64_t k) { return __dadd_rn(tstart, __dmul_rn((double)k, dt)); };
    const double e_last = edge(n_edges - 1);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_events; i += (int64_t)gridDim.x * blockDim.x) {
        const double t = times[i];
        if (!(t >= tstart) || !(t <= e_last)) continue;  // also drops NaN
        int64_t b = (int64_t)floor((t - tstart) / dt);
        if (b > n_bins - 1) b = n_bins - 1;
        if (b < 0) b = 0;
        // the division can be off by one ulp-bin: settle against the actual float64 edges, as searchsorted would
        while (b > 0 && t < edge(b)) --b;
        while (b < n_bins - 1 && t >= edge(b + 1)) ++b;
        atomicAdd(&counts[b], 1ull);
    }
}

int launch_bin_events(ipn_ctx* ctx, int64_t n_events, const double* times, double tstart, double dt, int64_t n_edges,
                      int64_t* counts) {
    IPN_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)(n_edges - 1) * sizeof(int64_t), ctx->stream));
    if (n_events > 0) {
        int64_t blocks = (n_events + 255) / 256;
        if (blocks > (int64_t)ctx->sm_count * 16) blocks = (int64_t)ctx->sm_count * 16;
        bin_events_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(n_events, times, tstart, dt, n_edges,
                                                                reinterpret_cast<unsigned long long*>(counts));
        ctx->launches += 1;
        IPN_CUDA_CHECK(cudaGetLastError());
    }
    return IPN_OK;
}

}  // namespace ipn
