// Per-call and per-draw constants of the delay-grid log-likelihood, the finalisation pass, the sky->delay
// mapping and the dispatcher between the two contraction back-ends.
#include <math.h>

#include "loglik_common.cuh"

namespace ipn {

// ---- sum n, sum n log e, sum e, sum lgamma(n+1) over the bins (fp64) ---------------------------------
__global__ void bin_const_kernel(int64_t T, const double* __restrict__ exposure, const int64_t* __restrict__ counts,
                                 double* __restrict__ out4) {
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < T; i += (int64_t)gridDim.x * blockDim.x) {
        const double n = (double)counts[i];
        const double e = exposure[i];
        v[0] += n;
        if (n > 0.0) v[1] += n * log(e);  // e == 0 with n > 0 -> -inf, as Stan's poisson_lpmf(n > 0 | 0)
        v[2] += e;
        v[3] += lgamma(n + 1.0);
    }
    __shared__ double red[4][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        double x = v[q];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) red[q][warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double x = lane < nw ? red[q][lane] : 0.0;
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) atomicAdd(&out4[q], x);
        }
    }
}

__global__ void draw_const_kernel(int64_t S, const double* __restrict__ amp, const double* __restrict__ bkg,
                                  const double* __restrict__ binc, DrawConst* __restrict__ dc) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= S) return;
    const double A = amp[s], B = bkg[s];
    DrawConst d;
    // n log(e B) - e B - lgamma(n+1) summed over bins; N == 0 keeps log B out (B > 0 is validated upstream)
    d.C = binc[1] + (binc[0] > 0.0 ? binc[0] * log(B) : 0.0) - B * binc[2] - binc[3];
    d.kap = B / kLn2;
    const double c = log2(A / B);  // A == 0 -> -inf -> u == 0: background-only likelihood
    d.c_hi = (float)c;
    d.c_lo = isfinite(c) ? (float)(c - (double)d.c_hi) : 0.0f;
    dc[s] = d;
}

// ll[g*S + s] (+)= C_s + ln2 * R[(s-S0)*Gpad + g]
__global__ void finalize_kernel(int64_t S0, int64_t Sb, int64_t S, int64_t G, int64_t Gpad, const double* __restrict__ R,
                                const DrawConst* __restrict__ dc, double* __restrict__ ll, int accumulate) {
    // tile transpose through shared memory so both the R reads (g fastest) and ll writes (s fastest) coalesce
    __shared__ double tile[32][33];
    const int64_t g0 = (int64_t)blockIdx.x * 32, s0 = (int64_t)blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t sl = s0 + r, g = g0 + threadIdx.x;
        tile[r][threadIdx.x] = (sl < Sb && g < G) ? R[sl * Gpad + g] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t g = g0 + r, sl = s0 + threadIdx.x;
        if (g < G && sl < Sb) {
            const int64_t s = S0 + sl;
            const double v = dc[s].C + kLn2 * tile[threadIdx.x][r];
            double* p = &ll[g * S + s];
            *p = accumulate ? (*p + v) : v;
        }
    }
}

// delays[d*P + p] = ghat[p] . (sc_pos[0] - sc_pos[d]) / c      (functions.stan:65-69,164-169)
__global__ void sky_delay_kernel(int64_t D, const double* __restrict__ sc_pos, int64_t P, const double* __restrict__ ghat,
                                 double* __restrict__ delays) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t d = blockIdx.y;
    if (p >= P) return;
    const double inv_c = 1.0 / 299792.46;
    const double dx = sc_pos[0] - sc_pos[d * 3 + 0];
    const double dy = sc_pos[1] - sc_pos[d * 3 + 1];
    const double dz = sc_pos[2] - sc_pos[d * 3 + 2];
    // same association as Stan's dot_product followed by * inv(c)
    delays[d * P + p] = (ghat[p * 3 + 0] * dx + ghat[p * 3 + 1] * dy + ghat[p * 3 + 2] * dz) * inv_c;
}

int prepare_bin_constants(ipn_ctx* ctx, int64_t T, const double* exposure, const int64_t* counts, double* d_binc) {
    IPN_CUDA_CHECK(cudaMemsetAsync(d_binc, 0, 4 * sizeof(double), ctx->stream));
    int64_t blocks = (T + 255) / 256;
    if (blocks > ctx->sm_count * 4) blocks = ctx->sm_count * 4;
    if (blocks < 1) blocks = 1;
    bin_const_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(T, exposure, counts, d_binc);
    ctx->launches += 1;
    IPN_CUDA_CHECK(cudaGetLastError());
    return IPN_OK;
}

int prepare_draw_constants(ipn_ctx* ctx, int64_t S, const double* amp, const double* bkg, const double* d_binc,
                           DrawConst* d_dc) {
    draw_const_kernel<<<(int)((S + 127) / 128), 128, 0, ctx->stream>>>(S, amp, bkg, d_binc, d_dc);
    ctx->launches += 1;
    IPN_CUDA_CHECK(cudaGetLastError());
    return IPN_OK;
}

int finalize_loglik(ipn_ctx* ctx, int64_t S0, int64_t Sb, int64_t S, int64_t G, int64_t Gpad, const double* d_R,
                    const DrawConst* d_dc, double* d_ll, bool accumulate) {
    dim3 grid((unsigned)((G + 31) / 32), (unsigned)((Sb + 31) / 32));
    dim3 block(32, 8);
    finalize_kernel<<<grid, block, 0, ctx->stream>>>(S0, Sb, S, G, Gpad, d_R, d_dc, d_ll, accumulate ? 1 : 0);
    ctx->launches += 1;
    IPN_CUDA_CHECK(cudaGetLastError());
    return IPN_OK;
}

int launch_sky_delays(ipn_ctx* ctx, int64_t D, const double* sc_pos, int64_t P, const double* ghat, double* delays) {
    dim3 grid((unsigned)((P + 255) / 256), (unsigned)D);
    sky_delay_kernel<<<grid, 256, 0, ctx->stream>>>(D, sc_pos, P, ghat, delays);
    ctx->launches += 1;
    IPN_CUDA_CHECK(cudaGetLastError());
    return IPN_OK;
}

int run_loglik_grid(ipn_ctx* ctx, int64_t T, const double* time, const double* exposure, const int64_t* counts,
                    int64_t S, int64_t J, const double* omega, const double* a, const double* b, const double* amp,
                    const double* bkg, int64_t G, const double* delays, double* ll, bool accumulate) {
    if (J > IPN_MAX_J) {
        set_error("J = %lld exceeds IPN_MAX_J = %d", (long long)J, IPN_MAX_J);
        return IPN_ERR_UNSUPPORTED;
    }
    int rc = ctx->binprep.ensure(4 * sizeof(double));
    if (rc) return rc;
    rc = ctx->draws.ensure((size_t)S * sizeof(DrawConst));
    if (rc) return rc;
    double* d_binc = ctx->binprep.as<double>();
    DrawConst* d_dc = ctx->draws.as<DrawConst>();
    rc = prepare_bin_constants(ctx, T, exposure, counts, d_binc);
    if (rc) return rc;
    rc = prepare_draw_constants(ctx, S, amp, bkg, d_binc, d_dc);
    if (rc) return rc;
    const bool use_tc = ctx->loglik_impl == 2 || (ctx->loglik_impl == 0 && tc_supported(J));
    ctx->last_impl = use_tc ? 2 : 1;
    if (use_tc) return run_loglik_grid_tc(ctx, T, time, exposure, counts, S, J, omega, a, b, d_dc, G, delays, ll, accumulate);
    return run_loglik_grid_simt(ctx, T, time, exposure, counts, S, J, omega, a, b, d_dc, G, delays, ll, accumulate);
}

}  // namespace ipn
