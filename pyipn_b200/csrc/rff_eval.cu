// Direct evaluation of the RFF intensity  amp_s * exp(sum_j a_sj cos(w_sj tau) + b_sj sin(w_sj tau)),
// tau = time_i - shift_s.   Drop-in arithmetic for pyipn/rff.py:5-29 and the draw loop pyipn/fit.py:651-691.
//
// One CTA handles (draw s, a tile of RFF_TILE bins); the draw's omega/a/b vectors are staged in shared
// memory once per CTA and broadcast to all threads.  Phases are formed in fp64 (|omega tau| reaches 1e3-1e4
// rad for sky-scale delays, SURVEY.md section 7 hard part 3), reduced to [-pi, pi] in fp64 with a two-term
// 2*pi, and the sine/cosine pair is then evaluated with fp32 minimax polynomials on the reduced argument
// plus a first-order correction for the fp32 rounding of the argument.  The J-term sum is accumulated in
// fp64 so the result meets the 1e-6 relative tolerance on lambda with a wide margin.
#include "ipn_common.cuh"

namespace ipn {

constexpr int RFF_THREADS = 256;
constexpr int RFF_BINS_PER_THREAD = 2;
constexpr int RFF_TILE = RFF_THREADS * RFF_BINS_PER_THREAD;

// sin and cos of r (|r| <= pi + eps) given as fp32 hi + lo.  Quadrant fold to |y| <= pi/4 (Cody-Waite, exact
// because |q| <= 2), then near-minimax polynomials fitted in this repo (Chebyshev-node interpolation in
// y^2; truncation error 1.3e-11 / 8.4e-13).  Emulated fp32 accuracy: max 8.6e-8, rms 1.9e-8, mean < 1.1e-9.
__device__ __forceinline__ void sincos_reduced(float rh, float rl, float& s, float& c) {
    const float qf = rintf(rh * 0.636619772367581343f);
    const int q = (int)qf;
    float y = fmaf(qf, -1.57079637050628662109375f, rh);
    y = fmaf(qf, 4.37113900018624283e-8f, y);
    y += rl;
    const float y2 = y * y;
    // sin(y) = y + y^3 P(y^2)
    float ps = fmaf(y2, 2.7237618209228658e-06f, -0.00019839989971813914f);
    ps = fmaf(ps, y2, 0.008333331692152657f);
    ps = fmaf(ps, y2, -0.16666666663377125f);
    const float sy = fmaf(ps * y2, y, y);
    // cos(y) = 1 - y^2/2 + y^4 Q(y^2)
    float pc = fmaf(y2, -2.7290681690617153e-07f, 2.4800519567960354e-05f);
    pc = fmaf(pc, y2, -0.0013888887519541376f);
    pc = fmaf(pc, y2, 0.04166666666392175f);
    const float cy = fmaf(fmaf(pc, y2, -0.5f), y2, 1.0f);
    const float ss = (q & 1) ? cy : sy;
    const float cc = (q & 1) ? sy : cy;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
}

template <bool kFullFp64>
__global__ void __launch_bounds__(RFF_THREADS)
rff_eval_kernel(int64_t T, const double* __restrict__ time, int64_t S, int J, const double* __restrict__ omega,
                const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ shift,
                const double* __restrict__ amp, double* __restrict__ out, int64_t tiles_per_draw) {
    extern __shared__ double sm[];
    double* s_w = sm;
    double* s_a = sm + J;
    double* s_b = sm + 2 * J;
    const int64_t n_items = S * tiles_per_draw;
    int64_t cur_s = -1;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int64_t s = item / tiles_per_draw;
        const int64_t tile = item - s * tiles_per_draw;
        if (s != cur_s) {
            __syncthreads();
            for (int j = threadIdx.x; j < J; j += RFF_THREADS) {
                s_w[j] = omega[s * J + j];
                s_a[j] = a[s * J + j];
                s_b[j] = b[s * J + j];
            }
            __syncthreads();
            cur_s = s;
        }
        const double sh = shift[s];
        const double am = amp[s];
        double tau[RFF_BINS_PER_THREAD];
        double acc[RFF_BINS_PER_THREAD];
        int64_t idx[RFF_BINS_PER_THREAD];
#pragma unroll
        for (int u = 0; u < RFF_BINS_PER_THREAD; ++u) {
            idx[u] = tile * RFF_TILE + u * RFF_THREADS + threadIdx.x;
            tau[u] = (idx[u] < T) ? (time[idx[u]] - sh) : 0.0;
            acc[u] = 0.0;
        }
        for (int j = 0; j < J; ++j) {
            const double w = s_w[j], aj = s_a[j], bj = s_b[j];
#pragma unroll
            for (int u = 0; u < RFF_BINS_PER_THREAD; ++u) {
                const double ph = w * tau[u];
                if (kFullFp64) {
                    double sn, cs;
                    sincos(ph, &sn, &cs);
                    acc[u] = fma(aj, cs, fma(bj, sn, acc[u]));
                } else {
                    // fp64 reduction to [-pi, pi]
                    const double kq = rint(ph * 0.15915494309189534561);
                    double r = fma(kq, -6.283185307179586232, ph);
                    r = fma(kq, -2.4492935982947064e-16, r);
                    const float rh = (float)r;
                    const float rl = (float)(r - (double)rh);
                    float sn, cs;
                    sincos_reduced(rh, rl, sn, cs);
                    acc[u] = fma(aj, (double)cs, fma(bj, (double)sn, acc[u]));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < RFF_BINS_PER_THREAD; ++u)
            if (idx[u] < T) out[s * T + idx[u]] = am * exp(acc[u]);
    }
}

int launch_rff_eval(ipn_ctx* ctx, int64_t T, const double* time, int64_t S, int64_t J, const double* omega,
                    const double* a, const double* b, const double* shift, const double* amp, double* out) {
    const int64_t tiles = (T + RFF_TILE - 1) / RFF_TILE;
    const int64_t items = S * tiles;
    if (items == 0) return IPN_OK;
    const int64_t max_grid = (int64_t)ctx->sm_count * 8;
    const int grid = (int)(items < max_grid ? items : max_grid);
    const size_t smem = 3 * (size_t)J * sizeof(double);
    const bool full64 = getenv("IPN_RFF_FP64") != nullptr;
    if (full64)
        rff_eval_kernel<true><<<grid, RFF_THREADS, smem, ctx->stream>>>(T, time, S, (int)J, omega, a, b, shift, amp,
                                                                         out, tiles);
    else
        rff_eval_kernel<false><<<grid, RFF_THREADS, smem, ctx->stream>>>(T, time, S, (int)J, omega, a, b, shift, amp,
                                                                          out, tiles);
    ctx->launches += 1;
    IPN_CUDA_CHECK(cudaGetLastError());
    return IPN_OK;
}

}  // namespace ipn
