// Direct evaluation of the RFF intensity  amp_s * exp(sum_j a_sj cos(w_sj tau) + b_sj sin(w_sj tau)),
// tau = time_i - shift_s.   Drop-in arithmetic for pyipn/rff.py:5-29 and the draw loop pyipn/fit.py:651-691.
//
// One CTA handles (draw s, a tile of RFF_TILE bins); the draw's parameters are staged in shared memory once per CTA
// and broadcast to all threads.  |omega tau| reaches 1e4-1e5 rad for sky-scale delays (SURVEY.md section 7 hard part
// 3), so the phase needs ~48 bits -- but not fp64 arithmetic: on B200 the fp64 conversions (F2F.F64, 16 lanes/clk/SM)
// of a per-term fp64 reduction cost more than everything else together (first version: 82 issue cycles per term).
// Instead the phase is formed in TURNS as an fp32 double-float product:
//   w' = omega / 2 pi = wh + wl,  tau = th + tl  (split once per draw / once per bin, in fp64)
//   ph = fl(wh th),  e = fma(wh, th, -ph) (exact),  cross = wh tl + wl th + e          |cross| <~ 2^-23 |ph|
//   k = rint(ph) (magic-number add, exact),  fh = ph - k (exact),  phase mod 1 = fh + cross
// then a quadrant fold (q = rint(4 fh), y = fh - q/4 + cross, |y| <= 1/8 + eps), y -> radians with a hi/lo 2 pi and
// the fp32 minimax sine / cosine below.  Emulated accuracy at |phase| = 1e5 rad: max 8.5e-8, rms 1.8e-8, mean 3e-11
// -- the same as with the fp64 reduction.  The J-term sum is a Kahan sum in fp32, folded to fp64 for the final exp.
// Terms with |ph| >= 2^22 turns (2.6e7 rad) take an fp64 side path.
#include "rff_terms.cuh"  // rff_term, rff_term_df (sincos_turns from loglik_common.cuh)

namespace ipn {

constexpr int RFF_THREADS = 256;
constexpr int RFF_BINS_PER_THREAD = 4;
constexpr int RFF_TILE = RFF_THREADS * RFF_BINS_PER_THREAD;

template <bool kFullFp64>
__global__ void __launch_bounds__(RFF_THREADS)
rff_eval_kernel(int64_t T, const double* __restrict__ time, int64_t S0, int64_t S, int J, const double* __restrict__ omega,
                const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ shift,
                const double* __restrict__ amp, double* __restrict__ out, int64_t tiles_per_draw,
                const TcDraw* __restrict__ only) {
    // the float4 table comes first: behind 3 J doubles it would sit at 8 mod 16 bytes for odd J (misaligned LDS.128)
    extern __shared__ __align__(16) unsigned char sm_raw[];
    float4* s_f = reinterpret_cast<float4*>(sm_raw);      // [J] (wh, wl, a, b) with w in turns per unit time
    double* s_w = reinterpret_cast<double*>(s_f + J);     // [J] omega (fp64: side paths)
    double* s_a = s_w + J;
    double* s_b = s_w + 2 * J;
    __shared__ unsigned int s_wmax;                       // bits of max_j |wh|
    __shared__ float s_amp2;                              // sum_j a_j^2 + b_j^2 of the draw
    int64_t cur_s = -1;
    // Unmasked: items (draw, tile) strided over the CTAs.  Masked run behind the contraction route (loglik_tc.cu:
    // run_rff_eval_tc; only the draws it left out, TcDraw::hifi == 2): draw-major per CTA -- CTA c takes draws c, c + grid, ...
    // and one look at a draw's mode skips all its tiles.
    const int64_t n_it = only ? ((S + gridDim.x - 1) / gridDim.x) * tiles_per_draw : S * tiles_per_draw;
    for (int64_t it = only ? 0 : (int64_t)blockIdx.x; it < n_it; it += only ? 1 : (int64_t)gridDim.x) {
        int64_t sl, tile;
        if (only) {
            sl = (int64_t)blockIdx.x + (it / tiles_per_draw) * (int64_t)gridDim.x;
            tile = it % tiles_per_draw;
            if (sl >= S) break;
            if (only[sl].hifi != 2) {
                it += tiles_per_draw - 1 - tile;  // on to this CTA's next draw
                continue;
            }
        } else {
            sl = it / tiles_per_draw;
            tile = it - sl * tiles_per_draw;
        }
        const int64_t s = S0 + sl;  // draws S0 .. S0 + S - 1
        if (s != cur_s) {
            __syncthreads();
            if (threadIdx.x == 0) s_wmax = 0u;
            __syncthreads();
            for (int j = threadIdx.x; j < J; j += RFF_THREADS) {
                const double w = omega[s * J + j], aj = a[s * J + j], bj = b[s * J + j];
                s_w[j] = w;
                s_a[j] = aj;
                s_b[j] = bj;
                const double wt = w * 0.15915494309189534561;  // turns
                const float wh = (float)wt;
                s_f[j] = make_float4(wh, (float)(wt - (double)wh), (float)aj, (float)bj);
                atomicMax(&s_wmax, __float_as_uint(fabsf(wh)));
            }
            __syncthreads();
            if (threadIdx.x < 32) {  // deterministic: lane-strided partial sums, butterfly reduction
                double q = 0.0;
                for (int j = threadIdx.x; j < J; j += 32) q += s_a[j] * s_a[j] + s_b[j] * s_b[j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
                if (threadIdx.x == 0) s_amp2 = (float)q;
            }
            __syncthreads();
            cur_s = s;
        }
        const double sh = shift[s];
        const double am = amp[s];
        double tau[RFF_BINS_PER_THREAD];
        int64_t idx[RFF_BINS_PER_THREAD];
#pragma unroll
        for (int u = 0; u < RFF_BINS_PER_THREAD; ++u) {
            idx[u] = tile * RFF_TILE + u * RFF_THREADS + threadIdx.x;
            tau[u] = (idx[u] < T) ? (time[idx[u]] - sh) : 0.0;
        }
        if (kFullFp64) {
            double acc[RFF_BINS_PER_THREAD];
#pragma unroll
            for (int u = 0; u < RFF_BINS_PER_THREAD; ++u) acc[u] = 0.0;
            for (int j = 0; j < J; ++j) {
#pragma unroll
                for (int u = 0; u < RFF_BINS_PER_THREAD; ++u) {
                    double sn, cs;
                    sincos(s_w[j] * tau[u], &sn, &cs);
                    acc[u] = fma(s_a[j], cs, fma(s_b[j], sn, acc[u]));
                }
            }
#pragma unroll
            for (int u = 0; u < RFF_BINS_PER_THREAD; ++u)
                if (idx[u] < T) out[s * T + idx[u]] = am * exp(acc[u]);
            continue;
        }
        float th[RFF_BINS_PER_THREAD], tl[RFF_BINS_PER_THREAD], sum[RFF_BINS_PER_THREAD], comp[RFF_BINS_PER_THREAD];
        double side[RFF_BINS_PER_THREAD];
        const float wmax = __uint_as_float(s_wmax);
        bool big = false;
#pragma unroll
        for (int u = 0; u < RFF_BINS_PER_THREAD; ++u) {
            th[u] = (float)tau[u];
            tl[u] = (float)(tau[u] - (double)th[u]);
            sum[u] = comp[u] = 0.0f;
            side[u] = 0.0;
            big = big || !(fabsf(th[u]) * wmax < 4194304.0f);
        }
        // Accuracy tiers by the draw's amplitude A2 = sum a^2 + b^2 (the error is absolute in log lambda and scales with it):
        //   A2 <= 12   plain fp32 terms + Kahan sum: measured max 4.7e-7 at A2 = 4.5, i.e. 7.6e-7 at 12 (prior: median 6.7)
        //   A2 <= 100  products and their sum as fp32 double-floats (exact FMA / TwoSum error terms): what remains is the
        //              sin/cos error, 1.8e-8 rms per term
        //   beyond     fp64 sincos and accumulation (also taken when the fp32 phase reduction does not apply)
        const float amp2 = s_amp2;
        if (big || !(amp2 <= 100.0f)) {
            // (the magic-number rounding below needs |ph| < 2^22 turns = 2.6e7 rad)
            for (int j = 0; j < J; ++j) {
#pragma unroll
                for (int u = 0; u < RFF_BINS_PER_THREAD; ++u) {
                    double sn, cs;
                    sincos(s_w[j] * tau[u], &sn, &cs);
                    side[u] = fma(s_a[j], cs, fma(s_b[j], sn, side[u]));
                }
            }
        } else if (amp2 > 12.0f) {
            float lo[RFF_BINS_PER_THREAD];
#pragma unroll
            for (int u = 0; u < RFF_BINS_PER_THREAD; ++u) lo[u] = 0.0f;
            for (int j = 0; j < J; ++j) {
                const float4 f = s_f[j];
#pragma unroll
                for (int u = 0; u < RFF_BINS_PER_THREAD; ++u) {
                    rff_term_df(f, th[u], tl[u], sum[u], comp[u], lo[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < RFF_BINS_PER_THREAD; ++u) side[u] = (double)lo[u];
        } else {
            for (int j = 0; j < J; ++j) {
                const float4 f = s_f[j];
#pragma unroll
                for (int u = 0; u < RFF_BINS_PER_THREAD; ++u) {
                    rff_term(f, th[u], tl[u], sum[u], comp[u]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < RFF_BINS_PER_THREAD; ++u)
            if (idx[u] < T) out[s * T + idx[u]] = am * exp(((double)sum[u] - (double)comp[u]) + side[u]);
    }
}

// max_i |time[i] - (time[0] + i dt)| with dt = (time[T - 1] - time[0]) / (T - 1): are the bins uniform?  (atomicMax on the bit
// pattern of a non-negative double)
__global__ void __launch_bounds__(256) rff_uniform_kernel(int64_t T, const double* __restrict__ time, double* __restrict__ res) {
    const double t0 = time[0], tl = time[T - 1];
    const double dt = (tl - t0) / (double)(T - 1);
    double m = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < T; i += (int64_t)gridDim.x * blockDim.x) {
        const double d = fabs(time[i] - (t0 + (double)i * dt));
        m = (d > m || d != d) ? (d != d ? __longlong_as_double(0x7ff0000000000000ll) : d) : m;  // NaN -> inf
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned long long*>(res + 2), (unsigned long long)__double_as_longlong(m));
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        res[0] = t0;
        res[1] = tl;
    }
}

int rff_uniform_dev(ipn_ctx* ctx, int64_t T, const double* d_time, double* t0, double* dt, bool* uniform) {
    *uniform = false;
    if (T < 2) return IPN_OK;
    int rc = ctx->rffs.ensure(64);
    if (rc) return rc;
    double* res = ctx->rffs.as<double>();
    IPN_CUDA_CHECK(cudaMemsetAsync(res, 0, 24, ctx->stream));
    const int grid = (int)std::min<int64_t>((T + 255) / 256, (int64_t)ctx->sm_count * 4);
    rff_uniform_kernel<<<grid, 256, 0, ctx->stream>>>(T, d_time, res);
    ctx->launches += 1;
    double h[3];
    IPN_CUDA_CHECK(cudaMemcpyAsync(h, res, 24, cudaMemcpyDeviceToHost, ctx->stream));
    IPN_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    *t0 = h[0];
    *dt = (h[1] - h[0]) / (double)(T - 1);
    *uniform = *dt > 0.0 && std::isfinite(*dt) && h[2] <= 1e-12 * (std::fabs(h[0]) + std::fabs(h[1]));
    return IPN_OK;
}

// draws S0 .. S0 + Sb - 1; tcdraw != nullptr: only those with TcDraw::hifi == 2 (the contraction route has done the others)
int launch_rff_eval_masked(ipn_ctx* ctx, int64_t T, const double* time, int64_t S0, int64_t Sb, int64_t J, const double* omega,
                           const double* a, const double* b, const double* shift, const double* amp, double* out,
                           const void* tcdraw) {
    const int64_t tiles = (T + RFF_TILE - 1) / RFF_TILE;
    const int64_t items = Sb * tiles;
    if (items == 0) return IPN_OK;
    const int64_t max_grid = (int64_t)ctx->sm_count * 8;
    const int grid = (int)(items < max_grid ? items : max_grid);
    const size_t smem = 3 * (size_t)J * sizeof(double) + (size_t)J * sizeof(float4);
    bool full64 = false;
#ifdef IPN_TC_HOOKS
    full64 = getenv("IPN_RFF_FP64") != nullptr;  // hooks build only: every draw through the fp64 side path
#endif
    const TcDraw* only = static_cast<const TcDraw*>(tcdraw);
    if (full64)
        rff_eval_kernel<true><<<grid, RFF_THREADS, smem, ctx->stream>>>(T, time, S0, Sb, (int)J, omega, a, b, shift, amp, out, tiles, only);
    else
        rff_eval_kernel<false><<<grid, RFF_THREADS, smem, ctx->stream>>>(T, time, S0, Sb, (int)J, omega, a, b, shift, amp, out, tiles, only);
    ctx->launches += 1;
    IPN_CUDA_CHECK(cudaGetLastError());
    return IPN_OK;
}

int launch_rff_eval(ipn_ctx* ctx, int64_t T, const double* time, int64_t S, int64_t J, const double* omega,
                    const double* a, const double* b, const double* shift, const double* amp, double* out) {
    return launch_rff_eval_masked(ctx, T, time, 0, S, J, omega, a, b, shift, amp, out, nullptr);
}

}  // namespace ipn
