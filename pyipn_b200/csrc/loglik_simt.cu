// FP32-pipe implementation of the delay-grid log-likelihood (the contraction form the north star names).
//
//   x[i,g] = c_s + sum_k Phi_s[i,k] W_s[k,g]        k < 2J   (angle-addition identity)
//   R[g]   = sum_i n_i log2(1 + 2^x) - kappa_i 2^x
//
// Work item = (draw s, tile of BM = 64 bins).  A persistent CTA (128 threads, 3 CTAs/SM so that one CTA's
// MUFU/fp64 epilogue overlaps another's FFMA main loop) builds the Phi tile once in shared memory
// (fp64 phase -> fp32 sincos, k-major so fragment loads are conflict-free LDS.128), then sweeps all delay
// tiles of BN = 128 delays, streaming the draw's W table (fp32, L2-resident: 2J x G x 4 B = 3.3 MB at the
// bench shape) through a 4-stage cp.async ring in chunks of KC = 8 k-rows.  Each thread owns an 8 x 8
// register tile (64 FFMA per 4 LDS.128).  The epilogue turns the 64 exponents into Poisson terms (2 MUFU +
// ~10 FP32 ops each), sums its 8 bins in fp64, shuffle-reduces over the 8 bin-threads of the warp and issues
// one fp64 RED per delay into R[s][g].
//
// PRECISE mode streams a second table with the fp32 rounding residual of W first (W = hi + lo, 2x the FFMAs):
// the rounding error of W is the same for every bin of a (delay, draw) pair, so it does not average out over
// the 1e5-bin sum the way the per-bin rounding of Phi and of the accumulation does (error budget).
#include "loglik_common.cuh"

namespace ipn {

namespace simt {
constexpr int BM = 64;
constexpr int BN = 128;
constexpr int KC = 8;
constexpr int NST = 4;
constexpr int NTHREADS = 128;
}  // namespace simt

struct SimtParams {
    int64_t T;
    const double* time;
    const double* exposure;
    const int64_t* counts;
    int64_t S0;  // first draw of the batch
    int Sb;      // draws in the batch
    int J;
    int KP;      // 2J rounded up to a multiple of KC
    const double* omega;  // [S][J]
    const DrawConst* dc;  // [S]
    const float* W;       // [Sb][KT][Gpad]   KT = KP (fast) or 2 KP (precise: lo rows, then hi rows)
    int64_t G, Gpad;
    double* R;            // [Sb][Gpad]
    int64_t n_bin_tiles;
};

// ---- delay table:  W[s][2j][g] = log2e (a cos(w D) - b sin(w D)),  W[s][2j+1][g] = log2e (a sin(w D) + b cos(w D)) ----
template <bool PRECISE>
__global__ void wtab_kernel(int64_t S0, int Sb, int J, int KP, const double* __restrict__ omega,
                            const double* __restrict__ a, const double* __restrict__ b, int64_t G, int64_t Gpad,
                            const double* __restrict__ delays, float* __restrict__ W) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int jrow = blockIdx.y;  // 0 .. KP/2-1 (pairs of rows)
    const int sl = blockIdx.z;
    if (g >= Gpad) return;
    const int KT = PRECISE ? 2 * KP : KP;
    double wc = 0.0, ws = 0.0;
    if (jrow < J && g < G) {
        const int64_t s = S0 + sl;
        const double w = omega[s * J + jrow], aj = a[s * J + jrow], bj = b[s * J + jrow];
        double sn, cs;
        sincos(w * delays[g], &sn, &cs);
        wc = kLog2e * (aj * cs - bj * sn);
        ws = kLog2e * (aj * sn + bj * cs);
    }
    float* base = W + (int64_t)sl * KT * Gpad;
    const float wch = (float)wc, wsh = (float)ws;
    if (PRECISE) {
        base[(int64_t)(2 * jrow) * Gpad + g] = (float)(wc - (double)wch);
        base[(int64_t)(2 * jrow + 1) * Gpad + g] = (float)(ws - (double)wsh);
        base[(int64_t)(KP + 2 * jrow) * Gpad + g] = wch;
        base[(int64_t)(KP + 2 * jrow + 1) * Gpad + g] = wsh;
    } else {
        base[(int64_t)(2 * jrow) * Gpad + g] = wch;
        base[(int64_t)(2 * jrow + 1) * Gpad + g] = wsh;
    }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

template <bool PRECISE>
__global__ void __launch_bounds__(simt::NTHREADS, 3) loglik_simt_kernel(const SimtParams p) {
    using namespace simt;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int KP = p.KP;
    const int KT = PRECISE ? 2 * KP : KP;
    float* phi = reinterpret_cast<float*>(smem_raw);                  // [KP][BM]
    float* wst = phi + (size_t)KP * BM;                               // [NST][KC][BN]
    double* s_omega = reinterpret_cast<double*>(wst + NST * KC * BN);  // [J]
    double* s_t = s_omega + p.J;                                       // [BM]
    float* s_n = reinterpret_cast<float*>(s_t + BM);                   // [BM]
    float* s_khi = s_n + BM;
    float* s_klo = s_khi + BM;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int ty = lane & 7;               // bin-direction thread (8 of them, in-warp)
    const int tx = warp * 4 + (lane >> 3);  // delay-direction thread (16 of them)

    const int NCH = KT / KC;                      // chunks per delay tile
    const int NT = (int)(p.Gpad / BN);            // delay tiles
    const int total_chunks = NCH * NT;
    const int64_t n_items = (int64_t)p.Sb * p.n_bin_tiles;

    int cur_sl = -1;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int sl = (int)(item / p.n_bin_tiles);
        const int64_t btile = item - (int64_t)sl * p.n_bin_tiles;
        const int64_t s = p.S0 + sl;
        __syncthreads();  // previous item fully consumed (phi, meta, ring)
        if (sl != cur_sl) {
            for (int j = tid; j < p.J; j += NTHREADS) s_omega[j] = p.omega[s * p.J + j];
            cur_sl = sl;
        }
        const DrawConst dcs = p.dc[s];
        if (tid < BM) {
            const int64_t i = btile * BM + tid;
            const bool valid = i < p.T;
            const double kap = valid ? p.exposure[i] * dcs.kap : 0.0;
            const float khi = (float)kap;
            s_t[tid] = valid ? p.time[i] : 0.0;
            s_n[tid] = valid ? (float)p.counts[i] : 0.0f;
            s_khi[tid] = khi;
            s_klo[tid] = (float)(kap - (double)khi);
        }
        __syncthreads();
        // ---- Phi tile: rows 2j = cos(w_j t_i), 2j+1 = sin(w_j t_i); padded rows / invalid bins are zero ----
        {
            const int bin = tid & (BM - 1);
            const bool valid = btile * BM + bin < p.T;
            const double t = s_t[bin];
            for (int j = tid / BM; j < KP / 2; j += NTHREADS / BM) {
                float sn = 0.0f, cs = 0.0f;
                if (valid && j < p.J) sincos_phase(s_omega[j] * t, sn, cs);
                phi[(2 * j) * BM + bin] = cs;
                phi[(2 * j + 1) * BM + bin] = sn;
            }
        }
        // (visibility of phi is guaranteed by the __syncthreads inside the chunk loop before first use)

        const float* Wd = p.W + (int64_t)sl * KT * p.Gpad;
        auto load_chunk = [&](int c) {
            const int gt = c / NCH, kc = c - gt * NCH;
            const float* src = Wd + (int64_t)(kc * KC) * p.Gpad + (int64_t)gt * BN;
            float* dst = wst + (c % NST) * (KC * BN);
#pragma unroll
            for (int r = 0; r < (KC * BN / 4) / NTHREADS; ++r) {
                const int piece = tid + r * NTHREADS;
                const int row = piece / (BN / 4), col4 = piece % (BN / 4);
                cp_async16(dst + row * BN + col4 * 4, src + (int64_t)row * p.Gpad + col4 * 4);
            }
        };
#pragma unroll
        for (int c = 0; c < NST - 1; ++c) {
            if (c < total_chunks) load_chunk(c);
            cp_async_commit();
        }

        float acc[8][8];
        for (int c = 0; c < total_chunks; ++c) {
            cp_async_wait<NST - 2>();
            __syncthreads();
            if (c + NST - 1 < total_chunks) load_chunk(c + NST - 1);
            cp_async_commit();
            const int gt = c / NCH, kc = c - gt * NCH;
            if (kc == 0) {
                const float init = PRECISE ? dcs.c_lo : 0.0f;
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc[r][q] = init;
            }
            int krow = kc * KC;
            if (PRECISE && krow >= KP) krow -= KP;
            const float* ph = phi + (size_t)krow * BM + ty * 4;
            const float* wp = wst + (c % NST) * (KC * BN) + tx * 4;
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(ph + kk * BM);
                const float4 a1 = *reinterpret_cast<const float4*>(ph + kk * BM + 32);
                const float4 b0 = *reinterpret_cast<const float4*>(wp + kk * BN);
                const float4 b1 = *reinterpret_cast<const float4*>(wp + kk * BN + 64);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc[r][q] = fmaf(av[r], bv[q], acc[r][q]);
            }
            if (kc == NCH - 1) {
                // ---- epilogue for delay tile gt ----
                double cs[8];
                float lo[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    cs[q] = 0.0;
                    lo[q] = 0.0f;
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int bin = (r < 4) ? (ty * 4 + r) : (32 + ty * 4 + (r - 4));
                    const float n = s_n[bin], khi = s_khi[bin], klo = s_klo[bin];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float l;
                        cs[q] += (double)poisson_term(exp2_hilo_clamped(acc[r][q] + dcs.c_hi, 0.0f), n, n * 1.44269504f, khi, klo, l);
                        lo[q] += l;
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) cs[q] += (double)lo[q];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    cs[q] += __shfl_xor_sync(0xffffffffu, cs[q], 1);
                    cs[q] += __shfl_xor_sync(0xffffffffu, cs[q], 2);
                    cs[q] += __shfl_xor_sync(0xffffffffu, cs[q], 4);
                }
                if (ty == 0) {
                    double* Rrow = p.R + (int64_t)sl * p.Gpad + (int64_t)gt * BN;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int col = (q < 4) ? (tx * 4 + q) : (64 + tx * 4 + (q - 4));
                        if ((int64_t)gt * BN + col < p.G) atomicAdd(Rrow + col, cs[q]);
                    }
                }
            }
        }
        cp_async_wait<0>();
    }
}

int run_loglik_grid_simt(ipn_ctx* ctx, int64_t T, const double* time, const double* exposure, const int64_t* counts,
                         int64_t S, int64_t J, const double* omega, const double* a, const double* b,
                         const DrawConst* d_dc, int64_t G, const double* delays, double* ll, bool accumulate) {
    using namespace simt;
    const bool precise = ctx->precise_w != 0;
    const int KP = (int)(((2 * J + KC - 1) / KC) * KC);
    const int KT = precise ? 2 * KP : KP;
    const int64_t Gpad = ((G + BN - 1) / BN) * BN;
    const int64_t n_bin_tiles = (T + BM - 1) / BM;

    // draws per batch: bound the W scratch to ~1 GiB unless the caller chose a batch size
    const size_t w_per_draw = (size_t)KT * Gpad * sizeof(float);
    int64_t Sbatch = ctx->draw_batch > 0 ? ctx->draw_batch : (int64_t)((1ull << 30) / w_per_draw);
    if (Sbatch < 1) Sbatch = 1;
    if (Sbatch > S) Sbatch = S;
    if (Sbatch > 65535) Sbatch = 65535;  // gridDim.z of the table kernel
    ctx->last_draw_batch = Sbatch;

    int rc = ctx->wtab.ensure((size_t)Sbatch * w_per_draw);
    if (rc) return rc;
    rc = ctx->accum.ensure((size_t)Sbatch * Gpad * sizeof(double));
    if (rc) return rc;
    float* d_W = ctx->wtab.as<float>();
    double* d_R = ctx->accum.as<double>();

    const size_t smem = (size_t)KP * BM * sizeof(float) + (size_t)NST * KC * BN * sizeof(float) +
                        (size_t)J * sizeof(double) + BM * sizeof(double) + 3 * BM * sizeof(float);
    // per launch, not cached: the attribute belongs to the (function, device) pair and contexts on several GPUs share this code
    if (precise)
        IPN_CUDA_CHECK(cudaFuncSetAttribute(loglik_simt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    else
        IPN_CUDA_CHECK(cudaFuncSetAttribute(loglik_simt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (smem > 200 * 1024) {
        set_error("J = %lld needs %zu bytes of shared memory per CTA (limit 200 KiB)", (long long)J, smem);
        return IPN_ERR_UNSUPPORTED;
    }
    int ctas_per_sm = (int)((227 * 1024) / (smem + 1024));
    if (ctas_per_sm > 3) ctas_per_sm = 3;
    if (ctas_per_sm < 1) ctas_per_sm = 1;

    for (int64_t S0 = 0; S0 < S; S0 += Sbatch) {
        const int Sb = (int)((S - S0 < Sbatch) ? (S - S0) : Sbatch);
        IPN_CUDA_CHECK(cudaMemsetAsync(d_R, 0, (size_t)Sb * Gpad * sizeof(double), ctx->stream));
        dim3 tgrid((unsigned)((Gpad + 127) / 128), (unsigned)(KP / 2), (unsigned)Sb);
        if (precise)
            wtab_kernel<true><<<tgrid, 128, 0, ctx->stream>>>(S0, Sb, (int)J, KP, omega, a, b, G, Gpad, delays, d_W);
        else
            wtab_kernel<false><<<tgrid, 128, 0, ctx->stream>>>(S0, Sb, (int)J, KP, omega, a, b, G, Gpad, delays, d_W);
        ctx->launches += 1;
        IPN_CUDA_CHECK(cudaGetLastError());

        SimtParams p;
        p.T = T; p.time = time; p.exposure = exposure; p.counts = counts;
        p.S0 = S0; p.Sb = Sb; p.J = (int)J; p.KP = KP; p.omega = omega; p.dc = d_dc; p.W = d_W;
        p.G = G; p.Gpad = Gpad; p.R = d_R; p.n_bin_tiles = n_bin_tiles;
        const int64_t items = (int64_t)Sb * n_bin_tiles;
        int64_t grid = (int64_t)ctx->sm_count * ctas_per_sm;
        if (grid > items) grid = items;
        if (grid > 0) {
            main_kernel_begin(ctx);
            if (precise)
                loglik_simt_kernel<true><<<(int)grid, NTHREADS, smem, ctx->stream>>>(p);
            else
                loglik_simt_kernel<false><<<(int)grid, NTHREADS, smem, ctx->stream>>>(p);
            main_kernel_end(ctx);
            ctx->launches += 1;
            IPN_CUDA_CHECK(cudaGetLastError());
        }
        rc = finalize_loglik(ctx, S0, Sb, S, G, Gpad, d_R, d_dc, ll, accumulate);
        if (rc) return rc;
    }
    return IPN_OK;
}

}  // namespace ipn
