// ipn_microbench(ctx, 2): the int8 tensor-pipe peak of this GPU as the likelihood kernel can use it, measured in-process.
// One CTA per SM, one elected thread issuing tcgen05.mma kind::i8 (M = 128, N = 32, K = 32, A from tensor memory, B from
// shared memory, four accumulators cycled) back to back -- the instruction mix of loglik_tc_kernel's MMA issuers with
// nothing else on the SM.  Timed with CUDA events over the whole grid, so the figure is at the clock the chip sustains
// under this load (what bench.py's roofline.frac is quoted against), not cycles.  tools/umma_bench.cu is the stand-alone
// study (other N, A in shared memory, interference) this was cut from.
#include "ipn_common.cuh"

namespace ipn {

namespace {

__device__ __forceinline__ uint32_t up_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int UP_N = 32;
constexpr int UP_MMAS_PER_ROUND = 64;

__global__ void __launch_bounds__(64, 1) umma_peak_kernel(int rounds, int* sink) {
    __shared__ __align__(128) uint8_t sA[128 * 32];        // one K = 32 slab of A (copied into tensor memory)
    __shared__ __align__(128) uint8_t sB[4 * UP_N * 32];   // four B tiles, cycled
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < (int)sizeof(sA); i += blockDim.x) sA[i] = (uint8_t)(i * 7 + 1);
    for (int i = threadIdx.x; i < (int)sizeof(sB); i += blockDim.x) sB[i] = (uint8_t)(i * 5 + 3);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(up_smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(up_smem_u32(&tslot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tslot;
    if (warp == 1) {
        // D = S32, A = B = signed 8-bit, K-major, N = 32, M = 128; K-major no-swizzle descriptors (LBO = rows x 16 B, SBO = 128 B)
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(UP_N >> 3) << 17) | (8u << 24);
        const uint64_t ad = (uint64_t)((up_smem_u32(sA) >> 4) & 0x3FFFu) | ((uint64_t)((128 * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
        const uint64_t bd = (uint64_t)((up_smem_u32(sB) >> 4) & 0x3FFFu) | ((uint64_t)((UP_N * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
        uint32_t elected;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
        if (elected) {
            asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tm), "l"(ad) : "memory");
            for (int r = 0; r < rounds; ++r) {
#pragma unroll
                for (int i = 0; i < UP_MMAS_PER_ROUND; ++i) {
                    const uint64_t b = bd + (uint64_t)((i & 3) * (UP_N * 32 / 16));
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(
                                     tm + 64 + (i & 3) * UP_N),
                                 "r"(tm), "l"(b), "r"(idesc)
                                 : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(up_smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        uint32_t done = 0;
        for (long long spin = 0; !done && spin < (1ll << 34); ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done)
                         : "r"(up_smem_u32(&bar))
                         : "memory");
        if (!done && sink) sink[0] = -1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u) : "memory");
}

}  // namespace

// result: int8 TOP/s (2 ops per multiply-add) over the whole chip
int run_umma_peak(ipn_ctx* ctx, double* result) {
    int rc = ctx->misc.ensure(256);
    if (rc) return rc;
    const int rounds = 4096;  // 262144 MMAs per SM: ~2.2 ms at 16 cycles each
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        IPN_CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
        umma_peak_kernel<<<ctx->sm_count, 64, 0, ctx->stream>>>(rounds, ctx->misc.as<int>());
        ctx->launches += 1;
        IPN_CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
        IPN_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        IPN_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    const double macs = (double)ctx->sm_count * rounds * UP_MMAS_PER_ROUND * 128.0 * UP_N * 32.0;
    *result = 2.0 * macs / (best * 1e-3) / 1e12;
    return IPN_OK;
}

}  // namespace ipn
