// Micro-benchmark of tcgen05.mma kind::i8 issue/throughput on B200 (M = 128, K = 32) as a function of N, of where
// the A operand lives (shared memory vs tensor memory) and of how many independent accumulators are cycled.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_bench tools/umma_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}

// out[0] = cycles from first issue to completion, out[1] = cycles spent issuing
template <int N, bool TS, int NACC, bool VARY, int NOISE, int VARYA = 0, int LDW = 0>
__global__ void __launch_bounds__(320, 1) bench(int iters, long long* out, float* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    uint8_t* sA = smem;                 // 128 rows x 32 B
    uint8_t* sB = smem + 4096;          // 21 x (N rows x 32 B) when VARY
    for (int i = threadIdx.x; i < 4096 + 21 * N * 32; i += blockDim.x) smem[i] = (uint8_t)(i * 7 + 1);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tslot;
    if (warp == 1) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint64_t ad = make_desc(smem_u32(sA), 128 * 16, 128);
        const uint64_t bd = make_desc(smem_u32(sB), N * 16, 128);
        const uint32_t a_t = tm;                 // A in TMEM: 8 columns at 0
        const uint32_t acc0 = tm + 224;          // accumulators after the (up to 28) A tiles
        long long t0 = 0, t1 = 0, t2 = 0;
        if (elect_one()) {
            if (TS) asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(a_t), "l"(ad) : "memory");
            t0 = clock64();
            for (int it = 0; it < iters; it += 63) {
#pragma unroll
                for (int r = 0; r < 63; ++r) {
                    const uint64_t b = VARY ? bd + (uint64_t)((r % 21) * (N * 32 / 16)) : bd;
                    // VARYA = 1: a different A tile every MMA; 2: same A tile for 3 consecutive MMAs, then the next
                    const uint32_t at = VARYA == 1 ? a_t + (r % 28) * 8 : (VARYA == 2 ? a_t + ((r / 3) % 28) * 8 : a_t);
                    if (TS) mma_ts(acc0 + (r % NACC) * N, at, b, idesc);
                    else mma_ss(acc0 + (r % NACC) * N, ad, b, idesc);
                }
            }
            t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        // wait
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        }
        t2 = clock64();
        if (t0 != 0 && blockIdx.x == 0) {
            out[0] = t2 - t0;
            out[1] = t1 - t0;
        }
    }
    if (LDW > 0 && warp >= 2 && warp < 2 + LDW) {
        // TMEM read traffic: every warp streams x8 loads from its lane quarter while the MMA warp issues
        const uint32_t base = tm + ((uint32_t)((warp & 3) * 32) << 16) + 224;
        int acc = 0;
        long long t0 = clock64();
        const int nld = iters * 2;
        for (int it = 0; it < nld; ++it) {
            int v0, v1, v2, v3, v4, v5, v6, v7;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3), "=r"(v4), "=r"(v5), "=r"(v6), "=r"(v7)
                         : "r"(base + (it & 15) * 8));
            if ((it & 3) == 3) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += v0 ^ v7;
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        long long t1 = clock64();
        if (acc == 0x12345678) sink[1] = 1.f;
        if (warp == 2 && (threadIdx.x & 31) == 0 && blockIdx.x == 0) out[2] = (t1 - t0) * 1000 / nld;  // milli-cycles per LDTM.x8 of this warp
    } else if (warp >= 2 && warp < 2 + NOISE) {
        // compute noise on every SMSP while the MMA warp issues
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (float)(threadIdx.x + i);
        for (int it = 0; it < iters * 12; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], 0.999f, 0.001f);
        }
        float sacc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) sacc += v[i];
        if (sacc == 1.2345f) sink[0] = sacc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

template <int N, bool TS, int NACC, bool VARY = false, int NOISE = 0, int VARYA = 0, int LDW = 0>
void run(int grid, long long* d) {
    const int iters = 63 * 32;
    static float* sink = nullptr;
    if (!sink) cudaMalloc(&sink, 16);
    cudaFuncSetAttribute(bench<N, TS, NACC, VARY, NOISE, VARYA, LDW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
    bench<N, TS, NACC, VARY, NOISE, VARYA, LDW><<<grid, 320, 4096 + 21 * N * 32 + 1024>>>(iters, d, sink);
    bench<N, TS, NACC, VARY, NOISE, VARYA, LDW><<<grid, 320, 4096 + 21 * N * 32 + 1024>>>(iters, d, sink);
    long long h[3] = {0, 0, 0};
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("N=%3d A=%s nacc=%d varyB=%d noise=%d varyA=%d ldw=%d (%.1f cyc per LDTM.x8 per warp) grid=%3d : %6.1f cyc/MMA total, %5.1f cyc/MMA issue, %6.0f MAC/clk/SM  %s\n", N, TS ? "tmem" : "smem", NACC, (int)VARY, NOISE, VARYA, LDW, LDW ? h[2] / 1000.0 : 0.0, grid,
           (double)h[0] / iters, (double)h[1] / iters, 128.0 * N * 32.0 * iters / (double)h[0], e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* d;
    cudaMalloc(&d, 64);
    for (int grid : {1, 148}) {
        run<32, true, 1>(grid, d);  run<32, true, 2>(grid, d);  run<32, true, 4>(grid, d);  run<32, true, 6>(grid, d);
        run<32, false, 1>(grid, d); run<32, false, 4>(grid, d);
        run<64, true, 1>(grid, d);  run<64, true, 2>(grid, d);  run<64, true, 4>(grid, d);
        run<64, false, 1>(grid, d); run<64, false, 4>(grid, d);
        run<128, true, 1>(grid, d); run<128, true, 2>(grid, d); run<128, false, 2>(grid, d);
        run<256, true, 1>(grid, d); run<256, false, 1>(grid, d);
        run<32, true, 4, true, 0>(grid, d);   // B descriptor changes every MMA
        run<32, true, 4, false, 8>(grid, d);  // 8 FFMA-bound warps (2 per SMSP) competing for issue slots
        run<32, true, 4, true, 8>(grid, d);
        run<64, true, 4, true, 8>(grid, d);
        run<32, true, 4, true, 0, 1>(grid, d);  // a different A tile (TMEM columns) every MMA
        run<32, true, 4, true, 0, 2>(grid, d);  // A tile reused by 3 consecutive MMAs
        run<32, true, 2, true, 0, 1>(grid, d);
        run<32, true, 4, true, 0, 1, 4>(grid, d);  // + 4 warps streaming tcgen05.ld
        run<32, true, 4, true, 0, 1, 8>(grid, d);  // + 8 warps streaming tcgen05.ld
    }
    return cudaDeviceSynchronize() != cudaSuccess;
}
