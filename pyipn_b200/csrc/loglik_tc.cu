// tcgen05 implementation of the delay-grid log-likelihood: exact sliced-integer contraction on the INT8 tensor
// pipe (UTCIMMA) with int32 accumulators in TMEM, fused Poisson epilogue.
//
// Why tensor cores at all (north star: "only if ncu shows the contraction beats the FP32/SFU pipe"):
// the FP32-pipe kernel (loglik_simt.cu) tops out at 200 FFMA per (grid-point x bin) -- 1.86e11 units/s at the
// nominal FP32 peak -- and misses the 1e-3 nat tolerance at 1e5 bins because 200-400 fp32 additions leave a
// random ~1.4e-6 error in every exponent.  Integer tensor-core accumulation is EXACT, so the contraction can
// be done to any precision by slicing the operands into signed 8-bit digits (an Ozaki-style splitting):
//
//   Phi (|.| <= 1)   = (p1 2^24 + p2 2^16 + p3 2^8 + p4) 2^-30                 4 digit planes  (|error| <= 2^-31, per bin)
//   W sigma (|.|<=1) = (w1 2^32 + w2 2^24 + w3 2^16 + w4 2^8 + w5) 2^-38        5 digit planes  (sigma = power of two)
//   x = sum_k Phi W  = kappa [ S2 + 2^-8 S3 + 2^-16 S4 + 2^-24 S5 + 2^-32 S6 ],  kappa = 2^-12 / sigma
//   S2 = w1.p1   S3 = w2.p1 + w1.p2   S4 = w3.p1 + w2.p2 + w1.p3   S5 = w4.p1 + w3.p2 + w2.p3 + w1.p4   S6 = w5.p1 + w4.p2
//                                                                                                  (12 int8 GEMMs)
// Which products are kept follows from how their errors add up over the bins.  A rounding error of Phi is random from bin
// to bin, so it enters a sum over 1e5 bins with sqrt(sum r_i^2) (r_i = d ll / d x_i, the Poisson residual weight); a
// rounding error of W -- and every dropped product of a fixed W digit with a SMOOTH Phi digit (p1, p2) -- is the same for
// all bins of a (delay, draw), so it enters with the coherent sum_i r_i Phi_ik, which for a model that is off by 1e4 nats
// (every posterior draw but the best few) is 100 times larger.  Measured on the bench draws before the fifth W digit and
// the w4.p2 product existed: up to 1.6e-3 nats.  The level-6 products with the pseudo-random low Phi digits (w3.p3, w2.p4)
// and everything below are dropped (< 1e-8 in x, random).  The additive constant log2(A/B) of the
// draw rides along as extra K rows (Phi = 1, W = c sigma / n_rows), so x comes out of the tensor core complete.
//
// Mapping: MMA M = 128 trial delays, N = 32 time bins, K = 32 per UTCIMMA.  The first four W digits of a work item
// (draw, 128 delays) are staged once through shared memory into TENSOR MEMORY (tcgen05.cp, 224 columns, two planes at a
// time) and used as the A operand from there; the fifth digit plane stays in shared memory for the item (one product, SS
// form).  Why tensor memory: with both operands in shared memory every MMA re-reads the 128-row A operand (4 KB + 1 KB of B
// per MMA at 128 B/clk: measured 43 cycles against 18 with A in tensor memory).  The Phi digits (B operand, 28 KB per tile)
// stream through a 4-stage bulk-copy ring, the
// tiles' Poisson constants through a separate 6-deep ring.  4 level accumulators x 32 columns x 2 buffers sit in
// TMEM columns [256, 512), the level-6 accumulator single-buffered in the 32 columns left, [224, 256).  An epilogue thread
// owns ONE delay (TMEM lane) and walks the tile's bins (columns): a bin's counts / kappa are warp-uniform, bins are
// partitioned so that tiles are homogeneous and tiles without counts skip the log entirely (and the two kinds of tile
// are interleaved, tc_bin_position), and the sum over bins is a per-thread compensated sum -- no shuffles, no per-tile
// atomics.
// Warp roles: 0..15 = epilogue, 16 = bulk-copy producer, 17 / 18 = MMA issuers of the even / odd tiles (17 also does the
// tcgen05.cp of the W digits).  Three instantiations by epilogue arithmetic (plain fp32 / compensated fp32 / fp64), each
// working on the draws whose brightness asks for it (tc_draw_kernel decides per draw), and a fourth with
// a store epilogue, which turns the same contraction into ipn_rff_eval on uniform bins (run_rff_eval_tc below).
// Experiment / test hooks (IPN_TC_DEBUG wait-cycle counters, IPN_TC_SKIP_MATH, IPN_TC_SKIP_MMA, IPN_TC_ONE_ISSUER,
// IPN_TC_EPI_DELAY, IPN_TC_GRID, IPN_TC_CHUNK_TILES) exist only in a build with -DIPN_TC_HOOKS (libipn_b200_hooks.so, which
// tests/test_gpu_loglik.py and tools/ load to check that results do not depend on pacing): the shipped library contains
// neither the environment look-ups nor the branches.
#include "loglik_common.cuh"
#include "loglik_epilogue.cuh"
#include "tc_perm.cuh"
#include "tc_images.cuh"

namespace ipn {

// namespace tc (tile geometry), TcDraw, tc_draw_kernel / tc_wimg_kernel / tc_pimg_kernel: tc_images.cuh

struct TcParams {
    const uint8_t* Wimg;  // [Sb][n_dt][a_bytes]
    const uint8_t* Pimg;  // [Sb][n_bt][img_bytes]: digit planes (b_dig bytes) followed by the constants block
    const TcDraw* td;     // [Sb]
    double* R;            // [Sb][Gpad]
    int Sb, n_dt, n_bt, nchunk, tiles_per_chunk, ksteps;
    int64_t G, Gpad;
    uint32_t a_bytes, b_dig, img_bytes, a_plane, b_plane;  // sizes / plane strides in bytes
    const int* flags;         // [0] = #bins with counts, [1] = epilogue mode option, [2..3] = sum n_i^2 (double)
    // store mode (EPI_STORE): out[sl * out_T + p * out_Q + q] = sgn[sl] 2^x for "bin" p (column) and "delay" q (row), see run_rff_eval_tc
    double* out;
    int64_t out_T;
    int out_Q;
    const double* sgn;        // [Sb] +1, -1 or 0 (sign of the draw's amplitude)
    // the fields below are only read by a -DIPN_TC_HOOKS build (IPN_HOOK())
    int one_issuer;           // experiment switch (IPN_TC_ONE_ISSUER): warp 13 issues every tile, warp 14 only takes part in the hand-shakes
    int epi_delay;            // test hook (IPN_TC_EPI_DELAY, cycles): stall every epilogue unit -> protocol must survive a slow consumer
    int skip_mma;             // experiment switch (IPN_TC_SKIP_MMA): issue one MMA per tile only -> exposes the epilogue-bound period
    int skip_math;            // experiment switch (IPN_TC_SKIP_MATH): epilogue only drains TMEM -> exposes the MMA-bound period
    unsigned long long* dbg;  // optional [grid][8] cycle counters of the pipeline wait sites (NULL = off)
};

#ifdef IPN_TC_HOOKS
#define IPN_HOOK(field) (p.field)
#else
#define IPN_HOOK(field) 0
#endif

// ------------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Where a timed-out wait leaves its identity before trapping (host-mapped, see trap_record_device_ptr()).
__device__ int* g_trap_record = nullptr;
enum WaitId { W_A_FULL = 1, W_A_EMPTY, W_A_TMEM_FULL, W_A_TMEM_FREE, W_B_FULL, W_B_EMPTY, W_C_FULL, W_C_EMPTY, W_T_FULL, W_T_EMPTY, W_W5_FULL, W_T6_EMPTY };

__device__ __noinline__ void mbar_timeout(int id, uint32_t parity, uint32_t seq, const volatile uint32_t* prog) {
    int* r = g_trap_record;
    if (r && atomicCAS(r, 0, id) == 0) {
        r[1] = (int)blockIdx.x;
        r[2] = (int)(threadIdx.x >> 5);
        r[3] = (int)parity;
        r[4] = (int)seq;
        for (int w = 0; w < 24; ++w) r[8 + w] = (int)prog[w];  // what every warp of the CTA was last waiting for
        __threadfence_system();
    }
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int id, uint32_t seq, volatile uint32_t* prog) {
    // bounded spin: a protocol bug must surface as a launch failure (trap), never as a hung GPU
    const uint32_t addr = smem_u32(bar);
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 27); ++spin) {
        if (spin == 1) prog[threadIdx.x >> 5] = ((uint32_t)id << 24) | (seq & 0xFFFFFFu);  // not ready at once: note what we wait for
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)  // no suspend-time hint: with one ptxas emits NANOSLEEP back-off, whose wake-up
                                       // is not tied to the barrier and adds hundreds of cycles to every hand-over
            : "memory");
        if (done) return;
    }
    mbar_timeout(id, parity, seq, prog);
}
// Polling with back-off for the producer: it runs 2-4 stages ahead, so wake-up latency is irrelevant, while a
// tight try_wait loop would steal issue slots from the epilogue warps of its sub-partition.
template <int SLEEP_NS = 256>
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, int id, uint32_t seq, volatile uint32_t* prog) {
    prog[threadIdx.x >> 5] = ((uint32_t)id << 24) | (seq & 0xFFFFFFu);
    const uint32_t addr = smem_u32(bar);
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        __nanosleep(SLEEP_NS);
    }
    mbar_timeout(id, parity, seq, prog);
}
// Wait with optional cycle accounting (IPN_TC_DEBUG)
__device__ __forceinline__ void mbar_wait_timed(uint64_t* bar, uint32_t parity, unsigned long long& acc, bool on, int id, uint32_t seq, volatile uint32_t* prog) {
    if (!on) {
        mbar_wait(bar, parity, id, seq, prog);
        return;
    }
    const long long t0 = clock64();
    mbar_wait(bar, parity, id, seq, prog);
    acc += (unsigned long long)(clock64() - t0);
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
template <bool ACCUM>
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    if (ACCUM)
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.eq.u32 p, 1, 1;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc)
            : "memory");
    else
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.eq.u32 p, 1, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc)
            : "memory");
}
// A operand from tensor memory (TS form)
template <bool ACCUM>
__device__ __forceinline__ void tc_mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
    if (ACCUM)
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.eq.u32 p, 1, 1;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t"
            "}" ::"r"(d_tmem),
            "r"(a_tmem), "l"(bdesc), "r"(idesc)
            : "memory");
    else
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.eq.u32 p, 1, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t"
            "}" ::"r"(d_tmem),
            "r"(a_tmem), "l"(bdesc), "r"(idesc)
            : "memory");
}
// shared memory (matrix descriptor: 128 rows x 32 bytes, K-major core matrices) -> 128 lanes x 8 columns of TMEM
__device__ __forceinline__ void tc_cp_128x256b(uint32_t dst_tmem, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(dst_tmem), "l"(sdesc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, int (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (version 1 = Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// epi_units8 / epi_units8_comp / epi_units8_hifi: loglik_epilogue.cuh

__device__ __forceinline__ void tc_ld8(uint32_t taddr, int (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}

// ------------------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------------------
// One instantiation per epilogue mode (EPI_FP32 / EPI_COMP / EPI_FP64): each handles only the draws flagged for its mode
// (every role skips the others' items), so the common fp32 instantiation carries no compensated or fp64 code in its
// instruction-cache footprint.
enum { EPI_FP32 = 0, EPI_COMP = 1, EPI_FP64 = 2, N_EPI = 3, EPI_STORE = 3 };  // EPI_STORE: the contraction with a store epilogue (ipn_rff_eval)
// (register budget: 96 per thread -- 19 warps x 32 x 104 does not launch, the allocation granule is 512 registers per warp)
template <int KSTEPS, int MODE>
__global__ void __launch_bounds__(tc::NTHREADS, 1) loglik_tc_kernel(const TcParams p) {
    using namespace tc;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sW5 = smem;                                              // fifth W digit plane of the current item (resident)
    uint8_t* sA = smem + p.a_plane;                                   // staging of W digit planes 1-4, two at a time
    uint8_t* sB0 = sA + 2 * (size_t)p.a_plane;                        // [NBST][b_dig]  Phi digit planes
    uint8_t* sC0 = sB0 + NBST * (size_t)p.b_dig;                      // [NCST][CONST_BYTES]  Poisson constants
    uint64_t* bars = reinterpret_cast<uint64_t*>(sC0 + NCST * CONST_BYTES);
    uint64_t* a_full = bars + 0;
    uint64_t* a_empty = bars + 1;
    uint64_t* t_empty = bars + 2;            // [NBUF]
    uint64_t* t_full = t_empty + NBUF;       // [NTF]
    uint64_t* b_full = t_full + NTF;         // [NBFULL]: indexed by tile % NBFULL so that each barrier belongs to ONE issuer
    uint64_t* b_empty = b_full + NBFULL;     // [NBST]
    uint64_t* c_full = b_empty + NBST;       // [NCST]
    uint64_t* c_empty = c_full + NCST;       // [NCST]
    uint64_t* a_tmem_full = c_empty + NCST;  // W digits of the current item are in tensor memory (issuer 0 -> issuer 1)
    uint64_t* a_tmem_free = a_tmem_full + 1; // both issuers' MMAs of the item have finished reading them (-> issuer 0, producer)
    uint64_t* w5_full = a_tmem_free + 1;     // fifth W digit plane of the item is in shared memory (-> both issuers)
    uint64_t* t6_empty = w5_full + 1;        // [2] level-6 accumulator drained by the epilogue of tile n - 1 (-> issuer of tile n, n % 2)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t6_empty + 2);
    volatile uint32_t* prog = tmem_slot + 4;  // [32] last wait of every warp (time-out diagnostics; the record keeps the first 24)
    static_assert(NTHREADS / 32 <= 32, "prog[] holds one word per warp");

    // warp index made provably warp-uniform so the role branches (and everything inside them) stay on the
    // uniform datapath: UTCIMMA takes its descriptors from uniform registers
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    constexpr int DRAW_MODE = MODE == EPI_STORE ? (int)EPI_FP32 : MODE;  // which draws (TcDraw::hifi) this instantiation works on

    // In automatic mode every instantiation is launched and the draws' modes are only known on the device: leave at once --
    // before barriers, tensor-memory allocation and the role loops -- if no draw of the batch is ours (a single-draw grid is
    // ~1 ms of work, and two empty instantiations that set everything up and tore it down again were 3 % of it).
    {
        __shared__ int s_any;
        if (warp == 0) {
            int any = 0;
            for (int i = lane; i < p.Sb; i += 32) any |= p.td[i].hifi == DRAW_MODE;
            any = __any_sync(0xffffffffu, any);
            if (lane == 0) s_any = any;
        }
        __syncthreads();
        if (!s_any) return;
    }

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        mbar_init(a_tmem_full, 1);
        mbar_init(a_tmem_free, 2);
        mbar_init(w5_full, 1);
        mbar_init(t6_empty + 0, EPI_ARRIVALS);
        mbar_init(t6_empty + 1, EPI_ARRIVALS);
        for (int i = 0; i < NBUF; ++i) mbar_init(t_empty + i, EPI_ARRIVALS);
        for (int i = 0; i < NTF; ++i) mbar_init(t_full + i, 1);
        for (int i = 0; i < NBFULL; ++i) mbar_init(b_full + i, 1);
        for (int i = 0; i < NBST; ++i) mbar_init(b_empty + i, 1);
        for (int i = 0; i < NCST; ++i) {
            mbar_init(c_full + i, 1);
            mbar_init(c_empty + i, EPI_ARRIVALS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int items_per_draw = p.nchunk * p.n_dt;
    const int64_t n_items = (int64_t)p.Sb * items_per_draw;

    if (warp == PRODUCER_WARP) {
        // ===== producer: bulk copies global -> shared, completion on the *_full barriers =====
        if (lane == 0) {
            uint32_t tile_seq = 0, item_seq = 0;
            unsigned long long w_be = 0, w_ce = 0;
            const bool dbg = IPN_HOOK(dbg) != nullptr;
            for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x, ++item_seq) {
                const int sl = (int)(item / items_per_draw);
                if (p.td[sl].hifi != DRAW_MODE) {
                    --item_seq;  // not this instantiation's draw: the item does not exist for any role
                    continue;
                }
                const int rem = (int)(item - (int64_t)sl * items_per_draw);
                const int chunk = rem / p.n_dt, dt = rem - chunk * p.n_dt;
                // W digit planes 1-4 through the staging buffer, two planes at a time (a_full / a_empty change phase twice per
                // item); plane 5 into its resident buffer once every MMA of the previous item has completed
                const uint8_t* wimg = p.Wimg + ((int64_t)sl * p.n_dt + dt) * p.a_bytes;
                mbar_wait_relaxed(a_empty, 1, W_A_EMPTY, item_seq, prog);  // second half of the previous item copied into TMEM
                mbar_expect_tx(a_full, 2 * p.a_plane);
                bulk_g2s(sA, wimg, 2 * p.a_plane, a_full);
                mbar_wait_relaxed(a_tmem_free, (item_seq & 1) ^ 1, W_A_TMEM_FREE, item_seq, prog);
                mbar_expect_tx(w5_full, p.a_plane);
                bulk_g2s(sW5, wimg + 4 * (size_t)p.a_plane, p.a_plane, w5_full);
                mbar_wait_relaxed(a_empty, 0, W_A_EMPTY, item_seq, prog);  // first half of this item copied into TMEM
                mbar_expect_tx(a_full, 2 * p.a_plane);
                bulk_g2s(sA, wimg + 2 * (size_t)p.a_plane, 2 * p.a_plane, a_full);
                const int bt0 = chunk * p.tiles_per_chunk;
                const int bt1 = min(p.n_bt, bt0 + p.tiles_per_chunk);
                for (int bt = bt0; bt < bt1; ++bt, ++tile_seq) {
                    const uint8_t* img = p.Pimg + ((int64_t)sl * p.n_bt + bt) * p.img_bytes;
                    const uint32_t st = tile_seq % NBST, ph = (tile_seq / NBST) & 1;
                    mbar_wait_relaxed(b_empty + st, ph ^ 1, W_B_EMPTY, tile_seq, prog);  // released by the MMA commit
                    mbar_expect_tx(b_full + tile_seq % NBFULL, p.b_dig);
                    bulk_g2s(sB0 + (size_t)st * p.b_dig, img, p.b_dig, b_full + tile_seq % NBFULL);
                    const uint32_t cs = tile_seq % NCST, cph = (tile_seq / NCST) & 1;
                    mbar_wait_relaxed(c_empty + cs, cph ^ 1, W_C_EMPTY, tile_seq, prog);  // released by the epilogue warps
                    mbar_expect_tx(c_full + cs, CONST_BYTES);
                    bulk_g2s(sC0 + (size_t)cs * CONST_BYTES, img + p.b_dig, CONST_BYTES, c_full + cs);
                }
            }
            if (dbg) {
                p.dbg[blockIdx.x * 8 + 0] = w_be;
                p.dbg[blockIdx.x * 8 + 1] = w_ce;
            }
        }
    } else if (warp == MMA_WARP || warp == MMA_WARP2) {
        // ===== MMA issuers: two warps on different sub-partitions, one per accumulator buffer (even / odd tiles) =====
        // Measured: with the epilogue math running on its sub-partition a single issuing warp needs ~2200 cycles for the
        // 63 MMAs of a tile (the tensor pipe needs 1000); two issuers halve that.  Each walks the whole (warp-uniform)
        // schedule and one elected lane issues the tiles of its parity.
        const uint32_t which = warp == MMA_WARP ? 0u : 1u;
        // instruction descriptor: D = S32, A = B = signed 8-bit, both K-major, N = 32, M = 128
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
        constexpr uint32_t a_lbo = TM * 16, b_lbo = TN * 16;
        // descriptor template: LBO / SBO / version bits; the 14-bit start-address field is added per use
        const uint64_t a_tmpl = make_desc(0, a_lbo, 128), b_tmpl = make_desc(0, b_lbo, 128);
        const uint32_t a_base = smem_u32(sA) >> 4;
        const uint32_t w5_base = smem_u32(sW5) >> 4;
        const uint32_t b_base0 = smem_u32(sB0) >> 4;
        uint32_t tile_seq = 0, item_seq = 0;
        unsigned long long w_af = 0, w_bf = 0, w_te = 0;
        const bool dbg = IPN_HOOK(dbg) != nullptr;
        const long long t_start = clock64();
        for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x, ++item_seq) {
            if (p.td[(int)(item / items_per_draw)].hifi != DRAW_MODE) {
                --item_seq;
                continue;
            }
            const int rem = (int)(item % items_per_draw);
            const int chunk = rem / p.n_dt;
            const int bt0 = chunk * p.tiles_per_chunk;
            const int bt1 = min(p.n_bt, bt0 + p.tiles_per_chunk);
            if (which == 0) {
                // W digit planes 1-4: shared -> tensor memory, 8 columns (32 K-bytes x 128 lanes) per copy, two planes per
                // staging-buffer fill.  PTX orders cp -> mma of one thread but not mma -> cp, hence the a_tmem_free hand-shake
                // (a pipeline drain per item, i.e. per >= 16 tiles).
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    mbar_wait_timed(a_full, half, w_af, dbg, W_A_FULL, item_seq, prog);
                    if (half == 0) mbar_wait(a_tmem_free, (item_seq & 1) ^ 1, W_A_TMEM_FREE, item_seq, prog);  // all MMAs of the previous item have read its W digits
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int q = 2 * half; q < 2 * half + 2; ++q)
#pragma unroll
                            for (int ks = 0; ks < KSTEPS; ++ks)
                                tc_cp_128x256b(tmem_base + (q * KSTEPS + ks) * 8,
                                               a_tmpl + (uint64_t)(a_base + (q & 1) * (p.a_plane >> 4) + ks * (2 * a_lbo >> 4)));
                        tc_commit(a_empty);                     // staging buffer free once the copies have completed
                        if (half == 1) tc_commit(a_tmem_full);  // ... and issuer 1 may start on this item
                    }
                    __syncwarp();
                }
            } else {
                mbar_wait(a_tmem_full, item_seq & 1, W_A_TMEM_FULL, item_seq, prog);
                tc_fence_after();
            }
            mbar_wait(w5_full, item_seq & 1, W_W5_FULL, item_seq, prog);  // the fifth digit plane is in shared memory
            for (int bt = bt0; bt < bt1; ++bt, ++tile_seq) {
                if ((IPN_HOOK(one_issuer) ? 0u : (tile_seq & 1)) != which) continue;
                const uint32_t st = tile_seq % NBST;
                const uint32_t tb = tile_seq % NBUF, tph = (tile_seq / NBUF) & 1;
                mbar_wait_timed(b_full + tile_seq % NBFULL, (tile_seq / NBFULL) & 1, w_bf, dbg, W_B_FULL, tile_seq, prog);
                mbar_wait_timed(t_empty + tb, tph ^ 1, w_te, dbg, W_T_EMPTY, tile_seq, prog);
                tc_fence_after();
                const uint32_t d0 = tmem_base + ACC_COL0 + tb * (NLEV * TN);
                const uint32_t b_base = b_base0 + st * (p.b_dig >> 4);
                if (elect_one()) {
                    // products (W plane q, Phi plane pp), level = q + pp.  K steps outermost so that consecutive MMAs
                    // accumulate into DIFFERENT level accumulators (no back-to-back read-modify-write of one tile).
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ++ks) {
                        if (IPN_HOOK(skip_mma) && ks > 0) break;
#pragma unroll
                        for (int pp = 0; pp < PP; ++pp) {
#pragma unroll
                            for (int lev = 0; lev < NLEV; ++lev) {
                                const int q = lev - pp;
                                if (q < 0 || q >= PWT) continue;
                                const bool first = (pp == 0) || (lev - (pp - 1) >= PWT);  // first product of this level
                                const uint32_t at = tmem_base + (q * KSTEPS + ks) * 8;
                                const uint64_t bd = b_tmpl + (uint64_t)(b_base + pp * (p.b_plane >> 4) + ks * (2 * b_lbo >> 4));
                                if (first && ks == 0)
                                    tc_mma_i8_ts<false>(d0 + lev * TN, at, bd, idesc);
                                else
                                    tc_mma_i8_ts<true>(d0 + lev * TN, at, bd, idesc);
                            }
                        }
                    }
                }
                __syncwarp();
                // level 6 (w4.p2 from tensor memory, w5.p1 with both operands in shared memory) into the single-buffered
                // accumulator: the epilogue of the PREVIOUS tile must have drained it.  t6_empty[n % 2] belongs to the issuer of
                // the tiles n = this parity and completes one phase per tile n - 1, from tile 1 on.
                if (tile_seq > 0) mbar_wait_timed(t6_empty + (tile_seq & 1), ((tile_seq - 1) >> 1) & 1, w_te, dbg, W_T6_EMPTY, tile_seq, prog);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d6 = tmem_base + L6_COL0;
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ++ks) {
                        if (IPN_HOOK(skip_mma) && ks > 0) break;
                        const uint64_t bd1 = b_tmpl + (uint64_t)(b_base + ks * (2 * b_lbo >> 4));
#ifndef IPN_TC_L6
#define IPN_TC_L6 2  // experiment switch: 2 = both level-6 products (default), 1 = w4.p2 only, 0 = w5.p1 only
#endif
                        if (IPN_TC_L6 != 0) {
                            const uint64_t bd2 = bd1 + (uint64_t)(p.b_plane >> 4);
                            if (ks == 0)
                                tc_mma_i8_ts<false>(d6, tmem_base + (3 * KSTEPS + ks) * 8, bd2, idesc);
                            else
                                tc_mma_i8_ts<true>(d6, tmem_base + (3 * KSTEPS + ks) * 8, bd2, idesc);
                        }
                        if (IPN_TC_L6 == 2)
                            tc_mma_i8<true>(d6, a_tmpl + (uint64_t)(w5_base + ks * (2 * a_lbo >> 4)), bd1, idesc);
                        else if (IPN_TC_L6 == 0) {
                            if (ks == 0)
                                tc_mma_i8<false>(d6, a_tmpl + (uint64_t)(w5_base + ks * (2 * a_lbo >> 4)), bd1, idesc);
                            else
                                tc_mma_i8<true>(d6, a_tmpl + (uint64_t)(w5_base + ks * (2 * a_lbo >> 4)), bd1, idesc);
                        }
                    }
                    tc_commit(b_empty + st);  // operands of this stage consumed
                    tc_commit(t_full + tile_seq % NTF);  // accumulators of this tile complete
                }
                __syncwarp();
            }
            if (elect_one()) tc_commit(a_tmem_free);  // arrives when this issuer's MMAs of the item have completed
            __syncwarp();
        }
        if (dbg && lane == 0 && which == 0 && tile_seq > 0) {  // (an instantiation without draws of its mode leaves the counters alone)
            p.dbg[blockIdx.x * 8 + 2] = w_af;
            p.dbg[blockIdx.x * 8 + 3] = w_bf;
            p.dbg[blockIdx.x * 8 + 4] = w_te;
            p.dbg[blockIdx.x * 8 + 7] = (unsigned long long)(clock64() - t_start);
        }
    } else {
        // ===== epilogue: TMEM -> registers -> Poisson terms -> per-delay compensated sum =====
        // Work unit = (tile, group of 8 columns).  The three warps of a lane quarter take the groups round-robin
        // (global group index mod 3), so at any time they sit in different phases (barrier wait / TMEM load / math)
        // and cover one another's latencies; every (quarter, group) arrives once per tile on t_empty / c_empty.
        const int quarter = warp & 3;        // TMEM lanes this warp may read
        const int slot = warp >> 2;          // 0..2
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        uint32_t tile_seq = 0;
        unsigned long long w_cf = 0, w_tf = 0;
        const bool dbg = IPN_HOOK(dbg) != nullptr;
        for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int sl = (int)(item / items_per_draw);
            if (p.td[sl].hifi != DRAW_MODE) continue;
            const int rem = (int)(item - (int64_t)sl * items_per_draw);
            const int chunk = rem / p.n_dt, dt = rem - chunk * p.n_dt;
            const int bt0 = chunk * p.tiles_per_chunk;
            const int bt1 = min(p.n_bt, bt0 + p.tiles_per_chunk);
            const float ks = p.td[sl].ks;
            const float ks24 = ks * 5.9604644775390625e-8f;
            // ks = 2^-(12 + e): the integer exponent assembly shifts by sh = 20 + e (exponent_parts, loglik_common.cuh)
            const int sh = 8 - (((__float_as_int(ks) >> 23) & 0xff) - 127);
            const float inv = ks * 0.00390625f;  // 2^-sh
            // Kahan pair + small-correction sum carried across the whole item; folded into fp64 once at the end
            float ssum = 0.0f, scomp = 0.0f, slo = 0.0f;
            float qsum = 0.0f, qcomp = 0.0f;  // compensated epilogue: the kappa u sum (the n y0 sum is in ssum / scomp)
            double hacc = 0.0;
            const double ksd = (double)ks;
            const double sgn = MODE == EPI_STORE ? p.sgn[sl] : 0.0;
            const bool zero_plain = MODE == EPI_COMP && (p.td[sl].pad & 1);  // tiles without counts take the plain fp32 unit
            for (int bt = bt0; bt < bt1; ++bt, ++tile_seq) {
                const uint32_t tb = tile_seq % NBUF, tph = (tile_seq / NBUF) & 1;
                const uint32_t cs = tile_seq % NCST, cph = (tile_seq / NCST) & 1;
                // groups of this tile that belong to this warp: (tile_seq * GROUPS + cg) % EPI_SLOTS == slot
                int cg = (int)((slot + EPI_SLOTS - (tile_seq * GROUPS) % EPI_SLOTS) % EPI_SLOTS);
                if (cg >= GROUPS) continue;  // not this warp's tile; it never touches that tile's barriers (see NTF)
                mbar_wait_timed(c_full + cs, cph, w_cf, dbg, W_C_FULL, tile_seq, prog);  // the tile's Poisson constants are in shared memory
                mbar_wait_timed(t_full + tile_seq % NTF, (tile_seq / NTF) & 1, w_tf, dbg, W_T_FULL, tile_seq, prog);
                tc_fence_after();
                const float4* cbase = reinterpret_cast<const float4*>(sC0 + (size_t)cs * CONST_BYTES);
                const float* unit_flag = reinterpret_cast<const float*>(cbase + TN);  // one per 8 bins: the unit holds bins with counts
                for (; cg < GROUPS; cg += EPI_SLOTS) {
                    const float4* cst = cbase + cg * GCOLS;
                    const uint32_t t0 = tmem_base + ACC_COL0 + lane_addr + tb * (NLEV * TN) + cg * GCOLS;
                    // The accumulators leave this block as TWO integers per bin, H = 256 S2 + S3 and L = 256 S4 + S5 (levels_high /
                    // levels_low, the operands of the exponent assembly): 32 registers for the 16 bins instead of 64, so the
                    // second unit's values do not crowd the registers while the first unit's eight bins are interleaved.
                    static_assert(NLEV == 4, "levels 2 .. 5, level 6 folded into level 5");
                    int hi[NH][8], lo[NH][8];
                    // Two rounds of tensor-memory loads so that no more than 64 accumulator registers are live at once.
                    // First the lowest level and level 6 (weight 2^-8 of it, single-buffered): fold, and hand the level-6
                    // columns back at once -- the issuer of the NEXT tile waits for them.
                    {
                        int acc6[GCOLS];
                        if (NH == 2) {
                            int v[16];
                            tc_ld16(t0 + (NLEV - 1) * TN, v);
                            tc_ld16(tmem_base + L6_COL0 + lane_addr + cg * GCOLS, reinterpret_cast<int(&)[16]>(acc6));
                            tc_wait_ld();
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                lo[0][j] = fold_level6(v[j], acc6[j]);
                                lo[NH - 1][j] = fold_level6(v[8 + j], acc6[8 + j]);
                            }
                        } else {
#pragma unroll
                            for (int h = 0; h < NH; ++h) {
                                tc_ld8(t0 + (NLEV - 1) * TN + h * 8, lo[h]);
                                tc_ld8(tmem_base + L6_COL0 + lane_addr + cg * GCOLS + h * 8, reinterpret_cast<int(&)[8]>(acc6[h * 8]));
                            }
                            tc_wait_ld();
#pragma unroll
                            for (int h = 0; h < NH; ++h)
#pragma unroll
                                for (int j = 0; j < 8; ++j) lo[h][j] = fold_level6(lo[h][j], acc6[h * 8 + j]);
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(t6_empty + ((tile_seq + 1) & 1));
                    }
                    if (NH == 2) {
                        // one 16-column load per level instead of two 8-column ones
                        int v2[16], v3[16], v4[16];
                        tc_ld16(t0, v2);
                        tc_ld16(t0 + TN, v3);
                        tc_ld16(t0 + 2 * TN, v4);
                        tc_wait_ld();
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            hi[0][j] = levels_high(v2[j], v3[j]);
                            hi[NH - 1][j] = levels_high(v2[8 + j], v3[8 + j]);
                            lo[0][j] = levels_low(v4[j], lo[0][j]);
                            lo[NH - 1][j] = levels_low(v4[8 + j], lo[NH - 1][j]);
                        }
                    } else {
                        int v2[NH][8], v3[NH][8], v4[NH][8];
#pragma unroll
                        for (int h = 0; h < NH; ++h) {
                            tc_ld8(t0 + h * 8, v2[h]);
                            tc_ld8(t0 + TN + h * 8, v3[h]);
                            tc_ld8(t0 + 2 * TN + h * 8, v4[h]);
                        }
                        tc_wait_ld();
#pragma unroll
                        for (int h = 0; h < NH; ++h)
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                hi[h][j] = levels_high(v2[h][j], v3[h][j]);
                                lo[h][j] = levels_low(v4[h][j], lo[h][j]);
                            }
                    }
                    // this (quarter, group) of the accumulator buffer is in registers: hand it back before the math
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(t_empty + tb);
                    if (IPN_HOOK(epi_delay)) {
                        const long long d0 = clock64();
                        while (clock64() - d0 < (long long)IPN_HOOK(epi_delay) * (1 + ((warp * 7 + tile_seq) & 3))) {}
                    }
                    if (IPN_HOOK(skip_math) == 1 || (IPN_HOOK(skip_math) == 2 && quarter == (MMA_WARP & 3))) {
                        int x = 0;
#pragma unroll
                        for (int h = 0; h < NH; ++h)
#pragma unroll
                            for (int lev = 0; lev < NLEV; ++lev) x ^= (lev & 1 ? hi : lo)[h][(h + lev) & 7];
                        slo += __int_as_float(x) * 0.0f;
                    } else {
#pragma unroll
                        for (int h = 0; h < NH; ++h) {
                            const bool has_counts = unit_flag[(cg * GCOLS) / 8 + h] != 0.0f;  // warp-uniform
                            if (MODE == EPI_STORE) {
                                float u8[8];
                                epi_units8_store(hi[h], lo[h], sh, inv, ks24, u8);
                                // bin i = p Q + q: p = the column's position among the "bins", q = this thread's "delay" row;
                                // for one column the 32 lanes of the warp write 256 consecutive bytes
                                const int64_t q = (int64_t)dt * TM + quarter * 32 + lane;
                                const int64_t p0 = (int64_t)bt * TN + cg * GCOLS + 8 * h;
                                double* orow = p.out + (int64_t)sl * p.out_T;
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const int64_t i = (p0 + j) * p.out_Q + q;
                                    if (q < p.out_Q && i < p.out_T) orow[i] = sgn * (double)u8[j];
                                }
                            } else if (MODE == EPI_FP64)
                                epi_units8_hifi(hi[h], lo[h], cst + 8 * h, ksd, hacc);
                            else if (MODE == EPI_COMP && has_counts)
                                epi_units8_comp<true>(hi[h], lo[h], cst + 8 * h, sh, inv, ks24, ssum, scomp, qsum, qcomp, slo);
                            else if (MODE == EPI_COMP && zero_plain)  // bins without counts: - kappa u is small unless the model is grossly off
                                epi_units8<false>(hi[h], lo[h], cst + 8 * h, sh, inv, ks24, ssum, scomp, slo);
                            else if (MODE == EPI_COMP)
                                epi_units8_comp<false>(hi[h], lo[h], cst + 8 * h, sh, inv, ks24, ssum, scomp, qsum, qcomp, slo);
                            else if (has_counts)
                                epi_units8<true>(hi[h], lo[h], cst + 8 * h, sh, inv, ks24, ssum, scomp, slo);
                            else
                                epi_units8<false>(hi[h], lo[h], cst + 8 * h, sh, inv, ks24, ssum, scomp, slo);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(c_empty + cs);
                }
            }
            if (MODE != EPI_STORE) {
                double acc = ((double)ssum - (double)scomp) - ((double)qsum - (double)qcomp) + (double)slo + hacc;
                const int64_t g = (int64_t)dt * TM + quarter * 32 + lane;
                if (g < p.G) atomicAdd(p.R + (int64_t)sl * p.Gpad + g, acc);
            }
        }
        if (dbg && warp == 0 && lane == 0 && tile_seq > 0) {
            p.dbg[blockIdx.x * 8 + 5] = w_cf;
            p.dbg[blockIdx.x * 8 + 6] = w_tf;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// the 1024-point fixed-point unit circle of tc_pimg_kernel, built on first use (8 KB, lives as long as the context)
static int ensure_phitab(ipn_ctx* ctx) {
    if (ctx->phitab.p) return IPN_OK;
    int rc = ctx->phitab.ensure(kPhiTabSize * sizeof(PhiTabEntry));
    if (rc) return rc;
    tc_phitab_kernel<<<kPhiTabSize / 256, 256, 0, ctx->stream>>>(ctx->phitab.as<PhiTabEntry>());
    ctx->launches += 1;
    IPN_CUDA_CHECK(cudaGetLastError());
    return IPN_OK;
}

bool tc_supported(int64_t J) { return 2 * J + tc::MIN_CROWS <= tc::KPAD_MAX; }

// Store mode of the runner (ipn_rff_eval on uniform bins, run_rff_eval_tc below): the caller supplies the bin permutation and
// the flag block (no counts to sort), a delay grid PER DRAW, and instead of reducing over the bins the epilogue stores 2^x.
struct TcStore {
    double* out;           // [S][out_T]
    int64_t out_T;
    int out_Q;             // bins i = p * out_Q + q ("bin" p, "delay" q)
    const double* sgn;     // [S]
    const int* perm;       // [T] identity
    const int* flags;      // [8]: n_nz = 0, mode option 2 (plain fp32 unless the exponent is unbounded), sum n^2 = 0
    const double *time, *shift, *amp;  // the caller's arrays, for the draws the direct kernel has to take (unbounded exponent)
};

int launch_rff_eval_masked(ipn_ctx* ctx, int64_t T, const double* time, int64_t S0, int64_t Sb, int64_t J, const double* omega,
                           const double* a, const double* b, const double* shift, const double* amp, double* out,
                           const void* tcdraw);  // rff_eval.cu: draws S0 .. S0 + Sb - 1 with TcDraw::hifi == 2 only

static int run_tc_grid(ipn_ctx* ctx, int64_t T, const double* time, const double* exposure, const int64_t* counts,
                       int64_t S, int64_t J, const double* omega, const double* a, const double* b, const DrawConst* d_dc,
                       int64_t G, const double* delays, int64_t delay_stride, double* ll, bool accumulate, const TcStore* st) {
    using namespace tc;
    if (!tc_supported(J)) {
        set_error("tcgen05 path supports J <= %d (got %lld)", (KPAD_MAX - MIN_CROWS) / 2, (long long)J);
        return IPN_ERR_UNSUPPORTED;
    }
    const int Kpad = (int)(((2 * J + MIN_CROWS + 31) / 32) * 32);
    const int ksteps = Kpad / 32;
    int rc0;
    const int n_dt = (int)((G + TM - 1) / TM);
    const int64_t Gpad = (int64_t)n_dt * TM;
    const int n_bt = (int)((T + TN - 1) / TN);
    const uint32_t a_plane = (uint32_t)Kpad * TM, b_plane = (uint32_t)Kpad * TN;
    const uint32_t a_bytes = PW * a_plane, b_dig = PP * b_plane, b_bytes = b_dig + CONST_BYTES;  // b_bytes = global tile image
    if ((rc0 = ensure_phitab(ctx))) return rc0;
    // fifth W plane (resident) + two-plane staging buffer + Phi ring + constants ring + barriers + TMEM slot + wait log
    const size_t smem = 3 * (size_t)a_plane + NBST * (size_t)b_dig + (size_t)NCST * CONST_BYTES + (7 + NBUF + NTF + NBFULL + NBST + 2 * NCST) * 8 + 16 + 32 * 4;

    // draws per batch: bound the operand-image scratch (T/64 Phi tiles x ~44 KB per draw) to ~6 GiB
    const size_t p_per_draw = (size_t)n_bt * b_bytes, w_per_draw = (size_t)n_dt * a_bytes;
    int64_t Sbatch = ctx->draw_batch > 0 ? ctx->draw_batch : (int64_t)((6ull << 30) / (p_per_draw + w_per_draw + 1));
    if (Sbatch < 1) Sbatch = 1;
    if (Sbatch > S) Sbatch = S;
    if (Sbatch > 65535) Sbatch = 65535;
    ctx->last_draw_batch = Sbatch;

    int rc;
    if ((rc = ctx->wtab.ensure((size_t)Sbatch * w_per_draw))) return rc;
    if ((rc = ctx->misc.ensure((size_t)Sbatch * p_per_draw))) return rc;
    if (!st && (rc = ctx->accum.ensure((size_t)Sbatch * Gpad * sizeof(double)))) return rc;
    if ((rc = ctx->tcdraw.ensure((size_t)Sbatch * sizeof(TcDraw)))) return rc;
    const int* d_nnz;   // [0] n_nz, [1] mode, [2..3] sum n^2 as double (8-byte aligned)
    const int* d_perm;
    if (st) {
        d_nnz = st->flags;
        d_perm = st->perm;
    } else {
        const int n_pblk = (int)((T + 1023) / 1024);
        const size_t perm_ints = (((size_t)T + 8 + 1) / 2) * 2;  // flags + permutation, padded so the doubles behind stay aligned
        if ((rc = ctx->perm.ensure(perm_ints * sizeof(int) + (size_t)(n_pblk + 1) * (sizeof(double) + 4 * sizeof(int)) + 64))) return rc;
        int* nnz = ctx->perm.as<int>();
        int* perm = nnz + 8;
        double* d_blk_sq = reinterpret_cast<double*>(nnz + perm_ints);
        int* d_blk_cnt = reinterpret_cast<int*>(d_blk_sq + n_pblk + 1);
        if (n_pblk > 0) tc_perm_count_kernel<<<n_pblk, 1024, 0, ctx->stream>>>(T, counts, d_blk_cnt, d_blk_sq);
        tc_perm_kernel<<<n_pblk > 0 ? n_pblk : 1, 1024, 0, ctx->stream>>>(T, counts, n_pblk, d_blk_cnt, d_blk_sq, perm, nnz,
                                                                        ctx->hifi_mode);  // T == 0: only writes the flags
        ctx->launches += n_pblk > 0 ? 2 : 1;
        d_nnz = nnz;
        d_perm = perm;
    }
    uint8_t* d_W = ctx->wtab.as<uint8_t>();
    uint8_t* d_P = ctx->misc.as<uint8_t>();
    double* d_R = ctx->accum.as<double>();
    TcDraw* d_td = ctx->tcdraw.as<TcDraw>();

    void (*kern[N_EPI + 1])(const TcParams) = {nullptr, nullptr, nullptr, nullptr};  // by epilogue mode
    switch (ksteps) {
#define IPN_TC_CASE(K)                            \
    case K:                                       \
        kern[0] = loglik_tc_kernel<K, EPI_FP32>;  \
        kern[1] = loglik_tc_kernel<K, EPI_COMP>;  \
        kern[2] = loglik_tc_kernel<K, EPI_FP64>;  \
        kern[3] = loglik_tc_kernel<K, EPI_STORE>; \
        break;
        IPN_TC_CASE(1) IPN_TC_CASE(2) IPN_TC_CASE(3) IPN_TC_CASE(4) IPN_TC_CASE(5) IPN_TC_CASE(6)
        default: IPN_TC_CASE(7)
#undef IPN_TC_CASE
    }
    // which instantiations can have work: auto -> fp32, compensated and (rarely) fp64; forced modes -> one
    // (a draw whose exponent is not bounded goes to the fp64 epilogue whatever the option says, so that one always runs)
    const int hm = ctx->hifi_mode;
    const bool run_mode[N_EPI + 1] = {!st && (hm == 0 || hm == 2), !st && (hm == 0 || hm == 3), !st, st != nullptr};
    // once per context (= per device): the time-out record's address, and the shared-memory attribute of every instantiation
    // that runs (a single-draw grid is ~1 ms of device work: a dozen microseconds of driver calls per launch set show)
    if (!ctx->tc_trap_set) {
        int* trap = trap_record_device_ptr();
        IPN_CUDA_CHECK(cudaMemcpyToSymbolAsync(g_trap_record, &trap, sizeof(trap), 0, cudaMemcpyHostToDevice, ctx->stream));
        ctx->tc_trap_set = true;
    }
    for (int m = 0; m <= N_EPI; ++m)
        if (run_mode[m] && !ctx->tc_attr_set[ksteps][m]) {
            IPN_CUDA_CHECK(cudaFuncSetAttribute(kern[m], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ctx->tc_attr_set[ksteps][m] = true;
        }

    for (int64_t S0 = 0; S0 < S; S0 += Sbatch) {
        const int Sb = (int)((S - S0 < Sbatch) ? (S - S0) : Sbatch);
        if (!st) IPN_CUDA_CHECK(cudaMemsetAsync(d_R, 0, (size_t)Sb * Gpad * sizeof(double), ctx->stream));
        tc_draw_kernel<<<(Sb + 127) / 128, 128, 0, ctx->stream>>>(S0, Sb, (int)J, Kpad - 2 * (int)J, a, b, d_dc, d_nnz, ctx->n_terms, d_td);
        tc_wimg_kernel<<<dim3(Kpad / 16, n_dt, Sb), 128, 0, ctx->stream>>>(S0, (int)J, Kpad, omega, a, b, d_dc, d_td, G, delays,
                                                                            delay_stride, n_dt, a_bytes, a_plane, d_W);
        if (n_bt > 0)
            tc_pimg_kernel<<<dim3((n_bt + PIMG_TILES - 1) / PIMG_TILES, Sb), 256, 0, ctx->stream>>>(S0, (int)J, Kpad, T, time, exposure, counts, d_perm, d_nnz,
                                                                     omega, d_dc, d_td, ctx->phitab.as<PhiTabEntry>(), n_bt, b_bytes, b_plane, d_P);
        ctx->launches += 3;
        IPN_CUDA_CHECK(cudaGetLastError());

        if (n_bt > 0) {
            // split the bins of a (draw, delay tile) into chunks when there are too few items to balance the SMs
            int nchunk = 1;
            const int64_t base_items = (int64_t)Sb * n_dt;
            if (base_items < 32 * (int64_t)ctx->sm_count) {
                nchunk = (int)((32 * (int64_t)ctx->sm_count + base_items - 1) / base_items);
                const int min_tiles = 512 / TN;  // keep >= 512 bins per item
                const int max_chunks = (n_bt + min_tiles - 1) / min_tiles;
                if (nchunk > max_chunks) nchunk = max_chunks;
                if (nchunk < 1) nchunk = 1;
            }
            // ... and when there are many tiles, so that an item lasts ~0.1 ms: the 32 delay tiles that share a chunk of Phi
            // start up to one item apart (148 CTAs is not a multiple of 32), and the chunk must still be in L2 for the late
            // ones (ncu: 600-tile items re-read half of Phi from HBM, 2.2x the minimum traffic)
            int max_tiles = 4096 / TN;
#ifdef IPN_TC_HOOKS
            if (getenv("IPN_TC_CHUNK_TILES")) max_tiles = atoi(getenv("IPN_TC_CHUNK_TILES"));
#endif
            if ((n_bt + nchunk - 1) / nchunk > max_tiles) nchunk = (n_bt + max_tiles - 1) / max_tiles;
            const int tiles_per_chunk = (n_bt + nchunk - 1) / nchunk;
            nchunk = (n_bt + tiles_per_chunk - 1) / tiles_per_chunk;
            TcParams p;
            p.Wimg = d_W; p.Pimg = d_P; p.td = d_td; p.R = d_R;
            p.Sb = Sb; p.n_dt = n_dt; p.n_bt = n_bt; p.nchunk = nchunk; p.tiles_per_chunk = tiles_per_chunk; p.ksteps = ksteps;
            p.G = G; p.Gpad = Gpad; p.a_bytes = a_bytes; p.b_dig = b_dig; p.img_bytes = b_bytes; p.a_plane = a_plane; p.b_plane = b_plane;
            const int64_t items = (int64_t)Sb * nchunk * n_dt;
            int grid = (int)(items < ctx->sm_count ? items : ctx->sm_count);
            p.flags = d_nnz;
            p.out = st ? st->out + S0 * st->out_T : nullptr;
            p.out_T = st ? st->out_T : 0;
            p.out_Q = st ? st->out_Q : 0;
            p.sgn = st ? st->sgn + S0 : nullptr;
            p.skip_math = p.skip_mma = p.epi_delay = p.one_issuer = 0;
            p.dbg = nullptr;
            bool want_dbg = false;
#ifdef IPN_TC_HOOKS
            if (getenv("IPN_TC_GRID")) grid = (int)std::min<int64_t>(items, atoi(getenv("IPN_TC_GRID")));  // several CTA waves
            p.skip_math = getenv("IPN_TC_SKIP_MATH") ? atoi(getenv("IPN_TC_SKIP_MATH")) : 0;
            p.skip_mma = getenv("IPN_TC_SKIP_MMA") != nullptr;
            p.epi_delay = getenv("IPN_TC_EPI_DELAY") ? atoi(getenv("IPN_TC_EPI_DELAY")) : 0;
            p.one_issuer = getenv("IPN_TC_ONE_ISSUER") != nullptr;
            want_dbg = getenv("IPN_TC_DEBUG") != nullptr;
#endif
            if (want_dbg) {
                if ((rc = ctx->dbg.ensure((size_t)grid * 8 * sizeof(unsigned long long)))) return rc;
                IPN_CUDA_CHECK(cudaMemsetAsync(ctx->dbg.p, 0, (size_t)grid * 8 * sizeof(unsigned long long), ctx->stream));
                p.dbg = ctx->dbg.as<unsigned long long>();
            }
            main_kernel_begin(ctx);
            for (int m = 0; m <= N_EPI; ++m) {
                if (!run_mode[m]) continue;
                // the per-draw modes are decided on the device, so in auto mode all three are launched; an instantiation
                // without draws of its mode exits at once (~15 us)
                kern[m]<<<grid, NTHREADS, smem, ctx->stream>>>(p);
                ctx->launches += 1;
            }
            main_kernel_end(ctx);
            IPN_CUDA_CHECK(cudaGetLastError());
            if (want_dbg) {
                std::vector<unsigned long long> h((size_t)grid * 8);
                IPN_CUDA_CHECK(cudaMemcpyAsync(h.data(), ctx->dbg.p, h.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
                IPN_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
                double m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int c = 0; c < grid; ++c)
                    for (int k = 0; k < 8; ++k) m[k] += (double)h[(size_t)c * 8 + k] / grid;
                const double tiles = (double)items * tiles_per_chunk / grid;
                fprintf(stderr,
                        "[ipn tc] per-CTA mean cycles: total %.3g (%.0f/tile, %.0f tiles) | producer waits b_empty %.0f c_empty %.0f | "
                        "mma waits a_full %.0f b_full %.0f t_empty %.0f | epilogue(w2) waits c_full %.0f t_full %.0f  [per tile]\n",
                        m[7], m[7] / tiles, tiles, m[0] / tiles, m[1] / tiles, m[2] / tiles, m[3] / tiles, m[4] / tiles,
                        m[5] / tiles, m[6] / tiles);
            }
        }
        if (!st) {
            rc = finalize_loglik(ctx, S0, Sb, S, G, Gpad, d_R, d_dc, ll, accumulate);
            if (rc) return rc;
        } else {
            // draws of this batch with an unbounded exponent (TcDraw::hifi == 2): the direct kernel, those draws only
            rc = launch_rff_eval_masked(ctx, st->out_T, st->time, S0, Sb, J, omega, a, b, st->shift, st->amp, st->out, d_td);
            if (rc) return rc;
        }
    }
    return IPN_OK;
}

int run_loglik_grid_tc(ipn_ctx* ctx, int64_t T, const double* time, const double* exposure, const int64_t* counts,
                       int64_t S, int64_t J, const double* omega, const double* a, const double* b, const DrawConst* d_dc,
                       int64_t G, const double* delays, double* ll, bool accumulate) {
    return run_tc_grid(ctx, T, time, exposure, counts, S, J, omega, a, b, d_dc, G, delays, 0, ll, accumulate, nullptr);
}

// ------------------------------------------------------------------------------------------------------
// ipn_rff_eval on UNIFORM bins (the only way the reference calls it: mid-points of np.arange edges) as the same contraction.
// With Q = 128 and bin i = p Q + q the argument splits into  t_i - shift_s = (t_0 + p Q dt) - (shift_s - q dt)  -- a "bin"
// time t'_p = t_0 + p Q dt (T / Q of them) and a per-draw "delay" grid D_sq = shift_s - q dt (Q of them) -- so the
// log-intensity of every bin of a draw is ONE (T / Q x 2J) . (2J x Q) product:  log2 lambda[p, q] = log2|amp| + sum_k Phi[p, k]
// W[k, q]: (T / Q + Q) J sine / cosine pairs per draw instead of T J, 2J multiply-adds per (draw, bin) on the tensor pipe
// instead of ~4300 FP32-pipe instructions, and the digit-plane arithmetic keeps the exponent exact to ~1e-9, so
// lambda = 2^x carries only the 7e-8 of the fp32 exp2 (tolerance: 1e-6 relative).  One work item = one draw (T / Q / 32 tiles
// of the kernel above with the store epilogue).  Draws whose exponent is not bounded by 2^120 (tc_draw_kernel) are left to
// the direct kernel (rff_eval.cu), which handles overflow to inf as the reference does.
// ------------------------------------------------------------------------------------------------------
constexpr int RFF_Q = 128;

__global__ void rff_tc_bins_kernel(int64_t Tp, double t0, double step, double* __restrict__ time_p, double* __restrict__ expo_p,
                                   int64_t* __restrict__ counts_p, int* __restrict__ perm, int* __restrict__ flags) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < Tp) {
        time_p[p] = t0 + (double)p * step;
        expo_p[p] = 0.0;
        counts_p[p] = 0;
        perm[p] = (int)p;
    }
    if (p == 0) {
        flags[0] = 0;  // no bin holds counts: the tiles' constants are not used
        flags[1] = 2;  // IPN_OPT_HIFI = 2: TcDraw::hifi = 0 unless the draw's exponent is unbounded (then 2)
        flags[2] = flags[3] = 0;  // sum n^2 = 0.0
    }
}

__global__ void rff_tc_draws_kernel(int64_t S, double dt, const double* __restrict__ shift, const double* __restrict__ amp,
                                    double* __restrict__ delays, double* __restrict__ sgn, DrawConst* __restrict__ dc) {
    const int64_t s = blockIdx.x;
    for (int q = threadIdx.x; q < RFF_Q; q += blockDim.x) delays[s * RFF_Q + q] = shift[s] - (double)q * dt;
    if (threadIdx.x == 0) {
        const double am = amp[s];
        sgn[s] = am > 0.0 ? 1.0 : (am < 0.0 ? -1.0 : 0.0);
        DrawConst d;
        d.C = 0.0;
        d.kap = 0.0;
        const double c = am != 0.0 ? log2(fabs(am)) : 0.0;  // amp = 0: x = log2 f, multiplied by sgn = 0 at the store
        d.c_hi = (float)c;
        d.c_lo = (float)(c - (double)d.c_hi);
        dc[s] = d;
    }
}

bool rff_tc_supported(int64_t T, int64_t J) { return tc_supported(J) && T >= 8 * RFF_Q; }

// time[i] = t0 + i dt (checked by the caller), all arrays on the device
int run_rff_eval_tc(ipn_ctx* ctx, int64_t T, const double* time, double t0, double dt, int64_t S, int64_t J, const double* omega,
                    const double* a, const double* b, const double* shift, const double* amp, double* out) {
    const int64_t Tp = (T + RFF_Q - 1) / RFF_Q;
    // scratch: time' | exposure' | counts' (Tp each) | delays (S x Q) | sgn (S) | DrawConst (S) | flags (8) | perm (Tp)
    const size_t n8 = (size_t)3 * Tp + (size_t)S * RFF_Q + (size_t)S;
    const size_t bytes = n8 * 8 + (size_t)S * sizeof(DrawConst) + 8 * sizeof(int) + (size_t)Tp * sizeof(int) + 64;
    int rc = ctx->rffs.ensure(bytes);
    if (rc) return rc;
    double* time_p = ctx->rffs.as<double>();
    double* expo_p = time_p + Tp;
    int64_t* counts_p = reinterpret_cast<int64_t*>(expo_p + Tp);
    double* delays = reinterpret_cast<double*>(counts_p + Tp);
    double* sgn = delays + (size_t)S * RFF_Q;
    DrawConst* dc = reinterpret_cast<DrawConst*>(sgn + S);
    int* flags = reinterpret_cast<int*>(dc + S);
    int* perm = flags + 8;
    rff_tc_bins_kernel<<<(unsigned)((Tp + 255) / 256), 256, 0, ctx->stream>>>(Tp, t0, (double)RFF_Q * dt, time_p, expo_p, counts_p, perm, flags);
    rff_tc_draws_kernel<<<(unsigned)S, 128, 0, ctx->stream>>>(S, dt, shift, amp, delays, sgn, dc);
    ctx->launches += 2;
    IPN_CUDA_CHECK(cudaGetLastError());
    TcStore st{out, T, RFF_Q, sgn, perm, flags, time, shift, amp};
    return run_tc_grid(ctx, Tp, time_p, expo_p, counts_p, S, J, omega, a, b, dc, RFF_Q, delays, RFF_Q, nullptr, false, &st);
}

}  // namespace ipn
