// Bin permutations of the tcgen05 likelihood kernel (loglik_tc.cu): plain CUDA C++ without inline PTX, kept in a header of
// their own so that tests/emu/perm_emu.cpp can compile these very kernels for the host (a thread-per-CUDA-thread shim with
// real barriers and warp collectives) and check them against a stable sort -- the CPU suite's substitute for a first run
// on the device.
#pragma once
#ifndef IPN_HOST_EMU
#include "ipn_common.cuh"
#endif
#include "loglik_common.cuh"  // tc_bin_position

namespace ipn {

// Bin permutation: bins with counts > 0 first, zero-count bins last (the Poisson sum does not care about the
// order).  Tiles of the Phi image are then homogeneous, and the epilogue of an all-zero tile skips the log
// entirely -- a per-TILE (not per-thread) decision.  Stable partition in two passes over blocks of 1024 bins: per-block
// counts (and sum n^2 for the epilogue choice), then every block adds up the counts of the blocks before it and scatters.
// No atomics, so the permutation and the flags are deterministic.  (A single-block version took 100 us at T = 1e5 --
// 5 % of a single-draw grid.)
__global__ void __launch_bounds__(1024) tc_perm_count_kernel(int64_t T, const int64_t* __restrict__ counts, int* __restrict__ blk_cnt,
                                                             double* __restrict__ blk_sq) {
    __shared__ int wsum[32];
    __shared__ double wsq[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t i = (int64_t)blockIdx.x * 1024 + tid;
    const double n = (i < T) ? (double)counts[i] : 0.0;
    int cnt = n > 0.0;
    double sq = n * n;
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (lane == 0) {
        wsum[warp] = cnt;
        wsq[warp] = sq;
    }
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        double q = 0.0;
        for (int w = 0; w < 32; ++w) {
            t += wsum[w];
            q += wsq[w];
        }
        blk_cnt[blockIdx.x] = t;
        blk_sq[blockIdx.x] = q;
    }
}

__global__ void __launch_bounds__(1024) tc_perm_kernel(int64_t T, const int64_t* __restrict__ counts, int n_blk,
                                                       const int* __restrict__ blk_cnt, const double* __restrict__ blk_sq,
                                                       int* __restrict__ perm, int* __restrict__ n_nz_out, int hifi_mode) {
    __shared__ int wsum[32], wtot[32];
    __shared__ double wsq[32];
    __shared__ int s_before, s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // counts of the blocks before this one, and of all blocks (integer sums: exact, order-independent)
    int before = 0, total = 0;
    double sq = 0.0;
    for (int b = tid; b < n_blk; b += 1024) {
        const int c = blk_cnt[b];
        total += c;
        if (b < (int)blockIdx.x) before += c;
        sq += blk_sq[b];  // integers below 2^53: exact as well
    }
    for (int o = 16; o > 0; o >>= 1) {
        before += __shfl_xor_sync(0xffffffffu, before, o);
        total += __shfl_xor_sync(0xffffffffu, total, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (lane == 0) {
        wsum[warp] = before;
        wtot[warp] = total;
        wsq[warp] = sq;
    }
    __syncthreads();
    if (tid == 0) {
        int bsum = 0, t = 0;
        double q = 0.0;
        for (int w = 0; w < 32; ++w) {
            bsum += wsum[w];
            t += wtot[w];
            q += wsq[w];
        }
        s_before = bsum;
        s_total = t;
        if (blockIdx.x == 0) {
            n_nz_out[0] = t;
            n_nz_out[1] = hifi_mode;
            *reinterpret_cast<double*>(n_nz_out + 2) = q;  // sum n_i^2, read by tc_draw_kernel for the per-draw epilogue choice
        }
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * 1024 + tid;
    const int flag = (i < T) ? (counts[i] > 0) : 0;
    int incl = flag;
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    __syncthreads();  // wsum is reused
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += wsum[w];
    const int nz_before = s_before + woff + incl - flag;
    // rank within the class -> position: tiles of the two classes interleaved (tc_bin_position, loglik_common.cuh)
    if (i < T) perm[tc_bin_position(flag != 0, flag ? nz_before : i - nz_before, s_total, T, 32)] = (int)i;
}

}  // namespace ipn
