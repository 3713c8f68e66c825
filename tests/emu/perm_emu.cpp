// The bin-permutation kernels of the tcgen05 likelihood path (pyipn_b200/csrc/tc_perm.cuh), compiled for the HOST: test
// infrastructure.  A small shim gives every CUDA thread of a block its own OS thread, __syncthreads a real barrier and the
// warp collectives (__shfl_xor_sync, __ballot_sync, ...) a per-warp exchange, so the kernels run here from their own
// source, block after block, exactly as written -- and are checked against std::stable_sort.  This is how the count-sorted
// permutation of the opt-in product epilogue (written after the round's GPU minutes were spent) got its first execution.
//
//   perm_emu <T> <seed> <mean counts per bin>   ->  "ok ..." or a description of the first mismatch (exit code 1)
#include <algorithm>
#include <barrier>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <random>
#include <thread>
#include <vector>

#define IPN_HOST_EMU 1
#define __global__
#define __device__
#define __forceinline__ inline
#define __shared__ static  // one block runs at a time: a function-local static IS the block's shared memory
#define __launch_bounds__(...)

struct Dim3 {
    unsigned x = 1, y = 1, z = 1;
};
static thread_local Dim3 threadIdx, blockIdx;
static Dim3 blockDim, gridDim;

namespace shim {
struct Warp {
    std::barrier<> bar{32};
    uint64_t slot[32];
};
static std::unique_ptr<std::barrier<>> block_bar;
static std::vector<std::unique_ptr<Warp>> warps;
static thread_local Warp* my_warp;
static thread_local int lane;

template <class T>
inline T exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "");
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    my_warp->slot[lane] = raw;
    my_warp->bar.arrive_and_wait();
    const uint64_t got = my_warp->slot[src_lane];
    my_warp->bar.arrive_and_wait();  // nobody overwrites a slot before everyone has read
    T r;
    std::memcpy(&r, &got, sizeof(T));
    return r;
}

template <class Kernel, class... Args>
void launch(Kernel k, unsigned grid, unsigned block, Args... args) {
    gridDim.x = grid;
    blockDim.x = block;
    for (unsigned b = 0; b < grid; ++b) {
        block_bar = std::make_unique<std::barrier<>>(block);
        warps.clear();
        for (unsigned w = 0; w < block / 32; ++w) warps.push_back(std::make_unique<Warp>());
        std::vector<std::thread> th;
        th.reserve(block);
        for (unsigned t = 0; t < block; ++t)
            th.emplace_back([=] {
                threadIdx.x = t;
                blockIdx.x = b;
                my_warp = warps[t / 32].get();
                lane = (int)(t % 32);
                k(args...);
            });
        for (auto& x : th) x.join();
    }
}
}  // namespace shim

inline void __syncthreads() { shim::block_bar->arrive_and_wait(); }
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int o) { return shim::exchange(v, shim::lane ^ o); }
template <class T>
inline T __shfl_up_sync(unsigned, T v, int o) { return shim::exchange(v, shim::lane >= o ? shim::lane - o : shim::lane); }
template <class T>
inline T __shfl_sync(unsigned, T v, int src) { return shim::exchange(v, src); }
inline unsigned __ballot_sync(unsigned, bool p) {
    unsigned m = 0;
    shim::my_warp->slot[shim::lane] = p ? 1u : 0u;
    shim::my_warp->bar.arrive_and_wait();
    for (int l = 0; l < 32; ++l) m |= (unsigned)shim::my_warp->slot[l] << l;
    shim::my_warp->bar.arrive_and_wait();
    return m;
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }

#include "../../pyipn_b200/csrc/tc_perm.cuh"

static int cls_of(int64_t n) { return n <= 0 ? 4 : (n >= 4 ? 3 : (int)n - 1); }

int main(int argc, char** argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: perm_emu <T> <seed> <mean counts per bin>\n");
        return 64;
    }
    const int64_t T = std::atoll(argv[1]);
    std::mt19937_64 rng(std::atoll(argv[2]));
    std::poisson_distribution<int> pois(std::atof(argv[3]));
    std::vector<int64_t> counts(T);
    double sumsq = 0.0;
    int64_t n_nz = 0;
    for (auto& c : counts) {
        c = pois(rng);
        sumsq += (double)c * (double)c;
        n_nz += c > 0;
    }
    const int n_blk = (int)((T + 1023) / 1024);
    std::vector<int> blk_cnt((size_t)(n_blk + 1) * 4, -7), perm((size_t)T + 1, -1);
    std::vector<double> blk_sq((size_t)n_blk + 1, -7.0);
    alignas(8) int flags[4] = {-1, -1, -1, -1};

    auto check = [&](const char* what, const std::vector<int>& want, int mode) -> int {
        for (int64_t i = 0; i < T; ++i)
            if (perm[i] != want[i]) {
                std::printf("%s: perm[%lld] = %d, expected %d\n", what, (long long)i, perm[i], want[i]);
                return 1;
            }
        double sq;
        std::memcpy(&sq, flags + 2, 8);
        if (flags[0] != (int)n_nz || flags[1] != mode || sq != sumsq) {
            std::printf("%s: flags (%d, %d, %.17g), expected (%lld, %d, %.17g)\n", what, flags[0], flags[1], sq, (long long)n_nz, mode, sumsq);
            return 1;
        }
        if (perm[T] != -1) {
            std::printf("%s: wrote past the end\n", what);
            return 1;
        }
        return 0;
    };

    // the partition of the default epilogues: bins with counts first, stable
    std::vector<int> want(T);
    for (int64_t i = 0; i < T; ++i) want[i] = (int)i;
    std::stable_sort(want.begin(), want.end(), [&](int a, int b) { return (counts[a] > 0) > (counts[b] > 0); });
    if (n_blk > 0) shim::launch(ipn::tc_perm_count_kernel, n_blk, 1024, T, counts.data(), blk_cnt.data(), blk_sq.data());
    shim::launch(ipn::tc_perm_kernel, n_blk > 0 ? n_blk : 1, 1024, T, counts.data(), n_blk, blk_cnt.data(), blk_sq.data(), perm.data(),
                 flags, 0);
    if (check("tc_perm_kernel", want, 0)) return 1;

    // the count-sorted permutation of the product epilogue: classes 1, 2, 3, >= 4, then 0, each in its original order
    std::fill(perm.begin(), perm.end(), -1);
    std::fill(blk_cnt.begin(), blk_cnt.end(), -7);
    std::stable_sort(want.begin(), want.end(), [&](int a, int b) { return cls_of(counts[a]) < cls_of(counts[b]); });
    // (want was already a stable partition; re-sorting from identity keeps the reference independent of it)
    for (int64_t i = 0; i < T; ++i) want[i] = (int)i;
    std::stable_sort(want.begin(), want.end(), [&](int a, int b) { return cls_of(counts[a]) < cls_of(counts[b]); });
    if (n_blk > 0) shim::launch(ipn::tc_perm_sorted_count_kernel, n_blk, 1024, T, counts.data(), blk_cnt.data(), blk_sq.data());
    shim::launch(ipn::tc_perm_sorted_kernel, n_blk > 0 ? n_blk : 1, 1024, T, counts.data(), n_blk, blk_cnt.data(), blk_sq.data(),
                 perm.data(), flags, 4);
    if (check("tc_perm_sorted_kernel", want, 4)) return 1;
    std::printf("ok T=%lld blocks=%d n_nz=%lld\n", (long long)T, n_blk, (long long)n_nz);
    return 0;
}
