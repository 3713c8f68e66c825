// The bin-permutation kernels of the tcgen05 likelihood path (pyipn_b200/csrc/tc_perm.cuh), compiled for the HOST: test
// infrastructure.  A small shim gives every CUDA thread of a block its own OS thread, __syncthreads a real barrier and the
// warp collectives (__shfl_xor_sync, __ballot_sync, ...) a per-warp exchange, so the kernels run here from their own
// source, block after block, exactly as written -- and are checked against std::stable_sort.
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

#include "cuda_shim.h"

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

    // expected order, built tile by tile (independently of tc_bin_position's closed form) from the two classes -- bins with
    // counts and bins without, each in its original order
    std::vector<int> cls_c, cls_z;
    for (int64_t i = 0; i < T; ++i) (counts[i] > 0 ? cls_c : cls_z).push_back((int)i);
    std::vector<int> want;
#if IPN_TC_INTERLEAVE >= 2
    // mixed tiles while both classes last: 16 bins with counts in one half of the tile and 16 without in the other (mode 3:
    // the halves swap every other pair of tiles); then the rest of the bins with counts, then the rest of those without
    const size_t M = std::min(cls_c.size(), cls_z.size()) / 16;
    for (size_t j = 0; j < M; ++j) {
        const bool swap = IPN_TC_INTERLEAVE == 3 && ((j >> 1) & 1);
        const std::vector<int>& lo = swap ? cls_z : cls_c;
        const std::vector<int>& hi = swap ? cls_c : cls_z;
        want.insert(want.end(), lo.begin() + 16 * j, lo.begin() + 16 * j + 16);
        want.insert(want.end(), hi.begin() + 16 * j, hi.begin() + 16 * j + 16);
    }
    want.insert(want.end(), cls_c.begin() + 16 * M, cls_c.end());
    want.insert(want.end(), cls_z.begin() + 16 * M, cls_z.end());
#else
    // homogeneous tiles: the last tile of the first class is filled up with the first bins of the second; of
    // the FULL tiles of both classes the first m = min(#full C, #full Z) are laid out C Z Z C | C Z Z C | ..., the rest follows
    // in natural order (all remaining C tiles, then the remaining Z tiles, a partial one last)
    std::vector<std::vector<int>> tc_tiles, tz_tiles;
    size_t zi = 0;
    for (size_t k = 0; k < cls_c.size(); k += 32) tc_tiles.emplace_back(cls_c.begin() + k, cls_c.begin() + std::min(cls_c.size(), k + 32));
    if (!tc_tiles.empty())
        while (tc_tiles.back().size() < 32 && zi < cls_z.size()) tc_tiles.back().push_back(cls_z[zi++]);
    for (; zi < cls_z.size(); zi += 32) tz_tiles.emplace_back(cls_z.begin() + zi, cls_z.begin() + std::min(cls_z.size(), zi + 32));
    size_t full_c = 0, full_z = 0;
    for (auto& t : tc_tiles) full_c += t.size() == 32;
    for (auto& t : tz_tiles) full_z += t.size() == 32;
    const size_t m = IPN_TC_INTERLEAVE ? std::min(full_c, full_z) : 0;
    std::vector<std::vector<int>> order(2 * m);
    for (size_t j = 0; j < m; ++j) {
        order[4 * (j / 2) + ((j & 1) ? 3 : 0)] = tc_tiles[j];
        order[4 * (j / 2) + 1 + (j & 1)] = tz_tiles[j];
    }
    for (size_t j = m; j < tc_tiles.size(); ++j) order.push_back(tc_tiles[j]);
    for (size_t j = m; j < tz_tiles.size(); ++j) order.push_back(tz_tiles[j]);
    for (auto& t : order) want.insert(want.end(), t.begin(), t.end());
#endif
    if ((int64_t)want.size() != T) {
        std::printf("expected order has %zu entries for T = %lld\n", want.size(), (long long)T);
        return 1;
    }
    if (n_blk > 0) shim::launch(ipn::tc_perm_count_kernel, n_blk, 1024, T, counts.data(), blk_cnt.data(), blk_sq.data());
    shim::launch(ipn::tc_perm_kernel, n_blk > 0 ? n_blk : 1, 1024, T, counts.data(), n_blk, blk_cnt.data(), blk_sq.data(), perm.data(),
                 flags, 0);
    if (check("tc_perm_kernel", want, 0)) return 1;

    std::printf("ok T=%lld blocks=%d n_nz=%lld\n", (long long)T, n_blk, (long long)n_nz);
    return 0;
}
