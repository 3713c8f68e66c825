// A thread-per-CUDA-thread shim: just enough of the CUDA execution model to run kernels that contain no inline PTX on the
// host, from their own source.  Every CUDA thread of a block is an OS thread, __syncthreads is a real barrier, the warp
// collectives exchange values through a per-warp buffer, __shared__ variables are function-local statics (blocks run one
// after the other).  Test infrastructure only (tests/emu/*.cpp).
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "../../include/ipn.h"  // IPN_MAX_J

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

struct float2 {
    float x, y;
};
struct float4 {
    float x, y, z, w;
};
struct alignas(16) int4 {
    int x, y, z, w;
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float __int_as_float(int i) {
    float f;
    std::memcpy(&f, &i, 4);
    return f;
}
static inline int __float_as_int(float f) {
    int i;
    std::memcpy(&i, &f, 4);
    return i;
}
static inline int __float2int_rn(float x) { return (int)std::nearbyintf(x); }
// every CUDA thread is an OS thread here: shared-memory atomics must be real ones
static inline unsigned atomicMax(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
using std::isfinite;
template <class T>
static inline T min(T a, T b) { return b < a ? b : a; }

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

// launch<<<(gx, gy, gz), block>>>: blocks one after the other, the threads of a block concurrently (block % 32 == 0)
template <class Kernel, class... Args>
void launch3(Kernel k, Dim3 grid, unsigned block, Args... args) {
    gridDim = grid;
    blockDim.x = block;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                block_bar = std::make_unique<std::barrier<>>(block);
                warps.clear();
                for (unsigned w = 0; w < block / 32; ++w) warps.push_back(std::make_unique<Warp>());
                std::vector<std::thread> th;
                th.reserve(block);
                for (unsigned t = 0; t < block; ++t)
                    th.emplace_back([=] {
                        threadIdx.x = t;
                        blockIdx.x = bx;
                        blockIdx.y = by;
                        blockIdx.z = bz;
                        my_warp = warps[t / 32].get();
                        lane = (int)(t % 32);
                        k(args...);
                    });
                for (auto& x : th) x.join();
            }
}
template <class Kernel, class... Args>
void launch(Kernel k, unsigned grid, unsigned block, Args... args) {
    launch3(k, Dim3{grid, 1, 1}, block, args...);
}
}  // namespace shim

// The collectives below assume all 32 lanes of the warp take part (true for every call site in the kernels compiled here:
// whole-warp code, or a branch that a whole warp takes).
inline void __syncthreads() { shim::block_bar->arrive_and_wait(); }
inline unsigned __activemask() { return 0xffffffffu; }
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
inline bool __all_sync(unsigned m, bool p) { return __ballot_sync(m, p) == 0xffffffffu; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }

// MUFU.LG2 is the one instruction of the headers compiled here that has no host counterpart; tests/emu/tc_emu.cpp brings
// its own stand-in with an injectable error, everything else gets the correctly rounded logarithm.
#ifndef IPN_SHIM_OWN_LG2
namespace ipn {
static inline float lg2_approx(float x) { return (float)std::log2((double)x); }
}  // namespace ipn
#endif
