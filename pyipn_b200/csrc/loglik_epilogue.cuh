// Arithmetic of the tcgen05 delay-grid kernel that is independent of the hardware plumbing: the signed 8-bit digit split
// of the operand images and the fused Poisson epilogue (exponent assembly from the four level sums, Poisson terms,
// compensated sums).  Kept apart from loglik_tc.cu so that tests/emu/tc_emu.cpp can compile exactly these lines with g++
// (IPN_HOST_EMU) and check the error budget of DESIGN.md section 3 at full size on the CPU.
#pragma once
#include "loglik_common.cuh"

namespace ipn {

__device__ __forceinline__ void split_digits4(int q, int8_t& d1, int8_t& d2, int8_t& d3, int8_t& d4) {
    int r = q;
    const int e4 = ((r + 128) & 255) - 128;
    r = (r - e4) >> 8;
    const int e3 = ((r + 128) & 255) - 128;
    r = (r - e3) >> 8;
    const int e2 = ((r + 128) & 255) - 128;
    r = (r - e2) >> 8;
    d1 = (int8_t)r;
    d2 = (int8_t)e2;
    d3 = (int8_t)e3;
    d4 = (int8_t)e4;
}

// Five signed 8-bit digits of a 40-bit integer (W sigma 2^38): q = d1 2^32 + d2 2^24 + d3 2^16 + d4 2^8 + d5, d2..d5 in [-128, 127]
__device__ __forceinline__ void split_digits5(long long q, int8_t& d1, int8_t& d2, int8_t& d3, int8_t& d4, int8_t& d5) {
    const int e5 = (int)(((q + 128) & 255) - 128);
    const long long r = (q - e5) >> 8;  // |r| <= 2^30 + 1
    d5 = (int8_t)e5;
    split_digits4((int)r, d1, d2, d3, d4);
}

// The level-6 sum (W's fifth digit times Phi's first and W's fourth times Phi's second, weight 2^-8 of level 5) brought into
// level-5 units right after the epilogue threads have loaded it: an arithmetic shift, i.e. rounded DOWN -- the half unit this
// loses on average is added back where the low levels are converted (kLevel6Half in exponent_parts and in the fp64 epilogue), so
// the fold costs two integer instructions per bin and leaves a zero-mean error of 2^-37 / sigma in x.
__device__ __forceinline__ int fold_level6(int s5, int s6) { return s5 + (s6 >> 8); }

// Eight consecutive bins (TMEM columns) of one delay (TMEM lane): exponent assembly, Poisson terms, compensated sum.
// HC = the tile holds bins with counts > 0 (log needed); all-zero tiles only need -kappa u.
// The units take the levels either as four arrays (S2 .. S5) or already combined into H = 256 S2 + S3 and L = 256 S4 + S5
// (levels_high / levels_low): the same integers either way.
#define IPN_COMBINE_LEVELS8(H, L)                    \
    int H[8], L[8];                                  \
    _Pragma("unroll") for (int j = 0; j < 8; ++j) { \
        H[j] = levels_high(s2[j], s3[j]);            \
        L[j] = levels_low(s4[j], s5[j]);             \
    }
template <bool HC>
__device__ __forceinline__ void epi_units8(const int (&H)[8], const int (&L)[8],
                                           const float4* __restrict__ cst, int sh, float inv, float ks24, float& ssum, float& scomp, float& slo) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 c = cst[j];  // warp-uniform (n, kappa_hi, kappa_lo, n log2e)
        // x = ks (S2 + 2^-8 S3 + 2^-16 S4 + 2^-24 S5): integer part and fraction from the integers (exponent_parts)
        int k;
        const float xf = exponent_parts_hl(H[j], L[j], sh, inv, ks24, k);
        const float u = exp2_parts(xf, k);
        if (HC) {
            float l;
            v[j] = poisson_term(u, c.x, c.w, c.y, c.z, l);
            slo += l;
        } else {
            v[j] = -fmaf(c.y, u, c.z * u);
        }
    }
    // the terms are rounded at their own size, and so is every partial sum: a pairwise sum of all eight before one Kahan step
    // (|partial| <= 8 |v|) turned out to be most of this epilogue's error on data with bright bins (emulated on the bench
    // draws: 1.1e-4 nats rms against 6.8e-5 with a Kahan step per PAIR, 5.9e-5 with one per term) -- so: pairs
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        const float yk = (v[j] + v[j + 1]) - scomp;
        const float tt = ssum + yk;
        scomp = (tt - ssum) - yk;
        ssum = tt;
    }
}

template <bool HC>
__device__ __forceinline__ void epi_units8(const int (&s2)[8], const int (&s3)[8], const int (&s4)[8], const int (&s5)[8],
                                           const float4* __restrict__ cst, int sh, float inv, float ks24, float& ssum, float& scomp, float& slo) {
    IPN_COMBINE_LEVELS8(H, L)
    epi_units8<HC>(H, L, cst, sh, inv, ks24, ssum, scomp, slo);
}

// Store epilogue (ipn_rff_eval on uniform bins through the same contraction): no Poisson terms, just u = 2^x of the eight bins.
__device__ __forceinline__ void epi_units8_store(const int (&H)[8], const int (&L)[8], int sh, float inv, float ks24, float (&u)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int k;
        const float xf = exponent_parts_hl(H[j], L[j], sh, inv, ks24, k);
        u[j] = exp2_parts(xf, k);
    }
}
__device__ __forceinline__ void epi_units8_store(const int (&s2)[8], const int (&s3)[8], const int (&s4)[8], const int (&s5)[8],
                                                 int sh, float inv, float ks24, float (&u)[8]) {
    IPN_COMBINE_LEVELS8(H, L)
    epi_units8_store(H, L, sh, inv, ks24, u);
}

// Compensated fp32 epilogue for moderately bright data (tens of counts per bin).  What limits the plain version there is
// not the exponent or the logarithm but three roundings of O(n)-sized quantities: n*y0, kappa*u and their difference
// (6e-8 |v| each, |v| ~ n), and the uncompensated pairwise sum of such terms.  Here the two products are formed as fp32
// double-floats (exact FMA error terms), the small parts go to the `slo` sum, and every product takes its own Kahan step.  ~11 more instructions per (delay, bin); remaining error ~3.7e-8 sqrt(sum n_i^2)
// nats rms (measured: 1e-3 max at sum n_i^2 = 1.1e8) from the 2^(-y0) refinement of the logarithm, hence the hand-over to
// the fp64 epilogue at sum n_i^2 = 4e7.
// The two products are NOT subtracted from one another per bin: n y0 goes into one Kahan sum (ssum, scomp), kappa u into a
// second one (qsum, qcomp), and only their rounding errors (exact, by FMA) and the small corrections meet in `slo`.  That saves
// the TwoSum of the difference (3 instructions per bin less than summing d = n y0 - kappa u with one Kahan step) and gives
// two independent dependency chains.  The caller forms (ssum - scomp) - (qsum - qcomp) + slo in float64 per work item.
template <bool HC>
__device__ __forceinline__ void epi_units8_comp(const int (&H)[8], const int (&L)[8],
                                                const float4* __restrict__ cst, int sh, float inv, float ks24, float& ssum, float& scomp,
                                                float& qsum, float& qcomp, float& slo) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 c = cst[j];  // warp-uniform (n, kappa_hi, kappa_lo, n log2e)
        int k;
        const float xf = exponent_parts_hl(H[j], L[j], sh, inv, ks24, k);
        float ul;  // u = 2^x as hi + lo
        const float u = exp2_parts_df(xf, k, ul);
        // kappa u = q + (eq + klo u + khi ul), q = fl(khi u)
        const float q = c.y * u;
        const float qe = fmaf(c.y, ul, fmaf(c.z, u, fmaf(c.y, u, -q)));
        {
            const float yk = q - qcomp;  // Kahan step of the kappa u sum
            const float tt = qsum + yk;
            qcomp = (tt - qsum) - yk;
            qsum = tt;
        }
        if (HC) {
            const float w = 1.0f + u;
            const float ew = (u - (w - 1.0f)) + ul;
            const float y0 = lg2_approx(w);
            const float tm = -y0 + kMagic;
            const float t = exp2_core(-y0 - (tm - kMagic), tm);
            const float r = fmaf(ew, t, fmaf(w, t, -1.0f));
            // n log2(1 + u) = pr + (epr + nl r), pr = fl(n y0)
            const float pr = c.x * y0;
            const float epr = fmaf(c.x, y0, -pr);
            slo += fmaf(c.w, r, epr) - qe;
            const float yk = pr - scomp;  // Kahan step of the n y0 sum
            const float tt = ssum + yk;
            scomp = (tt - ssum) - yk;
            ssum = tt;
        } else {
            slo -= qe;
        }
    }
}

template <bool HC>
__device__ __forceinline__ void epi_units8_comp(const int (&s2)[8], const int (&s3)[8], const int (&s4)[8], const int (&s5)[8],
                                                const float4* __restrict__ cst, int sh, float inv, float ks24, float& ssum, float& scomp,
                                                float& qsum, float& qcomp, float& slo) {
    IPN_COMBINE_LEVELS8(H, L)
    epi_units8_comp<HC>(H, L, cst, sh, inv, ks24, ssum, scomp, qsum, qcomp, slo);
}

// fp64 epilogue for bright / coarsely binned data (hundreds of counts per bin): the fp32 terms above carry
// ~6e-8 n_i relative noise each, which exceeds 1e-3 nats once sum n_i^2 is large.  4-5x slower, selected per draw.
__device__ __forceinline__ void epi_units8_hifi(const int (&H)[8], const int (&L)[8], const float4* __restrict__ cst, double ksd, double& acc) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 c = cst[j];
        // S2 + 2^-8 S3 = 2^-8 H and 256 S4 + S5 = L, all exact in float64
        const double x = ksd * (0.00390625 * (double)H[j] + 5.9604644775390625e-8 * ((double)L[j] + (double)kLevel6Half));
        const double u = exp2(x);
        const double kap = (double)c.y + (double)c.z;
        acc += (double)c.x * (log1p(u) * kLog2e) - kap * u;
    }
}

__device__ __forceinline__ void epi_units8_hifi(const int (&s2)[8], const int (&s3)[8], const int (&s4)[8],
                                                const int (&s5)[8], const float4* __restrict__ cst, double ksd, double& acc) {
    IPN_COMBINE_LEVELS8(H, L)
    epi_units8_hifi(H, L, cst, ksd, acc);
}

}  // namespace ipn
