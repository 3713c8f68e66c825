// Pieces shared by the FP32-pipe and the tcgen05 delay-grid log-likelihood kernels.
//
// Arithmetic restated from pyipn/stan_models/functions.stan:23-35 (see DESIGN.md "Likelihood algebra"):
//   mu_i = e_i (A f_i + B) = e_i B (1 + u_i),  u_i = (A/B) f(t_i - Delta) = 2^(x_i),
//   x_i  = log2(A/B) + log2(e) * sum_j [a_j cos(w_j (t_i - Delta)) + b_j sin(w_j (t_i - Delta))]
//        = c_s + sum_k Phi[i,k] W[k,g]                     (angle addition; Phi depends on bin, W on delay)
//   ll   = C_s + ln2 * R,   R = sum_i [ n_i log2(1 + u_i) - kappa_i u_i ],  kappa_i = e_i B / ln2
//   C_s  = sum_i [ n_i log(e_i) ] + N log B - B sum_i e_i - sum_i lgamma(n_i + 1)          (fp64, once per draw)
// Every constant factor that multiplies a sum over ~1e5 bins is applied in fp64 (or as an fp32 hi/lo pair):
// a single fp32 rounding of such a factor is a 6e-8 relative bias on a sum of ~5e4 nats, i.e. > 1e-3 nats.
//
// IPN_HOST_EMU: tests/emu/epilogue_emu.cpp compiles the arithmetic below with g++ (it supplies __device__, fmaf-exact
// int/float bit casts and a stand-in for MUFU.LG2) so that the error budget of DESIGN.md section 3 is checked on the CPU.
#pragma once
#ifndef IPN_HOST_EMU
#include "ipn_common.cuh"
#endif

namespace ipn {

constexpr double kLog2e = 1.4426950408889634074;
constexpr double kLn2 = 0.69314718055994530942;

// per-call bin constants: [0] = sum n, [1] = sum n log e, [2] = sum e, [3] = sum lgamma(n+1)
struct DrawConst {
    double C;      // additive constant of the draw (natural log units)
    double kap;    // B / ln2
    float c_hi;    // log2(A/B) split in two fp32
    float c_lo;
};

// per-draw scale and epilogue choice of the tcgen05 path (tc_draw_kernel, tc_images.cuh)
struct TcDraw {
    float ks;     // kappa = 2^-12 / sigma
    float sigma;  // power of two
    int hifi;     // epilogue mode of this draw: 0 fp32, 1 compensated fp32, 2 fp64 (EPI_* in loglik_tc.cu)
    int pad;
};

// ---- order of the bins in the Phi image of the tcgen05 path ----
// Bins with counts cost the epilogue 2.4x the instructions of bins without (the logarithm), and with twelve digit-plane products
// a tile of either kind costs the tensor pipe the same: tiles WITH counts are epilogue-bound (the tensor pipe idles), tiles
// WITHOUT are tensor-bound (the epilogue warps idle).  So the two kinds are interleaved: the bins are partitioned (counts
// first, original order within each class, so every 32-bin tile is homogeneous up to the one boundary tile), and the TILES
// are laid out  C Z Z C | C Z Z C | ...  (C: with counts, Z: without) while both kinds last -- any two consecutive tiles hold
// one of each, and each of the four epilogue warps of a lane quarter, which take units of every other tile, alternates between
// the kinds as well.  Only full tiles take part; what is left of the longer class follows in its natural order.
// The epilogue decides "logarithm or not" per 8-bin UNIT (the flag block of the Phi image holds one flag per 8 rows: the
// unit's first bin holds counts -- within any 8 rows of any layout here the bins with counts come first).
// Layouts that were measured against this one (tools/gpu_mixed.sh, profiles/r02_tile_layouts_ab.txt):
//   -DIPN_TC_INTERLEAVE=0: everything in natural order (the round-1 layout): 2 % slower
//   -DIPN_TC_INTERLEAVE=2: the kinds mixed WITHIN every tile while both last -- rows 0..15 of tile j hold the bins with counts
//      of rank 16 j .. 16 j + 15, rows 16..31 the bins without counts of the same ranks, so every tile costs the epilogue the
//      same: 4-6 % SLOWER.  Of the four epilogue warps of a sub-partition two then always hold the 16-column group with the
//      logarithm and two the one without; the pair with the logarithm becomes the critical path (the issuer waits 1100 cycles
//      per tile for their accumulator buffer, against 340-400 in the shipped layout) while the other two idle.
//   -DIPN_TC_INTERLEAVE=3: as 2, the halves swapping every other pair of tiles (every warp alternates): 3-5 % slower.
// tc_bin_position: position in the permuted order of the bin with rank r within its class (n_nz bins with counts among T).
#ifndef IPN_TC_INTERLEAVE
#define IPN_TC_INTERLEAVE 1
#endif
__device__ __forceinline__ long long tc_bin_position(bool has_counts, long long r, long long n_nz, long long T, int TN) {
    if (IPN_TC_INTERLEAVE >= 2) {
        const long long H = TN / 2, n_z = T - n_nz;
        const long long M = (n_nz < n_z ? n_nz : n_z) / H;  // mixed tiles
        if (r < M * H) {
            const long long j = r / H;
            const bool upper = IPN_TC_INTERLEAVE == 3 ? (has_counts == (((j >> 1) & 1) != 0)) : !has_counts;
            return j * TN + (upper ? H : 0) + r % H;
        }
        return M * TN + (has_counts ? 0 : n_nz - M * H) + (r - M * H);
    }
    const long long Tc = (n_nz + TN - 1) / TN;   // tiles that start with a bin with counts
    const long long hole = Tc * TN - n_nz;       // rows of the last of them that bins without counts fill up
    const long long n_z = T - n_nz;
    const long long Fz = n_z > hole ? (n_z - hole) / TN : 0;  // full tiles without counts
    const long long Fc = n_z >= hole ? Tc : Tc - 1;           // full tiles with counts
    long long m = Fc < Fz ? Fc : Fz;                          // tiles of each kind that are interleaved
    if (!IPN_TC_INTERLEAVE || m < 0) m = 0;
    long long tile, row;
    if (has_counts) {
        const long long j = r / TN;
        row = r - j * TN;
        tile = j < m ? 4 * (j / 2) + ((j & 1) ? 3 : 0) : m + j;
    } else if (r < hole) {
        const long long j = Tc - 1;
        row = (n_nz - j * TN) + r;
        tile = j < m ? 4 * (j / 2) + ((j & 1) ? 3 : 0) : m + j;
    } else {
        const long long rr = r - hole, j = rr / TN;
        row = rr - j * TN;
        tile = j < m ? 4 * (j / 2) + 1 + (j & 1) : Tc + j;
    }
    return tile * TN + row;
}

#ifndef IPN_HOST_EMU
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#endif

// 2^x on the FP32 pipe.  MUFU.EX2 is not usable here: measured on B200 its relative error has a MEAN of
// -8e-9 (tools/mufu_probe.cu), and a bias b on every bin becomes b * sum_i kappa_i u_i ~ 1e-3 nats over 1e5
// bins.  For the same reason the polynomial must not have a systematic error either: a base-2 polynomial with
// fp32 coefficients cannot represent ln2 (2.7e-9 relative) and that alone is ~1e-4 nats.  So: z = xf ln2 with
// ln2 as an fp32 hi/lo pair (one rounding), then e^z with EXACT leading coefficients 1, 1, 1/2 and fitted
// c3..c7 (|z| <= 0.375; max error 5.3e-10, mean over any eighth of the interval < 2.6e-10, fp32 emulation:
// mean -3e-11, rms 2.6e-8, max 7.2e-8).
constexpr float kMagic = 12582912.0f;        // 1.5 * 2^23: adding it rounds to the nearest integer
constexpr int kMagicBits = 0x4B400000;

// 2^xf * 2^xi for |xf| <= 0.54; t = float whose low mantissa bits hold the integer part (xi + magic)
__device__ __forceinline__ float exp2_core(float xf, float t) {
    const float z = fmaf(xf, 0.693145751953125f, xf * 1.4286068203094172e-06f);
    float p = fmaf(z, 0.00019602585234679282f, 0.0013943274971097708f);
    p = fmaf(p, z, 0.008333928883075714f);
    p = fmaf(p, z, 0.041666384786367416f);
    p = fmaf(p, z, 0.166666641831398f);
    p = fmaf(p, z, 0.5f);
    p = fmaf(p, z, 1.0f);
    p = fmaf(p, z, 1.0f);
    return __int_as_float(__float_as_int(p) + ((__float_as_int(t) - kMagicBits) << 23));
}

// The same with the rounding error of the last Horner step recovered: 2^xf 2^xi = hi + lo.  That step (1 + z q, a result in
// [0.7, 1.5] rounded to 2^-24) is most of the 2.6e-8 rms of exp2_core; 1 - hi is exact there, so one more FMA returns what the
// rounding dropped.  What remains (~1.2e-8 rms) are the roundings of xf, of z and of q.  Used by the compensated epilogue,
// whose other terms are exact: on grossly mismatched data the error of u is what is left (it enters the sum with sqrt(sum r_i^2)).
__device__ __forceinline__ float exp2_core_df(float xf, float t, float& lo) {
    const float z = fmaf(xf, 0.693145751953125f, xf * 1.4286068203094172e-06f);
    float p = fmaf(z, 0.00019602585234679282f, 0.0013943274971097708f);
    p = fmaf(p, z, 0.008333928883075714f);
    p = fmaf(p, z, 0.041666384786367416f);
    p = fmaf(p, z, 0.166666641831398f);
    p = fmaf(p, z, 0.5f);
    p = fmaf(p, z, 1.0f);
    const float hi = fmaf(p, z, 1.0f);
    const float l = fmaf(p, z, 1.0f - hi);
    const float sc = __int_as_float(((__float_as_int(t) - kMagicBits) << 23) + 0x3f800000);  // 2^xi, |xi| <= 120
    lo = l * sc;
    return hi * sc;
}

// ---- exponent assembly of the tcgen05 epilogues in integers ----
// x = ks [S2 + 2^-8 S3 + 2^-24 (256 S4 + S5)] with ks = 2^-(12 + e) a power of two: H = 256 S2 + S3 is an exact 32-bit
// integer (|S2| < 2^20, |S3| < 2^22) and x = H 2^-sh + L ks 2^-24 with sh = 20 + e (14 ... 30: tc_draw_kernel keeps e in
// [-6, 10]).  So the integer part of x is a rounding shift of H and the fraction one int -> float conversion: 9 instructions
// (LEA, IADD, SHF, LEA, LEA, 2 I2FP, FMUL, FFMA) where the floating-point route needed 12 (two magic-number conversions,
// three FFMA, five FADD for rint and the remainder), and the exponent goes into the result with one LEA as before.
// Returns the fraction xf (|xf| <= 1/2 + |low levels|), k = the integer part.
// kLevel6Half: S5 arrives with the level-6 sum folded in by an arithmetic shift (fold_level6, loglik_epilogue.cuh), i.e. half a
// level-5 unit too low on average; the half goes back in through the addend of the conversion's FMA (no extra instruction).
constexpr float kLevel6Half = 0.5f;
// The two integers the assembly works on.  The epilogue threads form them right behind the tensor-memory loads, which halves the
// registers the accumulators of a work unit occupy while the unit before it is in flight (loglik_tc.cu).
__device__ __forceinline__ int levels_high(int s2, int s3) { return s2 * 256 + s3; }
__device__ __forceinline__ int levels_low(int s4, int s5) { return s4 * 256 + s5; }
__device__ __forceinline__ float exponent_parts_hl(int H, int L, int sh, float inv, float ks24, int& k) {
    k = (H + (1 << (sh - 1))) >> sh;
    const int fr = H - (k << sh);
    // the half unit rides in the FMA of the LOW levels: their product has a full mantissa of pseudo-random bits, so the tiny
    // addend survives the rounding on average; behind the exact fr 2^-sh (few significant bits) it would be rounded away every
    // time (measured in the emulation: a bias of -0.5 ks 2^-24 in x)
    return fmaf((float)fr, inv, fmaf((float)L, ks24, kLevel6Half * ks24));  // inv = 2^-sh
}
__device__ __forceinline__ float exponent_parts(int s2, int s3, int s4, int s5, int sh, float inv, float ks24, int& k) {
    return exponent_parts_hl(levels_high(s2, s3), levels_low(s4, s5), sh, inv, ks24, k);
}
// e^(xf ln 2) 2^k, |xf| <= 0.54, |k| <= 120
__device__ __forceinline__ float exp2_parts(float xf, int k) {
    const float z = fmaf(xf, 0.693145751953125f, xf * 1.4286068203094172e-06f);
    float p = fmaf(z, 0.00019602585234679282f, 0.0013943274971097708f);
    p = fmaf(p, z, 0.008333928883075714f);
    p = fmaf(p, z, 0.041666384786367416f);
    p = fmaf(p, z, 0.166666641831398f);
    p = fmaf(p, z, 0.5f);
    p = fmaf(p, z, 1.0f);
    p = fmaf(p, z, 1.0f);
    return __int_as_float(__float_as_int(p) + (k << 23));
}
// ... as hi + lo (see exp2_core_df)
__device__ __forceinline__ float exp2_parts_df(float xf, int k, float& lo) {
    const float z = fmaf(xf, 0.693145751953125f, xf * 1.4286068203094172e-06f);
    float p = fmaf(z, 0.00019602585234679282f, 0.0013943274971097708f);
    p = fmaf(p, z, 0.008333928883075714f);
    p = fmaf(p, z, 0.041666384786367416f);
    p = fmaf(p, z, 0.166666641831398f);
    p = fmaf(p, z, 0.5f);
    p = fmaf(p, z, 1.0f);
    const float hi = fmaf(p, z, 1.0f);
    const float l = fmaf(p, z, 1.0f - hi);
    const float sc = __int_as_float((k << 23) + 0x3f800000);
    lo = l * sc;
    return hi * sc;
}

__device__ __forceinline__ float exp2_poly(float x) {
    x = fminf(fmaxf(x, -125.0f), 127.0f);
    const float t = x + kMagic;
    return exp2_core(x - (t - kMagic), t);
}

// 2^(xh + xl) with xh exact (e.g. an integer multiple of a power of two) and xl a small correction (|xl| < 1), for
// |xh + xl| <= 120: no clamps.
// The tcgen05 path bounds the exponent per draw instead (tc_draw_kernel: |c| + log2e sum_j |(a_j, b_j)| <= 120, ~40 for a draw
// of the prior; a draw beyond the bound goes to the fp64 epilogue, whose exp2 needs none), which saves the two FMNMX of
// every (delay, bin).  The bound also keeps 1 + u below 2^121, so the log refinement below needs no clamp of its own.
// -DIPN_EPI_CLAMP restores the clamped form (A/B timing).
__device__ __forceinline__ float exp2_hilo_bounded(float xh, float xl) {
    // the integer part is taken from xh + xl, not from xh alone: xl (the level-3 and lower sums) reaches a few tenths for
    // draws with large amplitudes (sigma <= 1/2), and (xh - rint(xh)) + xl would then leave the polynomial's interval --
    // a systematic error of 1e-8 .. 1e-7 in u that showed as +2e-2 nats on bright, grossly mismatched data
    const float t = (xh + xl) + kMagic;
    return exp2_core((xh - (t - kMagic)) + xl, t);
}
__device__ __forceinline__ float exp2_hilo_df(float xh, float xl, float& lo) {
    const float t = (xh + xl) + kMagic;
    return exp2_core_df((xh - (t - kMagic)) + xl, t, lo);
}
// the clamped form, for callers without a per-draw bound (the FP32-pipe kernel)
__device__ __forceinline__ float exp2_hilo_clamped(float xh, float xl) {
    return exp2_hilo_bounded(fminf(fmaxf(xh, -125.0f), 120.0f), xl);
}
__device__ __forceinline__ float exp2_hilo(float xh, float xl) {
#ifdef IPN_EPI_CLAMP
    xh = fminf(fmaxf(xh, -125.0f), 120.0f);
#endif
    return exp2_hilo_bounded(xh, xl);
}

// One bin's contribution in base-2 units:  n log2(1 + u) - (khi + klo) u  as a main part (ONE rounding of the
// exact n y0 - kappa u) plus a small correction `lo` that the caller sums separately in fp32:
//   w = fl(1 + u),  e_w = (1 + u) - w  recovered exactly (Fast2Sum; valid for every u < 2^24),
//   y0 = MUFU.LG2(w)  (measured mean error +4e-8 on [1,2): unusable alone, same argument as for EX2),
//   t = 2^(-y0) from the unbiased exp2_core,  r = w t - 1  (|r| < 1e-6),
//   log2(1 + u) = y0 + log2e (r + e_w t) + O(1e-13).
// y0 + log2e r must NOT be rounded back into one fp32: whenever the correction is below half an ulp of y0 the
// sum returns y0, i.e. it would silently drop exactly the bias the correction is there to remove.  The e_w term
// also makes the tiny-u case exact (w = 1, y0 = 0, r = 0, e_w = u  ->  log2e u).
//   nl = n log2e (precomputed per bin).
__device__ __forceinline__ float poisson_term(float u, float n, float nl, float khi, float klo, float& lo) {
    const float ku = fmaf(khi, u, klo * u);
    const float w = 1.0f + u;
    const float ew = u - (w - 1.0f);
    const float y0 = lg2_approx(w);
    const float tm = -y0 + kMagic;
    const float t = exp2_core(-y0 - (tm - kMagic), tm);
    const float r = fmaf(ew, t, fmaf(w, t, -1.0f));
    lo = nl * r;
    return fmaf(n, y0, -ku);
}

// sin and cos of r (|r| <= pi + eps) given as fp32 hi + lo; see rff_eval.cu for the derivation of the constants.
__device__ __forceinline__ void sincos_reduced_f32(float rh, float rl, float& s, float& c) {
    const float qf = rintf(rh * 0.636619772367581343f);
    const int q = (int)qf;
    float y = fmaf(qf, -1.57079637050628662109375f, rh);
    y = fmaf(qf, 4.37113900018624283e-8f, y);
    y += rl;
    const float y2 = y * y;
    float ps = fmaf(y2, 2.7237618209228658e-06f, -0.00019839989971813914f);
    ps = fmaf(ps, y2, 0.008333331692152657f);
    ps = fmaf(ps, y2, -0.16666666663377125f);
    const float sy = fmaf(ps * y2, y, y);
    float pc = fmaf(y2, -2.7290681690617153e-07f, 2.4800519567960354e-05f);
    pc = fmaf(pc, y2, -0.0013888887519541376f);
    pc = fmaf(pc, y2, 0.04166666666392175f);
    const float cy = fmaf(fmaf(pc, y2, -0.5f), y2, 1.0f);
    const float ss = (q & 1) ? cy : sy;
    const float cc = (q & 1) ? sy : cy;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
}

// Near-minimax polynomials fitted in this repo (Chebyshev-node interpolation in y^2; truncation error 1.3e-11 /
// 8.4e-13) on |y| <= pi/4.
// pi/2-quadrant sine / cosine of 2 pi (fh + cross) turns, |fh| <= 1/2 exact, cross a small correction
__device__ __forceinline__ void sincos_turns(float fh, float cross, float& s, float& c) {
    const float t4 = fmaf(fh, 4.0f, 12582912.0f);  // 1.5 * 2^23 + rint(4 fh): quadrant index in the low mantissa bits
    const float qf = t4 - 12582912.0f;
    const int q = __float_as_int(t4);
    float y = fmaf(qf, -0.25f, fh) + cross;         // first part exact (Sterbenz)
    y = fmaf(y, 6.2831854820251465f, y * -1.7484555e-07f);  // radians, 2 pi as fp32 hi + lo
    const float y2 = y * y;
    float ps = fmaf(y2, 2.7237618209228658e-06f, -0.00019839989971813914f);
    ps = fmaf(ps, y2, 0.008333331692152657f);
    ps = fmaf(ps, y2, -0.16666666663377125f);
    const float sy = fmaf(ps * y2, y, y);
    float pc = fmaf(y2, -2.7290681690617153e-07f, 2.4800519567960354e-05f);
    pc = fmaf(pc, y2, -0.0013888887519541376f);
    pc = fmaf(pc, y2, 0.04166666666392175f);
    const float cy = fmaf(fmaf(pc, y2, -0.5f), y2, 1.0f);
    const float ss = (q & 1) ? cy : sy;
    const float cc = (q & 1) ? sy : cy;
    // signs by XOR into the sign bit: bit 1 of q for the sine, bit 1 of q + 1 for the cosine
    s = __int_as_float(__float_as_int(ss) ^ ((q << 30) & 0x80000000));
    c = __int_as_float(__float_as_int(cc) ^ (((q + 1) << 30) & 0x80000000));
}

// (sin, cos) of 2 pi w t for w = wh + wl (turns per unit time) and t = th + tl, all fp32 double-float splits of fp64 values:
// ph = fl(wh th), exact FMA error term, cross terms, k = rint(ph) by magic-number add (needs |ph| < 2^22), ph - k exact.
// Same accuracy as the fp64 reduction below (max 8.5e-8, rms 1.8e-8 at |phase| = 1e5 rad) without any fp64 instruction.
__device__ __forceinline__ bool sincos_phase_df(float wh, float wl, float th, float tl, float& s, float& c) {
    const float ph = wh * th;
    if (!(fabsf(ph) < 4194304.0f)) return false;  // caller takes the fp64 route
    const float e = fmaf(wh, th, -ph);
    const float cross = fmaf(wh, tl, fmaf(wl, th, e));
    const float k = (ph + 12582912.0f) - 12582912.0f;
    sincos_turns(ph - k, cross, s, c);
    return true;
}

// ---- Phi in fixed point: (sin, cos)(2 pi w t) as integers scaled by 2^30 (the four 8-bit digit planes of the tcgen05 path) ----
// A float result (24 bits, 2e-8 rms from the polynomial) limits Phi to ~2^-23; what the digit planes need is 2^-31.  Instead
// of a longer polynomial in higher precision: a table of the 1024 points k/1024 of the unit circle in 2^-30 fixed point
// (built once per context in float64, 8 KB, staged in shared memory by the image kernel) plus a first-order rotation by
// the remainder, |theta| <= 2 pi / 2048 = 3.07e-3:
//   sin(a + theta) = s + (s dc + c ds),  cos(a + theta) = c + (c dc - s ds),  ds = theta - theta^3 / 6,  dc = -theta^2 / 2
// (theta^4 / 24 = 3.7e-12 and theta^5 / 120 are below 2^-31 and dropped).  The bracket is at most 3.1e-3, so its fp32
// evaluation is good to ~2e-10 absolute; it is rounded to an integer by a magic-number add and added to the table entry in
// the same IADD3.  About 30 instructions per (sin, cos) pair including the phase reduction -- fewer than the 37 of the fp32
// polynomial route it replaces -- and no fp64 instruction.
constexpr int kPhiTabSize = 1024;
struct PhiTabEntry {
    int s, c;  // llrint(sin(2 pi k / 1024) 2^30), llrint(cos(2 pi k / 1024) 2^30)
};

// fh in [-1/2, 1/2] turns (exact), cross a small correction in turns (sincos_phase_df's split)
__device__ __forceinline__ void sincos_turns_q30(float fh, float cross, const PhiTabEntry* __restrict__ tab, int& qs, int& qc) {
    // table point nearest to fh + cross (at |phase| ~ 1e4 turns the correction is itself a table step wide, so it must take part
    // in the choice; the rounding of this sum only decides between neighbouring points, the remainder below is exact)
    const float t = fmaf(fh + cross, 1024.0f, kMagic);         // low mantissa bits: rint(1024 (fh + cross))
    const float kf = t - kMagic;
    const PhiTabEntry e = tab[__float_as_int(t) & (kPhiTabSize - 1)];  // kMagicBits is a multiple of 1024: two's-complement index
    const float fl = fmaf(kf, -0.0009765625f, fh) + cross;     // first part exact, |fl| <= 2^-11 (+ cross)
    const float th = fmaf(fl, 6.2831854820251465f, fl * -1.7484555e-07f);  // radians, 2 pi as fp32 hi + lo
    const float th2 = th * th;
    const float ds = fmaf(th2 * th, -0.16666667f, th);
    const float dc = -0.5f * th2;
    const float sf = (float)e.s, cf = (float)e.c;              // table values scaled by 2^30 (24 bits are plenty for the bracket)
    const float rs = fmaf(sf, dc, cf * ds) + kMagic;           // |bracket| <= 3.4e6 < 2^22: the magic-number add rounds it
    const float rc = fmaf(cf, dc, -(sf * ds)) + kMagic;
    qs = e.s + (__float_as_int(rs) - kMagicBits);
    qc = e.c + (__float_as_int(rc) - kMagicBits);
}

// (sin, cos)(2 pi (wh + wl)(th + tl)) scaled by 2^30; false if the phase is too large for the fp32 split (caller: fp64 route)
__device__ __forceinline__ bool sincos_phase_q30(float wh, float wl, float th, float tl, const PhiTabEntry* __restrict__ tab,
                                                 int& qs, int& qc) {
    const float ph = wh * th;
    if (!(fabsf(ph) < 4194304.0f)) return false;
    const float e = fmaf(wh, th, -ph);
    const float cross = fmaf(wh, tl, fmaf(wl, th, e));
    const float k = (ph + kMagic) - kMagic;
    sincos_turns_q30(ph - k, cross, tab, qs, qc);
    return true;
}

// the same for callers that have bounded |wh th| < 2^22 for all their frequencies at once (tc_pimg_kernel: per bin)
__device__ __forceinline__ void sincos_phase_q30_bounded(float wh, float wl, float th, float tl, const PhiTabEntry* __restrict__ tab,
                                                         int& qs, int& qc) {
    const float ph = wh * th;
    const float e = fmaf(wh, th, -ph);
    const float cross = fmaf(wh, tl, fmaf(wl, th, e));
    const float k = (ph + kMagic) - kMagic;
    sincos_turns_q30(ph - k, cross, tab, qs, qc);
}

// The four signed 8-bit digits of q (|q| <= 2^30 + 2^22) packed into one word, lowest digit in byte 0: adding 128 to each of
// the three low bytes turns the signed digits into unsigned ones with the carries resolved by the addition itself, and the
// XOR takes the bias off again (split_digits4, loglik_epilogue.cuh, gives the same digits one at a time).
__device__ __forceinline__ unsigned pack_digits4(int q) { return ((unsigned)q + 0x00808080u) ^ 0x00808080u; }

// fp64 phase -> (sin, cos) in fp32 with ~2e-8 rms error regardless of |phase| (< 2^40 or so).
__device__ __forceinline__ void sincos_phase(double ph, float& s, float& c) {
    const double kq = rint(ph * 0.15915494309189534561);
    double r = fma(kq, -6.283185307179586232, ph);
    r = fma(kq, -2.4492935982947064e-16, r);
    const float rh = (float)r;
    const float rl = (float)(r - (double)rh);
    sincos_reduced_f32(rh, rl, s, c);
}

#ifndef IPN_HOST_EMU
int prepare_bin_constants(ipn_ctx* ctx, int64_t T, const double* exposure, const int64_t* counts, double* d_binc);
int prepare_draw_constants(ipn_ctx* ctx, int64_t S, const double* amp, const double* bkg, const double* d_binc,
                           DrawConst* d_dc);
int finalize_loglik(ipn_ctx* ctx, int64_t S0, int64_t Sb, int64_t S, int64_t G, int64_t Gpad, const double* d_R,
                    const DrawConst* d_dc, double* d_ll, bool accumulate);

// FP32-pipe implementation (loglik_simt.cu)
int run_loglik_grid_simt(ipn_ctx* ctx, int64_t T, const double* time, const double* exposure, const int64_t* counts,
                         int64_t S, int64_t J, const double* omega, const double* a, const double* b,
                         const DrawConst* d_dc, int64_t G, const double* delays, double* ll, bool accumulate);

// tcgen05 sliced-integer implementation (loglik_tc.cu)
bool tc_supported(int64_t J);
int run_loglik_grid_tc(ipn_ctx* ctx, int64_t T, const double* time, const double* exposure, const int64_t* counts,
                       int64_t S, int64_t J, const double* omega, const double* a, const double* b, const DrawConst* d_dc,
                       int64_t G, const double* delays, double* ll, bool accumulate);
#endif  // IPN_HOST_EMU

}  // namespace ipn
