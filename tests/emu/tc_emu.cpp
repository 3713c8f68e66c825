// Host emulation of the ARITHMETIC of the tcgen05 delay-grid log-likelihood (pyipn_b200/csrc/loglik_tc.cu): test infrastructure.
//
// It compiles pyipn_b200/csrc/loglik_common.cuh and loglik_epilogue.cuh -- the very lines the kernel is built from -- with
// g++ (IPN_HOST_EMU) and restates only what lives in __global__ kernels around them (operand quantisation, level sums,
// item / warp ownership of the partial sums, fp64 fold), line by line, citing the kernel source.  Every float operation
// is IEEE fp32 with fused fmaf exactly as on the device (compile with -ffp-contract=off; nvcc may additionally contract
// a few `a*b + c` into FFMA, which only removes roundings).  The one thing that cannot be compiled here is MUFU.LG2:
// `lg2_approx` below is a stand-in (correctly rounded log2 plus an injected bias and uniform noise, argv), and the
// tests check that the result does not depend on what is injected -- which is the claim DESIGN.md section 3 makes about
// the refinement step in poisson_term.
//
// Nothing of the product path runs through this file and it never touches a GPU: tests/test_emu.py builds and runs it
// so that the error budget (1e-3 nats at 1e5 bins) is checked in the CPU suite at FULL size.
//
//   tc_emu funcs                          -> JSON with error statistics of exp2_poly / exp2_hilo / sincos_* / poisson_term
//   tc_emu grid <in.bin> <bias> <noise>   -> for every delay: ll by the fp32, compensated fp32 and fp64 epilogues
//   tc_emu images <in.bin> <out>          -> the restated operand images in the device's byte order (cf. images_emu.cpp)
//   tc_emu rff <in.bin> <out.bin>         -> rff_eval_kernel's arithmetic (rff_terms.cuh): lambda (S, T) float64 + the tier of every draw
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#define IPN_HOST_EMU 1
#define __device__
#define __forceinline__ inline
struct float4 {
    float x, y, z, w;
};
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

namespace ipn {
static double g_lg2_bias = 4e-8, g_lg2_noise = 6e-8;  // measured MUFU.LG2 on [1,2): mean +4e-8 (tools/mufu_probe.cu)
static std::mt19937_64 g_rng(12345);
static inline float lg2_approx(float x) {
    if (x == 1.0f) return 0.0f;  // MUFU.LG2(1) is exactly 0 on the hardware (tools/mufu_probe.cu); tiny u relies on it
    const double u = (double)(g_rng() >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
    return (float)(std::log2((double)x) + g_lg2_bias + g_lg2_noise * u);
}
}  // namespace ipn

#include "../../pyipn_b200/csrc/loglik_epilogue.cuh"
#include "../../pyipn_b200/csrc/rff_terms.cuh"

using namespace ipn;

// ---------------------------------------------------------------------------------------------------------------
// function-level statistics
// ---------------------------------------------------------------------------------------------------------------
struct Stat {
    double sum = 0, sq = 0, mx = 0;
    long n = 0;
    void add(double e) {
        sum += e;
        sq += e * e;
        if (std::fabs(e) > mx) mx = std::fabs(e);
        ++n;
    }
    void print(const char* name, bool last = false) const {
        std::printf("\"%s\": {\"mean\": %.6e, \"rms\": %.6e, \"max\": %.6e, \"n\": %ld}%s\n", name, sum / n, std::sqrt(sq / n), mx, n,
                    last ? "" : ",");
    }
};

static int run_funcs() {
    std::mt19937_64 rng(20261020);
    auto unif = [&](double lo, double hi) { return lo + (hi - lo) * ((double)(rng() >> 11) * (1.0 / 9007199254740992.0)); };
    Stat e_poly, e_hilo, e_sin, e_cos, e_sin_df, e_cos_df, e_sin_q30, e_cos_q30, e_log_b0, e_log_b1, e_tiny;
    std::vector<PhiTabEntry> phitab(kPhiTabSize);
    for (int k = 0; k < kPhiTabSize; ++k) {
        const double ang = 6.283185307179586476925 * (double)k / (double)kPhiTabSize;
        phitab[k].s = (int)std::llrint(std::sin(ang) * 1073741824.0);
        phitab[k].c = (int)std::llrint(std::cos(ang) * 1073741824.0);
    }
    const int N = 2000000;
    for (int i = 0; i < N; ++i) {
        const float x = (float)unif(-30.0, 30.0);
        e_poly.add((double)exp2_poly(x) / std::exp2((double)x) - 1.0);
        // exp2_hilo: xh a multiple of 2^-12 (what kappa S2 is at sigma = 1), |xl| < 0.04
        const float xh = (float)(std::nearbyint(unif(-30.0, 30.0) * 4096.0) / 4096.0);
        const float xl = (float)unif(-0.04, 0.04);
        e_hilo.add((double)exp2_hilo(xh, xl) / std::exp2((double)xh + (double)xl) - 1.0);
    }
    for (int i = 0; i < N; ++i) {
        // phases as in the Phi image: omega t with |omega t| up to 1e5 rad
        const double w = unif(-1000.0, 1000.0), t = unif(0.0, 100.0);
        float s, c;
        sincos_phase(w * t, s, c);
        e_sin.add((double)s - std::sin(w * t));
        e_cos.add((double)c - std::cos(w * t));
        const double wt = w * 0.15915494309189534561;
        const float wh = (float)wt, wl = (float)(wt - (double)wh);
        const float th = (float)t, tl = (float)(t - (double)th);
        if (sincos_phase_df(wh, wl, th, tl, s, c)) {
            e_sin_df.add((double)s - std::sin(w * t));
            e_cos_df.add((double)c - std::cos(w * t));
        }
        int qs, qc;  // the 2^-30 fixed-point route of the Phi image (table + first-order rotation)
        if (sincos_phase_q30(wh, wl, th, tl, phitab.data(), qs, qc)) {
            e_sin_q30.add((double)qs / 1073741824.0 - std::sin(w * t));
            e_cos_q30.add((double)qc / 1073741824.0 - std::cos(w * t));
        }
    }
    // log2(1 + u) reconstruction y0 + lo/nl: independent of the seed error
    for (int pass = 0; pass < 2; ++pass) {
        g_lg2_bias = pass ? -2e-7 : 4e-8;
        g_lg2_noise = pass ? 3e-7 : 6e-8;
        Stat& st = pass ? e_log_b1 : e_log_b0;
        for (int i = 0; i < N; ++i) {
            const float u = (float)std::exp2(unif(-12.0, 12.0));
            float lo;
            // n = 1, kappa = 0: main = y0, lo = log2e r
            const float y0 = poisson_term(u, 1.0f, 1.44269504f, 0.0f, 0.0f, lo);
            st.add(((double)y0 + (double)lo) - std::log2(1.0 + (double)u));
        }
    }
    for (int i = 0; i < N; ++i) {
        const float u = (float)std::exp2(unif(-60.0, -26.0));  // w == 1: the e_w term carries everything
        float lo;
        const float y0 = poisson_term(u, 1.0f, 1.44269504f, 0.0f, 0.0f, lo);
        e_tiny.add(((double)y0 + (double)lo) / ((double)u * 1.4426950408889634) - 1.0);
    }
    // integer exponent assembly + 2^x from parts (the fp32 epilogues' route), on random level sums
    Stat e_parts, e_parts_df;
    long digit_mismatch = 0, position_errors = 0;
    for (int i = 0; i < N; ++i) {
        const int sh = 17 + (int)(rng() % 8);                            // 17 .. 24: sigma = 2^-3 .. 2^4 (prior draws: 2^-1 .. 2^1)
        const float ks = std::ldexp(1.0f, 8 - sh), inv = ks * 0.00390625f, ks24 = ks * 5.9604644775390625e-8f;
        const double xr = unif(-40.0, 40.0);                             // target exponent
        long long h = std::llrint(std::ldexp(xr, sh));                   // H = 256 S2 + S3
        if (h > 230000000ll) h = 230000000ll;
        if (h < -230000000ll) h = -230000000ll;
        const int s3 = (int)(rng() % 2000001) - 1000000, s2 = (int)((h - s3) / 256);
        // the low levels as the kernel sees them: |ks 2^-24 (256 S4 + S5)| stays below a few hundredths for every draw the fp32
        // epilogues take (bounded exponent), and |S4| < 7.4e6, |S5| < 1.2e7 by the digit ranges
        double lt = unif(-0.04, 0.04) / (double)ks24;
        if (lt > 1.8e9) lt = 1.8e9;
        if (lt < -1.8e9) lt = -1.8e9;
        const int s5 = (int)(rng() % 2000001) - 1000000, s4 = (int)((lt - s5) / 256.0);
        const double x = (double)ks * ((double)s2 + (double)s3 / 256.0 + ((double)s4 * 256.0 + (double)s5 + 0.5) / 16777216.0);
        if (std::fabs(x) > 100.0) continue;
        int k;
        const float xf = exponent_parts(s2, s3, s4, s5, sh, inv, ks24, k);
        e_parts.add((double)exp2_parts(xf, k) / std::exp2(x) - 1.0);
        float lo;
        const float hi = exp2_parts_df(xf, k, lo);
        e_parts_df.add(((double)hi + (double)lo) / std::exp2(x) - 1.0);
        // packed digit split against the one-at-a-time split
        const int q = (int)(rng() % 2147483647u) - 1073741824;
        int8_t d1, d2, d3, d4;
        split_digits4(q, d1, d2, d3, d4);
        const unsigned w = pack_digits4(q);
        digit_mismatch += (int8_t)(w >> 24) != d1 || (int8_t)(w >> 16) != d2 || (int8_t)(w >> 8) != d3 || (int8_t)w != d4;
    }
    // the interleaved bin order is a permutation of [0, T) for any split of T bins into the two classes
    for (int it = 0; it < 400; ++it) {
        const long long T = 1 + (long long)(rng() % 5000), n_nz = (long long)(rng() % (T + 1));
        std::vector<char> hit((size_t)T, 0);
        for (long long r = 0; r < n_nz; ++r) {
            const long long p = tc_bin_position(true, r, n_nz, T, 32);
            if (p < 0 || p >= T || hit[(size_t)p]++) ++position_errors;
        }
        for (long long r = 0; r < T - n_nz; ++r) {
            const long long p = tc_bin_position(false, r, n_nz, T, 32);
            if (p < 0 || p >= T || hit[(size_t)p]++) ++position_errors;
        }
    }
    std::printf("{\n");
    std::printf("\"digit_pack_mismatches\": %ld,\n\"bin_position_errors\": %ld,\n", digit_mismatch, position_errors);
    e_parts.print("exp2_parts_rel");
    e_parts_df.print("exp2_parts_df_rel");
    e_poly.print("exp2_poly_rel");
    e_hilo.print("exp2_hilo_rel");
    e_sin.print("sin_phase_abs");
    e_cos.print("cos_phase_abs");
    e_sin_df.print("sin_phase_df_abs");
    e_cos_df.print("cos_phase_df_abs");
    e_sin_q30.print("sin_phase_q30_abs");
    e_cos_q30.print("cos_phase_q30_abs");
    e_log_b0.print("log2_1pu_abs_mufu_like");
    e_log_b1.print("log2_1pu_abs_bad_seed");
    e_tiny.print("log2_1pu_tiny_rel", true);
    std::printf("}\n");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// grid emulation
// ---------------------------------------------------------------------------------------------------------------
namespace cfg {  // loglik_tc.cu namespace tc
constexpr int TN = 32, GCOLS = 16, GROUPS = TN / GCOLS, EPI_SLOTS = 4, MIN_CROWS = 8, MAX_TILES = 4096 / TN;
}

template <class T>
static bool rd(FILE* f, T* p, size_t n) {
    return std::fread(p, sizeof(T), n, f) == n;
}

static int run_grid(const char* path, double bias, double noise, const char* dump_path = nullptr) {
    g_lg2_bias = bias;
    g_lg2_noise = noise;
    FILE* f = std::fopen(path, "rb");
    if (!f) return 2;
    int64_t hdr[3];
    double ab[2];
    if (!rd(f, hdr, 3) || !rd(f, ab, 2)) return 2;
    const int64_t T = hdr[0], J = hdr[1], G = hdr[2];
    const double A = ab[0], B = ab[1];
    std::vector<double> time(T), expo(T), omega(J), a(J), b(J), delays(G);
    std::vector<int64_t> counts(T);
    if (!rd(f, time.data(), T) || !rd(f, expo.data(), T) || !rd(f, counts.data(), T) || !rd(f, omega.data(), J) ||
        !rd(f, a.data(), J) || !rd(f, b.data(), J) || !rd(f, delays.data(), G))
        return 2;
    std::fclose(f);

    // bin_const_kernel / draw_const_kernel (loglik_prep.cu)
    double binc[4] = {0, 0, 0, 0}, sumsq = 0;
    for (int64_t i = 0; i < T; ++i) {
        const double n = (double)counts[i];
        binc[0] += n;
        if (n > 0.0) binc[1] += n * std::log(expo[i]);
        binc[2] += expo[i];
        binc[3] += std::lgamma(n + 1.0);
        sumsq += n * n;
    }
    DrawConst dc;
    dc.C = binc[1] + (binc[0] > 0.0 ? binc[0] * std::log(B) : 0.0) - B * binc[2] - binc[3];
    dc.kap = B / kLn2;
    const double c = std::log2(A / B);
    dc.c_hi = (float)c;
    dc.c_lo = std::isfinite(c) ? (float)(c - (double)dc.c_hi) : 0.0f;

    // run_loglik_grid_tc: padded K, tiles, chunks
    const int Kpad = (int)(((2 * J + cfg::MIN_CROWS + 31) / 32) * 32);
    const int ncrow = Kpad - 2 * (int)J;
    const int n_bt = (int)((T + cfg::TN - 1) / cfg::TN);
    int nchunk = (n_bt + cfg::MAX_TILES - 1) / cfg::MAX_TILES;
    if (nchunk < 1) nchunk = 1;
    const int tiles_per_chunk = (n_bt + nchunk - 1) / nchunk;

    // tc_draw_kernel: sigma, kappa, automatic epilogue mode
    double mx = 0.0, sumw = 0.0;
    for (int j = 0; j < J; ++j) {
        const double wj = kLog2e * std::sqrt(a[j] * a[j] + b[j] * b[j]);
        mx = std::fmax(mx, wj);
        sumw += wj;
    }
    const double cabs = std::fabs((double)dc.c_hi + (double)dc.c_lo);
    int e = 10;
    const double lim_c = (double)ncrow * (1.0 - 1e-6);
    while (e > -6 && (std::ldexp(mx, e) > 1.0 || (std::isfinite(cabs) && std::ldexp(cabs, e) > lim_c))) --e;
    const float sigma_f = (float)std::ldexp(1.0, e), ks = (float)std::ldexp(1.0, -12 - e);
    const double ysc = std::isfinite(cabs) ? std::fmax(1.0, (double)dc.c_hi) : 1.0;
    const double crit = ysc * ysc * sumsq;
    // a draw whose exponent could leave [-120, 120] is evaluated by the fp64 epilogue whatever the option says (the fp32
    // epilogues evaluate 2^x without clamps)
    const double xmax = (std::isfinite(cabs) ? cabs : 0.0) + sumw * (1.0 + 1e-6) + 1e-3;
    const int auto_mode = xmax > 120.0 ? 2 : (crit > 4.0e7 ? 2 : (crit > 2.0e5 ? 1 : 0));  // n_terms = 1

    // tc_perm_kernel: bins with counts and bins without, each class in its original order, tiles of the two classes interleaved
    int64_t n_nz = 0;
    for (int64_t i = 0; i < T; ++i) n_nz += counts[i] > 0;
    std::vector<int> perm((size_t)T, -1);
    {
        int64_t rc = 0, rz = 0;
        for (int64_t i = 0; i < T; ++i) {
            const bool f = counts[i] > 0;
            perm[(size_t)tc_bin_position(f, f ? rc++ : rz++, n_nz, T, cfg::TN)] = (int)i;
        }
    }

    // tc_phitab_kernel: the 1024-point unit circle in 2^-30 fixed point
    std::vector<PhiTabEntry> phitab(kPhiTabSize);
    for (int k = 0; k < kPhiTabSize; ++k) {
        const double ang = 6.283185307179586476925 * (double)k / (double)kPhiTabSize;
        phitab[k].s = (int)std::llrint(std::sin(ang) * 1073741824.0);
        phitab[k].c = (int)std::llrint(std::cos(ang) * 1073741824.0);
    }
    // tc_pimg_kernel: Phi digit planes [4][Tpad][Kpad] (2^-30 fixed point) and per-bin constants
    const int64_t Tpad = (int64_t)n_bt * cfg::TN;
    std::vector<int8_t> P((size_t)4 * Tpad * Kpad, 0);
    std::vector<float4> cst(Tpad);
    std::vector<float> wh(J), wl(J);
    for (int j = 0; j < J; ++j) {
        const double wt = omega[j] * 0.15915494309189534561;
        wh[j] = (float)wt;
        wl[j] = (float)(wt - (double)wh[j]);
    }
    const bool src = std::isfinite(dc.c_hi);
    const bool exact_phi = std::getenv("IPN_EMU_EXACT_PHI") != nullptr;
    const bool no_w5 = std::getenv("IPN_EMU_NO_W5") != nullptr;  // experiment: drop the fifth W digit (W to 2^-31 only)
    const bool no_p2w4 = std::getenv("IPN_EMU_NO_P2W4") != nullptr;
    const bool all_l6 = std::getenv("IPN_EMU_ALL_L6") != nullptr;  // experiment: the two dropped level-6 products as well
    for (int64_t pos = 0; pos < Tpad; ++pos) {
        const bool valid = pos < T;
        const int64_t i = valid ? perm[pos] : 0;
        const double t = valid ? time[i] : 0.0;
        const float th = (float)t, tl = (float)(t - (double)(float)t);
        for (int k = 0; k < Kpad; k += 2) {
            int q0 = 0, q1 = 0;
            if (valid) {
                if (k < 2 * J) {
                    if (exact_phi) {  // experiment (IPN_EMU_EXACT_PHI): float64 sin / cos, isolates the fixed-point route's share
                        q1 = (int)std::llrint(std::sin(omega[k >> 1] * t) * 1073741824.0);
                        q0 = (int)std::llrint(std::cos(omega[k >> 1] * t) * 1073741824.0);
                    } else if (!sincos_phase_q30(wh[k >> 1], wl[k >> 1], th, tl, phitab.data(), q1, q0)) {
                        q1 = (int)std::llrint(std::sin(omega[k >> 1] * t) * 1073741824.0);
                        q0 = (int)std::llrint(std::cos(omega[k >> 1] * t) * 1073741824.0);
                    }
                } else {
                    q0 = q1 = 1073741824;
                }
            }
            int8_t d0[4], d1[4];
            split_digits4(q0, d0[0], d0[1], d0[2], d0[3]);
            split_digits4(q1, d1[0], d1[1], d1[2], d1[3]);
            for (int pl = 0; pl < 4; ++pl) {
                P[((size_t)pl * Tpad + pos) * Kpad + k] = d0[pl];
                P[((size_t)pl * Tpad + pos) * Kpad + k + 1] = d1[pl];
            }
        }
        const double kap = (valid && src) ? expo[i] * dc.kap : 0.0;
        const float khi = (float)kap;
        float4 cc;
        cc.x = (valid && src) ? (float)counts[i] : 0.0f;
        cc.y = khi;
        cc.z = (float)(kap - (double)khi);
        cc.w = cc.x * 1.44269504f;
        cst[pos] = cc;
    }
    const float ks24 = ks * 5.9604644775390625e-8f;
    const int sh = 8 - (((__float_as_int(ks) >> 23) & 0xff) - 127);
    const float inv = ks * 0.00390625f;
    const double ksd = (double)ks, sigma = (double)sigma_f;
    const double cs = (double)dc.c_hi + (double)dc.c_lo;

    // tc_wimg_kernel: W digit planes of one delay (rows beyond G are zero)
    auto w_digits = [&](int64_t g, std::vector<int8_t>& W) {
        const double dl = g < G ? delays[g] : 0.0;
        for (int k = 0; k < Kpad; k += 2) {
            long long q0 = 0, q1 = 0;
            if (g < G) {
                if (k < 2 * J) {
                    const int j = k >> 1;
                    const double sn = std::sin(omega[j] * dl), cn = std::cos(omega[j] * dl);
                    q0 = std::llrint(std::fmin(std::fmax(kLog2e * (a[j] * cn - b[j] * sn) * sigma, -1.0), 1.0) * 274877906944.0);
                    q1 = std::llrint(std::fmin(std::fmax(kLog2e * (a[j] * sn + b[j] * cn) * sigma, -1.0), 1.0) * 274877906944.0);
                } else if (std::isfinite(cs)) {
                    const long long tot = std::llrint(cs * sigma * 274877906944.0);
                    const long long base = tot / ncrow, rem = tot - base * ncrow;
                    const int i0 = k - 2 * (int)J, i1 = i0 + 1;
                    const long long ar = rem < 0 ? -rem : rem, sg = rem < 0 ? -1 : 1;
                    q0 = base + (i0 < ar ? sg : 0);
                    q1 = base + (i1 < ar ? sg : 0);
                }
            }
            split_digits5(q0, W[0 * Kpad + k], W[1 * Kpad + k], W[2 * Kpad + k], W[3 * Kpad + k], W[4 * Kpad + k]);
            split_digits5(q1, W[0 * Kpad + k + 1], W[1 * Kpad + k + 1], W[2 * Kpad + k + 1], W[3 * Kpad + k + 1], W[4 * Kpad + k + 1]);
        }
    };
    std::vector<int8_t> W((size_t)5 * Kpad);

    if (dump_path) {
        // the images in the device's byte order ([plane][k / 16][row][k % 16], constants and flag block behind the Phi planes),
        // to be compared with what the kernels themselves produce (tests/emu/images_emu.cpp)
        const int TM = 128, n_dt = (int)((G + TM - 1) / TM);
        const size_t a_plane = (size_t)Kpad * TM, b_plane = (size_t)Kpad * cfg::TN;
        const size_t a_bytes = 5 * a_plane, b_bytes = 4 * b_plane + cfg::TN * 16 + 16;
        std::vector<uint8_t> Wi((size_t)n_dt * a_bytes, 0), Pi((size_t)n_bt * b_bytes, 0);
        for (int64_t g = 0; g < (int64_t)n_dt * TM; ++g) {
            w_digits(g, W);
            for (int pl = 0; pl < 5; ++pl)
                for (int k = 0; k < Kpad; ++k)
                    Wi[(size_t)(g / TM) * a_bytes + pl * a_plane + ((size_t)(k / 16) * TM + g % TM) * 16 + k % 16] = (uint8_t)W[pl * Kpad + k];
        }
        for (int64_t pos = 0; pos < Tpad; ++pos) {
            uint8_t* img = &Pi[(size_t)(pos / cfg::TN) * b_bytes];
            const int row = (int)(pos % cfg::TN);
            for (int pl = 0; pl < 4; ++pl)
                for (int k = 0; k < Kpad; ++k)
                    img[pl * b_plane + ((size_t)(k / 16) * cfg::TN + row) * 16 + k % 16] = (uint8_t)P[((size_t)pl * Tpad + pos) * Kpad + k];
            std::memcpy(img + 4 * b_plane + (size_t)row * 16, &cst[pos], 16);
            if (row % 8 == 0) {  // one flag per 8 rows: the unit's first bin holds counts
                const float fl = (pos < T && counts[perm[pos]] > 0) ? 1.0f : 0.0f;
                std::memcpy(img + 4 * b_plane + (size_t)cfg::TN * 16 + (row / 8) * 4, &fl, 4);
            }
        }
        FILE* g = std::fopen(dump_path, "wb");
        if (!g) return 2;
        const int ih[4] = {Kpad, n_dt, n_bt, auto_mode};
        const float fh[2] = {ks, sigma_f};
        std::fwrite(ih, 4, 4, g);
        std::fwrite(fh, 4, 2, g);
        std::fwrite(Wi.data(), 1, Wi.size(), g);
        std::fwrite(Pi.data(), 1, Pi.size(), g);
        std::fclose(g);
        return 0;
    }

    std::printf("{\"auto_mode\": %d, \"xmax\": %.6g, \"sigma\": %g, \"Kpad\": %d, \"n_nz\": %lld, \"nchunk\": %d, \"sumsq\": %.17g,\n\"ll\": [\n",
                auto_mode, xmax, sigma, Kpad, (long long)n_nz, nchunk, sumsq);
    std::vector<int> S2(cfg::TN), S3(cfg::TN), S4(cfg::TN), S5(cfg::TN);
    for (int64_t g = 0; g < G; ++g) {
        w_digits(g, W);
        double R[3] = {0.0, 0.0, 0.0};  // per epilogue mode (fp32, compensated, fp64), the float64 atomicAdd target
        for (int chunk = 0; chunk < nchunk; ++chunk) {
            const int bt0 = chunk * tiles_per_chunk, bt1 = std::min(n_bt, bt0 + tiles_per_chunk);
            // per item: one (ssum, scomp, slo) per epilogue warp of the lane quarter (slot), for each fp32 mode
            float ssum[2][cfg::EPI_SLOTS] = {}, scomp[2][cfg::EPI_SLOTS] = {}, slo[2][cfg::EPI_SLOTS] = {};
            float qsum[cfg::EPI_SLOTS] = {}, qcomp[cfg::EPI_SLOTS] = {};  // compensated epilogue: the kappa u sum
            double hacc[cfg::EPI_SLOTS] = {};
            unsigned tile_seq = 0;
            for (int bt = bt0; bt < bt1; ++bt, ++tile_seq) {
                // the eleven digit-plane products of the tile (int32, exact), loglik_tc.cu header
                for (int r = 0; r < cfg::TN; ++r) {
                    const size_t pos = (size_t)bt * cfg::TN + r;
                    const int8_t* p1 = &P[((size_t)0 * Tpad + pos) * Kpad];
                    const int8_t* p2 = &P[((size_t)1 * Tpad + pos) * Kpad];
                    const int8_t* p3 = &P[((size_t)2 * Tpad + pos) * Kpad];
                    const int8_t* p4 = &P[((size_t)3 * Tpad + pos) * Kpad];
                    const int8_t *w1 = &W[0], *w2 = &W[Kpad], *w3 = &W[2 * Kpad], *w4 = &W[3 * Kpad], *w5 = &W[4 * Kpad];
                    int s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0;
                    for (int k = 0; k < Kpad; ++k) {
                        s2 += w1[k] * p1[k];
                        s3 += w2[k] * p1[k] + w1[k] * p2[k];
                        s4 += w3[k] * p1[k] + w2[k] * p2[k] + w1[k] * p3[k];
                        s5 += w4[k] * p1[k] + w3[k] * p2[k] + w2[k] * p3[k] + w1[k] * p4[k];
                        s6 += w5[k] * p1[k] + (no_p2w4 ? 0 : w4[k] * p2[k]) + (all_l6 ? w3[k] * p3[k] + w2[k] * p4[k] : 0);
                    }
                    S2[r] = s2;
                    S3[r] = s3;
                    S4[r] = s4;
                    S5[r] = no_w5 ? s5 : fold_level6(s5, s6);
                }
                for (int cg = 0; cg < cfg::GROUPS; ++cg) {
                    const int slot = (int)((tile_seq * cfg::GROUPS + cg) % cfg::EPI_SLOTS);
                    for (int h = 0; h < cfg::GCOLS / 8; ++h) {
                        const int o = cg * cfg::GCOLS + 8 * h;
                        int t2[8], t3[8], t4[8], t5[8];
                        for (int j = 0; j < 8; ++j) {
                            t2[j] = S2[o + j];
                            t3[j] = S3[o + j];
                            t4[j] = S4[o + j];
                            t5[j] = S5[o + j];
                        }
                        const float4* cp = &cst[(size_t)bt * cfg::TN + o];
                        // the unit's first bin holds counts (tc_pimg_kernel's flag, one per 8 rows)
                        const size_t p0 = (size_t)bt * cfg::TN + o;
                        const bool has_counts = (int64_t)p0 < T && counts[perm[p0]] > 0;
                        if (has_counts) {
                            epi_units8<true>(t2, t3, t4, t5, cp, sh, inv, ks24, ssum[0][slot], scomp[0][slot], slo[0][slot]);
                            epi_units8_comp<true>(t2, t3, t4, t5, cp, sh, inv, ks24, ssum[1][slot], scomp[1][slot], qsum[slot], qcomp[slot], slo[1][slot]);
                        } else {
                            epi_units8<false>(t2, t3, t4, t5, cp, sh, inv, ks24, ssum[0][slot], scomp[0][slot], slo[0][slot]);
                            // n_terms = 1 (TcDraw::pad bit 0): tiles without counts take the plain unit in the compensated epilogue too
                            epi_units8<false>(t2, t3, t4, t5, cp, sh, inv, ks24, ssum[1][slot], scomp[1][slot], slo[1][slot]);
                        }
                        epi_units8_hifi(t2, t3, t4, t5, cp, ksd, hacc[slot]);
                    }
                }
            }
            for (int s = 0; s < cfg::EPI_SLOTS; ++s) {
                for (int m = 0; m < 2; ++m) R[m] += ((double)ssum[m][s] - (double)scomp[m][s]) + (double)slo[m][s];
                R[1] -= (double)qsum[s] - (double)qcomp[s];
                R[2] += hacc[s];
            }
        }
        // finalize_kernel
        std::printf("[");
        for (int m = 0; m < 3; ++m) {
            const double v = dc.C + kLn2 * R[m];
            if (std::isfinite(v))
                std::printf("%.17g%s", v, m < 2 ? ", " : "");
            else  // JSON spelling python's json module accepts
                std::printf("%s%s", std::isnan(v) ? "NaN" : (v < 0 ? "-Infinity" : "Infinity"), m < 2 ? ", " : "");
        }
        std::printf("]%s\n", g + 1 < G ? "," : "");
    }
    std::printf("]}\n");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// rff_eval_kernel (rff_eval.cu): tier choice, per-bin split of tau, J-term loop, final exp
// ---------------------------------------------------------------------------------------------------------------
static int run_rff(const char* path, const char* out_path) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return 2;
    int64_t hdr[3];
    if (!rd(f, hdr, 3)) return 2;
    const int64_t T = hdr[0], S = hdr[1], J = hdr[2];
    std::vector<double> time(T), omega(S * J), a(S * J), b(S * J), shift(S), amp(S), out(S * T);
    if (!rd(f, time.data(), T) || !rd(f, omega.data(), S * J) || !rd(f, a.data(), S * J) || !rd(f, b.data(), S * J) ||
        !rd(f, shift.data(), S) || !rd(f, amp.data(), S))
        return 2;
    std::fclose(f);
    std::vector<double> tiers(S);
    std::vector<float4> sf(J);
    for (int64_t s = 0; s < S; ++s) {
        const double *w = &omega[s * J], *aj = &a[s * J], *bj = &b[s * J];
        float wmax = 0.0f;
        double q = 0.0;
        for (int j = 0; j < J; ++j) {
            const double wt = w[j] * 0.15915494309189534561;
            const float wh = (float)wt;
            sf[j] = float4{wh, (float)(wt - (double)wh), (float)aj[j], (float)bj[j]};
            wmax = std::fmax(wmax, std::fabs(wh));
            q += aj[j] * aj[j] + bj[j] * bj[j];
        }
        const float amp2 = (float)q;
        int tier_max = 0;
        for (int64_t i = 0; i < T; ++i) {
            const double tau = time[i] - shift[s];
            const float th = (float)tau, tl = (float)(tau - (double)th);
            // the kernel decides `big` per thread over its four bins; per bin here (the fp64 path is the accurate one)
            const bool big = !(std::fabs(th) * wmax < 4194304.0f);
            float sum = 0.0f, comp = 0.0f, lo = 0.0f;
            double side = 0.0;
            int tier;
            if (big || !(amp2 <= 100.0f)) {
                tier = 2;
                for (int j = 0; j < J; ++j) side = std::fma(aj[j], std::cos(w[j] * tau), std::fma(bj[j], std::sin(w[j] * tau), side));
            } else if (amp2 > 12.0f) {
                tier = 1;
                for (int j = 0; j < J; ++j) rff_term_df(sf[j], th, tl, sum, comp, lo);
                side = (double)lo;
            } else {
                tier = 0;
                for (int j = 0; j < J; ++j) rff_term(sf[j], th, tl, sum, comp);
            }
            if (tier > tier_max) tier_max = tier;
            out[s * T + i] = amp[s] * std::exp(((double)sum - (double)comp) + side);
        }
        tiers[s] = tier_max;
    }
    FILE* g = std::fopen(out_path, "wb");
    if (!g) return 2;
    std::fwrite(out.data(), sizeof(double), out.size(), g);
    std::fwrite(tiers.data(), sizeof(double), tiers.size(), g);
    std::fclose(g);
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && !std::strcmp(argv[1], "funcs")) return run_funcs();
    if (argc >= 5 && !std::strcmp(argv[1], "grid")) return run_grid(argv[2], std::atof(argv[3]), std::atof(argv[4]));
    if (argc >= 4 && !std::strcmp(argv[1], "images")) return run_grid(argv[2], 0.0, 0.0, argv[3]);
    if (argc >= 4 && !std::strcmp(argv[1], "rff")) return run_rff(argv[2], argv[3]);
    std::fprintf(stderr, "usage: tc_emu funcs | tc_emu grid <in.bin> <lg2 bias> <lg2 noise> | tc_emu images <in.bin> <out.bin> | tc_emu rff <in.bin> <out.bin>\n");
    return 64;
}
