// Operand images of the tcgen05 likelihood kernel (loglik_tc.cu): tile geometry, the per-draw scale / epilogue choice and the
// kernels that build the W and Phi digit planes and the per-bin Poisson constants.  Plain CUDA C++ without inline PTX, in a
// header of their own so that tests/emu/images_emu.cpp can compile these very kernels for the host and hold the restatement
// the full-size emulation (tests/emu/tc_emu.cpp) uses against them byte for byte.
#pragma once
#include "loglik_epilogue.cuh"  // split_digits4, DrawConst, sincos_phase_df
#include "tc_perm.cuh"

namespace ipn {

namespace tc {
constexpr int TM = 128;       // delays per tile (MMA M, TMEM lanes)
#ifndef IPN_TC_TN
#define IPN_TC_TN 32
#endif
constexpr int TN = IPN_TC_TN; // bins per tile (MMA N, TMEM columns per level)
constexpr int NLEV = 4;       // level accumulators S2..S5
constexpr int PW = 5;         // W digit planes (W sigma to 2^-39)
constexpr int PWT = 4;        // ... of which the first four live in tensor memory; the fifth stays in shared memory (one product)
#ifndef IPN_TC_PP
#define IPN_TC_PP 4
#endif
constexpr int PP = IPN_TC_PP; // Phi digit planes: 4 = Phi to 2^-31 (default), 3 = to 2^-23 (-DIPN_TC_PP=3: the round-1 layout, kept for A/B timing)
static_assert(PP == 3 || PP == 4, "Phi has three or four digit planes");
#ifndef IPN_TC_EPI_WARPS
#define IPN_TC_EPI_WARPS 16
#endif
#ifndef IPN_TC_GCOLS
#define IPN_TC_GCOLS 16
#endif
constexpr int EPI_WARPS = IPN_TC_EPI_WARPS;  // 4 per TMEM lane quarter (and per SM sub-partition) rotating over column groups:
                                             // 96 registers per thread, measured 3 % faster than 12 warps at 126
constexpr int EPI_SLOTS = EPI_WARPS / 4;
constexpr int GCOLS = IPN_TC_GCOLS;          // columns per epilogue work unit (batches of 8)
constexpr int NH = GCOLS / 8;
constexpr int GROUPS = TN / GCOLS; // work units per tile and lane quarter
constexpr int EPI_ARRIVALS = 4 * GROUPS;  // one arrival per (lane quarter, group) and tile
constexpr int NTHREADS = (3 + EPI_WARPS) * 32;
// The warp scheduler favours the highest warp id among eligible warps: the two light but latency-critical roles
// get the top ids so the epilogue warps sharing their sub-partitions cannot starve them.
constexpr int PRODUCER_WARP = EPI_WARPS;      // 16
constexpr int MMA_WARP = EPI_WARPS + 1;       // 17: issues the even tiles (accumulator buffer 0) and the TMEM copies of W
constexpr int MMA_WARP2 = EPI_WARPS + 2;      // 18: issues the odd tiles (accumulator buffer 1) from another sub-partition
constexpr int KPAD_MAX = 224;
constexpr int MIN_CROWS = 8;  // K rows reserved for the additive constant
constexpr int CONST_BYTES = TN * 16 + 16;  // per-bin (n, kappa_hi, kappa_lo, n log2e) + one block of flags, one per 8 bins
static_assert(TN % 16 == 0 && TN / 8 <= 4, "the flag block holds four floats: one per 8 bins of a tile; tc_bin_position mixes halves of a tile");
// A parity wait on an mbarrier is only sound if the waiter sees EVERY phase of that barrier (it cannot tell "two phases
// away" from "done").  A warp of epilogue slot s owns the groups with (GROUPS tile + g) % EPI_SLOTS == s, so it skips tiles
// (every other one with 4 slots and 2 groups, every third with 3 slots); with two MMA issuers tiles may also complete out
// of order.  Hence every barrier an epilogue warp waits on is indexed by tile % (a multiple of the ownership period): all
// uses of one barrier then belong to the same warps, and the others never look at it.  6 serves both 3 and 4 slots.
// (five slots -- 20 epilogue warps, an untried experiment -- need a multiple of 5: 10)
constexpr int cgcd(int a, int b) { return b == 0 ? a : cgcd(b, a % b); }
constexpr int OWN_PERIOD = EPI_SLOTS / cgcd(GROUPS, EPI_SLOTS);  // tiles after which the group -> slot assignment repeats
constexpr int NCST = (6 % OWN_PERIOD == 0) ? 6 : 2 * OWN_PERIOD;  // depth of the shared-memory ring of those constant blocks
constexpr int NTF = NCST;                  // "accumulators complete" barriers: tile % NTF
static_assert(NCST % OWN_PERIOD == 0 && NTF % OWN_PERIOD == 0 && NTF % 2 == 0, "epilogue barriers must not be shared between owner sets");
constexpr int NBST = 128 / TN;             // depth of the Phi digit-plane ring (4 tiles of 32 bins x 4 planes: 112 KB)
constexpr int NBUF = 256 / (NLEV * TN);    // accumulator buffers in TMEM columns [256, 512)
constexpr int NBFULL = (NBST % 2 == 0) ? NBST : 2 * NBST;  // "stage filled" barriers: tile % NBFULL.  The two MMA issuers take
                                           // alternate tiles, and a parity wait is only sound if the waiter sees every phase of
                                           // its barrier: with an odd ring depth a stage alternates between the issuers, so the
                                           // barriers (not the stages) are doubled and each one belongs to one issuer
static_assert(NBUF >= 2 && NBUF % 2 == 0 && NTF >= NBUF, "each accumulator buffer belongs to one issuer");
constexpr int ACC_COL0 = 256;              // TMEM: W digits 1-4 in columns [0, 224), level-6 accumulator in [224, 256), levels 2-5 x 2 buffers in [256, 512)
constexpr int L6_COL0 = 224;               // the single-buffered level-6 accumulator (TN columns)
static_assert(L6_COL0 + TN <= ACC_COL0 && PWT * (KPAD_MAX / 4) <= L6_COL0, "tensor-memory map");
}  // namespace tc

// struct TcDraw: loglik_common.cuh

// ------------------------------------------------------------------------------------------------------
// operand images (global memory, exactly the shared-memory byte order the UMMA descriptors expect:
// K-major, no swizzle: [plane][k/16][row][k%16], i.e. 8x16-byte core matrices contiguous (SBO = 128 B),
// next 16-byte K chunk after all rows (LBO = rows * 16 B))
// ------------------------------------------------------------------------------------------------------
// split_digits4: loglik_epilogue.cuh

// sigma / kappa of every draw: sigma = largest power of two with sigma*max|W| <= 1 and sigma*|c| <= n_crows
__global__ void tc_draw_kernel(int64_t S0, int Sb, int J, int n_crows, const double* __restrict__ a,
                               const double* __restrict__ b, const DrawConst* __restrict__ dc, const int* __restrict__ flags,
                               int n_terms, TcDraw* __restrict__ td) {
    const int sl = blockIdx.x * blockDim.x + threadIdx.x;
    if (sl >= Sb) return;
    const int64_t s = S0 + sl;
    double mx = 0.0, sumw = 0.0;
    for (int j = 0; j < J; ++j) {
        const double aj = a[s * J + j], bj = b[s * J + j];
        const double wj = kLog2e * sqrt(aj * aj + bj * bj);
        mx = fmax(mx, wj);
        sumw += wj;
    }
    const double c = fabs((double)dc[s].c_hi + (double)dc[s].c_lo);  // inf for amp == 0 (handled through the consts)
    int e = 10;  // sigma = 2^e, start high and walk down (e in [-6, 10]: the epilogue's integer exponent assembly shifts by 20 + e)
    const double lim_c = (double)n_crows * (1.0 - 1e-6);
    while (e > -6 && (ldexp(mx, e) > 1.0 || (isfinite(c) && ldexp(c, e) > lim_c))) --e;
    TcDraw t;
    t.sigma = (float)ldexp(1.0, e);
    t.ks = (float)ldexp(1.0, -12 - e);
    // fp32 epilogue error model: every bin's term n y - kappa u is rounded at its own size (and nothing
    // in fp32 can undo that), so rms |dll| ~ 1.5e-7 y sqrt(sum n_i^2) with y the typical log2(1 + u) of the bins that hold the
    // counts (measured on the bench data, sum n^2 = 4.5e5: 1.1e-4 rms, but 7e-4 ... 1.0e-3 in the tail of 200 checked grid
    // points; compensated fp32: 4e-5 rms, 1.4e-4 max).  With the other sources at ~4e-5 and a tail of ~8 sigma over a grid of
    // 1e7 points, the plain epilogue (1.0e-7 sqrt(.) with its Kahan step per pair) is only taken where it predicts < 5e-5 nats: n_terms y^2 sum n^2 < 2e5, i.e. faint
    // data; the compensated fp32 epilogue up to 4e7, the fp64 one beyond (bright, coarsely binned data).  n_terms: the
    // detector passes a sky-grid call adds up.  flags[1] = IPN_OPT_HIFI: 0 auto, 1 always fp64, 2 always plain fp32, 3
    // always compensated fp32.
    const double sumsq = *reinterpret_cast<const double*>(flags + 2);
    const double ysc = isfinite(c) ? fmax(1.0, (double)dc[s].c_hi) : 1.0;
    // t.hifi = epilogue mode of the draw: 0 fp32, 1 compensated fp32, 2 fp64 (EPI_* below)
    const double crit = ysc * ysc * sumsq * (double)(n_terms > 1 ? n_terms : 1);
    t.hifi = flags[1] == 1 ? 2 : (flags[1] == 2 ? 0 : (flags[1] == 3 ? 1 : (crit > 4.0e7 ? 2 : (crit > 2.0e5 ? 1 : 0))));
    const bool bounded = (isfinite(c) ? c : 0.0) + sumw * (1.0 + 1e-6) + 1e-3 <= 120.0;
    // the fp32 epilogues evaluate 2^x without clamps (exp2_hilo): a draw whose exponent could leave [-120, 120] -- amplitudes
    // far outside anything the model's priors produce -- is evaluated by the fp64 epilogue whatever the option says
    if (!bounded) t.hifi = 2;
    // bit 0: in the compensated epilogue, tiles WITHOUT counts take the plain fp32 unit (28 instead of 37 instructions per
    // bin): their terms are - kappa u, far below the size at which fp32 rounding matters unless the model puts many counts
    // where the data have none -- which is what most pixels of a sky grid do for the far detectors, so not there
    t.pad = n_terms <= 1 ? 1 : 0;
    td[sl] = t;
}

// W digit planes of one (draw, tile of 128 delays).  One thread = one (delay row, 16-wide K chunk).
__global__ void __launch_bounds__(128) tc_wimg_kernel(int64_t S0, int J, int Kpad, const double* __restrict__ omega,
                                                      const double* __restrict__ a, const double* __restrict__ b,
                                                      const DrawConst* __restrict__ dc, const TcDraw* __restrict__ td,
                                                      int64_t G, const double* __restrict__ delays, int64_t delay_stride,
                                                      int n_dt, uint32_t a_bytes, uint32_t a_plane, uint8_t* __restrict__ Wimg) {
    const int row = threadIdx.x;  // 0..127
    const int kc = blockIdx.x;    // K chunk of 16
    const int dt = blockIdx.y;
    const int sl = blockIdx.z;
    const int64_t s = S0 + sl;
    const int64_t g = (int64_t)dt * tc::TM + row;
    const double sigma = (double)td[sl].sigma;
    const double cs = (double)dc[s].c_hi + (double)dc[s].c_lo;
    const int ncrow = Kpad - 2 * J;
    alignas(16) int8_t d[tc::PW][16];
    const double dl = (g < G) ? delays[s * delay_stride + g] : 0.0;  // delay_stride = 0: one grid for all draws, G: one per draw
    for (int kk = 0; kk < 16; kk += 2) {
        const int k = kc * 16 + kk;
        long long q0 = 0, q1 = 0;
        if (g < G) {
            if (k < 2 * J) {
                const int j = k >> 1;
                const double w = omega[s * J + j], aj = a[s * J + j], bj = b[s * J + j];
                double sn, cn;
                sincos(w * dl, &sn, &cn);
                // saturate instead of wrapping if a caller of the _dev entry point passes amplitudes beyond the supported range
                q0 = llrint(fmin(fmax(kLog2e * (aj * cn - bj * sn) * sigma, -1.0), 1.0) * 274877906944.0);  // 2^38
                q1 = llrint(fmin(fmax(kLog2e * (aj * sn + bj * cn) * sigma, -1.0), 1.0) * 274877906944.0);
            } else if (isfinite(cs)) {
                // additive constant c spread exactly over the ncrow constant rows
                const long long tot = llrint(cs * sigma * 274877906944.0);
                const long long base = tot / ncrow, rem = tot - base * ncrow;  // |rem| < ncrow, same sign as tot
                const int i0 = k - 2 * J, i1 = i0 + 1;
                const long long ar = rem < 0 ? -rem : rem, sg = rem < 0 ? -1 : 1;
                q0 = base + (i0 < ar ? sg : 0);
                q1 = base + (i1 < ar ? sg : 0);
            }
        }
        split_digits5(q0, d[0][kk], d[1][kk], d[2][kk], d[3][kk], d[4][kk]);
        split_digits5(q1, d[0][kk + 1], d[1][kk + 1], d[2][kk + 1], d[3][kk + 1], d[4][kk + 1]);
    }
    uint8_t* img = Wimg + ((int64_t)sl * n_dt + dt) * a_bytes;
#pragma unroll
    for (int pl = 0; pl < tc::PW; ++pl)
        *reinterpret_cast<int4*>(img + (size_t)pl * a_plane + ((size_t)kc * tc::TM + row) * 16) =
            *reinterpret_cast<const int4*>(d[pl]);
}


// The 1024-point unit circle in 2^-30 fixed point (loglik_common.cuh: sincos_turns_q30), built once per context in float64.
__global__ void tc_phitab_kernel(PhiTabEntry* __restrict__ tab) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= kPhiTabSize) return;
    const double ang = 6.283185307179586476925 * (double)k / (double)kPhiTabSize;
    PhiTabEntry e;
    e.s = (int)llrint(sin(ang) * 1073741824.0);
    e.c = (int)llrint(cos(ang) * 1073741824.0);
    tab[k] = e;
}

// Phi digit planes + Poisson constants of PIMG_TILES consecutive tiles (TN permuted bins each) of one draw.
// 256 threads: row = tid % 32, eight threads per row interleave the 16-wide K chunks.
// Phi = (p1 2^24 + p2 2^16 + p3 2^8 + p4) 2^-30 with the sine / cosine formed directly in 2^-30 fixed point (PP = 4), or the
// round-1 variant (p1 2^16 + p2 2^8 + p3) 2^-22 from fp32 polynomials (PP = 3); p1 is the same digit in both.
constexpr int PIMG_TILES = 4;  // tiles per CTA: the 8 KB table and the draw's frequencies are staged once for all of them

// one 16-wide K chunk (8 sine / cosine pairs) of one bin -> 16 packed digit words -> the four planes' 16-byte rows
template <bool BOUNDED>
__device__ __forceinline__ void pimg_chunk_q30(int kc, int J, bool valid, float th, float tl, double t, const float2* __restrict__ s_w,
                                               const PhiTabEntry* __restrict__ s_tab, const double* __restrict__ omega_s,
                                               uint8_t* __restrict__ img, uint32_t b_plane, int row) {
    unsigned w[16];
#pragma unroll
    for (int kk = 0; kk < 16; kk += 2) {
        const int k = kc * 16 + kk;
        int q0 = 0, q1 = 0;
        if (valid) {
            if (k < 2 * J) {
                const float2 wv = s_w[k >> 1];
                if (BOUNDED) {
                    sincos_phase_q30_bounded(wv.x, wv.y, th, tl, s_tab, q1, q0);
                } else if (!sincos_phase_q30(wv.x, wv.y, th, tl, s_tab, q1, q0)) {
                    double sn, cn;
                    sincos(omega_s[k >> 1] * t, &sn, &cn);
                    q0 = (int)llrint(cn * 1073741824.0);
                    q1 = (int)llrint(sn * 1073741824.0);
                }
            } else {
                q0 = q1 = 1073741824;  // constant rows: Phi = 1
            }
        }
        w[kk] = pack_digits4(q0);      // cos row
        w[kk + 1] = pack_digits4(q1);  // sin row
    }
    // byte transpose: plane pl (0 = most significant digit) takes byte 3 - pl of each of the 16 words
#pragma unroll
    for (int pl = 0; pl < 4; ++pl) {
        const int sh = 24 - 8 * pl;
        int4 o;
        o.x = (int)(((w[0] >> sh) & 0xffu) | (((w[1] >> sh) & 0xffu) << 8) | (((w[2] >> sh) & 0xffu) << 16) | (((w[3] >> sh) & 0xffu) << 24));
        o.y = (int)(((w[4] >> sh) & 0xffu) | (((w[5] >> sh) & 0xffu) << 8) | (((w[6] >> sh) & 0xffu) << 16) | (((w[7] >> sh) & 0xffu) << 24));
        o.z = (int)(((w[8] >> sh) & 0xffu) | (((w[9] >> sh) & 0xffu) << 8) | (((w[10] >> sh) & 0xffu) << 16) | (((w[11] >> sh) & 0xffu) << 24));
        o.w = (int)(((w[12] >> sh) & 0xffu) | (((w[13] >> sh) & 0xffu) << 8) | (((w[14] >> sh) & 0xffu) << 16) | (((w[15] >> sh) & 0xffu) << 24));
        *reinterpret_cast<int4*>(img + (size_t)pl * b_plane + ((size_t)kc * tc::TN + row) * 16) = o;
    }
}

__global__ void __launch_bounds__(256) tc_pimg_kernel(int64_t S0, int J, int Kpad, int64_t T, const double* __restrict__ time,
                                                      const double* __restrict__ exposure, const int64_t* __restrict__ counts,
                                                      const int* __restrict__ perm, const int* __restrict__ n_nz,
                                                      const double* __restrict__ omega, const DrawConst* __restrict__ dc,
                                                      const TcDraw* __restrict__ td, const PhiTabEntry* __restrict__ phitab, int n_bt,
                                                      uint32_t b_bytes, uint32_t b_plane, uint8_t* __restrict__ Pimg) {
    const int row = threadIdx.x & (tc::TN - 1);
    const int kg = threadIdx.x / tc::TN;
    const int sl = blockIdx.y;
    const int64_t s = S0 + sl;
    // omega / 2 pi of this draw as fp32 hi + lo: the phases are formed without fp64 arithmetic (sincos_phase_df / _q30)
    __shared__ float2 s_w[IPN_MAX_J];
    __shared__ PhiTabEntry s_tab[tc::PP == 4 ? kPhiTabSize : 1];
    __shared__ unsigned int s_wmax;  // bits of max_j |omega_j / 2 pi| (fp32 hi part)
    if (threadIdx.x == 0) s_wmax = 0u;
    __syncthreads();
    for (int j = threadIdx.x; j < J; j += 256) {
        const double wt = omega[s * J + j] * 0.15915494309189534561;
        const float wh = (float)wt;
        s_w[j] = make_float2(wh, (float)(wt - (double)wh));
        atomicMax(&s_wmax, (unsigned)__float_as_int(fabsf(wh)));
    }
    if (tc::PP == 4)
        for (int k = threadIdx.x; k < kPhiTabSize; k += 256) s_tab[k] = phitab[k];
    __syncthreads();
    const float wmax = __int_as_float((int)s_wmax);
    const DrawConst dcs = dc[s];
    for (int bt = blockIdx.x * PIMG_TILES; bt < n_bt && bt < (blockIdx.x + 1) * PIMG_TILES; ++bt) {
        const int64_t pos = (int64_t)bt * tc::TN + row;
        const bool valid = pos < T;
        const int64_t i = valid ? perm[pos] : 0;
        const double t = valid ? time[i] : 0.0;
        const float th = (float)t, tl = (float)(t - (double)(float)t);
        uint8_t* img = Pimg + ((int64_t)sl * n_bt + bt) * b_bytes;
        if (tc::PP == 4) {
            // every phase of this bin within the fp32 reduction's range (|turns| < 2^22, with a margin for the rounding of the
            // product)?  Then the eight pairs of a chunk are straight-line code; the float64 route stays out of the loop.
            const bool bounded = fabsf(th) * wmax < 4.0e6f;
            for (int kc = kg; kc < Kpad / 16; kc += 256 / tc::TN) {
                if (bounded)
                    pimg_chunk_q30<true>(kc, J, valid, th, tl, t, s_w, s_tab, omega + s * J, img, b_plane, row);
                else
                    pimg_chunk_q30<false>(kc, J, valid, th, tl, t, s_w, s_tab, omega + s * J, img, b_plane, row);
            }
        } else {
            for (int kc = kg; kc < Kpad / 16; kc += 256 / tc::TN) {
                alignas(16) int8_t d[tc::PP][16];
#pragma unroll
                for (int kk = 0; kk < 16; kk += 2) {
                    const int k = kc * 16 + kk;
                    int q0 = 0, q1 = 0;
                    if (valid) {
                        if (k < 2 * J) {
                            float sn, cn;
                            const float2 w = s_w[k >> 1];
                            if (!sincos_phase_df(w.x, w.y, th, tl, sn, cn)) sincos_phase(omega[s * J + (k >> 1)] * t, sn, cn);
                            q0 = __float2int_rn(cn * 4194304.0f);  // cos row
                            q1 = __float2int_rn(sn * 4194304.0f);  // sin row
                        } else {
                            q0 = q1 = 4194304;  // constant rows: Phi = 1
                        }
                    }
                    int8_t dummy;
                    split_digits4(q0, dummy, d[0][kk], d[1][kk], d[2][kk]);
                    split_digits4(q1, dummy, d[0][kk + 1], d[1][kk + 1], d[2][kk + 1]);
                }
#pragma unroll
                for (int pl = 0; pl < tc::PP; ++pl)
                    *reinterpret_cast<int4*>(img + (size_t)pl * b_plane + ((size_t)kc * tc::TN + row) * 16) =
                        *reinterpret_cast<const int4*>(d[pl]);
            }
        }
        if (kg == 0) {
            const bool src = isfinite(dcs.c_hi);  // amp == 0: the likelihood is the constant C_s, every term is zero
            const double kap = (valid && src) ? exposure[i] * dcs.kap : 0.0;
            const float khi = (float)kap;
            float4 c;
            c.x = (valid && src) ? (float)counts[i] : 0.0f;
            c.y = khi;
            c.z = (float)(kap - (double)khi);
            c.w = c.x * 1.44269504f;  // n log2e, multiplies the log correction term
            *reinterpret_cast<float4*>(img + (size_t)tc::PP * b_plane + (size_t)row * 16) = c;
            // one flag per 8 rows (= per epilogue unit): the unit holds at least one bin with counts iff its first bin does (within
            // any 8 rows the bins with counts come first, tc_bin_position); the unit's first thread has that bin in i
            if ((row & 7) == 0)
                reinterpret_cast<float*>(img + (size_t)tc::PP * b_plane + (size_t)tc::TN * 16)[row >> 3] = (valid && counts[i] > 0) ? 1.0f : 0.0f;
        }
    }
}

}  // namespace ipn
