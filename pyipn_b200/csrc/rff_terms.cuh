// One term  a_j cos(w_j tau) + b_j sin(w_j tau)  of the RFF exponent in the two fp32 accuracy tiers of rff_eval_kernel
// (rff_eval.cu has the derivation), added to a Kahan sum.  Kept in a header of their own so that tests/emu/tc_emu.cpp can
// compile exactly these lines with g++ (IPN_HOST_EMU) and hold them against the fixtures minted from the reference's rff.py.
//   f = (wh, wl, a, b) with w = omega / 2 pi in turns per unit time as an fp32 hi/lo pair; tau = th + tl likewise.
#pragma once
#include "loglik_common.cuh"

namespace ipn {

// plain fp32 tier (A2 = sum a^2 + b^2 <= 12)
__device__ __forceinline__ void rff_term(const float4 f, float th, float tl, float& sum, float& comp) {
    const float ph = f.x * th;
    const float e = fmaf(f.x, th, -ph);
    const float cross = fmaf(f.x, tl, fmaf(f.y, th, e));
    const float k = (ph + 12582912.0f) - 12582912.0f;
    float sn, cs;
    sincos_turns(ph - k, cross, sn, cs);
    const float v = fmaf(f.w, sn, f.z * cs);
    const float yk = v - comp;  // Kahan
    const float t = sum + yk;
    comp = (t - sum) - yk;
    sum = t;
}

// double-float tier (12 < A2 <= 100): the two products and their sum carry exact error terms, collected in `lo`
__device__ __forceinline__ void rff_term_df(const float4 f, float th, float tl, float& sum, float& comp, float& lo) {
    const float ph = f.x * th;
    const float e = fmaf(f.x, th, -ph);
    const float cross = fmaf(f.x, tl, fmaf(f.y, th, e));
    const float k = (ph + 12582912.0f) - 12582912.0f;
    float sn, cs;
    sincos_turns(ph - k, cross, sn, cs);
    const float pc = f.z * cs, epc = fmaf(f.z, cs, -pc);  // a cos, exact error
    const float ps = f.w * sn, eps = fmaf(f.w, sn, -ps);  // b sin, exact error
    const float v = pc + ps;                              // TwoSum
    const float bb = v - pc;
    const float ev = (pc - (v - bb)) + (ps - bb);
    lo += (epc + eps) + ev;
    const float yk = v - comp;  // Kahan
    const float t = sum + yk;
    comp = (t - sum) - yk;
    sum = t;
}

}  // namespace ipn
