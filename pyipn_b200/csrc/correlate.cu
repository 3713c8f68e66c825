// Classical IPN chi^2 cross-correlation over a grid of integer shifts (drop-in arithmetic for `correlate`,
// pyipn/correlation.py:247-372; SURVEY.md section 8f row N1 -- the reference's own delay-grid scan).
//
// One thread per trial shift.  The re-binning of light curve 1 onto the bins of light curve 2 is a running
// (fT1, l) recurrence with fractional bin sharing, kept exactly as the reference orders it (float64, no
// reassociation) so chi^2 agrees to the last bits and the arg-min is identical.  Light curve 2's per-bin sums are
// shift-independent and are staged once per CTA in shared memory.
#include "ipn_common.cuh"

namespace ipn {

__global__ void __launch_bounds__(128) correlate_kernel(const double* __restrict__ t1, const double* __restrict__ c1,
                                                        const double* __restrict__ e1, const double* __restrict__ t2,
                                                        const double* __restrict__ c2, const double* __restrict__ e2,
                                                        int64_t n1, int64_t i_beg_2, int64_t i_end_2, int64_t i_beg_1,
                                                        int64_t n_max_1, double fscale, double dt_1, double dt_2,
                                                        int64_t n_sum_2, double* __restrict__ arr_dt,
                                                        double* __restrict__ arr_chi) {
    extern __shared__ double sh[];  // [2][n_k]: scaled counts / errors of light curve 2 per (summed) bin
    const int64_t n_k = (i_end_2 - i_beg_2) / n_sum_2 + 1;
    double* s_c2 = sh;
    double* s_e2 = sh + n_k;
    for (int64_t kk = threadIdx.x; kk < n_k; kk += blockDim.x) {
        const int64_t k = i_beg_2 + kk * n_sum_2;
        double fCnt2 = 0.0, fErr2 = 0.0;
        for (int64_t s = k; s < k + n_sum_2; ++s) {
            fCnt2 += c2[s];
            fErr2 += e2[s];
        }
        s_c2[kk] = fCnt2 * fscale;
        s_e2[kk] = fErr2 * (fscale * fscale);
    }
    __syncthreads();
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= n_max_1) return;
    const int64_t i = i_beg_1 + n;
    const double nDOF = (double)(i_end_2 - i_beg_2) / (double)n_sum_2;
    const double step2 = dt_2 * (double)n_sum_2;
    double fRij = 0.0, fT1 = dt_1, fT2 = 0.0, fdCnt1 = 0.0, fdErr1 = 0.0;
    int64_t l = i;
    for (int64_t kk = 0; kk < n_k; ++kk) {
        double fCnt1 = fdCnt1, fErr1 = fdErr1;
        const double lim = fT2 + step2;
        while (fT1 <= lim && l < n1) {
            fCnt1 += c1[l];
            fErr1 += e1[l];
            fT1 += dt_1;
            ++l;
        }
        const double a1 = l < n1 ? c1[l] : 0.0, b1 = l < n1 ? e1[l] : 0.0;  // the reference would read out of bounds here
        const double fdT = (fT1 - lim) / dt_1;
        fdCnt1 = fdT * a1;
        fCnt1 += (1.0 - fdT) * a1;
        fdErr1 = fdT * b1;
        fErr1 += (1.0 - fdT) * b1;
        ++l;
        fT1 += dt_1;
        fT2 += step2;
        double fDif = fCnt1 - s_c2[kk];
        fDif *= fDif;
        fDif /= fErr1 + s_e2[kk];
        fRij += fDif;
    }
    fRij /= nDOF - 1.0;
    arr_dt[n] = t2[i_beg_2] - t1[i];
    arr_chi[n] = fRij;
}

int launch_correlate(ipn_ctx* ctx, const double* t1, const double* c1, const double* e1, const double* t2, const double* c2,
                     const double* e2, int64_t n1, int64_t i_beg_2, int64_t i_end_2, int64_t i_beg_1, int64_t n_max_1,
                     double fscale, double dt_1, double dt_2, int64_t n_sum_2, double* arr_dt, double* arr_chi) {
    const int64_t n_k = (i_end_2 - i_beg_2) / n_sum_2 + 1;
    const size_t smem = 2 * (size_t)n_k * sizeof(double);
    if (smem > 200 * 1024) {
        set_error("correlation window of %lld bins does not fit in shared memory", (long long)n_k);
        return IPN_ERR_UNSUPPORTED;
    }
    IPN_CUDA_CHECK(cudaFuncSetAttribute(correlate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (int)((n_max_1 + 127) / 128);
    correlate_kernel<<<blocks, 128, smem, ctx->stream>>>(t1, c1, e1, t2, c2, e2, n1, i_beg_2, i_end_2, i_beg_1, n_max_1, fscale,
                                                         dt_1, dt_2, n_sum_2, arr_dt, arr_chi);
    ctx->launches += 1;
    IPN_CUDA_CHECK(cudaGetLastError());
    return IPN_OK;
}

}  // namespace ipn
