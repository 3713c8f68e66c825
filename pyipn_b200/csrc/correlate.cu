// Classical IPN chi^2 cross-correlation over a grid of integer shifts (drop-in arithmetic for `correlate`,
// pyipn/correlation.py:247-372; SURVEY.md section 8f row N1 -- the reference's own delay-grid scan).
//
// The reference re-bins light curve 1 onto the (summed) bins of light curve 2 with a running (fT1, fT2, l) recurrence and
// fractional sharing of the bin that straddles an edge, once per trial shift.  That recurrence does not depend on the shift:
// for every bin k of light curve 2 it consumes the same NUMBER of whole bins of light curve 1 and the same FRACTION fdT_k of
// the next one, only the starting index moves with the shift.  So:
//   * the plan (first whole bin, number of whole bins, fraction, per k) is worked out ONCE on the host in the reference's own
//     float64 operations (corr_plan_host) together with light curve 2's scaled per-bin sums -- no thread walks a `while`
//     loop, every thread of the grid executes the same instruction stream;
//   * one CTA = 128 consecutive shifts; the bins of light curve 1 that a stretch of the plan needs (window of the stretch +
//     127) are staged once in shared memory with coalesced loads, and thread n reads s_c1[offset + n]: consecutive threads,
//     consecutive addresses, no bank conflicts, every staged value used by up to 128 shifts;
//   * per (shift, k) the additions are made in the reference's order (carry, whole bins in sequence, fraction), so chi^2
//     agrees with the numba function to the last bits (tests: 1e-11 relative, arg-min exact).
// Bins of light curve 1 beyond its end read as zero (the reference would index out of bounds there).
#include <vector>

#include "ipn_common.cuh"

namespace ipn {

struct CorrStep {
    int l0;      // first whole bin of light curve 1 for this bin of light curve 2, relative to the shift's start bin
    int cnt;     // number of whole bins
    double fdT;  // fraction of the straddling bin (l0 + cnt) that belongs to the NEXT bin of light curve 2
    double c2;   // light curve 2: scaled counts of the (summed) bin
    double e2;   // ... and scaled error term
};

constexpr int CORR_THREADS = 128;
constexpr int CORR_WINDOW = 6144;  // bins of light curve 1 staged per stretch of the plan (+ CORR_THREADS - 1), 2 x 49 KB

__global__ void __launch_bounds__(CORR_THREADS)
correlate_kernel(const double* __restrict__ t1, const double* __restrict__ c1, const double* __restrict__ e1,
                 const double* __restrict__ t2, int64_t n1, int64_t i_beg_2, int64_t i_beg_1, int64_t n_max_1, int64_t n_k,
                 const CorrStep* __restrict__ plan, const int* __restrict__ stretch, int n_stretch, double dof_m1,
                 double* __restrict__ arr_dt, double* __restrict__ arr_chi) {
    extern __shared__ double sh[];
    double* s_c1 = sh;
    double* s_e1 = sh + (CORR_WINDOW + CORR_THREADS);
    const int64_t n0 = blockIdx.x * (int64_t)CORR_THREADS;
    const int tid = threadIdx.x;
    const int64_t n = n0 + tid;
    double fRij = 0.0, fdCnt1 = 0.0, fdErr1 = 0.0;
    for (int sidx = 0; sidx < n_stretch; ++sidx) {
        const int k0 = stretch[sidx], k1 = stretch[sidx + 1];
        const int base = plan[k0].l0;                                     // first bin of light curve 1 of the stretch (shift 0)
        const int len = plan[k1 - 1].l0 + plan[k1 - 1].cnt + 1 - base;    // ... through the last straddling bin
        __syncthreads();
        for (int j = tid; j < len + CORR_THREADS - 1; j += CORR_THREADS) {
            const int64_t l = i_beg_1 + n0 + base + j;
            s_c1[j] = l < n1 ? c1[l] : 0.0;
            s_e1[j] = l < n1 ? e1[l] : 0.0;
        }
        __syncthreads();
        for (int k = k0; k < k1; ++k) {
            const CorrStep st = plan[k];  // the same address for every thread: one broadcast load
            const double* pc = s_c1 + (st.l0 - base) + tid;
            const double* pe = s_e1 + (st.l0 - base) + tid;
            double fCnt1 = fdCnt1, fErr1 = fdErr1;
            for (int m = 0; m < st.cnt; ++m) {
                fCnt1 += pc[m];
                fErr1 += pe[m];
            }
            const double a1 = pc[st.cnt], b1 = pe[st.cnt];
            fdCnt1 = st.fdT * a1;
            fCnt1 += (1.0 - st.fdT) * a1;
            fdErr1 = st.fdT * b1;
            fErr1 += (1.0 - st.fdT) * b1;
            double fDif = fCnt1 - st.c2;
            fDif *= fDif;
            fDif /= fErr1 + st.e2;
            fRij += fDif;
        }
    }
    if (n < n_max_1) {
        arr_dt[n] = t2[i_beg_2] - t1[i_beg_1 + n];
        arr_chi[n] = fRij / dof_m1;
    }
}

// The shift-independent half of the reference's loop, in its own float64 operations (correlation.py:318-350): fT1 / fT2
// recurrences, the `while fT1 <= fT2 + dt_2 n_sum_2` count, fdT, and light curve 2's per-bin sums.
static void corr_plan_host(const double* c2, const double* e2, int64_t i_beg_2, int64_t i_end_2, double fscale, double dt_1,
                           double dt_2, int64_t n_sum_2, std::vector<CorrStep>& plan, std::vector<int>& stretch) {
    const double step2 = dt_2 * (double)n_sum_2;
    double fT1 = dt_1, fT2 = 0.0;
    int64_t l = 0;
    for (int64_t k = i_beg_2; k <= i_end_2; k += n_sum_2) {
        CorrStep st;
        double fCnt2 = 0.0, fErr2 = 0.0;
        for (int64_t s = k; s < k + n_sum_2; ++s) {
            fCnt2 += c2[s];
            fErr2 += e2[s];
        }
        st.c2 = fCnt2 * fscale;
        st.e2 = fErr2 * (fscale * fscale);
        const double lim = fT2 + step2;
        st.l0 = (int)l;
        int cnt = 0;
        while (fT1 <= lim) {
            fT1 += dt_1;
            ++l;
            ++cnt;
        }
        st.cnt = cnt;
        st.fdT = (fT1 - lim) / dt_1;
        ++l;
        fT1 += dt_1;
        fT2 += step2;
        plan.push_back(st);
    }
    // stretches of the plan whose window of light curve 1 fits the shared-memory stage
    stretch.push_back(0);
    int start = 0;
    for (int k = 0; k < (int)plan.size(); ++k) {
        const int len = plan[k].l0 + plan[k].cnt + 1 - plan[start].l0;
        if (len > CORR_WINDOW && k > start) {
            stretch.push_back(k);
            start = k;
        }
    }
    stretch.push_back((int)plan.size());
}

// c2 / e2 are HOST arrays here (the plan is built on the host); everything else is on the device.
int launch_correlate(ipn_ctx* ctx, const double* t1, const double* c1, const double* e1, const double* t2, const double* h_c2,
                     const double* h_e2, int64_t n1, int64_t i_beg_2, int64_t i_end_2, int64_t i_beg_1, int64_t n_max_1,
                     double fscale, double dt_1, double dt_2, int64_t n_sum_2, double* arr_dt, double* arr_chi) {
    std::vector<CorrStep> plan;
    std::vector<int> stretch;
    const double ratio = dt_2 * (double)n_sum_2 / dt_1;
    if (!(ratio > 0.0) || ratio > (double)(CORR_WINDOW - 2)) {
        set_error("correlate: one bin of light curve 2 spans %.3g bins of light curve 1 (supported: up to %d)", ratio, CORR_WINDOW - 2);
        return IPN_ERR_UNSUPPORTED;
    }
    corr_plan_host(h_c2, h_e2, i_beg_2, i_end_2, fscale, dt_1, dt_2, n_sum_2, plan, stretch);
    const int64_t n_k = (int64_t)plan.size();
    const size_t plan_bytes = plan.size() * sizeof(CorrStep), str_bytes = stretch.size() * sizeof(int);
    int rc = ctx->misc.ensure(plan_bytes + str_bytes + 64);
    if (rc) return rc;
    CorrStep* d_plan = ctx->misc.as<CorrStep>();
    int* d_stretch = reinterpret_cast<int*>(ctx->misc.as<char>() + plan_bytes);
    // pageable sources: the copies are staged by the runtime before the calls return, so the vectors may go out of scope
    IPN_CUDA_CHECK(cudaMemcpyAsync(d_plan, plan.data(), plan_bytes, cudaMemcpyHostToDevice, ctx->stream));
    IPN_CUDA_CHECK(cudaMemcpyAsync(d_stretch, stretch.data(), str_bytes, cudaMemcpyHostToDevice, ctx->stream));
    const size_t smem = 2 * (size_t)(CORR_WINDOW + CORR_THREADS) * sizeof(double);
    IPN_CUDA_CHECK(cudaFuncSetAttribute(correlate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const double nDOF = (double)(i_end_2 - i_beg_2) / (double)n_sum_2;
    const int blocks = (int)((n_max_1 + CORR_THREADS - 1) / CORR_THREADS);
    correlate_kernel<<<blocks, CORR_THREADS, smem, ctx->stream>>>(t1, c1, e1, t2, n1, i_beg_2, i_beg_1, n_max_1, n_k, d_plan, d_stretch,
                                                                   (int)stretch.size() - 1, nDOF - 1.0, arr_dt, arr_chi);
    ctx->launches += 1;
    IPN_CUDA_CHECK(cudaGetLastError());
    IPN_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));  // the plan's host buffers die with this frame
    return IPN_OK;
}

}  // namespace ipn
