// Shared internals of libipn_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/ipn.h"

namespace ipn {

void set_error(const char* fmt, ...);
int* trap_record_device_ptr();  // host-mapped int[16]; nullptr if it cannot be allocated

#define IPN_CUDA_CHECK(expr)                                                                      \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ipn::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return IPN_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

#define IPN_REQUIRE(cond, ...)              \
    do {                                    \
        if (!(cond)) {                      \
            ipn::set_error(__VA_ARGS__);    \
            return IPN_ERR_INVALID_ARGUMENT; \
        }                                   \
    } while (0)

// growable device scratch buffer owned by the context
struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace ipn

struct ipn_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint64_t launches = 0;
    double last_kernel_ms = 0.0;
    // event pairs around the dominant (contraction) kernel launches of the most recent compute call
    std::vector<cudaEvent_t> main_ev;
    size_t main_ev_used = 0;
    double last_main_ms = 0.0;
    int64_t last_main_launches = 0;
    int last_impl = 0;
    int64_t last_draw_batch = 0;
    // options
    int loglik_impl = 0;
    int precise_w = 1;
    int hifi_mode = 0;  // tcgen05 path epilogue: 0 auto, 1 force fp64, 2 force plain fp32, 3 force compensated fp32
    int64_t draw_batch = 0;
    int rff_impl = 0;   // ipn_rff_eval: 0 auto, 1 direct kernel, 2 tcgen05 contraction (uniform bins guaranteed by the caller)
    int last_rff_impl = 0;
    bool tc_attr_set[8][4] = {};  // tcgen05 path: shared-memory attribute set on this device for (K steps, epilogue mode)
    bool tc_trap_set = false;     // ... and the time-out record's address published to the device code
    int n_terms = 1;    // how many likelihood passes are summed into the caller's result (sky grid: D detectors): the 1e-3 nat
                        // tolerance is shared between them, so the per-draw epilogue choice is made for a 1 / n_terms share
    // scratch
    ipn::Scratch wtab;     // delay tables of the current draw batch
    ipn::Scratch binprep;  // per-call per-bin constants
    ipn::Scratch draws;    // per-draw constants
    ipn::Scratch accum;    // per-(grid point) accumulators
    ipn::Scratch io[12];   // staging for the host-pointer entry points
    ipn::Scratch misc;
    ipn::Scratch perm;     // bin permutation (counts > 0 first) of the tcgen05 path
    ipn::Scratch dbg;      // pipeline wait counters (IPN_TC_DEBUG)
    ipn::Scratch tcdraw;   // per-draw scale factors of the tcgen05 path
    ipn::Scratch rffs;     // ipn_rff_eval through the contraction: bin / delay tables, signs, draw constants
    ipn::Scratch phitab;   // 1024-point fixed-point unit circle of the Phi image kernel (built on first use)
    void* pinned = nullptr;  // pinned host bounce buffer
    size_t pinned_cap = 0;
};

namespace ipn {

// ---- kernels' host launchers (each returns IPN_OK / error, bumps ctx->launches) -----------------------
int launch_rff_eval(ipn_ctx* ctx, int64_t T, const double* time, int64_t S, int64_t J, const double* omega,
                    const double* a, const double* b, const double* shift, const double* amp, double* out);

// Accumulates into d_ll (G x S, must be zeroed or hold a previous detector's sum when accumulate = true).
int run_loglik_grid(ipn_ctx* ctx, int64_t T, const double* time, const double* exposure, const int64_t* counts,
                    int64_t S, int64_t J, const double* omega, const double* a, const double* b, const double* amp,
                    const double* bkg, int64_t G, const double* delays, double* ll, bool accumulate);

int launch_poisson_counts(ipn_ctx* ctx, int64_t T, const double* exposure, int64_t S, const double* rates,
                          const double* bkg, uint64_t seed, int64_t* out);

int launch_sky_delays(ipn_ctx* ctx, int64_t D, const double* sc_pos, int64_t P, const double* ghat, double* delays);

int run_microbench(ipn_ctx* ctx, int which, double* result);

// ipn_rff_eval on uniform bins through the tcgen05 contraction (loglik_tc.cu); rff_uniform_dev checks time[i] = t0 + i dt on the device
bool rff_tc_supported(int64_t T, int64_t J);
int run_rff_eval_tc(ipn_ctx* ctx, int64_t T, const double* time, double t0, double dt, int64_t S, int64_t J, const double* omega,
                    const double* a, const double* b, const double* shift, const double* amp, double* out);
int rff_uniform_dev(ipn_ctx* ctx, int64_t T, const double* d_time, double* t0, double* dt, bool* uniform);

int launch_bin_events(ipn_ctx* ctx, int64_t n_events, const double* times, double tstart, double dt, int64_t n_edges,
                      int64_t* counts);

int launch_correlate(ipn_ctx* ctx, const double* t1, const double* c1, const double* e1, const double* t2, const double* h_c2,
                     const double* h_e2, int64_t n1, int64_t i_beg_2, int64_t i_end_2, int64_t i_beg_1, int64_t n_max_1,
                     double fscale, double dt_1, double dt_2, int64_t n_sum_2, double* arr_dt, double* arr_chi);

// bracket a dominant-kernel launch with events (summed by ipn_ctx_synchronize / the host entry points)
int main_kernel_begin(ipn_ctx* ctx);
int main_kernel_end(ipn_ctx* ctx);

}  // namespace ipn
