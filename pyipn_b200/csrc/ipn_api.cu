// C ABI of libipn_b200 (include/ipn.h): context management, argument validation, host<->device staging.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "ipn_common.cuh"

namespace ipn {

static thread_local char g_err[1024] = "";

// Record a kernel writes (through a host-mapped pointer) just before it traps on a barrier time-out; host memory stays
// readable after the context has died with "unspecified launch failure", so the message can say which wait it was.
static volatile int* g_trap_host = nullptr;

int* trap_record_device_ptr() {
    static int* dev = nullptr;
    if (!dev) {
        int* host = nullptr;
        if (cudaHostAlloc(&host, 32 * sizeof(int), cudaHostAllocMapped) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        memset(host, 0, 32 * sizeof(int));
        if (cudaHostGetDevicePointer(&dev, host, 0) != cudaSuccess) {
            cudaGetLastError();
            dev = nullptr;
            return nullptr;
        }
        g_trap_host = host;
    }
    return dev;
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    if (g_trap_host && g_trap_host[0] != 0 && n >= 0 && n < (int)sizeof(g_err) - 1)
    {
        n += snprintf(g_err + n, sizeof(g_err) - n, " [kernel barrier time-out: wait id %d, block %d, warp %d, parity %d, sequence %d; last wait (id:sequence) per warp:",
                      g_trap_host[0], g_trap_host[1], g_trap_host[2], g_trap_host[3], g_trap_host[4]);
        for (int w = 0; w < 24 && n < (int)sizeof(g_err) - 24; ++w)
            n += snprintf(g_err + n, sizeof(g_err) - n, " %d:%d", (int)((unsigned)g_trap_host[8 + w] >> 24), g_trap_host[8 + w] & 0xFFFFFF);
        if (n < (int)sizeof(g_err) - 2) snprintf(g_err + n, sizeof(g_err) - n, "]");
    }
}

int Scratch::ensure(size_t bytes) {
    if (bytes <= cap) return IPN_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        p = nullptr;
        set_error("cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        return IPN_ERR_ALLOC;
    }
    cap = want;
    return IPN_OK;
}

void Scratch::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

static int ctx_guard(ipn_ctx* ctx) {
    if (!ctx) {
        set_error("ipn_ctx is NULL");
        return IPN_ERR_INVALID_ARGUMENT;
    }
    IPN_CUDA_CHECK(cudaSetDevice(ctx->device));
    return IPN_OK;
}

static cudaEvent_t next_main_event(ipn_ctx* ctx) {
    if (ctx->main_ev_used == ctx->main_ev.size()) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        ctx->main_ev.push_back(e);
    }
    return ctx->main_ev[ctx->main_ev_used++];
}
int main_kernel_begin(ipn_ctx* ctx) {
    cudaEvent_t e = next_main_event(ctx);
    if (e) IPN_CUDA_CHECK(cudaEventRecord(e, ctx->stream));
    return IPN_OK;
}
int main_kernel_end(ipn_ctx* ctx) { return main_kernel_begin(ctx); }

static int begin_timing(ipn_ctx* ctx) {
    ctx->main_ev_used = 0;
    IPN_CUDA_CHECK(cudaEventRecord(ctx->ev0, ctx->stream));
    return IPN_OK;
}
static int end_timing(ipn_ctx* ctx) {
    IPN_CUDA_CHECK(cudaEventRecord(ctx->ev1, ctx->stream));
    return IPN_OK;
}
static int finish_timing(ipn_ctx* ctx) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess)
        ctx->last_kernel_ms = ms;
    else
        cudaGetLastError();
    double tot = 0.0;
    for (size_t i = 0; i + 1 < ctx->main_ev_used; i += 2) {
        float m = 0.f;
        if (cudaEventElapsedTime(&m, ctx->main_ev[i], ctx->main_ev[i + 1]) == cudaSuccess)
            tot += m;
        else
            cudaGetLastError();
    }
    ctx->last_main_ms = tot;
    ctx->last_main_launches = (int64_t)(ctx->main_ev_used / 2);
    return IPN_OK;
}

static int h2d(ipn_ctx* ctx, Scratch& sc, const void* host, size_t bytes) {
    int rc = sc.ensure(bytes ? bytes : 8);
    if (rc) return rc;
    if (bytes) IPN_CUDA_CHECK(cudaMemcpyAsync(sc.p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return IPN_OK;
}

static bool all_finite(const double* v, int64_t n) {
    for (int64_t i = 0; i < n; ++i)
        if (!isfinite(v[i])) return false;
    return true;
}

}  // namespace ipn

using namespace ipn;

extern "C" {

const char* ipn_last_error(void) { return g_err; }
const char* ipn_version(void) { return "libipn_b200 0.1 (sm_100a)"; }

int ipn_ctx_create(int device, ipn_ctx** out) {
    if (!out) {
        set_error("out is NULL");
        return IPN_ERR_INVALID_ARGUMENT;
    }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); libipn_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return IPN_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        set_error("device %d out of range (0..%d)", device, n - 1);
        return IPN_ERR_INVALID_ARGUMENT;
    }
    cudaDeviceProp prop;
    IPN_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libipn_b200 only contains sm_100a code (B200)", device, prop.major, prop.minor);
        return IPN_ERR_NO_DEVICE;
    }
    IPN_CUDA_CHECK(cudaSetDevice(device));
    ipn_ctx* ctx = new ipn_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    // a BLOCKING stream: it keeps the implicit ordering with the legacy default stream (handle 0), so work a caller
    // enqueues there before / after a _dev call is ordered with the library's kernels without further ado
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamDefault) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        set_error("stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete ctx;
        return IPN_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return IPN_OK;
}

void ipn_ctx_destroy(ipn_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->wtab.release();
    ctx->binprep.release();
    ctx->draws.release();
    ctx->accum.release();
    ctx->misc.release();
    ctx->tcdraw.release();
    ctx->phitab.release();
    ctx->rffs.release();
    ctx->dbg.release();
    ctx->perm.release();
    for (auto& s : ctx->io) s.release();
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t e : ctx->main_ev) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int ipn_ctx_set_stream(ipn_ctx* ctx, void* cuda_stream) {
    int rc = ctx_guard(ctx);
    if (rc) return rc;
    // the handle is taken literally: NULL is CUDA's legacy default stream (what torch.cuda.current_stream().cuda_stream
    // is unless the caller entered a torch.cuda.stream(...) block), exactly as in every CUDA runtime call
    ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return IPN_OK;
}

int ipn_ctx_use_own_stream(ipn_ctx* ctx) {
    int rc = ctx_guard(ctx);
    if (rc) return rc;
    ctx->stream = ctx->own_stream;
    return IPN_OK;
}

int ipn_ctx_synchronize(ipn_ctx* ctx) {
    int rc = ctx_guard(ctx);
    if (rc) return rc;
    IPN_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    finish_timing(ctx);
    return IPN_OK;
}

int ipn_ctx_set_option(ipn_ctx* ctx, int key, int64_t value) {
    if (!ctx) {
        set_error("ipn_ctx is NULL");
        return IPN_ERR_INVALID_ARGUMENT;
    }
    switch (key) {
        case IPN_OPT_LOGLIK_IMPL:
            IPN_REQUIRE(value >= 0 && value <= 2, "IPN_OPT_LOGLIK_IMPL must be 0, 1 or 2");
            ctx->loglik_impl = (int)value;
            return IPN_OK;
        case IPN_OPT_PRECISE_W:
            ctx->precise_w = value != 0;
            return IPN_OK;
        case IPN_OPT_HIFI:
            IPN_REQUIRE(value >= 0 && value <= 3, "IPN_OPT_HIFI must be 0 (auto), 1 (fp64), 2 (fp32) or 3 (compensated fp32)");
            ctx->hifi_mode = (int)value;
            return IPN_OK;
        case IPN_OPT_RFF_IMPL:
            IPN_REQUIRE(value >= 0 && value <= 2, "IPN_OPT_RFF_IMPL must be 0 (auto), 1 (direct kernel) or 2 (tcgen05 contraction, uniform bins)");
            ctx->rff_impl = (int)value;
            return IPN_OK;
        case IPN_OPT_DRAW_BATCH:
            IPN_REQUIRE(value >= 0, "IPN_OPT_DRAW_BATCH must be >= 0");
            ctx->draw_batch = value;
            return IPN_OK;
        default:
            set_error("unknown option key %d", key);
            return IPN_ERR_INVALID_ARGUMENT;
    }
}

uint64_t ipn_ctx_launch_count(const ipn_ctx* ctx) { return ctx ? ctx->launches : 0; }
double ipn_ctx_last_kernel_ms(const ipn_ctx* ctx) { return ctx ? ctx->last_kernel_ms : 0.0; }
int ipn_ctx_last_loglik_impl(const ipn_ctx* ctx) { return ctx ? ctx->last_impl : 0; }
int ipn_ctx_last_rff_impl(const ipn_ctx* ctx) { return ctx ? ctx->last_rff_impl : 0; }
int64_t ipn_ctx_last_draw_batch(const ipn_ctx* ctx) { return ctx ? ctx->last_draw_batch : 0; }
double ipn_ctx_last_main_kernel_ms(const ipn_ctx* ctx, int64_t* launches) {
    if (!ctx) return 0.0;
    if (launches) *launches = ctx->last_main_launches;
    return ctx->last_main_ms;
}

// ---------------------------------------------------------------------------------------------------
static int check_rff_args(int64_t T, const void* time, int64_t S, int64_t J, const void* omega, const void* a,
                          const void* b, const void* shift, const void* amp, const void* out) {
    IPN_REQUIRE(T >= 0 && S >= 0 && J >= 1, "need T >= 0, S >= 0, J >= 1 (got T=%lld S=%lld J=%lld)", (long long)T,
                (long long)S, (long long)J);
    if (J > IPN_MAX_J) {
        set_error("J = %lld exceeds IPN_MAX_J = %d", (long long)J, IPN_MAX_J);
        return IPN_ERR_UNSUPPORTED;
    }
    if (T == 0 || S == 0) return IPN_OK;
    IPN_REQUIRE(time && omega && a && b && shift && amp && out, "NULL array argument");
    return IPN_OK;
}

// Are the bins uniform (time[i] = time[0] + i dt)?  The tolerance is what float64 arithmetic on np.arange edges leaves (a few
// ulp of the largest time) -- far below what matters: a deviation d enters the phase as omega d.
static bool uniform_host(int64_t T, const double* time, double* t0, double* dt) {
    if (T < 2) return false;
    *t0 = time[0];
    *dt = (time[T - 1] - time[0]) / (double)(T - 1);
    if (!(*dt > 0.0) || !std::isfinite(*dt)) return false;
    const double tol = 1e-12 * (std::fabs(time[0]) + std::fabs(time[T - 1]));
    for (int64_t i = 0; i < T; ++i)
        if (!(std::fabs(time[i] - (*t0 + (double)i * *dt)) <= tol)) return false;
    return true;
}

// device arrays; route by IPN_OPT_RFF_IMPL (uniform / t0 / dt already known for the host entry point)
static int rff_eval_device(ipn_ctx* ctx, int64_t T, const double* time, int64_t S, int64_t J, const double* omega, const double* a,
                           const double* b, const double* shift, const double* amp, double* out, int known_uniform, double t0,
                           double dt) {
    int rc;
    bool contraction = false;
    if (ctx->rff_impl != 1 && rff_tc_supported(T, J)) {
        if (known_uniform < 0) {  // not known yet: ask the device
            bool uni = false;
            if (ctx->rff_impl == 2) {
                double t2[2];
                IPN_CUDA_CHECK(cudaMemcpyAsync(&t2[0], time, 8, cudaMemcpyDeviceToHost, ctx->stream));
                IPN_CUDA_CHECK(cudaMemcpyAsync(&t2[1], time + (T - 1), 8, cudaMemcpyDeviceToHost, ctx->stream));
                IPN_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
                t0 = t2[0];
                dt = (t2[1] - t2[0]) / (double)(T - 1);
                uni = dt > 0.0;
            } else if ((rc = rff_uniform_dev(ctx, T, time, &t0, &dt, &uni))) {
                return rc;
            }
            known_uniform = uni ? 1 : 0;
        }
        contraction = known_uniform == 1;
    }
    if (ctx->rff_impl == 2 && !contraction) {
        set_error("IPN_OPT_RFF_IMPL = 2 needs uniform bins, T >= 1024 and J <= 108 (T = %lld, J = %lld)", (long long)T, (long long)J);
        return IPN_ERR_UNSUPPORTED;
    }
    ctx->last_rff_impl = contraction ? 2 : 1;
    if (contraction) return run_rff_eval_tc(ctx, T, time, t0, dt, S, J, omega, a, b, shift, amp, out);
    return launch_rff_eval(ctx, T, time, S, J, omega, a, b, shift, amp, out);
}

int ipn_rff_eval_dev(ipn_ctx* ctx, int64_t T, const double* time, int64_t S, int64_t J, const double* omega,
                     const double* a, const double* b, const double* shift, const double* amp, double* out) {
    int rc = ctx_guard(ctx);
    if (rc) return rc;
    rc = check_rff_args(T, time, S, J, omega, a, b, shift, amp, out);
    if (rc) return rc;
    if (T == 0 || S == 0) return IPN_OK;
    begin_timing(ctx);
    rc = rff_eval_device(ctx, T, time, S, J, omega, a, b, shift, amp, out, -1, 0.0, 0.0);
    end_timing(ctx);
    return rc;
}

int ipn_rff_eval(ipn_ctx* ctx, int64_t T, const double* time, int64_t S, int64_t J, const double* omega,
                 const double* a, const double* b, const double* shift, const double* amp, double* out) {
    int rc = ctx_guard(ctx);
    if (rc) return rc;
    rc = check_rff_args(T, time, S, J, omega, a, b, shift, amp, out);
    if (rc) return rc;
    if (T == 0 || S == 0) return IPN_OK;
    double t0 = 0.0, dt = 0.0;
    const int uni = (ctx->rff_impl != 1 && uniform_host(T, time, &t0, &dt)) ? 1 : 0;
    const size_t sj = (size_t)S * J * sizeof(double);
    if ((rc = h2d(ctx, ctx->io[0], time, (size_t)T * 8))) return rc;
    if ((rc = h2d(ctx, ctx->io[1], omega, sj))) return rc;
    if ((rc = h2d(ctx, ctx->io[2], a, sj))) return rc;
    if ((rc = h2d(ctx, ctx->io[3], b, sj))) return rc;
    if ((rc = h2d(ctx, ctx->io[4], shift, (size_t)S * 8))) return rc;
    if ((rc = h2d(ctx, ctx->io[5], amp, (size_t)S * 8))) return rc;
    if ((rc = ctx->io[6].ensure((size_t)S * T * 8))) return rc;
    begin_timing(ctx);
    rc = rff_eval_device(ctx, T, ctx->io[0].as<double>(), S, J, ctx->io[1].as<double>(), ctx->io[2].as<double>(),
                         ctx->io[3].as<double>(), ctx->io[4].as<double>(), ctx->io[5].as<double>(), ctx->io[6].as<double>(), uni, t0, dt);
    end_timing(ctx);
    if (rc) return rc;
    IPN_CUDA_CHECK(cudaMemcpyAsync(out, ctx->io[6].p, (size_t)S * T * 8, cudaMemcpyDeviceToHost, ctx->stream));
    IPN_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    finish_timing(ctx);
    return IPN_OK;
}

// ---------------------------------------------------------------------------------------------------
static int check_grid_shapes(int64_t T, int64_t S, int64_t J, int64_t G) {
    IPN_REQUIRE(T >= 0 && S >= 0 && G >= 0 && J >= 1, "need T, S, G >= 0 and J >= 1 (got T=%lld S=%lld J=%lld G=%lld)",
                (long long)T, (long long)S, (long long)J, (long long)G);
    if (J > IPN_MAX_J) {
        set_error("J = %lld exceeds IPN_MAX_J = %d", (long long)J, IPN_MAX_J);
        return IPN_ERR_UNSUPPORTED;
    }
    // kernel index types and launch geometry: bins are permuted through 32-bit indices, delay tiles (128 delays) ride in gridDim.y
    if (T > 0x7fffffffLL || G > 65535LL * 128) {
        set_error("grid too large for one call: T = %lld (max 2^31 - 1), G = %lld (max 8388480); split it", (long long)T, (long long)G);
        return IPN_ERR_UNSUPPORTED;
    }
    return IPN_OK;
}

static int check_grid_host_values(int64_t T, const double* time, const double* exposure, const int64_t* counts,
                                  int64_t S, int64_t J, const double* omega, const double* a, const double* b,
                                  const double* amp, const double* bkg) {
    IPN_REQUIRE(all_finite(time, T), "time contains a non-finite value");
    for (int64_t i = 0; i < T; ++i) {
        IPN_REQUIRE(counts[i] >= 0, "counts[%lld] = %lld is negative", (long long)i, (long long)counts[i]);
        IPN_REQUIRE(isfinite(exposure[i]) && exposure[i] >= 0.0, "exposure[%lld] is negative or non-finite", (long long)i);
    }
    IPN_REQUIRE(all_finite(omega, S * J) && all_finite(a, S * J) && all_finite(b, S * J),
                "omega/a/b contain a non-finite value");
    for (int64_t i = 0; i < S * J; ++i) {
        // |a_j|, |b_j| are log-intensity amplitudes (the Stan model has them ~ N(0,1)/sqrt(k)); beyond e^40 per term the
        // 8-bit digit planes of the delay table would saturate and the fp32 exponent is meaningless anyway
        if (a[i] * a[i] + b[i] * b[i] > 1600.0) {
            set_error("draw %lld: |a_j|, |b_j| of feature %lld exceed 40 (log-intensity amplitude)", (long long)(i / J), (long long)(i % J));
            return IPN_ERR_UNSUPPORTED;
        }
    }
    for (int64_t s = 0; s < S; ++s) {
        IPN_REQUIRE(isfinite(bkg[s]) && bkg[s] > 0.0, "bkg[%lld] must be finite and > 0", (long long)s);
        IPN_REQUIRE(isfinite(amp[s]) && amp[s] >= 0.0, "amp[%lld] must be finite and >= 0", (long long)s);
        IPN_REQUIRE(amp[s] == 0.0 || fabs(log2(amp[s] / bkg[s])) < 100.0, "amp/bkg ratio of draw %lld out of range",
                    (long long)s);
    }
    return IPN_OK;
}

int ipn_loglik_delay_grid_dev(ipn_ctx* ctx, int64_t T, const double* time, const double* exposure,
                              const int64_t* counts, int64_t S, int64_t J, const double* omega, const double* a,
                              const double* b, const double* amp, const double* bkg, int64_t G, const double* delays,
                              double* ll) {
    int rc = ctx_guard(ctx);
    if (rc) return rc;
    rc = check_grid_shapes(T, S, J, G);
    if (rc) return rc;
    if (S == 0 || G == 0) return IPN_OK;
    IPN_REQUIRE(omega && a && b && amp && bkg && delays && ll && (T == 0 || (time && exposure && counts)),
                "NULL array argument");
    begin_timing(ctx);
    rc = run_loglik_grid(ctx, T, time, exposure, counts, S, J, omega, a, b, amp, bkg, G, delays, ll, false);
    end_timing(ctx);
    return rc;
}

int ipn_loglik_delay_grid(ipn_ctx* ctx, int64_t T, const double* time, const double* exposure, const int64_t* counts,
                          int64_t S, int64_t J, const double* omega, const double* a, const double* b,
                          const double* amp, const double* bkg, int64_t G, const double* delays, double* ll) {
    int rc = ctx_guard(ctx);
    if (rc) return rc;
    rc = check_grid_shapes(T, S, J, G);
    if (rc) return rc;
    if (S == 0 || G == 0) return IPN_OK;
    IPN_REQUIRE(omega && a && b && amp && bkg && delays && ll && (T == 0 || (time && exposure && counts)),
                "NULL array argument");
    rc = check_grid_host_values(T, time, exposure, counts, S, J, omega, a, b, amp, bkg);
    if (rc) return rc;
    IPN_REQUIRE(all_finite(delays, G), "delays contain a non-finite value");
    const size_t sj = (size_t)S * J * 8;
    if ((rc = h2d(ctx, ctx->io[0], time, (size_t)T * 8))) return rc;
    if ((rc = h2d(ctx, ctx->io[1], exposure, (size_t)T * 8))) return rc;
    if ((rc = h2d(ctx, ctx->io[2], counts, (size_t)T * 8))) return rc;
    if ((rc = h2d(ctx, ctx->io[3], omega, sj))) return rc;
    if ((rc = h2d(ctx, ctx->io[4], a, sj))) return rc;
    if ((rc = h2d(ctx, ctx->io[5], b, sj))) return rc;
    if ((rc = h2d(ctx, ctx->io[6], amp, (size_t)S * 8))) return rc;
    if ((rc = h2d(ctx, ctx->io[7], bkg, (size_t)S * 8))) return rc;
    if ((rc = h2d(ctx, ctx->io[8], delays, (size_t)G * 8))) return rc;
    if ((rc = ctx->io[9].ensure((size_t)G * S * 8))) return rc;
    begin_timing(ctx);
    rc = run_loglik_grid(ctx, T, ctx->io[0].as<double>(), ctx->io[1].as<double>(), ctx->io[2].as<int64_t>(), S, J,
                         ctx->io[3].as<double>(), ctx->io[4].as<double>(), ctx->io[5].as<double>(),
                         ctx->io[6].as<double>(), ctx->io[7].as<double>(), G, ctx->io[8].as<double>(),
                         ctx->io[9].as<double>(), false);
    end_timing(ctx);
    if (rc) return rc;
    IPN_CUDA_CHECK(cudaMemcpyAsync(ll, ctx->io[9].p, (size_t)G * S * 8, cudaMemcpyDeviceToHost, ctx->stream));
    IPN_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    finish_timing(ctx);
    return IPN_OK;
}

// functions.stan:37-61 + 65-69 on device arrays: delays per (detector, direction), then one delay-grid pass per detector
// accumulating into d_ll (P x S).  nbins is a HOST array.  The per-detector passes share the 1e-3 nat tolerance of the sum,
// so the epilogue choice of every pass is made for a 1 / D share of it (ctx->n_terms).
static int sky_grid_device(ipn_ctx* ctx, int64_t D, const int64_t* nbins, int64_t Tmax, const double* d_time,
                           const double* d_exposure, const int64_t* d_counts, const double* d_sc_pos, int64_t P,
                           const double* d_ghat, int64_t S, int64_t J, const double* d_omega, const double* d_a,
                           const double* d_b, const double* d_amp, const double* d_bkg, double* d_ll) {
    int rc;
    if ((rc = ctx->io[11].ensure((size_t)D * P * 8))) return rc;
    rc = launch_sky_delays(ctx, D, d_sc_pos, P, d_ghat, ctx->io[11].as<double>());
    if (rc) return rc;
    ctx->n_terms = (int)D;
    for (int64_t d = 0; d < D; ++d) {
        // detector 0 is unshifted (rff_omega_ipn.stan:144): its delay column is identically zero, so a 1-point
        // grid would do; it is kept on the common path (its cost is 1/D of the call) to share one code path.
        rc = run_loglik_grid(ctx, nbins[d], d_time + d * Tmax, d_exposure + d * Tmax, d_counts + d * Tmax, S, J, d_omega, d_a, d_b,
                             d_amp + d * S, d_bkg + d * S, P, ctx->io[11].as<double>() + d * P, d_ll, d != 0);
        if (rc) break;
    }
    ctx->n_terms = 1;
    return rc;
}

int ipn_loglik_sky_grid(ipn_ctx* ctx, int64_t D, const int64_t* nbins, int64_t Tmax, const double* time,
                        const double* exposure, const int64_t* counts, const double* sc_pos, int64_t P,
                        const double* ghat, int64_t S, int64_t J, const double* omega, const double* a,
                        const double* b, const double* amp, const double* bkg, double* ll) {
    int rc = ctx_guard(ctx);
    if (rc) return rc;
    IPN_REQUIRE(D >= 1 && Tmax >= 0, "need D >= 1 and Tmax >= 0");
    rc = check_grid_shapes(Tmax, S, J, P);
    if (rc) return rc;
    if (S == 0 || P == 0) return IPN_OK;
    IPN_REQUIRE(nbins && sc_pos && ghat && omega && a && b && amp && bkg && ll && (Tmax == 0 || (time && exposure && counts)),
                "NULL array argument");
    IPN_REQUIRE(all_finite(sc_pos, D * 3) && all_finite(ghat, P * 3), "sc_pos / ghat contain a non-finite value");
    for (int64_t d = 0; d < D; ++d) {
        IPN_REQUIRE(nbins[d] >= 0 && nbins[d] <= Tmax, "nbins[%lld] = %lld outside 0..Tmax", (long long)d, (long long)nbins[d]);
        rc = check_grid_host_values(nbins[d], time + d * Tmax, exposure + d * Tmax, counts + d * Tmax, S, J, omega, a, b,
                                    amp + d * S, bkg + d * S);
        if (rc) return rc;
    }
    const size_t sj = (size_t)S * J * 8, dt = (size_t)D * Tmax * 8;
    if ((rc = h2d(ctx, ctx->io[0], time, dt))) return rc;
    if ((rc = h2d(ctx, ctx->io[1], exposure, dt))) return rc;
    if ((rc = h2d(ctx, ctx->io[2], counts, dt))) return rc;
    if ((rc = h2d(ctx, ctx->io[3], omega, sj))) return rc;
    if ((rc = h2d(ctx, ctx->io[4], a, sj))) return rc;
    if ((rc = h2d(ctx, ctx->io[5], b, sj))) return rc;
    if ((rc = h2d(ctx, ctx->io[6], amp, (size_t)D * S * 8))) return rc;
    if ((rc = h2d(ctx, ctx->io[7], bkg, (size_t)D * S * 8))) return rc;
    if ((rc = h2d(ctx, ctx->io[8], sc_pos, (size_t)D * 3 * 8))) return rc;
    if ((rc = h2d(ctx, ctx->io[10], ghat, (size_t)P * 3 * 8))) return rc;
    if ((rc = ctx->io[9].ensure((size_t)P * S * 8))) return rc;
    begin_timing(ctx);
    rc = sky_grid_device(ctx, D, nbins, Tmax, ctx->io[0].as<double>(), ctx->io[1].as<double>(), ctx->io[2].as<int64_t>(),
                         ctx->io[8].as<double>(), P, ctx->io[10].as<double>(), S, J, ctx->io[3].as<double>(), ctx->io[4].as<double>(),
                         ctx->io[5].as<double>(), ctx->io[6].as<double>(), ctx->io[7].as<double>(), ctx->io[9].as<double>());
    if (rc) return rc;
    end_timing(ctx);
    IPN_CUDA_CHECK(cudaMemcpyAsync(ll, ctx->io[9].p, (size_t)P * S * 8, cudaMemcpyDeviceToHost, ctx->stream));
    IPN_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    finish_timing(ctx);
    return IPN_OK;
}

int ipn_loglik_sky_grid_dev(ipn_ctx* ctx, int64_t D, const int64_t* nbins, int64_t Tmax, const double* d_time,
                            const double* d_exposure, const int64_t* d_counts, const double* d_sc_pos, int64_t P,
                            const double* d_ghat, int64_t S, int64_t J, const double* d_omega, const double* d_a,
                            const double* d_b, const double* d_amp, const double* d_bkg, double* d_ll) {
    int rc = ctx_guard(ctx);
    if (rc) return rc;
    IPN_REQUIRE(D >= 1 && Tmax >= 0, "need D >= 1 and Tmax >= 0");
    rc = check_grid_shapes(Tmax, S, J, P);
    if (rc) return rc;
    if (S == 0 || P == 0) return IPN_OK;
    IPN_REQUIRE(nbins && d_sc_pos && d_ghat && d_omega && d_a && d_b && d_amp && d_bkg && d_ll &&
                    (Tmax == 0 || (d_time && d_exposure && d_counts)),
                "NULL array argument");
    for (int64_t d = 0; d < D; ++d)
        IPN_REQUIRE(nbins[d] >= 0 && nbins[d] <= Tmax, "nbins[%lld] = %lld outside 0..Tmax", (long long)d, (long long)nbins[d]);
    begin_timing(ctx);
    rc = sky_grid_device(ctx, D, nbins, Tmax, d_time, d_exposure, d_counts, d_sc_pos, P, d_ghat, S, J, d_omega, d_a, d_b, d_amp,
                         d_bkg, d_ll);
    end_timing(ctx);
    return rc;
}

// ---------------------------------------------------------------------------------------------------
int ipn_poisson_counts_dev(ipn_ctx* ctx, int64_t T, const double* exposure, int64_t S, const double* rates,
                           const double* bkg, uint64_t seed, int64_t* out) {
    int rc = ctx_guard(ctx);
    if (rc) return rc;
    IPN_REQUIRE(T >= 0 && S >= 0, "need T >= 0 and S >= 0");
    if (T == 0 || S == 0) return IPN_OK;
    IPN_REQUIRE(exposure && rates && bkg && out, "NULL array argument");
    begin_timing(ctx);
    rc = launch_poisson_counts(ctx, T, exposure, S, rates, bkg, seed, out);
    end_timing(ctx);
    return rc;
}

int ipn_poisson_counts(ipn_ctx* ctx, int64_t T, const double* exposure, int64_t S, const double* rates,
                       const double* bkg, uint64_t seed, int64_t* out) {
    int rc = ctx_guard(ctx);
    if (rc) return rc;
    IPN_REQUIRE(T >= 0 && S >= 0, "need T >= 0 and S >= 0");
    if (T == 0 || S == 0) return IPN_OK;
    IPN_REQUIRE(exposure && rates && bkg && out, "NULL array argument");
    if ((rc = h2d(ctx, ctx->io[0], exposure, (size_t)T * 8))) return rc;
    if ((rc = h2d(ctx, ctx->io[1], rates, (size_t)S * T * 8))) return rc;
    if ((rc = h2d(ctx, ctx->io[2], bkg, (size_t)S * 8))) return rc;
    if ((rc = ctx->io[3].ensure((size_t)S * T * 8))) return rc;
    begin_timing(ctx);
    rc = launch_poisson_counts(ctx, T, ctx->io[0].as<double>(), S, ctx->io[1].as<double>(), ctx->io[2].as<double>(), seed,
                               ctx->io[3].as<int64_t>());
    end_timing(ctx);
    if (rc) return rc;
    IPN_CUDA_CHECK(cudaMemcpyAsync(out, ctx->io[3].p, (size_t)S * T * 8, cudaMemcpyDeviceToHost, ctx->stream));
    IPN_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    finish_timing(ctx);
    return IPN_OK;
}

int ipn_bin_events_dev(ipn_ctx* ctx, int64_t n_events, const double* times, double tstart, double dt, int64_t n_edges,
                       int64_t* counts) {
    int rc = ctx_guard(ctx);
    if (rc) return rc;
    IPN_REQUIRE(n_events >= 0 && n_edges >= 2 && dt > 0.0 && isfinite(tstart), "need n_events >= 0, n_edges >= 2, dt > 0");
    IPN_REQUIRE(counts && (n_events == 0 || times), "NULL array argument");
    return launch_bin_events(ctx, n_events, times, tstart, dt, n_edges, counts);
}

int ipn_bin_events(ipn_ctx* ctx, int64_t n_events, const double* times, double tstart, double dt, int64_t n_edges,
                   int64_t* counts) {
    int rc = ctx_guard(ctx);
    if (rc) return rc;
    IPN_REQUIRE(n_events >= 0 && n_edges >= 2 && dt > 0.0 && isfinite(tstart), "need n_events >= 0, n_edges >= 2, dt > 0");
    IPN_REQUIRE(counts && (n_events == 0 || times), "NULL array argument");
    if ((rc = h2d(ctx, ctx->io[0], times, (size_t)n_events * 8))) return rc;
    if ((rc = ctx->io[1].ensure((size_t)(n_edges - 1) * 8))) return rc;
    begin_timing(ctx);
    rc = launch_bin_events(ctx, n_events, ctx->io[0].as<double>(), tstart, dt, n_edges, ctx->io[1].as<int64_t>());
    end_timing(ctx);
    if (rc) return rc;
    IPN_CUDA_CHECK(cudaMemcpyAsync(counts, ctx->io[1].p, (size_t)(n_edges - 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    IPN_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    finish_timing(ctx);
    return IPN_OK;
}

int ipn_correlate(ipn_ctx* ctx, int64_t n1, const double* t1, const double* c1, const double* e1, int64_t n2,
                  const double* t2, const double* c2, const double* e2, int64_t i_beg_2, int64_t i_end_2, int64_t i_beg_1,
                  int64_t n_max_1, double fscale, double dt_1, double dt_2, int64_t n_sum_2, double* arr_dt,
                  double* arr_chi) {
    int rc = ctx_guard(ctx);
    if (rc) return rc;
    IPN_REQUIRE(n1 > 0 && n2 > 0 && t1 && c1 && e1 && t2 && c2 && e2, "empty or NULL light curve");
    IPN_REQUIRE(n_max_1 >= 0 && n_sum_2 >= 1, "need n_max_1 >= 0 and n_sum_2 >= 1");
    IPN_REQUIRE(0 <= i_beg_2 && i_beg_2 <= i_end_2 && i_end_2 + n_sum_2 - 1 < n2, "light curve 2 window [%lld, %lld] outside 0..%lld",
                (long long)i_beg_2, (long long)i_end_2, (long long)n2 - 1);
    IPN_REQUIRE(0 <= i_beg_1 && i_beg_1 + n_max_1 <= n1, "shift range outside light curve 1");
    IPN_REQUIRE(dt_1 > 0.0 && dt_2 > 0.0 && isfinite(fscale), "need dt_1, dt_2 > 0 and a finite scale");
    if (n_max_1 == 0) return IPN_OK;
    IPN_REQUIRE(arr_dt && arr_chi, "NULL output array");
    if ((rc = h2d(ctx, ctx->io[0], t1, (size_t)n1 * 8))) return rc;
    if ((rc = h2d(ctx, ctx->io[1], c1, (size_t)n1 * 8))) return rc;
    if ((rc = h2d(ctx, ctx->io[2], e1, (size_t)n1 * 8))) return rc;
    if ((rc = h2d(ctx, ctx->io[3], t2, (size_t)n2 * 8))) return rc;
    // light curve 2 stays on the host: its per-bin sums go into the shift-independent plan (correlate.cu)
    if ((rc = ctx->io[6].ensure((size_t)n_max_1 * 8))) return rc;
    if ((rc = ctx->io[7].ensure((size_t)n_max_1 * 8))) return rc;
    begin_timing(ctx);
    rc = launch_correlate(ctx, ctx->io[0].as<double>(), ctx->io[1].as<double>(), ctx->io[2].as<double>(), ctx->io[3].as<double>(),
                          c2, e2, n1, i_beg_2, i_end_2, i_beg_1, n_max_1, fscale, dt_1,
                          dt_2, n_sum_2, ctx->io[6].as<double>(), ctx->io[7].as<double>());
    end_timing(ctx);
    if (rc) return rc;
    IPN_CUDA_CHECK(cudaMemcpyAsync(arr_dt, ctx->io[6].p, (size_t)n_max_1 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    IPN_CUDA_CHECK(cudaMemcpyAsync(arr_chi, ctx->io[7].p, (size_t)n_max_1 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    IPN_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    finish_timing(ctx);
    return IPN_OK;
}

int ipn_microbench(ipn_ctx* ctx, int which, double* result) {
    int rc = ctx_guard(ctx);
    if (rc) return rc;
    IPN_REQUIRE(result != nullptr, "result is NULL");
    return run_microbench(ctx, which, result);
}

}  // extern "C"
