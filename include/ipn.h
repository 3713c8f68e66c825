/* libipn_b200 -- C ABI of the B200 (sm_100a) implementation of pyipn's RFF-intensity -> binned-Poisson
 * log-likelihood hot path.
 *
 * The reference (grburgess/pyipn) has no FFI layer: the boundary this header replaces is the set of Python
 * call signatures listed below (paths relative to the reference checkout).  The Python facade in
 * pyipn_b200/ binds these entry points with ctypes and keeps the reference's function names/argument
 * meaning; INTEGRATION.md shows the ctypes stub a pyipn maintainer would add.
 *
 * Conventions
 *  - every function returns 0 (IPN_OK) on success and a negative IPN_ERR_* code on failure; the message of
 *    the last failure on the calling thread is returned by ipn_last_error().
 *  - there is NO CPU fallback: without a usable sm_100 device ipn_ctx_create fails with IPN_ERR_NO_DEVICE.
 *  - "host" entry points take caller-owned host buffers (pageable or pinned), copy in, compute, copy out
 *    and return after the result is in the output buffer.  "_dev" entry points take device pointers that
 *    are already resident in HBM, enqueue on the context's stream and return without synchronising.
 *  - all arrays are dense, C-order, float64 / int64 exactly as numpy hands them over.
 *  - canonical folded form of one posterior draw (folding is done on the host in float64, see
 *    pyipn_b200/params.py):   log f(tau) = sum_{j<J} a_j cos(omega_j tau) + b_j sin(omega_j tau),  J = 2k,
 *        omega = [bw*omega1 ; bw*omega2],  a = [s1*beta1 ; s2*beta1],  b = [s1*beta2 ; s2*beta2]
 *    (pyipn/rff.py:8-14 single scale s1 = s2 = scale; pyipn/rff.py:22-27 multiscale;
 *     pyipn/stan_models/functions.stan:27-33).
 *  - one host thread per context at a time; contexts are independent (one per GPU / per process rank).
 */
#ifndef IPN_B200_H
#define IPN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPN_OK 0
#define IPN_ERR_INVALID_ARGUMENT (-1) /* NULL pointer, non-positive size, non-finite / out-of-range value     */
#define IPN_ERR_NO_DEVICE (-2)        /* no CUDA device, or the device is not compute capability 10.x         */
#define IPN_ERR_CUDA (-3)             /* a CUDA runtime call or a kernel failed; see ipn_last_error()          */
#define IPN_ERR_ALLOC (-4)            /* device or pinned-host allocation failed                               */
#define IPN_ERR_UNSUPPORTED (-5)      /* shape outside what the kernels are built for (e.g. J > IPN_MAX_J)     */

#define IPN_MAX_J 512 /* max number of folded frequencies per draw (J = 2k; the reference uses k = 50)      */

/* ipn_ctx_set_option keys */
#define IPN_OPT_LOGLIK_IMPL 1 /* 0 = auto (default), 1 = FP32-pipe contraction, 2 = tcgen05 sliced-integer contraction */
#define IPN_OPT_PRECISE_W 2   /* FP32-pipe path only: 1 = hi/lo split delay table (default), 0 = single fp32 table */
#define IPN_OPT_DRAW_BATCH 3  /* max draws per internal batch (bounds the delay-table scratch); 0 = auto     */
#define IPN_OPT_HIFI 4        /* tcgen05 path epilogue: 0 = auto (per draw: plain fp32; compensated fp32 when the fp32 error model
                                 predicts more than ~5e-5 nats -- n_terms * sum n_i^2 > 2e5 --; fp64 when > 4e7 or when the draw's
                                 exponent is not bounded by 2^120), 1 = fp64, 2 = plain fp32, 3 = compensated fp32             */

#define IPN_OPT_RFF_IMPL 5    /* ipn_rff_eval: 0 = auto (default: the tcgen05 contraction when the bins are uniform -- time[i] =
                                 time[0] + i dt, what np.arange edges give --, T >= 1024 and J <= 108, else the direct kernel; the
                                 _dev entry point checks the spacing on the device, one 24-byte read-back), 1 = direct kernel,
                                 2 = contraction without the check (the caller guarantees uniform bins)                      */

typedef struct ipn_ctx ipn_ctx;

/* ---- context -------------------------------------------------------------------------------------- */
int ipn_ctx_create(int device, ipn_ctx** out);
void ipn_ctx_destroy(ipn_ctx* ctx);
/* borrow an externally owned cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream).  The handle is taken
 * literally: NULL is CUDA's legacy default stream (torch's current stream outside a torch.cuda.stream(...) block).
 * A new context runs on a stream of its own (a blocking stream: implicitly ordered with the legacy default stream);
 * ipn_ctx_use_own_stream returns to it. */
int ipn_ctx_set_stream(ipn_ctx* ctx, void* cuda_stream);
int ipn_ctx_use_own_stream(ipn_ctx* ctx);
int ipn_ctx_synchronize(ipn_ctx* ctx);
int ipn_ctx_set_option(ipn_ctx* ctx, int key, int64_t value);
/* number of kernels of THIS library launched through the context so far (bench.py's gpu_launches)     */
uint64_t ipn_ctx_launch_count(const ipn_ctx* ctx);
/* CUDA-event time in ms spent between the first and last kernel of the most recent compute call        */
double ipn_ctx_last_kernel_ms(const ipn_ctx* ctx);
/* summed CUDA-event time of the dominant (contraction) kernel launches of that call, and how many there were  */
double ipn_ctx_last_main_kernel_ms(const ipn_ctx* ctx, int64_t* launches);
/* which contraction back-end the most recent log-likelihood call used: 1 = FP32 pipe, 2 = tcgen05 (0 = none yet) */
int ipn_ctx_last_loglik_impl(const ipn_ctx* ctx);
/* which kernel the most recent ipn_rff_eval call ran: 1 = direct, 2 = tcgen05 contraction on uniform bins (0 = none yet) */
int ipn_ctx_last_rff_impl(const ipn_ctx* ctx);
/* draws per internal batch of the most recent log-likelihood call (the operand-image scratch is bounded, so a call with many
 * draws is processed in batches; bench.py spreads its oracle spot check over every batch) */
int64_t ipn_ctx_last_draw_batch(const ipn_ctx* ctx);
const char* ipn_last_error(void);
const char* ipn_version(void);

/* ---- intensity: drop-in for RFF / RFF_multiscale / _expeced_rate / _expeced_rate_multiscale ----------
 * replaces  pyipn/rff.py:5-16, 19-29   and the draw loop  pyipn/fit.py:651-668, 671-691
 *   out[s*T + i] = amp[s] * exp( sum_j a[s,j] cos(omega[s,j] (time[i]-shift[s])) + b[s,j] sin(...) )
 * time (T), omega/a/b (S x J), shift/amp (S), out (S x T).  Relative error <= 1e-6 vs the reference. */
int ipn_rff_eval(ipn_ctx* ctx, int64_t T, const double* time, int64_t S, int64_t J, const double* omega,
                 const double* a, const double* b, const double* shift, const double* amp, double* out);
int ipn_rff_eval_dev(ipn_ctx* ctx, int64_t T, const double* d_time, int64_t S, int64_t J, const double* d_omega,
                     const double* d_a, const double* d_b, const double* d_shift, const double* d_amp,
                     double* d_out);

/* ---- delay-grid log-likelihood ---------------------------------------------------------------------
 * replaces  pyipn/stan_models/functions.stan:23-35 (partial_log_like_bw_multi_scale_fast) evaluated over a
 * grid of G trial delays x S posterior draws:
 *   ll[g*S + s] = sum_i poisson_lpmf( counts[i] | exposure[i] * (amp[s] * f_s(time[i] - delays[g]) + bkg[s]) )
 * including the -lgamma(n+1) term.  counts >= 0, exposure >= 0, bkg > 0, amp >= 0.
 * Absolute error <= 1e-3 nats vs the float64 reference composition.  ll is G x S (draw fastest). */
int ipn_loglik_delay_grid(ipn_ctx* ctx, int64_t T, const double* time, const double* exposure,
                          const int64_t* counts, int64_t S, int64_t J, const double* omega, const double* a,
                          const double* b, const double* amp, const double* bkg, int64_t G,
                          const double* delays, double* ll);
int ipn_loglik_delay_grid_dev(ipn_ctx* ctx, int64_t T, const double* d_time, const double* d_exposure,
                              const int64_t* d_counts, int64_t S, int64_t J, const double* d_omega,
                              const double* d_a, const double* d_b, const double* d_amp, const double* d_bkg,
                              int64_t G, const double* d_delays, double* d_ll);

/* ---- sky-grid log-likelihood ------------------------------------------------------------------------
 * replaces  partial_total_like  (functions.stan:37-61)  +  time_delay (functions.stan:164-169, c = 299792.46
 * km/s, functions.stan:65-69)  as wired in rff_omega_ipn.stan:63-68, 129-134, 144-145, over P sky directions:
 *   delay[p,d] = ghat[p] . (sc_pos[0] - sc_pos[d]) / c            (detector 0 unshifted)
 *   ll[p*S + s] = sum_d sum_{i < nbins[d]} poisson_lpmf(counts[d,i] | exposure[d,i]*(amp[d,s]*f_s(time[d,i]-delay[p,d]) + bkg[d,s]))
 * time/exposure/counts are padded D x Tmax (universe.py:441-529 layout), nbins (D), sc_pos (D x 3, km),
 * ghat (P x 3 unit vectors), amp/bkg (D x S; the Stan model fixes amp[0,:] = 1), ll (P x S). */
int ipn_loglik_sky_grid(ipn_ctx* ctx, int64_t D, const int64_t* nbins, int64_t Tmax, const double* time,
                        const double* exposure, const int64_t* counts, const double* sc_pos, int64_t P,
                        const double* ghat, int64_t S, int64_t J, const double* omega, const double* a,
                        const double* b, const double* amp, const double* bkg, double* ll);
/* the same on device-resident arrays (only nbins, D small integers, stays a host array); d_ll (P x S) is overwritten */
int ipn_loglik_sky_grid_dev(ipn_ctx* ctx, int64_t D, const int64_t* nbins, int64_t Tmax, const double* d_time,
                            const double* d_exposure, const int64_t* d_counts, const double* d_sc_pos, int64_t P,
                            const double* d_ghat, int64_t S, int64_t J, const double* d_omega, const double* d_a,
                            const double* d_b, const double* d_amp, const double* d_bkg, double* d_ll);

/* ---- posterior-predictive counts: drop-in for ppc_generator ------------------------------------------
 * replaces  pyipn/fit.py:694-705:  out[s*T + i] ~ Poisson((rates[s*T+i] + bkg[s]) * exposure[i]).
 * Counter-based Philox stream keyed by (seed, s, i): statistical, not stream, parity with numba's MT19937. */
int ipn_poisson_counts(ipn_ctx* ctx, int64_t T, const double* exposure, int64_t S, const double* rates,
                       const double* bkg, uint64_t seed, int64_t* out);
int ipn_poisson_counts_dev(ipn_ctx* ctx, int64_t T, const double* d_exposure, int64_t S, const double* d_rates,
                           const double* d_bkg, uint64_t seed, int64_t* d_out);

/* ---- photon arrival times by thinning (next-row N2) ---------------------------------------------------
 * replaces  source_poisson_generator / background_poisson_generator  pyipn/possion_gen.py:159-242 with the pulse shape
 * of :9-62 (sum of up to 8 Norris pulses K exp(2 sqrt(tr/td)) exp(-tr/(t-ts) - (t-ts)/td)) over a linear background
 * rate bkg_intercept + bkg_slope t (pass n_pulses = 0 for the background generator, 0/0 for a pure source).
 * out_times receives tstart followed by the sorted accepted times (the reference emits tstart itself first);
 * *out_count is the number written (or needed, if the call fails because capacity is too small).  Counter-based
 * Philox streams keyed by (seed, candidate index): the same point process as the reference, not the same draws. */
int ipn_thinning_events(ipn_ctx* ctx, double tstart, double tstop, int64_t n_pulses, const double* K,
                        const double* t_start, const double* t_rise, const double* t_decay, double bkg_slope,
                        double bkg_intercept, uint64_t seed, int64_t capacity, double* out_times, int64_t* out_count);

/* ---- event binning (next-row N3) ------------------------------------------------------------------------
 * replaces  LightCurve.get_binned_light_curve  pyipn/lightcurve.py:24-47 (np.arange edges + np.histogram):
 * the caller passes n_edges = len(np.arange(tstart, tstop, dt)) and dt = edges[1] - edges[0] (numpy fills a float arange
 * as start + i*delta with that delta); counts receives n_edges - 1 integers, bit-identical to np.histogram. */
int ipn_bin_events(ipn_ctx* ctx, int64_t n_events, const double* times, double tstart, double dt, int64_t n_edges,
                   int64_t* counts);
int ipn_bin_events_dev(ipn_ctx* ctx, int64_t n_events, const double* d_times, double tstart, double dt, int64_t n_edges,
                       int64_t* d_counts);

/* ---- classical chi^2 cross-correlation over integer shifts (next-row N1) ---------------------------------
 * replaces  correlate(...)  pyipn/correlation.py:247-372 (same argument meaning and order; dt_1, dt_2 in ms):
 * light curve 1 (n1 bins: bin starts t1, background-subtracted counts c1, raw counts e1 used as variances) is
 * shifted by i = i_beg_1 .. i_beg_1 + n_max_1 - 1 bins and re-binned onto bins i_beg_2..i_end_2 of light curve 2
 * (n2 bins) scaled by fscale.  arr_dt, arr_chi receive n_max_1 values; the arg-min scan is left to the caller. */
int ipn_correlate(ipn_ctx* ctx, int64_t n1, const double* t1, const double* c1, const double* e1, int64_t n2,
                  const double* t2, const double* c2, const double* e2, int64_t i_beg_2, int64_t i_end_2,
                  int64_t i_beg_1, int64_t n_max_1, double fscale, double dt_1, double dt_2, int64_t n_sum_2,
                  double* arr_dt, double* arr_chi);

/* ---- sky post-processing: credible regions on the pixel grid (next-row N4) ------------------------------
 * replaces the KDE -> multi-order-map -> MOC route  pyipn/fit.py:193-206 (_build_moc_map) and :328-347
 * (_contour_more_detectors: MOC.from_valued_healpix_cells(uniq, prob, cumul_to=level)) for likelihoods that are
 * already on a pixel grid (ipn_loglik_sky_grid):
 *   L[p] = logsumexp_s ll[p*S + s] - log S + log_weight[p]   (log_weight may be NULL: equal-area pixels, flat prior)
 *   prob[p] = exp(L[p] - *log_norm),  *log_norm = logsumexp_p L[p]
 *   order   = pixel indices by decreasing prob (stable: ties keep increasing pixel index)
 *   level[p] = cumulative prob in that order up to and including p; region(c) = {p : level[p] - prob[p] < c}
 * ll must not contain NaN or +inf and must have at least one finite entry (host entry point checks). */
int ipn_sky_credible(ipn_ctx* ctx, int64_t P, int64_t S, const double* ll, const double* log_weight, double* prob,
                     double* level, int64_t* order, double* log_norm);
int ipn_sky_credible_dev(ipn_ctx* ctx, int64_t P, int64_t S, const double* d_ll, const double* d_log_weight,
                         double* d_prob, double* d_level, int64_t* d_order, double* d_log_norm);

/* ---- device micro-benchmarks used by bench.py for the roofline denominators ------------------------
 * which: 0 = FP32 FMA loop (returns TFLOP/s), 1 = MUFU.EX2 loop (returns Tops/s), 2 = tcgen05.mma kind::i8 issue loop
 * (M = 128, N = 32, K = 32, A from tensor memory -- the hot kernel's MMA shape; returns int8 TOP/s, 2 ops per MAC). */
int ipn_microbench(ipn_ctx* ctx, int which, double* result);

#ifdef __cplusplus
}
#endif
#endif /* IPN_B200_H */
