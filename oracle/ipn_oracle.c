/* CPU restatement, in plain C, of the reference arithmetic on the hot path.  TEST / BENCH INFRASTRUCTURE ONLY: nothing
 * under pyipn_b200/ may load this; only tests/, __graft_entry__.smoke() and bench.py's CPU legs do (through
 * oracle/c_oracle.py).  It follows the same reference lines as oracle/ipn_oracle.py and is checked against that numpy
 * restatement and against the golden vectors minted from the reference's own pyipn/rff.py (tests/test_oracle.py).
 *
 *   orc_rff_multiscale    pyipn/rff.py:19-29        exp((s1 cos F1 + s2 cos F2) b1 + (s1 sin F1 + s2 sin F2) b2),
 *                                                    F_i = bw * time (x) omega_i     (RFF, rff.py:5-16, is s1 = s2 = scale)
 *   orc_expected_rate     pyipn/fit.py:651-691      per draw n: amplitude[n] * RFF*(time - dt[n], column n of (k, S) arrays)
 *   orc_loglik_folded     functions.stan:23-35      poisson_lpmf(counts | exposure .* (exp(logf) * A + B)) on the folded form
 *                                                    logf = sum_j a_j cos(w_j (t - dt)) + b_j sin(w_j (t - dt))
 *   orc_loglik_delay_grid SURVEY.md section 3E      ll[g * S + s] over delays x draws, direct evaluation, one pair at a time
 *   orc_loglik_sky_grid   functions.stan:37-61,164-169 + rff_omega_ipn.stan:63-68,129-134,144-145
 *
 * Scalar libm calls in evaluation order of the reference (numba compiles rff.py to the same scalar libm calls); the
 * draw / delay loops are the only parallelism (pthreads, one (delay, draw) pair per work item; this image has no OpenMP
 * runtime), as in the reference where more than one core means more than one process.
 *
 * Build: gcc -O2 -fPIC -shared -pthread -o oracle/_build/libipn_oracle.so oracle/ipn_oracle.c -lm   (oracle/Makefile) */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

#define ORC_C_KM_S 299792.46 /* functions.stan:65-69 */

static int g_threads = 0; /* 0 = all online cores */

int orc_max_threads(void) {
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
void orc_set_threads(int n) { g_threads = n; }

/* parallel for over [0, n): threads take the next index from a shared counter */
typedef void (*orc_body)(int64_t i, void* arg);
typedef struct {
    int64_t n;
    int64_t next;
    orc_body body;
    void* arg;
} orc_job;
static void* orc_worker(void* p) {
    orc_job* j = (orc_job*)p;
    for (;;) {
        const int64_t i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (i >= j->n) break;
        j->body(i, j->arg);
    }
    return 0;
}
static void orc_parallel_for(int64_t n, orc_body body, void* arg) {
    orc_job job = {n, 0, body, arg};
    int nt = orc_max_threads();
    if ((int64_t)nt > n) nt = (int)n;
    if (nt <= 1) {
        orc_worker(&job);
        return;
    }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nt);
    int started = 0;
    for (int t = 0; t < nt - 1; ++t)
        if (pthread_create(&th[started], 0, orc_worker, &job) == 0) ++started;
    orc_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(th[t], 0);
    free(th);
}

/* rff.py:19-29.  omega1/omega2/beta1/beta2 are read with a stride (column n of a (k, S) C-order array: stride S). */
void orc_rff_multiscale(int64_t T, const double* time, double shift, int64_t k, const double* omega1, const double* omega2,
                        const double* beta1, const double* beta2, int64_t stride, double bw, double scale1, double scale2,
                        double amplitude, double* out) {
    for (int64_t i = 0; i < T; ++i) {
        const double t = (time[i] - shift) * bw;
        double acc = 0.0;
        for (int64_t j = 0; j < k; ++j) {
            const double f1 = t * omega1[j * stride], f2 = t * omega2[j * stride];
            acc += (scale1 * cos(f1) + scale2 * cos(f2)) * beta1[j * stride] + (scale1 * sin(f1) + scale2 * sin(f2)) * beta2[j * stride];
        }
        out[i] = amplitude * exp(acc);
    }
}

typedef struct {
    int64_t T; const double* time; int64_t k, S; const double *omega1, *omega2, *beta1, *beta2, *bw, *scale; int multiscale;
    const double *amplitude, *dt; double* out;
} er_args;
static void er_body(int64_t n, void* p) {
    const er_args* A = (const er_args*)p;
    const double s1 = A->scale[n], s2 = A->multiscale ? A->scale[A->S + n] : A->scale[n];
    orc_rff_multiscale(A->T, A->time, A->dt[n], A->k, A->omega1 + n, A->omega2 + n, A->beta1 + n, A->beta2 + n, A->S, A->bw[n],
                       s1, s2, A->amplitude[n], A->out + n * A->T);
}

/* fit.py:651-691: out (N, T).  scale is (S) [multiscale == 0] or (2, S) [multiscale != 0]; omega*, beta* are (k, S). */
void orc_expected_rate(int64_t T, const double* time, int64_t k, int64_t S, const double* omega1, const double* omega2,
                       const double* beta1, const double* beta2, const double* bw, const double* scale, int multiscale,
                       const double* amplitude, const double* dt, int64_t N, double* out) {
    er_args A = {T, time, k, S, omega1, omega2, beta1, beta2, bw, scale, multiscale, amplitude, dt, out};
    orc_parallel_for(N, er_body, &A);
}

/* functions.stan:23-35 on the folded parameters; Stan's poisson_lpmf: n log(mu) - mu - lgamma(n + 1), with
 * mu == 0 and n == 0 -> 0, mu == 0 and n > 0 -> -inf. */
double orc_loglik_folded(int64_t T, const double* time, const double* exposure, const int64_t* counts, int64_t J,
                         const double* omega, const double* a, const double* b, double dt, double bkg, double amplitude) {
    double lp = 0.0, lg = 0.0;
    for (int64_t i = 0; i < T; ++i) {
        const double tau = time[i] - dt;
        double logf = 0.0;
        for (int64_t j = 0; j < J; ++j) {
            const double ph = tau * omega[j];
            logf += a[j] * cos(ph) + b[j] * sin(ph);
        }
        const double mu = exposure[i] * (exp(logf) * amplitude + bkg);
        const double n = (double)counts[i];
        lp += (counts[i] > 0 ? n * log(mu) : 0.0) - mu;
        lg += lgamma(n + 1.0);
    }
    return lp - lg;
}

typedef struct {
    int64_t T; const double *time, *exposure; const int64_t* counts; int64_t S, J; const double *omega, *a, *b, *amp, *bkg, *delays;
    double* ll;
} dg_args;
static void dg_body(int64_t gs, void* p) {
    const dg_args* A = (const dg_args*)p;
    const int64_t g = gs / A->S, s = gs - g * A->S;
    A->ll[gs] = orc_loglik_folded(A->T, A->time, A->exposure, A->counts, A->J, A->omega + s * A->J, A->a + s * A->J,
                                  A->b + s * A->J, A->delays[g], A->bkg[s], A->amp[s]);
}

/* ll[g * S + s]; omega/a/b are (S, J) row-major. */
void orc_loglik_delay_grid(int64_t T, const double* time, const double* exposure, const int64_t* counts, int64_t S, int64_t J,
                           const double* omega, const double* a, const double* b, const double* amp, const double* bkg,
                           int64_t G, const double* delays, double* ll) {
    dg_args A = {T, time, exposure, counts, S, J, omega, a, b, amp, bkg, delays, ll};
    orc_parallel_for(G * S, dg_body, &A);
}

typedef struct {
    int64_t D; const int64_t* nbins; int64_t Tmax; const double *time, *exposure; const int64_t* counts; const double *sc_pos, *ghat;
    int64_t S, J; const double *omega, *a, *b, *amp, *bkg; double* ll;
} sg_args;
static void sg_body(int64_t ps, void* q) {
    const sg_args* A = (const sg_args*)q;
    const int64_t p = ps / A->S, s = ps - p * A->S;
    double acc = 0.0;
    for (int64_t d = 0; d < A->D; ++d) {
        double dl = 0.0;
        for (int c = 0; c < 3; ++c) dl += A->ghat[p * 3 + c] * (A->sc_pos[c] - A->sc_pos[d * 3 + c]);
        dl *= 1.0 / ORC_C_KM_S;
        acc += orc_loglik_folded(A->nbins[d], A->time + d * A->Tmax, A->exposure + d * A->Tmax, A->counts + d * A->Tmax, A->J,
                                 A->omega + s * A->J, A->a + s * A->J, A->b + s * A->J, dl, A->bkg[d * A->S + s],
                                 A->amp[d * A->S + s]);
    }
    A->ll[ps] = acc;
}

/* ll[p * S + s] = sum_d loglik(detector d shifted by ghat_p . (p_0 - p_d) / c); time/exposure/counts padded (D, Tmax),
 * amp/bkg (D, S). */
void orc_loglik_sky_grid(int64_t D, const int64_t* nbins, int64_t Tmax, const double* time, const double* exposure,
                         const int64_t* counts, const double* sc_pos, int64_t P, const double* ghat, int64_t S, int64_t J,
                         const double* omega, const double* a, const double* b, const double* amp, const double* bkg,
                         double* ll) {
    sg_args A = {D, nbins, Tmax, time, exposure, counts, sc_pos, ghat, S, J, omega, a, b, amp, bkg, ll};
    orc_parallel_for(P * S, sg_body, &A);
}
