/*
 * oracle/xcorr_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * fp64 CPU restatement of the reference hot path (see xcorr_oracle.h for the
 * pinning story).  Each function names the reference lines it follows; the
 * FFTs go through the fftw3.h stand-in because FFTW3 is absent here.
 */
#define _GNU_SOURCE
#include "xcorr_oracle.h"
#include "fftw3.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* reference src/cross_correlation.c:52-67.
 * The running maximum starts from the SIGNED arr[0]; later entries compete
 * with their magnitude and must be strictly larger, so ties keep the earliest
 * index and a NaN never replaces the incumbent. */
size_t oracle_max_abs_index(const double *arr, size_t len)
{
    size_t best = 0;
    double best_val = arr[0];
    for (size_t i = 1; i < len; i++) {
        double mag = fabs(arr[i]);
        if (mag > best_val) {
            best_val = mag;
            best = i;
        }
    }
    return best;
}

/* reference src/cross_correlation.c:74-116: two passes, plain left-to-right
 * accumulation: means first (:82-95), then the centred products (:98-113),
 * then cov / sqrt(varx * vary) (:115).  n == 0 gives 0/0 = NaN like the
 * reference's empty pointer range. */
double oracle_pearson(const double *x, const double *y, size_t n)
{
    double sx = 0.0, sy = 0.0;
    for (size_t i = 0; i < n; i++) {
        sx += x[i];
        sy += y[i];
    }
    double mx = sx / (double)(long)n;
    double my = sy / (double)(long)n;
    double sxy = 0.0, sxx = 0.0, syy = 0.0;
    for (size_t i = 0; i < n; i++) {
        double dx = x[i] - mx;
        double dy = y[i] - my;
        sxy += dx * dy;
        sxx += dx * dx;
        syy += dy * dy;
    }
    return sxy / sqrt(sxx * syy);
}

/* reference src/cross_correlation.c:133-307 */
int oracle_cross_correlation(const double *source, const double *sample,
                             size_t sample_len, long *lag, double *coefficient,
                             struct oracle_extra *extra)
{
    const size_t L = sample_len;
    const size_t N = 2 * L;            /* :141 */
    const size_t nbins = N / 2 + 1;    /* :142 */
    int ret = -1;

    double *padded = fftw_alloc_real(N);              /* :159 */
    double *src = fftw_alloc_real(N);                 /* the shim never writes its input, but keep const-correct */
    fftw_complex *fa = fftw_alloc_complex(nbins);     /* :187 */
    fftw_complex *fb = fftw_alloc_complex(nbins);     /* :192 */
    double *corr = fftw_alloc_real(N);                /* :197 */
    if (!padded || !src || !fa || !fb || !corr) goto done;

    memcpy(src, source, N * sizeof(double));
    memcpy(padded, sample, L * sizeof(double));       /* :164 */
    memset(padded + L, 0, (N - L) * sizeof(double));  /* :165-166 */

    /* :204-229 -- two forward real transforms (the reference runs them on two
     * threads; the arithmetic is the same). */
    fftw_plan pa = fftw_plan_dft_r2c_1d((int)N, src, fa, FFTW_ESTIMATE);
    fftw_plan pb = fftw_plan_dft_r2c_1d((int)N, padded, fb, FFTW_ESTIMATE);
    if (!pa || !pb) { fftw_destroy_plan(pa); fftw_destroy_plan(pb); goto done; }
    fftw_execute(pa);
    fftw_execute(pb);
    fftw_destroy_plan(pa);
    fftw_destroy_plan(pb);

    /* :232-233 -- F(source) * conj(F(sample)) per bin, C99 complex product */
    for (size_t k = 0; k < nbins; k++)
        fa[k] = fa[k] * conj(fb[k]);

    /* :237-239 -- unnormalised inverse real transform */
    fftw_plan pc = fftw_plan_dft_c2r_1d((int)N, fa, corr, FFTW_ESTIMATE);
    if (!pc) goto done;
    fftw_execute(pc);
    fftw_destroy_plan(pc);

    /* :242 */
    size_t idx = oracle_max_abs_index(corr, N);
    if (extra) {
        double second = 0.0;
        for (size_t i = 0; i < N; i++) {
            if (i == idx) continue;
            double m = fabs(corr[i]);
            if (m > second) second = m;
        }
        extra->raw_index = (long)idx;
        extra->peak = corr[idx];
        extra->second = second;
        extra->r0 = corr[0];
    }

    /* :256-271 -- fold the index into a signed lag and pick the windows */
    const double *wx, *wy;
    size_t wn;
    long folded;
    if ((long)idx >= (long)L) {
        folded = ((long)idx % (long)L) - (long)L;     /* :259 */
        wx = source;                                  /* :260 */
        wy = padded - folded;                         /* :262 */
        wn = (size_t)((long)L + folded);              /* :261,263 */
    } else {
        folded = (long)idx;
        wx = source + folded;                         /* :267 */
        wy = padded;                                  /* :269 */
        wn = L;                                       /* :268,270 */
    }
    *lag = folded;
    *coefficient = oracle_pearson(wx, wy, wn);        /* :272-273 */

    if (*coefficient != *coefficient) goto done;      /* :276 NaN gate */
    ret = 0;                                          /* :298 */

done:
    if (padded) fftw_free(padded);
    if (src) fftw_free(src);
    if (fa) fftw_free(fa);
    if (fb) fftw_free(fb);
    if (corr) fftw_free(corr);
    return ret;
}

/* reference src/audiosync.c:254 with include/audiosync/audiosync.h:24 */
int oracle_accept(int ret, double coefficient)
{
    return ret == 0 && coefficient >= 0.95;
}

/* reference src/audiosync.c:255 with include/audiosync/audiosync.h:21 */
long oracle_frames_to_ms(long lag_frames)
{
    return (long)round((double)lag_frames * (1000.0 / 48000.0));
}

/* reference src/audiosync.c:50-57: {3, 6, 10, 15, 20, 30} s at 48 kHz */
const size_t ORACLE_INTERV_SAMPLE[ORACLE_N_INTERVALS] = {
    3 * 48000, 6 * 48000, 10 * 48000, 15 * 48000, 20 * 48000, 30 * 48000,
};

/* reference src/audiosync.c:226-259 on complete buffers (no reader threads) */
int oracle_interval_loop(const double *source, const double *sample,
                         int rets[ORACLE_N_INTERVALS],
                         long lags[ORACLE_N_INTERVALS],
                         double coefs[ORACLE_N_INTERVALS],
                         int succ[ORACLE_N_INTERVALS],
                         int *final_ret, long *final_lag)
{
    int done = 0;
    long lag = 0;
    *final_ret = -1;
    for (int i = 0; i < ORACLE_N_INTERVALS; i++) {
        double conf = 0.0;
        rets[i] = oracle_cross_correlation(source, sample, ORACLE_INTERV_SAMPLE[i],
                                           &lag, &conf, NULL);
        lags[i] = lag;
        coefs[i] = conf;
        succ[i] = oracle_accept(rets[i], conf);
        done = i + 1;
        if (rets[i] < 0) continue;                    /* :247-249 */
        if (succ[i]) {                                /* :254-258 */
            lag = oracle_frames_to_ms(lag);
            *final_ret = 0;
            break;
        }
    }
    *final_lag = lag;
    return done;
}

/* ------------------------------------------------------------ synthetic data
 * counter-based, all-integer, so host and device agree
 * bit for bit and every value is exact in fp32 and fp64. */

uint64_t synth_splitmix64(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static inline uint64_t synth_key(uint64_t seed, uint64_t pair_id, uint64_t stream)
{
    return seed ^ (pair_id * 0xD1342543DE82EF95ULL) ^ (stream * 0xA0761D6478BD642FULL);
}

static inline int64_t synth_q(uint64_t key, uint64_t i)
{
    return (int64_t)(synth_splitmix64(key + i) >> 40) - (int64_t)(1 << 23);
}

long synth_true_lag(uint64_t seed, uint64_t pair_id, size_t sample_len)
{
    uint64_t h = synth_splitmix64(synth_key(seed, pair_id, 2));
    return (long)(h % (uint64_t)(sample_len + 1)) - (long)(sample_len / 2);
}

void synth_pair_i32(uint64_t seed, uint64_t pair_id, size_t L,
                    int32_t *source, int32_t *sample)
{
    const uint64_t k0 = synth_key(seed, pair_id, 0);
    const uint64_t k1 = synth_key(seed, pair_id, 1);
    const long tl = synth_true_lag(seed, pair_id, L);
    const int64_t amp = (pair_id % 4 == 3) ? 768 : 102;
    const uint64_t half = L / 2;
    for (size_t i = 0; i < 2 * L; i++)
        source[i] = (int32_t)synth_q(k0, half + i);
    for (size_t j = 0; j < L; j++) {
        int64_t base = synth_q(k0, (uint64_t)((int64_t)half + tl + (int64_t)j));
        int64_t noise = synth_q(k1, j);
        int64_t scaled = (noise * amp) >> 10;   /* arithmetic shift = floor */
        sample[j] = (int32_t)(base + scaled);
    }
}

void synth_pair_f64(uint64_t seed, uint64_t pair_id, size_t L,
                    double *source, double *sample)
{
    int32_t *a = malloc(sizeof(int32_t) * 3 * L);
    if (!a) return;
    synth_pair_i32(seed, pair_id, L, a, a + 2 * L);
    const double s = 1.0 / 8388608.0;
    for (size_t i = 0; i < 2 * L; i++) source[i] = (double)a[i] * s;
    for (size_t j = 0; j < L; j++) sample[j] = (double)a[2 * L + j] * s;
    free(a);
}

void synth_pair_f32(uint64_t seed, uint64_t pair_id, size_t L,
                    float *source, float *sample)
{
    int32_t *a = malloc(sizeof(int32_t) * 3 * L);
    if (!a) return;
    synth_pair_i32(seed, pair_id, L, a, a + 2 * L);
    const float s = 1.0f / 8388608.0f;
    for (size_t i = 0; i < 2 * L; i++) source[i] = (float)a[i] * s;
    for (size_t j = 0; j < L; j++) sample[j] = (float)a[2 * L + j] * s;
    free(a);
}

struct batch_job {
    uint64_t seed, first;
    size_t count, L;
    int tid, nthreads;
    long *lags; double *coefs; int *rets; double *peaks; double *seconds;
    int failed;
};

static void *batch_worker(void *arg)
{
    struct batch_job *job = arg;
    double *src = malloc(sizeof(double) * 2 * job->L);
    double *smp = malloc(sizeof(double) * job->L);
    if (!src || !smp) { job->failed = 1; free(src); free(smp); return NULL; }
    for (size_t p = (size_t)job->tid; p < job->count; p += (size_t)job->nthreads) {
        struct oracle_extra ex;
        long lag = 0; double coef = 0.0;
        synth_pair_f64(job->seed, job->first + p, job->L, src, smp);
        int r = oracle_cross_correlation(src, smp, job->L, &lag, &coef, &ex);
        if (job->rets) job->rets[p] = r;
        if (job->lags) job->lags[p] = lag;
        if (job->coefs) job->coefs[p] = coef;
        if (job->peaks) job->peaks[p] = ex.peak;
        if (job->seconds) job->seconds[p] = ex.second;
    }
    free(src); free(smp);
    return NULL;
}

int oracle_synth_batch(uint64_t seed, uint64_t first_pair, size_t count,
                       size_t sample_len, int threads, long *lags,
                       double *coefs, int *rets, double *peaks, double *seconds)
{
    if (threads < 1) threads = 1;
    if ((size_t)threads > count) threads = (int)(count ? count : 1);
    pthread_t *th = calloc((size_t)threads, sizeof(*th));
    struct batch_job *jobs = calloc((size_t)threads, sizeof(*jobs));
    if (!th || !jobs) { free(th); free(jobs); return -1; }
    for (int t = 0; t < threads; t++) {
        struct batch_job j = { seed, first_pair, count, sample_len, t, threads,
                               lags, coefs, rets, peaks, seconds, 0 };
        jobs[t] = j;
        if (pthread_create(&th[t], NULL, batch_worker, &jobs[t]) != 0) {
            jobs[t].failed = 1;
            th[t] = 0;
            batch_worker(&jobs[t]);
        }
    }
    int bad = 0;
    for (int t = 0; t < threads; t++) {
        if (th[t]) pthread_join(th[t], NULL);
        bad |= jobs[t].failed;
    }
    free(th); free(jobs);
    return bad ? -1 : 0;
}

const char *oracle_backend(void) { return oracle_fft_backend(); }
