/*
 * oracle/xcorr_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp64) of the reference's FFT cross-correlation
 * hot path, used only as the checker in tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.  The product library
 * (libaudiosync_cuda.so) never links or calls anything declared here.
 *
 * Parity pin: the restatement is checked (tests/test_oracle.py) against
 *   (1) the 12 known-answer checks of the reference's own
 *       tests/test_cross_correlation.c and tests/test_pearson_coefficient.c,
 *   (2) the reference's unmodified src/cross_correlation.c compiled here into
 *       oracle/_ref/ (same FFT shim underneath), bit for bit, and
 *   (3) an independent NumPy (pocketfft) restatement, oracle/xcorr_numpy.py.
 * The FFT arithmetic itself lives in FFTW3, which is absent; see fftw3.h.
 */
#ifndef XCORR_ORACLE_H
#define XCORR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Extra observables the reference computes internally but never returns. */
struct oracle_extra {
    long   raw_index;   /* argmax index before folding, in [0, 2L)            */
    double peak;        /* results[raw_index] (signed, unnormalised, x N)     */
    double second;      /* largest |results[i]| over i != raw_index           */
    double r0;          /* results[0]                                         */
};

/* src/cross_correlation.c:52-67 */
size_t oracle_max_abs_index(const double *arr, size_t len);

/* src/cross_correlation.c:74-116 (pointer-range form -> (x, y, n) form) */
double oracle_pearson(const double *x, const double *y, size_t n);

/* src/cross_correlation.c:133-307.  `extra` may be NULL.  Returns 0 / -1
 * exactly as the reference does (outputs are written before the NaN gate). */
int oracle_cross_correlation(const double *source, const double *sample,
                             size_t sample_len, long *lag, double *coefficient,
                             struct oracle_extra *extra);

/* Caller-side acceptance, src/audiosync.c:246-258 with
 * include/audiosync/audiosync.h:21,24: success iff ret == 0 and
 * coef >= 0.95; lag_ms = round(lag * 1000 / 48000). */
int  oracle_accept(int ret, double coefficient);
long oracle_frames_to_ms(long lag_frames);

/* The interval schedule of src/audiosync.c:50-70 (sample lengths; the source
 * length is always twice that) and the loop of :226-259 run on complete
 * buffers: per interval i it records ret/lag/coef/success and stops at the
 * first success like the reference.  Returns the number of intervals
 * evaluated; final_ret and final_lag mirror audiosync_run's return and *lag
 * (including the "last frame lag on failure" quirk). */
#define ORACLE_N_INTERVALS 6
extern const size_t ORACLE_INTERV_SAMPLE[ORACLE_N_INTERVALS];
int oracle_interval_loop(const double *source, const double *sample,
                         int rets[ORACLE_N_INTERVALS],
                         long lags[ORACLE_N_INTERVALS],
                         double coefs[ORACLE_N_INTERVALS],
                         int succ[ORACLE_N_INTERVALS],
                         int *final_ret, long *final_lag);

/* ---- seeded all-integer synthetic pairs ---------- */
uint64_t synth_splitmix64(uint64_t x);
long     synth_true_lag(uint64_t seed, uint64_t pair_id, size_t sample_len);
/* integer samples (value = int * 2^-23); source has 2L, sample has L entries */
void synth_pair_i32(uint64_t seed, uint64_t pair_id, size_t sample_len,
                    int32_t *source, int32_t *sample);
void synth_pair_f64(uint64_t seed, uint64_t pair_id, size_t sample_len,
                    double *source, double *sample);
void synth_pair_f32(uint64_t seed, uint64_t pair_id, size_t sample_len,
                    float *source, float *sample);

/* Convenience for benchmarks/tests: generate pair ids [first, first+count)
 * and run oracle_cross_correlation on each with `threads` pthreads. */
int oracle_synth_batch(uint64_t seed, uint64_t first_pair, size_t count,
                       size_t sample_len, int threads, long *lags,
                       double *coefs, int *rets, double *peaks,
                       double *seconds);

const char *oracle_backend(void);

#ifdef __cplusplus
}
#endif
#endif
