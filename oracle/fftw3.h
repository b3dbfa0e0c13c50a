/*
 * oracle/fftw3.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A stand-in for the seven FFTW3 (double precision) entry points that the
 * reference's hot path links against.  FFTW3 itself is an external,
 * un-vendored dependency of the reference (`-lfftw3`, reference setup.py:12,
 * Dockerfile: libfftw3-dev from Debian buster, i.e. the 3.3.8 series) and is
 * not installed in this image, so the oracle supplies the published
 * semantics of those calls:
 *
 *   fftw_plan_dft_r2c_1d(n, in, out, flags): out[k] = sum_j in[j] e^{-2 pi i jk/n},
 *        k = 0..n/2, input preserved (out-of-place).
 *   fftw_plan_dft_c2r_1d(n, in, out, flags): out[j] = sum_k H[k] e^{+2 pi i jk/n}
 *        with H the Hermitian extension of in[0..n/2]; UNNORMALISED (a
 *        round trip scales by n); imaginary parts of bins 0 and n/2 ignored;
 *        the input array may be destroyed.
 *
 * Call sites in the reference: src/cross_correlation.c:34,39,43 (r2c plan,
 * execute, destroy), :159,187,192,197 (alloc), :237-239 (c2r), :301-304
 * (free); src/audiosync.c:189,277 (alloc/free of the source buffer).
 *
 * Only what those call sites use is declared.  Implementation:
 * oracle/fftw3_shim.c.  Every number produced through this header must be
 * labelled "shim", never "FFTW".
 */
#ifndef ORACLE_FFTW3_SHIM_H
#define ORACLE_FFTW3_SHIM_H

#include <stddef.h>
#include <complex.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double _Complex fftw_complex;
typedef struct oracle_fftw_plan_s *fftw_plan;

#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out,
                               unsigned flags);
fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex *in, double *out,
                               unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
double *fftw_alloc_real(size_t n);
fftw_complex *fftw_alloc_complex(size_t n);
void *fftw_malloc(size_t n);
void fftw_free(void *p);

/* Identifies the backend in reports ("shim-stockham-f64"). */
const char *oracle_fft_backend(void);

#ifdef __cplusplus
}
#endif
#endif
