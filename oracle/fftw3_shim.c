/*
 * oracle/fftw3_shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Double-precision implementation of the FFTW3 entry points declared in
 * oracle/fftw3.h (the seven symbols the reference's hot path uses, see the
 * call-site list there).  FFTW3 is an absent third-party dependency of the
 * reference; this file restates the *published* r2c / c2r semantics with an
 * independent algorithm:
 *
 *   - complex core: out-of-place Stockham autosort, radices 4, 2, 3, 5 and a
 *     generic O(p^2) butterfly for any other prime factor, so every length
 *     works (the reference's tests use n = 10, 12, 14, 2000);
 *   - real transforms of even n: one complex transform of length n/2 on the
 *     packed signal plus the usual split (r2c) / merge (c2r) step;
 *     odd n: a full-length complex transform;
 *   - twiddles: W_n^t built in long double from two ~sqrt(n)-sized tables
 *     (one sinl/cosl pair per table entry, one long-double complex product
 *     per twiddle) and rounded once to double; tables are cached per n for
 *     the life of the process so that plan creation stays cheap, as it is
 *     with FFTW_ESTIMATE.
 *
 * Optional second backend, for TIMING the reference only (bench.py's CPU legs):
 * with ORACLE_FFT_BACKEND=mkl in the environment and Intel MKL's DFTI entry
 * points reachable (PyTorch's libtorch_cpu.so exports them; ORACLE_MKL_LIB names
 * the library), plans of even length run on MKL -- an FFTW-class library -- one
 * thread per transform like FFTW's default plans, so the reference's CPU number
 * is not held back by this file's plain-C Stockham core.  The backend is
 * self-tested against a direct DFT when it is loaded and is never used by the
 * tests: golden vectors and parity checks stay on the independent core below.
 *
 * Nothing under /root/reference was consulted for this file beyond the call
 * sites; nothing in the product library links it.
 */
#define _GNU_SOURCE
#include "fftw3.h"

#include <dlfcn.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cpx;

/* ---------------------------------------------------------------- twiddles */

struct tw_table {
    int n;
    cpx *w;                 /* w[t] = exp(-2 pi i t / n), t in [0, n) */
    struct tw_table *next;
};

static pthread_mutex_t tw_mutex = PTHREAD_MUTEX_INITIALIZER;
static struct tw_table *tw_cache = NULL;

static const cpx *twiddles_for(int n)
{
    pthread_mutex_lock(&tw_mutex);
    for (struct tw_table *t = tw_cache; t; t = t->next) {
        if (t->n == n) {
            pthread_mutex_unlock(&tw_mutex);
            return t->w;
        }
    }
    struct tw_table *t = malloc(sizeof(*t));
    cpx *w = NULL;
    if (t && posix_memalign((void **)&w, 64, sizeof(cpx) * (size_t)(n > 0 ? n : 1)) != 0)
        w = NULL;
    if (!t || !w) {
        free(t);
        pthread_mutex_unlock(&tw_mutex);
        return NULL;
    }
    /* two-level construction: t = a * B + b */
    int B = (int)ceil(sqrt((double)n));
    if (B < 1) B = 1;
    int A = (n + B - 1) / B;
    long double *cr = malloc(sizeof(long double) * 2 * (size_t)(A + B));
    if (!cr) {
        free(t); free(w);
        pthread_mutex_unlock(&tw_mutex);
        return NULL;
    }
    long double *ci = cr + A, *fr = ci + A, *fi = fr + B;
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int a = 0; a < A; a++) {
        long double ang = -two_pi * (long double)((long long)a * B) / (long double)n;
        cr[a] = cosl(ang); ci[a] = sinl(ang);
    }
    for (int b = 0; b < B; b++) {
        long double ang = -two_pi * (long double)b / (long double)n;
        fr[b] = cosl(ang); fi[b] = sinl(ang);
    }
    for (int a = 0; a < A; a++) {
        for (int b = 0; b < B; b++) {
            long long idx = (long long)a * B + b;
            if (idx >= n) break;
            w[idx].re = (double)(cr[a] * fr[b] - ci[a] * fi[b]);
            w[idx].im = (double)(cr[a] * fi[b] + ci[a] * fr[b]);
        }
    }
    /* exact values on the axes */
    w[0].re = 1.0; w[0].im = 0.0;
    if (n % 2 == 0) { w[n / 2].re = -1.0; w[n / 2].im = 0.0; }
    if (n % 4 == 0) {
        w[n / 4].re = 0.0; w[n / 4].im = -1.0;
        w[3 * (n / 4)].re = 0.0; w[3 * (n / 4)].im = 1.0;
    }
    free(cr);
    t->n = n; t->w = w; t->next = tw_cache; tw_cache = t;
    pthread_mutex_unlock(&tw_mutex);
    return w;
}

/* ------------------------------------------------------------ complex core */

#define MAX_FACTORS 64

struct cfft {
    int n;
    int nfac;
    int fac[MAX_FACTORS];
    const cpx *w;           /* W_n^t */
};

static int cfft_init(struct cfft *c, int n)
{
    c->n = n; c->nfac = 0;
    int r = n;
    while (r % 4 == 0) { c->fac[c->nfac++] = 4; r /= 4; }
    while (r % 2 == 0) { c->fac[c->nfac++] = 2; r /= 2; }
    while (r % 3 == 0) { c->fac[c->nfac++] = 3; r /= 3; }
    while (r % 5 == 0) { c->fac[c->nfac++] = 5; r /= 5; }
    for (int p = 7; (long long)p * p <= r; p += 2)
        while (r % p == 0) { c->fac[c->nfac++] = p; r /= p; }
    if (r > 1) c->fac[c->nfac++] = r;
    c->w = twiddles_for(n);
    return c->w ? 0 : -1;
}

static inline cpx cmul(cpx a, cpx b)
{
    cpx r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re };
    return r;
}

/* multiply by the table twiddle, conjugated for the inverse direction */
static inline cpx twmul(cpx a, cpx w, int inverse)
{
    if (inverse) w.im = -w.im;
    return cmul(a, w);
}

/* One Stockham pass of radix R over `in` -> `out`.
 * ns = product of the radices already applied. */
static void pass_generic(const struct cfft *c, const cpx *in, cpx *out,
                         int R, int ns, int inverse)
{
    const int n = c->n, m = n / R, stride = n / (ns * R), pstep = n / R;
    cpx *v = malloc(sizeof(cpx) * 2 * (size_t)R);
    cpx *y = v + R;
    for (int j = 0; j < m; j++) {
        int k = j % ns;
        int ob = (j - k) * R + k;
        for (int q = 0; q < R; q++) {
            cpx x = in[j + q * m];
            if (k) x = twmul(x, c->w[(long long)q * k * stride], inverse);
            v[q] = x;
        }
        for (int a = 0; a < R; a++) {
            cpx s = { 0.0, 0.0 };
            for (int q = 0; q < R; q++) {
                cpx t = twmul(v[q], c->w[(long long)((a * q) % R) * pstep], inverse);
                s.re += t.re; s.im += t.im;
            }
            y[a] = s;
        }
        for (int a = 0; a < R; a++) out[ob + a * ns] = y[a];
    }
    free(v);
}

static void pass2(const struct cfft *c, const cpx *in, cpx *out, int ns, int inverse)
{
    const int n = c->n, m = n / 2, stride = n / (ns * 2);
    for (int jb = 0; jb < m; jb += ns) {
        for (int k = 0; k < ns; k++) {
            int j = jb + k;
            cpx a = in[j], b = in[j + m];
            if (k) b = twmul(b, c->w[(long long)k * stride], inverse);
            cpx *o = out + (size_t)jb * 2 + k;
            o[0].re = a.re + b.re; o[0].im = a.im + b.im;
            o[ns].re = a.re - b.re; o[ns].im = a.im - b.im;
        }
    }
}

static void pass3(const struct cfft *c, const cpx *in, cpx *out, int ns, int inverse)
{
    const int n = c->n, m = n / 3, stride = n / (ns * 3);
    const double s60 = inverse ? 0.86602540378443864676 : -0.86602540378443864676;
    for (int jb = 0; jb < m; jb += ns) {
        for (int k = 0; k < ns; k++) {
            int j = jb + k;
            cpx a = in[j], b = in[j + m], d = in[j + 2 * m];
            if (k) {
                b = twmul(b, c->w[(long long)k * stride], inverse);
                d = twmul(d, c->w[(long long)2 * k * stride], inverse);
            }
            cpx t1 = { b.re + d.re, b.im + d.im };
            cpx t2 = { a.re - 0.5 * t1.re, a.im - 0.5 * t1.im };
            cpx t3 = { s60 * (b.re - d.re), s60 * (b.im - d.im) };
            cpx *o = out + (size_t)jb * 3 + k;
            o[0].re = a.re + t1.re;      o[0].im = a.im + t1.im;
            o[ns].re = t2.re - t3.im;    o[ns].im = t2.im + t3.re;
            o[2 * ns].re = t2.re + t3.im; o[2 * ns].im = t2.im - t3.re;
        }
    }
}

static void pass4(const struct cfft *c, const cpx *in, cpx *out, int ns, int inverse)
{
    const int n = c->n, m = n / 4, stride = n / (ns * 4);
    for (int jb = 0; jb < m; jb += ns) {
        for (int k = 0; k < ns; k++) {
            int j = jb + k;
            cpx a = in[j], b = in[j + m], d = in[j + 2 * m], e = in[j + 3 * m];
            if (k) {
                b = twmul(b, c->w[(long long)k * stride], inverse);
                d = twmul(d, c->w[(long long)2 * k * stride], inverse);
                e = twmul(e, c->w[(long long)3 * k * stride], inverse);
            }
            cpx s0 = { a.re + d.re, a.im + d.im }, s1 = { a.re - d.re, a.im - d.im };
            cpx s2 = { b.re + e.re, b.im + e.im }, s3 = { b.re - e.re, b.im - e.im };
            /* forward: multiply s3 by -i ; inverse: by +i */
            cpx r3;
            if (!inverse) { r3.re = s3.im; r3.im = -s3.re; }
            else          { r3.re = -s3.im; r3.im = s3.re; }
            cpx *o = out + (size_t)jb * 4 + k;
            o[0].re = s0.re + s2.re;      o[0].im = s0.im + s2.im;
            o[ns].re = s1.re + r3.re;     o[ns].im = s1.im + r3.im;
            o[2 * ns].re = s0.re - s2.re; o[2 * ns].im = s0.im - s2.im;
            o[3 * ns].re = s1.re - r3.re; o[3 * ns].im = s1.im - r3.im;
        }
    }
}

static void pass5(const struct cfft *c, const cpx *in, cpx *out, int ns, int inverse)
{
    const int n = c->n, m = n / 5, stride = n / (ns * 5);
    const double c1 = 0.30901699437494742410;   /* cos(2pi/5) */
    const double c2 = -0.80901699437494742410;  /* cos(4pi/5) */
    const double sg = inverse ? 1.0 : -1.0;
    const double s1 = sg * 0.95105651629515357212; /* sin(2pi/5) */
    const double s2 = sg * 0.58778525229247312917; /* sin(4pi/5) */
    for (int jb = 0; jb < m; jb += ns) {
        for (int k = 0; k < ns; k++) {
            int j = jb + k;
            cpx x0 = in[j], x1 = in[j + m], x2 = in[j + 2 * m],
                x3 = in[j + 3 * m], x4 = in[j + 4 * m];
            if (k) {
                x1 = twmul(x1, c->w[(long long)k * stride], inverse);
                x2 = twmul(x2, c->w[(long long)2 * k * stride], inverse);
                x3 = twmul(x3, c->w[(long long)3 * k * stride], inverse);
                x4 = twmul(x4, c->w[(long long)4 * k * stride], inverse);
            }
            cpx a1 = { x1.re + x4.re, x1.im + x4.im }, b1 = { x1.re - x4.re, x1.im - x4.im };
            cpx a2 = { x2.re + x3.re, x2.im + x3.im }, b2 = { x2.re - x3.re, x2.im - x3.im };
            cpx p1 = { x0.re + c1 * a1.re + c2 * a2.re, x0.im + c1 * a1.im + c2 * a2.im };
            cpx p2 = { x0.re + c2 * a1.re + c1 * a2.re, x0.im + c2 * a1.im + c1 * a2.im };
            /* q = i * (s1 b1 + s2 b2) etc.; i*(u) = (-u.im, u.re) */
            cpx u1 = { s1 * b1.re + s2 * b2.re, s1 * b1.im + s2 * b2.im };
            cpx u2 = { s2 * b1.re - s1 * b2.re, s2 * b1.im - s1 * b2.im };
            cpx *o = out + (size_t)jb * 5 + k;
            o[0].re = x0.re + a1.re + a2.re; o[0].im = x0.im + a1.im + a2.im;
            o[ns].re = p1.re - u1.im;      o[ns].im = p1.im + u1.re;
            o[4 * ns].re = p1.re + u1.im;  o[4 * ns].im = p1.im - u1.re;
            o[2 * ns].re = p2.re - u2.im;  o[2 * ns].im = p2.im + u2.re;
            o[3 * ns].re = p2.re + u2.im;  o[3 * ns].im = p2.im - u2.re;
        }
    }
}

/* Transform n points held in `x` using `y` as scratch; returns the buffer
 * (x or y) that holds the result. */
static cpx *cfft_run(const struct cfft *c, cpx *x, cpx *y, int inverse)
{
    cpx *in = x, *out = y;
    int ns = 1;
    for (int f = 0; f < c->nfac; f++) {
        int R = c->fac[f];
        switch (R) {
        case 2: pass2(c, in, out, ns, inverse); break;
        case 3: pass3(c, in, out, ns, inverse); break;
        case 4: pass4(c, in, out, ns, inverse); break;
        case 5: pass5(c, in, out, ns, inverse); break;
        default: pass_generic(c, in, out, R, ns, inverse); break;
        }
        cpx *t = in; in = out; out = t;
        ns *= R;
    }
    return in;
}

/* ------------------------------------------------------------------ plans */

enum plan_kind { PLAN_R2C, PLAN_C2R };

struct oracle_fftw_plan_s {
    enum plan_kind kind;
    int n;                  /* logical (real) length */
    int h;                  /* complex transform length: n/2 (even n) or n */
    double *rbuf;           /* the real array    */
    fftw_complex *cbuf;     /* the complex array */
    struct cfft core;
    const cpx *wn;          /* W_n^t for the split/merge step (even n)   */
    cpx *work0, *work1;
    void *mkl;              /* committed DFTI descriptor when the MKL backend serves this plan */
};


/* ------------------------------------------------------- optional MKL backend */
/* Values of mkl_dfti.h (enum DFTI_CONFIG_PARAM / DFTI_CONFIG_VALUE); checked by the self-test. */
enum { DFTI_CONJUGATE_EVEN_STORAGE_ = 10, DFTI_PLACEMENT_ = 11, DFTI_PACKED_FORMAT_ = 21, DFTI_THREAD_LIMIT_ = 27,
       DFTI_REAL_ = 33, DFTI_DOUBLE_ = 36, DFTI_COMPLEX_COMPLEX_ = 39, DFTI_NOT_INPLACE_ = 44, DFTI_CCE_FORMAT_ = 57 };
typedef long (*dfti_create_t)(void **, int, long);
typedef long (*dfti_set_t)(void *, int, ...);
typedef long (*dfti_commit_t)(void *);
typedef long (*dfti_compute_t)(void *, void *, ...);
typedef long (*dfti_free_t)(void **);
static struct {
    int state;              /* 0 untried, 1 usable, -1 unavailable */
    dfti_create_t create; dfti_set_t set; dfti_commit_t commit;
    dfti_compute_t fwd, bwd; dfti_free_t release;
} mkl;
static pthread_mutex_t mkl_mutex = PTHREAD_MUTEX_INITIALIZER;

/* Committed descriptors are pooled per length: the reference creates and destroys its plans
 * on every call (FFTW_ESTIMATE plans are cheap, a DFTI commit for 2.88M points is not), so a
 * destroyed plan parks its descriptor here and the next plan of that length takes it over --
 * the counterpart of the twiddle cache of the built-in core. */
struct mkl_idle { int n; void *d; struct mkl_idle *next; };
static struct mkl_idle *mkl_pool = NULL;
static pthread_mutex_t mkl_pool_mutex = PTHREAD_MUTEX_INITIALIZER;

static void *mkl_pool_take(int n)
{
    void *d = NULL;
    pthread_mutex_lock(&mkl_pool_mutex);
    for (struct mkl_idle **pp = &mkl_pool; *pp; pp = &(*pp)->next) {
        if ((*pp)->n == n) {
            struct mkl_idle *e = *pp;
            *pp = e->next;
            d = e->d;
            free(e);
            break;
        }
    }
    pthread_mutex_unlock(&mkl_pool_mutex);
    return d;
}

static int mkl_pool_park(int n, void *d)
{
    struct mkl_idle *e = malloc(sizeof(*e));
    if (!e) return -1;
    e->n = n; e->d = d;
    pthread_mutex_lock(&mkl_pool_mutex);
    e->next = mkl_pool;
    mkl_pool = e;
    pthread_mutex_unlock(&mkl_pool_mutex);
    return 0;
}

static void *mkl_descriptor(int n)
{
    void *d = mkl_pool_take(n);
    if (d) return d;
    if (mkl.create(&d, DFTI_REAL_, (long)n) != 0 || !d) return NULL;
    if (mkl.set(d, DFTI_PLACEMENT_, DFTI_NOT_INPLACE_) != 0 ||
        mkl.set(d, DFTI_CONJUGATE_EVEN_STORAGE_, DFTI_COMPLEX_COMPLEX_) != 0 ||
        mkl.set(d, DFTI_PACKED_FORMAT_, DFTI_CCE_FORMAT_) != 0 ||
        mkl.set(d, DFTI_THREAD_LIMIT_, 1L) != 0 ||
        mkl.commit(d) != 0) {
        mkl.release(&d);
        return NULL;
    }
    return d;
}

/* r2c and c2r of a short signal against a direct DFT */
static int mkl_selftest(void)
{
    enum { N = 24 };
    double x[N], y[N];
    cpx X[N / 2 + 1];
    for (int j = 0; j < N; j++) x[j] = sin(0.7 * j) + 0.01 * j * j - 0.3 * (j % 5);
    void *d = mkl_descriptor(N);
    if (!d) return -1;
    int ok = mkl.fwd(d, x, X) == 0;
    for (int k = 0; ok && k <= N / 2; k++) {
        double re = 0.0, im = 0.0;
        for (int j = 0; j < N; j++) {
            re += x[j] * cos(2.0 * M_PI * j * k / N);
            im -= x[j] * sin(2.0 * M_PI * j * k / N);
        }
        if (fabs(re - X[k].re) > 1e-10 || fabs(im - X[k].im) > 1e-10) ok = 0;
    }
    ok = ok && mkl.bwd(d, X, y) == 0;
    for (int j = 0; ok && j < N; j++)
        if (fabs(y[j] - N * x[j]) > 1e-9) ok = 0;          /* unnormalised, like FFTW's c2r */
    mkl.release(&d);
    return ok ? 0 : -1;
}

static int mkl_ready(void)
{
    pthread_mutex_lock(&mkl_mutex);
    if (mkl.state == 0) {
        mkl.state = -1;
        const char *want = getenv("ORACLE_FFT_BACKEND");
        if (want && strcmp(want, "mkl") == 0) {
            const char *path = getenv("ORACLE_MKL_LIB");
            void *h = dlopen(path ? path : "libtorch_cpu.so", RTLD_NOW | RTLD_NOLOAD);
            if (!h) h = dlopen(path ? path : "libtorch_cpu.so", RTLD_NOW | RTLD_LOCAL);
            if (h) {
                mkl.create = (dfti_create_t)dlsym(h, "DftiCreateDescriptor_d_1d");
                mkl.set = (dfti_set_t)dlsym(h, "DftiSetValue");
                mkl.commit = (dfti_commit_t)dlsym(h, "DftiCommitDescriptor");
                mkl.fwd = (dfti_compute_t)dlsym(h, "DftiComputeForward");
                mkl.bwd = (dfti_compute_t)dlsym(h, "DftiComputeBackward");
                mkl.release = (dfti_free_t)dlsym(h, "DftiFreeDescriptor");
                if (mkl.create && mkl.set && mkl.commit && mkl.fwd && mkl.bwd && mkl.release &&
                    mkl_selftest() == 0)
                    mkl.state = 1;
            }
        }
    }
    const int r = mkl.state == 1;
    pthread_mutex_unlock(&mkl_mutex);
    return r;
}

static fftw_plan plan_new(enum plan_kind kind, int n, double *r, fftw_complex *cx)
{
    if (n <= 0 || !r || !cx) return NULL;
    fftw_plan p = calloc(1, sizeof(*p));
    if (!p) return NULL;
    p->kind = kind; p->n = n; p->rbuf = r; p->cbuf = cx;
    if (n % 2 == 0 && n >= 4 && mkl_ready()) {
        p->mkl = mkl_descriptor(n);
        if (p->mkl) return p;
    }
    p->h = (n % 2 == 0) ? n / 2 : n;
    if (cfft_init(&p->core, p->h) != 0) { free(p); return NULL; }
    p->wn = (n % 2 == 0) ? twiddles_for(n) : NULL;
    if (n % 2 == 0 && !p->wn) { free(p); return NULL; }
    if (posix_memalign((void **)&p->work0, 64, sizeof(cpx) * (size_t)p->h) != 0) p->work0 = NULL;
    if (posix_memalign((void **)&p->work1, 64, sizeof(cpx) * (size_t)p->h) != 0) p->work1 = NULL;
    if (!p->work0 || !p->work1) { free(p->work0); free(p->work1); free(p); return NULL; }
    return p;
}

fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags)
{
    (void)flags;
    return plan_new(PLAN_R2C, n, in, out);
}

fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex *in, double *out, unsigned flags)
{
    (void)flags;
    return plan_new(PLAN_C2R, n, out, in);
}

static void exec_r2c(const fftw_plan p)
{
    const int n = p->n, h = p->h;
    cpx *out = (cpx *)p->cbuf;
    if (n % 2) {
        for (int j = 0; j < n; j++) { p->work0[j].re = p->rbuf[j]; p->work0[j].im = 0.0; }
        cpx *z = cfft_run(&p->core, p->work0, p->work1, 0);
        for (int k = 0; k <= n / 2; k++) out[k] = z[k];
        return;
    }
    /* pack pairs of reals as complex points: z[j] = x[2j] + i x[2j+1] */
    memcpy(p->work0, p->rbuf, sizeof(double) * (size_t)n);
    cpx *z = cfft_run(&p->core, p->work0, p->work1, 0);
    /* split: X[k] = E[k] + W_n^k O[k],  E = (Z[k]+conj Z[h-k])/2,
     *        O = (Z[k]-conj Z[h-k])/(2i) */
    for (int k = 0; k <= h; k++) {
        cpx a = z[k == h ? 0 : k];
        cpx b = z[k == 0 ? 0 : h - k];
        cpx e = { 0.5 * (a.re + b.re), 0.5 * (a.im - b.im) };
        cpx o = { 0.5 * (a.im + b.im), -0.5 * (a.re - b.re) };
        cpx w;
        if (k == h) { w.re = -1.0; w.im = 0.0; } else w = p->wn[k];
        cpx wo = cmul(w, o);
        out[k].re = e.re + wo.re;
        out[k].im = e.im + wo.im;
    }
    out[0].im = 0.0;
    out[h].im = 0.0;
}

static void exec_c2r(const fftw_plan p)
{
    const int n = p->n, h = p->h;
    const cpx *in = (const cpx *)p->cbuf;
    if (n % 2) {
        p->work0[0].re = in[0].re; p->work0[0].im = 0.0;
        for (int k = 1; k <= n / 2; k++) {
            p->work0[k] = in[k];
            p->work0[n - k].re = in[k].re; p->work0[n - k].im = -in[k].im;
        }
        cpx *z = cfft_run(&p->core, p->work0, p->work1, 1);
        for (int j = 0; j < n; j++) p->rbuf[j] = z[j].re;
        return;
    }
    /* merge: Q[k] = (P[k] + conj P[h-k]) + i conj(W_n^k) (P[k] - conj P[h-k]) */
    for (int k = 0; k < h; k++) {
        cpx a = in[k], b = in[h - k];
        if (k == 0) { a.im = 0.0; b.im = 0.0; }
        cpx s = { a.re + b.re, a.im - b.im };
        cpx d = { a.re - b.re, a.im + b.im };
        cpx w = p->wn[k]; w.im = -w.im;
        cpx wd = cmul(w, d);
        p->work0[k].re = s.re - wd.im;
        p->work0[k].im = s.im + wd.re;
    }
    cpx *z = cfft_run(&p->core, p->work0, p->work1, 1);
    memcpy(p->rbuf, z, sizeof(double) * (size_t)n);
}

void fftw_execute(const fftw_plan p)
{
    if (!p) return;
    if (p->mkl) {
        if (p->kind == PLAN_R2C) mkl.fwd(p->mkl, p->rbuf, p->cbuf);
        else mkl.bwd(p->mkl, p->cbuf, p->rbuf);
        return;
    }
    if (p->kind == PLAN_R2C) exec_r2c(p); else exec_c2r(p);
}

void fftw_destroy_plan(fftw_plan p)
{
    if (!p) return;
    if (p->mkl && mkl_pool_park(p->n, p->mkl) != 0) mkl.release(&p->mkl);
    free(p->work0); free(p->work1); free(p);
}

void *fftw_malloc(size_t n)
{
    void *p = NULL;
    if (posix_memalign(&p, 64, n ? n : 1) != 0) return NULL;
    return p;
}

double *fftw_alloc_real(size_t n) { return fftw_malloc(n * sizeof(double)); }
fftw_complex *fftw_alloc_complex(size_t n) { return fftw_malloc(n * sizeof(fftw_complex)); }
void fftw_free(void *p) { free(p); }

const char *oracle_fft_backend(void)
{
    return mkl_ready() ? "mkl-dfti-f64 (one thread per transform)" : "shim-stockham-f64";
}
