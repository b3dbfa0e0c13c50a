// host_simd.cpp -- the one host-side inner loop of libaudiosync_cuda that is worth vectorising by
// hand: double -> float conversion of arriving audio while it is copied into the pinned staging
// ring (host narrowing, see audiosync_cuda_set_host_narrowing).  Round to nearest even, exactly the
// conversion the kernels apply when they load doubles.  The conversion also reports whether it was
// LOSSLESS -- every double the exact image of its float (audio decoded from 16/24-bit PCM or float
// samples always is; NaN, values beyond the float range or with more than 24 significant bits are
// not) -- which is what lets the library halve the PCIe bytes without changing a single result bit.
// Plain C++ (no CUDA), dispatched at run time on what the CPU supports.
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>

namespace asc {

static inline bool exact_scalar(float* dst, const double* src, size_t i) {
    const float f = (float)src[i];
    dst[i] = f;
    return (double)f == src[i];
}

// `stream`: non-temporal stores (the destination is a pinned buffer the copy engine reads next; the
// CPU never looks at it again, so it need not be read for ownership nor kept in cache).  Software
// prefetch of the source (512 / 2048 doubles ahead) was measured and costs 15 - 20 %.
__attribute__((target("avx512f"))) static bool narrow_avx512(float* __restrict__ dst, const double* __restrict__ src, size_t n,
                                                                bool stream) {
    size_t i = 0;
    bool ok = true;
    for (; i < n && (reinterpret_cast<uintptr_t>(dst + i) & 63u) != 0; i++) ok &= exact_scalar(dst, src, i);
    __mmask8 bad = 0;
    if (stream) {
        for (; i + 16 <= n; i += 16) {
            const __m512d x0 = _mm512_loadu_pd(src + i), x1 = _mm512_loadu_pd(src + i + 8);
            const __m256 a = _mm512_cvtpd_ps(x0), b = _mm512_cvtpd_ps(x1);
            bad |= _mm512_cmp_pd_mask(_mm512_cvtps_pd(a), x0, _CMP_NEQ_UQ) | _mm512_cmp_pd_mask(_mm512_cvtps_pd(b), x1, _CMP_NEQ_UQ);
            const __m512d ab = _mm512_insertf64x4(_mm512_castpd256_pd512(_mm256_castps_pd(a)), _mm256_castps_pd(b), 1);
            _mm512_stream_pd(reinterpret_cast<double*>(dst + i), ab);
        }
        _mm_sfence();
    } else {
        for (; i + 16 <= n; i += 16) {
            const __m512d x0 = _mm512_loadu_pd(src + i), x1 = _mm512_loadu_pd(src + i + 8);
            const __m256 a = _mm512_cvtpd_ps(x0), b = _mm512_cvtpd_ps(x1);
            bad |= _mm512_cmp_pd_mask(_mm512_cvtps_pd(a), x0, _CMP_NEQ_UQ) | _mm512_cmp_pd_mask(_mm512_cvtps_pd(b), x1, _CMP_NEQ_UQ);
            _mm256_storeu_ps(dst + i, a);
            _mm256_storeu_ps(dst + i + 8, b);
        }
    }
    ok &= bad == 0;
    for (; i < n; i++) ok &= exact_scalar(dst, src, i);
    return ok;
}

__attribute__((target("avx2"))) static bool narrow_avx2(float* __restrict__ dst, const double* __restrict__ src, size_t n, bool stream) {
    size_t i = 0;
    bool ok = true;
    for (; i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31u) != 0; i++) ok &= exact_scalar(dst, src, i);
    __m256d bad = _mm256_setzero_pd();
    for (; i + 8 <= n; i += 8) {
        const __m256d x0 = _mm256_loadu_pd(src + i), x1 = _mm256_loadu_pd(src + i + 4);
        const __m128 a = _mm256_cvtpd_ps(x0), b = _mm256_cvtpd_ps(x1);
        bad = _mm256_or_pd(bad, _mm256_or_pd(_mm256_cmp_pd(_mm256_cvtps_pd(a), x0, _CMP_NEQ_UQ),
                                             _mm256_cmp_pd(_mm256_cvtps_pd(b), x1, _CMP_NEQ_UQ)));
        const __m256 ab = _mm256_insertf128_ps(_mm256_castps128_ps256(a), b, 1);
        if (stream) _mm256_stream_ps(dst + i, ab);
        else _mm256_storeu_ps(dst + i, ab);
    }
    if (stream) _mm_sfence();
    ok &= _mm256_movemask_pd(bad) == 0;
    for (; i < n; i++) ok &= exact_scalar(dst, src, i);
    return ok;
}

static bool narrow_base(float* __restrict__ dst, const double* __restrict__ src, size_t n, bool) {
    bool ok = true;
    for (size_t i = 0; i < n; i++) ok &= exact_scalar(dst, src, i);
    return ok;
}

// dst[i] = (float)src[i] for i < n; returns true when every conversion was exact.
bool narrow_f64_to_f32(float* dst, const double* src, size_t n, bool stream) {
    typedef bool (*fn_t)(float*, const double*, size_t, bool);
    static const fn_t fn = [] {
        __builtin_cpu_init();
        if (__builtin_cpu_supports("avx512f")) return (fn_t)narrow_avx512;
        if (__builtin_cpu_supports("avx2")) return (fn_t)narrow_avx2;
        return (fn_t)narrow_base;
    }();
    return fn(dst, src, n, stream);
}

}  // namespace asc
