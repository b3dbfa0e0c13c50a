// fft_generic.cuh -- four-step transform kernels with RUNTIME radix lists.
//
// The static kernels of fft_kernels.cuh exist for the six lengths of the reference's interval
// schedule.  The reference itself plans FFTW at call time for whatever sample_len arrives
// (src/cross_correlation.c:34, :141-142, :237), so every other length must be O(N log N) too:
// these kernels run the same algorithm (packed real FFT of complex length M = M1 * M2, column
// pass / fused row pass / inverse column pass with the |r| argmax) for any M1, M2 whose factors
// come from {2,3,4,5,6,8,9,10,12,15,16}, with strides, digit maps and twiddles taken from
// tables instead of template arguments.
//
// Lengths that are not 2/3/5-smooth (or have no usable split) are EMBEDDED: the circular
// correlation of period N = 2L that the reference computes equals, for the lags j < 2L it
// looks at, the linear correlation of the periodically extended source
//     r[j] = sum_{n < L} source[(n + j) mod 2L] * sample[n],     n + j < 3L,
// so any transform length N' = 2M >= 3L gives the same r[0 .. 2L) -- the source is read as
// source[i mod 2L] for i < 3L and zero beyond, the argmax looks at indices < 2L only, and the
// peak is rescaled by N / N' (FFTW's c2r is unnormalised: the reference's values carry N).
// No Bluestein pass is needed.
//
// The arithmetic type is a template parameter: float (the product) or double -- the
// fp64-arithmetic validation mode (audiosync_cuda_set_precise), a second oracle for the fp32
// kernels (reference :187-239 computes in double complex throughout).
//
// Same executor convention as fft_kernels.cuh: the bodies also run on the CPU emulator.
#pragma once

#include "fft_kernels.cuh"

namespace asc {

constexpr int GEN_MAX_PASSES = 8;
constexpr int GEN_THREADS = 256;

struct GenAxis {
    int n;                         // axis length
    int npass;
    int radix[GEN_MAX_PASSES];
    int stride[GEN_MAX_PASSES];    // s(p) = n / (r(0) * ... * r(p)); in-place DIF order as RadixList
};

struct GenShape {
    long long L;        // sample_len
    long long M;        // complex transform length; N' = 2M real points; M == L unless embedded
    long long src_ext;  // real points of the (periodically extended) source: 2L exact, 3L embedded
    int M1, M2;
    GenAxis col, row;
};

template <typename T> struct GenTraits;
template <> struct GenTraits<float> {
    typedef cplx C;
    static constexpr int CT = 16;      // columns per tile: 128 bytes
    static constexpr int PADSH = 4;    // row padding: one point per 16 (32 banks)
};
template <> struct GenTraits<double> {
    typedef cplxd C;
    static constexpr int CT = 8;
    static constexpr int PADSH = 3;
};

// a + conj(b), a - conj(b), acc +- a conj(b) for the fp64 type (fp32: fft_kernels.cuh)
ASC_HD cplxd cconj_add(cplxd a, cplxd b) { return cmake(a.x + b.x, a.y - b.y); }
ASC_HD cplxd cconj_sub(cplxd a, cplxd b) { return cmake(a.x - b.x, a.y + b.y); }
ASC_HD cplxd cmulc_acc(cplxd a, cplxd b, cplxd acc) {
    return cmake(fma(a.x, b.x, fma(a.y, b.y, acc.x)), fma(a.y, b.x, fma(-a.x, b.y, acc.y)));
}
ASC_HD cplxd cmulc_nacc(cplxd a, cplxd b, cplxd acc) {
    return cmake(fma(-a.x, b.x, fma(-a.y, b.y, acc.x)), fma(-a.y, b.x, fma(a.x, b.y, acc.y)));
}
// split / conj-multiply / merge with the twiddle folded (see split_mul_merge_w2): 2 Q[k], 2 Q[M-k]
template <class C>
ASC_HD void gen_split_mul_merge(C a, C b, C c, C d, C w2, C& qk2, C& qmk2) {
    const C s = cconj_add(a, b), t = cconj_sub(a, b);
    const C s2 = cconj_add(c, d), t2 = cconj_sub(c, d);
    const C g = cmulc_acc(t, t2, cmulc(s, s2));
    const C x = cmulc(s, t2);
    const C h = cmulc_nacc(x, w2, cmulc(t, s2));
    qk2 = cadd(g, h);
    const C e = csub(g, h);
    qmk2 = cmake(e.x, -e.y);
}

template <class C>
ASC_HD C gen_tw2(const C* __restrict__ lo, const C* __restrict__ hi, unsigned a) {
    return cmul(ldg(lo + (a & TW2_MASK)), ldg(hi + (a >> TW2_BITS)));
}

// Packed point n = (real 2n, real 2n + 1) of a signal.  Source: x[i mod 2L] for i < src_ext, zero
// beyond; sample: y[i] for i < L, zero beyond (the zero pad of reference :159-166).
template <class C, typename InT>
ASC_HD C gen_load_point(const InT* __restrict__ x, long long n, int sig, long long L, long long src_ext) {
    typedef typename scalar_of<C>::type real;
    const long long i0 = 2 * n, i1 = 2 * n + 1;
    real a = (real)0, b = (real)0;
    if (sig == 0) {
        if (i0 < src_ext) a = (real)ldg(x + (i0 < 2 * L ? i0 : i0 - 2 * L));
        if (i1 < src_ext) b = (real)ldg(x + (i1 < 2 * L ? i1 : i1 - 2 * L));
    } else {
        if (i0 < L) a = (real)ldg(x + i0);
        if (i1 < L) b = (real)ldg(x + i1);
    }
    return cmake(a, b);
}

// Runs f(IC<R>) for the runtime radix r (one of the supported set).
template <class F>
ASC_HD void gen_dispatch_radix(int r, F&& f) {
    switch (r) {
        case 2: f(IC<2>{}); break;
        case 3: f(IC<3>{}); break;
        case 4: f(IC<4>{}); break;
        case 5: f(IC<5>{}); break;
        case 6: f(IC<6>{}); break;
        case 8: f(IC<8>{}); break;
        case 9: f(IC<9>{}); break;
        case 10: f(IC<10>{}); break;
        case 12: f(IC<12>{}); break;
        case 15: f(IC<15>{}); break;
        default: f(IC<16>{}); break;
    }
}
constexpr bool gen_radix_supported(int r) {
    return r == 2 || r == 3 || r == 4 || r == 5 || r == 6 || r == 8 || r == 9 || r == 10 || r == 12 || r == 15 || r == 16;
}

// --------------------------------------------------------------------- G_A
// Forward column pass of both signals: grid = (ceil(M2 / CT), 2, pairs).
template <typename T, typename InT>
struct GenColFwdKernel {
    typedef typename GenTraits<T>::C C;
    static constexpr int CT = GenTraits<T>::CT;
    static constexpr int THREADS = GEN_THREADS;
    static constexpr int MIN_CTAS = sizeof(T) == 4 ? 2 : 1;

    struct Params {
        const InT* sources;
        const InT* samples;
        C* planes;               // [pair][2][M]
        PairPeak* peaks;         // cleared here
        const C* wcol;           // exp(-2*pi*i*t/M1), t < M1
        const C* m_lo;           // W_M two-level tables
        const C* m_hi;
        const int* p2f_col;      // frequency held at position i after the DIF passes
        GenShape sh;
        long long src_pitch, smp_pitch;
    };
    static size_t smem_bytes(const GenShape& sh) { return (size_t)sh.M1 * CT * sizeof(C) + 16; }

    template <class Ex>
    static ASC_HD void run(Ex& ex, const Params& p, C* __restrict__ buf) {
        const GenShape& sh = p.sh;
        const int M1 = sh.M1, M2 = sh.M2;
        const int c0 = ex.bx() * CT;
        const int sig = ex.by();
        const long long pair = ex.bz();
        const InT* __restrict__ x = sig == 0 ? p.sources + pair * p.src_pitch : p.samples + pair * p.smp_pitch;
        C* __restrict__ out = p.planes + (pair * 2 + sig) * sh.M;
        const int P = sh.col.npass;
        for (int ps = 0; ps < P; ps++) {
            const int S = sh.col.stride[ps];
            const bool first = ps == 0, last = ps == P - 1;
            gen_dispatch_radix(sh.col.radix[ps], [&](auto RR) {
                constexpr int R = decltype(RR)::value;
                const int items = (M1 / R) * CT;
                const int tstep = M1 / (S * R);
                ex.phase([&](int tid) {
                    if (first && tid == 0 && ex.bx() == 0 && sig == 0) p.peaks[pair] = cleared_peak();
                    for (int w = tid; w < items; w += THREADS) {
                        const int c = w % CT, bf = w / CT;
                        const int n2 = c0 + c;
                        if (n2 >= M2) continue;
                        const int blk = bf / S, j = bf - blk * S;
                        const int i0 = blk * (S * R) + j;
                        C v[R];
                        static_for<0, R>([&](auto Q) {
                            constexpr int q = decltype(Q)::value;
                            if (first) v[q] = gen_load_point<C, InT>(x, (long long)(i0 + q * S) * M2 + n2, sig, sh.L, sh.src_ext);
                            else v[q] = buf[(i0 + q * S) * CT + c];
                        });
                        dft_reg<R, -1, C>(v);
                        if (!last) {
                            buf[i0 * CT + c] = v[0];
                            static_for<1, R>([&](auto K) {
                                constexpr int k = decltype(K)::value;
                                buf[(i0 + k * S) * CT + c] = cmul(v[k], ldg(p.wcol + j * k * tstep));
                            });
                        } else {
                            // S == 1: position i0 + k holds bin k1 = p2f[i0 + k]; times W_M^(n2 * k1)
                            static_for<0, R>([&](auto K) {
                                constexpr int k = decltype(K)::value;
                                const int k1 = ldg(p.p2f_col + i0 + k);
                                out[(long long)k1 * M2 + n2] = cmul(v[k], gen_tw2<C>(p.m_lo, p.m_hi, (unsigned)n2 * (unsigned)k1));
                            });
                        }
                    }
                });
            });
        }
    }
};

// --------------------------------------------------------------------- G_B
// Forward rows of both planes, split / conj-multiply / merge, inverse rows: grid = (M1/2 + 1, 1, pairs).
template <typename T>
struct GenRowFusedKernel {
    typedef typename GenTraits<T>::C C;
    static constexpr int PADSH = GenTraits<T>::PADSH;
    static constexpr int THREADS = GEN_THREADS;
    static constexpr int MIN_CTAS = sizeof(T) == 4 ? 2 : 1;

    struct Params {
        C* planes;
        const C* wrow;           // exp(-2*pi*i*t/M2), t < M2
        const C* m_lo;
        const C* m_hi;
        const int* p2f_row;      // frequency at position
        const int* f2p_row;      // position of frequency
        GenShape sh;
    };
    static ASC_HD int phys(int p) { return p + (p >> PADSH); }
    static ASC_HD int row_pitch(int M2) { return M2 + (M2 >> PADSH) + 1; }
    static size_t smem_bytes(const GenShape& sh) { return (size_t)4 * row_pitch(sh.M2) * sizeof(C) + 16; }

    template <class Ex>
    static ASC_HD void run(Ex& ex, const Params& p, C* __restrict__ buf) {
        const GenShape& sh = p.sh;
        const int M1 = sh.M1, M2 = sh.M2;
        const int RP = row_pitch(M2);
        const int r = ex.bx();
        const long long pair = ex.bz();
        const bool two = (r != 0) && (2 * r != M1);
        const int nrows = two ? 2 : 1;
        const int k1a = r, k1b = M1 - r;
        C* __restrict__ plane_s = p.planes + pair * 2 * sh.M;
        C* __restrict__ plane_p = plane_s + sh.M;
        const int P = sh.row.npass;

        // stage: slots 0,1 source rows (k1a, k1b); 2,3 sample rows
        ex.phase([&](int tid) {
            for (int slot = 0; slot < 4; slot++) {
                const int rr = slot & 1;
                if (rr == 1 && !two) continue;
                const C* __restrict__ g = (slot >= 2 ? plane_p : plane_s) + (long long)(rr ? k1b : k1a) * M2;
                C* __restrict__ row = buf + slot * RP;
                for (int e = tid; e < M2; e += THREADS) row[phys(e)] = ldg(g + e);
            }
        });
        // forward DIF on 2 * nrows rows (slots 0,1,2,3 or 0,2)
        for (int ps = 0; ps < P; ps++) {
            const int S = sh.row.stride[ps];
            gen_dispatch_radix(sh.row.radix[ps], [&](auto RR) {
                constexpr int R = decltype(RR)::value;
                const int per_row = M2 / R;
                const int items = per_row * 2 * nrows;
                const int tstep = M2 / (S * R);
                ex.phase([&](int tid) {
                    for (int w = tid; w < items; w += THREADS) {
                        const int bw = w / per_row, bf = w - bw * per_row;
                        C* __restrict__ row = buf + (two ? bw : 2 * bw) * RP;
                        const int blk = bf / S, j = bf - blk * S;
                        const int i0 = blk * (S * R) + j;
                        C v[R];
                        static_for<0, R>([&](auto Q) { v[decltype(Q)::value] = row[phys(i0 + decltype(Q)::value * S)]; });
                        dft_reg<R, -1, C>(v);
                        row[phys(i0)] = v[0];
                        static_for<1, R>([&](auto K) {
                            constexpr int k = decltype(K)::value;
                            row[phys(i0 + k * S)] = S > 1 ? cmul(v[k], ldg(p.wrow + j * k * tstep)) : v[k];
                        });
                    }
                });
            });
        }
        // split + conj-multiply (reference :232-233) + merge, into the source rows
        ex.phase([&](int tid) {
            const C wk1 = gen_tw2<C>(p.m_lo, p.m_hi, (unsigned)k1a);
            if (two) {
                C* __restrict__ zs_a = buf;
                C* __restrict__ zs_b = buf + RP;
                C* __restrict__ zp_a = buf + 2 * RP;
                C* __restrict__ zp_b = buf + 3 * RP;
                for (int e = tid; e < M2; e += THREADS) {
                    const int pa = phys(e), pb = phys(M2 - 1 - e);
                    const C w2 = cmul(ldg(p.wrow + ldg(p.p2f_row + e)), wk1);
                    C qk, qmk;
                    gen_split_mul_merge<C>(zs_a[pa], zs_b[pb], zp_a[pa], zp_b[pb], w2, qk, qmk);
                    zs_a[pa] = qk;
                    zs_b[pb] = qmk;
                }
            } else {
                C* __restrict__ zs = buf;
                C* __restrict__ zp = buf + 2 * RP;
                const int items = r == 0 ? M2 / 2 + 1 : (M2 + 1) / 2;
                for (int e = tid; e < items; e += THREADS) {
                    int ea, eb;      // positions
                    if (r == 0) {
                        ea = ldg(p.f2p_row + e);
                        eb = ldg(p.f2p_row + (e == 0 ? 0 : M2 - e));
                    } else {
                        ea = e;
                        eb = M2 - 1 - e;
                    }
                    const C w2 = cmul(ldg(p.wrow + ldg(p.p2f_row + ea)), wk1);
                    const int pa = phys(ea), pb = phys(eb);
                    C qk, qmk;
                    gen_split_mul_merge<C>(zs[pa], zs[pb], zp[pa], zp[pb], w2, qk, qmk);
                    zs[pa] = qk;
                    if (pa != pb) zs[pb] = qmk;
                }
            }
        });
        // inverse DIT on nrows rows (slots 0,1), passes P-1 .. 0
        for (int ps = P - 1; ps >= 0; ps--) {
            const int S = sh.row.stride[ps];
            gen_dispatch_radix(sh.row.radix[ps], [&](auto RR) {
                constexpr int R = decltype(RR)::value;
                const int per_row = M2 / R;
                const int items = per_row * nrows;
                const int tstep = M2 / (S * R);
                ex.phase([&](int tid) {
                    for (int w = tid; w < items; w += THREADS) {
                        const int rw = w / per_row, bf = w - rw * per_row;
                        C* __restrict__ row = buf + rw * RP;
                        const int blk = bf / S, j = bf - blk * S;
                        const int i0 = blk * (S * R) + j;
                        C v[R];
                        v[0] = row[phys(i0)];
                        static_for<1, R>([&](auto Q) {
                            constexpr int q = decltype(Q)::value;
                            const C xq = row[phys(i0 + q * S)];
                            v[q] = S > 1 ? cmulc(xq, ldg(p.wrow + j * q * tstep)) : xq;
                        });
                        dft_reg<R, +1, C>(v);
                        static_for<0, R>([&](auto K) { row[phys(i0 + decltype(K)::value * S)] = v[decltype(K)::value]; });
                    }
                });
            });
        }
        // natural order now: times conj W_M^(n2 * k1), back in place (row k1 of plane 0)
        ex.phase([&](int tid) {
            for (int w = tid; w < M2 * nrows; w += THREADS) {
                const int rw = w / M2, e = w - rw * M2;
                const int k1 = rw ? k1b : k1a;
                plane_s[(long long)k1 * M2 + e] = cmulc(buf[rw * RP + phys(e)], gen_tw2<C>(p.m_lo, p.m_hi, (unsigned)e * (unsigned)k1));
            }
        });
    }
};

// --------------------------------------------------------------------- G_C
// Inverse column pass.  fp32: |r| argmax epilogue over the indices < limit (= 2L), the correlation
// is never written.  fp64: r[0 .. limit) is written to `r_out` (the dead sample plane) and resolved
// by argmax_f64_kernel, which keeps full double keys.  grid = (pairs, ceil(M2 / CT)).
template <typename T>
struct GenColInvKernel {
    typedef typename GenTraits<T>::C C;
    static constexpr int CT = GenTraits<T>::CT;
    static constexpr int THREADS = GEN_THREADS;
    static constexpr int MIN_CTAS = sizeof(T) == 4 ? 2 : 1;

    struct Params {
        const C* planes;
        PairPeak* peaks;
        const C* wcol;
        const int* p2f_col;
        GenShape sh;
        T* r_out;                // fp64 only: [pair][2M] reals (aliases plane 1 of the pair)
    };
    static size_t smem_bytes(const GenShape& sh) {
        const size_t tile = (size_t)sh.M1 * CT * sizeof(C);
        return (tile > 512 ? tile : 512) + 16;
    }

    template <class Ex>
    static ASC_HD void run(Ex& ex, const Params& p, C* __restrict__ buf) {
        const GenShape& sh = p.sh;
        const int M1 = sh.M1, M2 = sh.M2;
        const int c0 = ex.by() * CT;
        const long long pair = ex.bx();
        const C* __restrict__ in = p.planes + pair * 2 * sh.M;
        const long long limit = 2 * sh.L;
        const int P = sh.col.npass;
        for (int ps = 0; ps < P - 1; ps++) {
            const int S = sh.col.stride[ps];
            const bool first = ps == 0;
            gen_dispatch_radix(sh.col.radix[ps], [&](auto RR) {
                constexpr int R = decltype(RR)::value;
                const int items = (M1 / R) * CT;
                const int tstep = M1 / (S * R);
                ex.phase([&](int tid) {
                    for (int w = tid; w < items; w += THREADS) {
                        const int c = w % CT, bf = w / CT;
                        const int n2 = c0 + c;
                        if (n2 >= M2) continue;
                        const int blk = bf / S, j = bf - blk * S;
                        const int i0 = blk * (S * R) + j;
                        C v[R];
                        static_for<0, R>([&](auto Q) {
                            constexpr int q = decltype(Q)::value;
                            if (first) v[q] = ldg(in + (long long)(i0 + q * S) * M2 + n2);
                            else v[q] = buf[(i0 + q * S) * CT + c];
                        });
                        dft_reg<R, +1, C>(v);
                        buf[i0 * CT + c] = v[0];
                        static_for<1, R>([&](auto K) {
                            constexpr int k = decltype(K)::value;
                            buf[(i0 + k * S) * CT + c] = cmulc(v[k], ldg(p.wcol + j * k * tstep));
                        });
                    }
                });
            });
        }
        // last pass (S == 1): packed point n = n1 * M2 + n2 carries r[2n] and r[2n + 1]
        const bool only = P == 1;
        gen_dispatch_radix(sh.col.radix[P - 1], [&](auto RR) {
            constexpr int R = decltype(RR)::value;
            const int items = (M1 / R) * CT;
            auto last_pass = [&](int tid, auto&& emit) {
                for (int w = tid; w < items; w += THREADS) {
                    const int c = w % CT, blk = w / CT;
                    const int n2 = c0 + c;
                    if (n2 >= M2) continue;
                    const int i0 = blk * R;
                    C v[R];
                    static_for<0, R>([&](auto Q) {
                        constexpr int q = decltype(Q)::value;
                        if (only) v[q] = ldg(in + (long long)(i0 + q) * M2 + n2);
                        else v[q] = buf[(i0 + q) * CT + c];
                    });
                    dft_reg<R, +1, C>(v);
                    static_for<0, R>([&](auto K) {
                        constexpr int k = decltype(K)::value;
                        const long long n = (long long)ldg(p.p2f_col + i0 + k) * M2 + n2;
                        emit(2 * n, v[k]);
                    });
                }
            };
            if constexpr (sizeof(T) == 4) {
                ex.phase_argmax(
                    [&](int tid) -> ArgmaxPair {
                        ArgmaxAcc acc;
                        last_pass(tid, [&](long long i_re, C val) {
                            if (i_re < limit) {
                                if (i_re == 0) acc.consider_seed(val.x);
                                else acc.consider(val.x, (uint32_t)i_re);
                            }
                            if (i_re + 1 < limit) acc.consider(val.y, (uint32_t)(i_re + 1));
                        });
                        return acc.result();
                    },
                    &p.peaks[pair].key, &p.peaks[pair].second_bits, buf);
            } else {
                T* __restrict__ r = p.r_out + (pair * 4 + 2) * sh.M;     // plane 1 of the pair, as reals
                ex.phase([&](int tid) {
                    last_pass(tid, [&](long long i_re, C val) {
                        if (i_re < limit) r[i_re] = val.x;
                        if (i_re + 1 < limit) r[i_re + 1] = val.y;
                    });
                });
            }
        });
    }
};

// Entry for kernels whose shared-memory buffer is not cplx-typed.
#if defined(__CUDACC__)
template <class K>
__global__ void __launch_bounds__(K::THREADS, K::MIN_CTAS)
gen_kernel_entry(const typename K::Params p) {
    extern __shared__ __align__(128) unsigned char asc_smem_gen[];
    pdl_prologue();
    DeviceExec ex;
    ex.x_ = (int)blockIdx.x; ex.y_ = (int)blockIdx.y; ex.z_ = (int)blockIdx.z;
    K::run(ex, p, reinterpret_cast<typename K::C*>(asc_smem_gen));
}
#endif

}  // namespace asc
