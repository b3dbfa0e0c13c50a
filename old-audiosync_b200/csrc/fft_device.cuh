// fft_device.cuh -- building blocks of the sm_100a transform kernels:
// compile-time twiddle constants, in-register DFT butterflies (radix 2/3/4/5
// by hand, composites by prime-factor / Cooley-Tukey recursion with every
// index resolved at compile time), radix lists and digit-reversal maps.
//
// Everything here is __host__ __device__ so that tests/emu can execute the
// exact kernel bodies on the CPU (phase by phase) and compare them with the
// oracle before any GPU time is spent.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>
#include <utility>

#define ASC_HD __host__ __device__ __forceinline__

namespace asc {

typedef float2 cplx;

ASC_HD cplx cmake(float a, float b) { cplx r; r.x = a; r.y = b; return r; }
ASC_HD cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
ASC_HD cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
ASC_HD cplx cconj(cplx a) { return cmake(a.x, -a.y); }
ASC_HD cplx cscale(cplx a, float s) { return cmake(a.x * s, a.y * s); }
// a * b
ASC_HD cplx cmul(cplx a, cplx b) {
    return cmake(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x));
}
// a * conj(b)
ASC_HD cplx cmulc(cplx a, cplx b) {
    return cmake(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -(a.x * b.y)));
}
// multiply by +i / -i
ASC_HD cplx cmul_pi(cplx a) { return cmake(-a.y, a.x); }
ASC_HD cplx cmul_ni(cplx a) { return cmake(a.y, -a.x); }

template <typename T>
ASC_HD T ldg(const T* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// --------------------------------------------------------------- static_for
template <int I>
using IC = std::integral_constant<int, I>;

template <int B, int E, class F>
ASC_HD void static_for(F&& f) {
    if constexpr (B < E) {
        f(IC<B>{});
        static_for<B + 1, E>(static_cast<F&&>(f));
    }
}

// ------------------------------------------------- compile-time trigonometry
// cos/sin of 2*pi*num/den evaluated in constant expressions only (the results
// become immediates in the SASS).  Range-reduced to [0, pi/4] with integer
// arithmetic, then a Taylor series in double: error < 1e-16.
namespace ct {
constexpr double PI = 3.14159265358979323846264338327950288;

constexpr double sin_small(double x) {  // |x| <= pi/4
    double x2 = x * x, term = x, sum = x;
    for (int k = 1; k < 12; k++) {
        term *= -x2 / (double)((2 * k) * (2 * k + 1));
        sum += term;
    }
    return sum;
}
constexpr double cos_small(double x) {
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int k = 1; k < 12; k++) {
        term *= -x2 / (double)((2 * k - 1) * (2 * k));
        sum += term;
    }
    return sum;
}
struct cd { double re, im; };
// exp(+2*pi*i*num/den), num any integer, den > 0
constexpr cd unit(long long num, long long den) {
    long long a = num % den;
    if (a < 0) a += den;
    // quadrant q = floor(4a/den); remainder angle = 2*pi*(4a - q*den)/(4*den)
    long long q = (4 * a) / den;
    long long rnum = 4 * a - q * den;          // in [0, den)
    double c = 0.0, s = 0.0;
    if (2 * rnum <= den) {                     // angle <= pi/4
        double th = 2.0 * PI * (double)rnum / (4.0 * (double)den);
        c = cos_small(th); s = sin_small(th);
    } else {                                   // use the complement to pi/2
        double th = 2.0 * PI * (double)(den - rnum) / (4.0 * (double)den);
        c = sin_small(th); s = cos_small(th);
    }
    if (rnum == 0) { c = 1.0; s = 0.0; }
    switch (q) {
        case 0: return cd{c, s};
        case 1: return cd{-s, c};
        case 2: return cd{-c, -s};
        default: return cd{s, -c};
    }
}
constexpr int gcd(int a, int b) { return b == 0 ? a : gcd(b, a % b); }
constexpr int first_factor(int r) {
    // split composites so that coprime (prime-factor) splits are preferred
    // and the sub-transforms stay in the hand-written set {2,3,4,5}.
    if (r % 4 == 0 && r != 4) return 4;
    if (r % 2 == 0) return 2;
    if (r % 3 == 0) return 3;
    if (r % 5 == 0) return 5;
    return r;
}
}  // namespace ct

// Multiply v by exp(DIR * 2*pi*i * T / R) with the constant folded; exact
// cases (1, -1, +-i, the eighth roots) use adds only.
template <int T_, int R, int DIR>
ASC_HD cplx mul_root(cplx v) {
    constexpr int t = ((T_ % R) + R) % R;
    if constexpr (t == 0) {
        return v;
    } else if constexpr (2 * t == R) {
        return cmake(-v.x, -v.y);
    } else if constexpr (4 * t == R) {
        return DIR > 0 ? cmul_pi(v) : cmul_ni(v);
    } else if constexpr (4 * t == 3 * R) {
        return DIR > 0 ? cmul_ni(v) : cmul_pi(v);
    } else {
        constexpr ct::cd w = ct::unit((long long)DIR * t, R);
        constexpr float wr = (float)w.re, wi = (float)w.im;
        if constexpr (8 * t == R || 8 * t == 3 * R || 8 * t == 5 * R || 8 * t == 7 * R) {
            // |wr| == |wi| == sqrt(1/2)
            constexpr float h = 0.70710678118654752440f;
            constexpr float sr = wr > 0 ? 1.f : -1.f, si = wi > 0 ? 1.f : -1.f;
            // (x + iy)(sr*h + i*si*h) = h*[(sr*x - si*y) + i(si*x + sr*y)]
            return cmake(h * (sr * v.x - si * v.y), h * (si * v.x + sr * v.y));
        } else {
            return cmake(fmaf(v.x, wr, -(v.y * wi)), fmaf(v.x, wi, v.y * wr));
        }
    }
}

// ------------------------------------------------------ register butterflies
// dft_reg<R, DIR>(v): v[k] <- sum_n v[n] exp(DIR*2*pi*i*n*k/R), natural order
// in and out, all indices compile-time (v stays in registers).
template <int R, int DIR>
struct DftReg;

template <int DIR>
struct DftReg<1, DIR> {
    static ASC_HD void run(cplx (&)[1]) {}
};

template <int DIR>
struct DftReg<2, DIR> {
    static ASC_HD void run(cplx (&v)[2]) {
        cplx a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <int DIR>
struct DftReg<3, DIR> {
    static ASC_HD void run(cplx (&v)[3]) {
        constexpr float s60 = (DIR > 0 ? 1.f : -1.f) * 0.86602540378443864676f;
        cplx a = v[0], b = v[1], c = v[2];
        cplx t1 = cadd(b, c);
        cplx t2 = cmake(fmaf(-0.5f, t1.x, a.x), fmaf(-0.5f, t1.y, a.y));
        cplx d = csub(b, c);
        cplx t3 = cmake(s60 * d.x, s60 * d.y);
        v[0] = cadd(a, t1);
        v[1] = cmake(t2.x - t3.y, t2.y + t3.x);
        v[2] = cmake(t2.x + t3.y, t2.y - t3.x);
    }
};

template <int DIR>
struct DftReg<4, DIR> {
    static ASC_HD void run(cplx (&v)[4]) {
        cplx s0 = cadd(v[0], v[2]), s1 = csub(v[0], v[2]);
        cplx s2 = cadd(v[1], v[3]), s3 = csub(v[1], v[3]);
        cplx r3 = DIR > 0 ? cmul_pi(s3) : cmul_ni(s3);
        v[0] = cadd(s0, s2);
        v[1] = cadd(s1, r3);
        v[2] = csub(s0, s2);
        v[3] = csub(s1, r3);
    }
};

template <int DIR>
struct DftReg<5, DIR> {
    static ASC_HD void run(cplx (&v)[5]) {
        constexpr float c1 = 0.30901699437494742410f;    // cos(2pi/5)
        constexpr float c2 = -0.80901699437494742410f;   // cos(4pi/5)
        constexpr float sg = DIR > 0 ? 1.f : -1.f;
        constexpr float s1 = sg * 0.95105651629515357212f;  // sin(2pi/5)
        constexpr float s2 = sg * 0.58778525229247312917f;  // sin(4pi/5)
        cplx x0 = v[0];
        cplx a1 = cadd(v[1], v[4]), b1 = csub(v[1], v[4]);
        cplx a2 = cadd(v[2], v[3]), b2 = csub(v[2], v[3]);
        cplx p1 = cmake(fmaf(c2, a2.x, fmaf(c1, a1.x, x0.x)), fmaf(c2, a2.y, fmaf(c1, a1.y, x0.y)));
        cplx p2 = cmake(fmaf(c1, a2.x, fmaf(c2, a1.x, x0.x)), fmaf(c1, a2.y, fmaf(c2, a1.y, x0.y)));
        cplx u1 = cmake(fmaf(s2, b2.x, s1 * b1.x), fmaf(s2, b2.y, s1 * b1.y));
        cplx u2 = cmake(fmaf(-s1, b2.x, s2 * b1.x), fmaf(-s1, b2.y, s2 * b1.y));
        v[0] = cmake(x0.x + a1.x + a2.x, x0.y + a1.y + a2.y);
        v[1] = cmake(p1.x - u1.y, p1.y + u1.x);
        v[4] = cmake(p1.x + u1.y, p1.y - u1.x);
        v[2] = cmake(p2.x - u2.y, p2.y + u2.x);
        v[3] = cmake(p2.x + u2.y, p2.y - u2.x);
    }
};

// Composite R = A * B.
//  gcd(A,B) == 1: prime-factor (Good-Thomas) map, no twiddles:
//      n = (B*n1 + A*n2) mod R,  k = the unique k with k%A == k1, k%B == k2.
//  otherwise: Cooley-Tukey, n = A*n2 + n1, k = B*k1 + k2, twiddle W_R^(n1*k2).
template <int R, int DIR>
struct DftReg {
    static constexpr int A = ct::first_factor(R);
    static constexpr int B = R / A;
    static_assert(A > 1 && A < R, "unsupported radix (prime > 5)");
    static constexpr bool PFA = ct::gcd(A, B) == 1;

    static constexpr int crt(int k1, int k2) {
        for (int k = 0; k < R; k++)
            if (k % A == k1 && k % B == k2) return k;
        return -1;
    }

    static ASC_HD void run(cplx (&v)[R]) {
        cplx t[R];   // t[n1 * B + k2]
        static_for<0, A>([&](auto N1) {
            constexpr int n1 = decltype(N1)::value;
            cplx u[B];
            static_for<0, B>([&](auto N2) {
                constexpr int n2 = decltype(N2)::value;
                constexpr int src = PFA ? (B * n1 + A * n2) % R : (A * n2 + n1);
                u[n2] = v[src];
            });
            DftReg<B, DIR>::run(u);
            static_for<0, B>([&](auto K2) {
                constexpr int k2 = decltype(K2)::value;
                if constexpr (PFA) t[n1 * B + k2] = u[k2];
                else t[n1 * B + k2] = mul_root<n1 * k2, R, DIR>(u[k2]);
            });
        });
        static_for<0, B>([&](auto K2) {
            constexpr int k2 = decltype(K2)::value;
            cplx u[A];
            static_for<0, A>([&](auto N1) {
                constexpr int n1 = decltype(N1)::value;
                u[n1] = t[n1 * B + k2];
            });
            DftReg<A, DIR>::run(u);
            static_for<0, A>([&](auto K1) {
                constexpr int k1 = decltype(K1)::value;
                constexpr int dst = PFA ? crt(k1, k2) : (B * k1 + k2);
                v[dst] = u[k1];
            });
        });
    }
};

template <int R, int DIR>
ASC_HD void dft_reg(cplx (&v)[R]) { DftReg<R, DIR>::run(v); }

// w[k] = exp(-2*pi*i*j*k/(S*R)) for k = 1..R-1 from the power-of-two table
// entries of one pass (tw points at the pass's table): k = 2^i is a load,
// any other k is the product w[hb(k)] * w[k - hb(k)] (at most 3 products deep
// for R <= 16, i.e. a few ulp).  Trades LSU wavefronts for FMA-pipe work.
template <int R>
ASC_HD void pass_twiddles(const cplx* __restrict__ tw, int S, int j, cplx (&w)[R]) {
    static_for<1, R>([&](auto K) {
        constexpr int k = decltype(K)::value;
        if constexpr ((k & (k - 1)) == 0) {
            constexpr int i = (k == 1) ? 0 : (k == 2) ? 1 : (k == 4) ? 2 : (k == 8) ? 3 : (k == 16) ? 4 : 5;
            w[k] = ldg(tw + i * S + j);
        } else {
            constexpr int hb = (k >= 16) ? 16 : (k >= 8) ? 8 : (k >= 4) ? 4 : 2;
            w[k] = cmul(w[hb], w[k - hb]);
        }
    });
}

// ---------------------------------------------------------------- radix list
// In-place decimation-in-frequency order: pass p has radix r(p) and
// sub-stride s(p) = n / (r(0) * ... * r(p)); after all passes position
// i = sum_p d_p * s(p) holds frequency k = sum_p d_p * (r(0)*...*r(p-1)).
template <int... Rs>
struct RadixList {
    static constexpr int count = sizeof...(Rs);
    static constexpr int n = (Rs * ... * 1);
    static constexpr int r(int p) {
        constexpr int a[] = {Rs...};
        return a[p];
    }
    static constexpr int stride(int p) {   // s(p)
        int prod = 1;
        for (int i = 0; i <= p; i++) prod *= r(i);
        return n / prod;
    }
    static constexpr int weight(int p) {   // r(0) * ... * r(p-1)
        int prod = 1;
        for (int i = 0; i < p; i++) prod *= r(i);
        return prod;
    }
    // Pass p's twiddle table holds only the power-of-two multiples
    //   tw[i*s + j] = exp(-2*pi*i * j * 2^i / (s*r)),  2^i < r,  j in [0,s);
    // the other multiples are products of two table values (pass_twiddles).
    static constexpr int npow(int radix) {
        int n = 0;
        for (int k = 1; k < radix; k *= 2) n++;
        return n;
    }
    static constexpr int tw_offset(int p) {
        int off = 0;
        for (int i = 0; i < p; i++) off += npow(r(i)) * stride(i);
        return off;
    }
    static constexpr int tw_total() { return tw_offset(count); }
    static constexpr int max_radix() {
        int m = 0;
        for (int i = 0; i < count; i++) m = r(i) > m ? r(i) : m;
        return m;
    }
    // position of frequency k after the DIF passes
    static ASC_HD int pos_of_freq(int k) {
        int pos = 0;
        static_for<0, count>([&](auto P) {
            constexpr int p = decltype(P)::value;
            constexpr int rp = r(p);
            int d = k % rp;
            k /= rp;
            pos += d * stride(p);
        });
        return pos;
    }
    // frequency held at position i after the DIF passes
    static ASC_HD int freq_of_pos(int i) {
        int k = 0;
        static_for<0, count>([&](auto P) {
            constexpr int p = decltype(P)::value;
            constexpr int sp = stride(p);
            int d = i / sp;
            i -= d * sp;
            k += d * weight(p);
        });
        return k;
    }
};

}  // namespace asc
