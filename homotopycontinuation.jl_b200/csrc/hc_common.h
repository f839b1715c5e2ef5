// Scalar numerics of the B200 path tracker: complex fp64, double-double, complex double-double.
// Device code (sm_100a).  All functions are HC_HD so that tests/host_sim can compile the very
// same kernel logic with g++ for CPU-side unit tests (never part of libhc_b200.so).
//
// Semantics follow the reference (file:line):
//   src/DoubleDouble.jl:13-66, 185-191, 223-231, 247-266, 311-327, 364-370  double-double
//   src/model_kit/operations.jl:184-248                                     op_* kernels
//   Julia Base.FastMath.div_fast / inv_fast                                 unscaled complex division
#pragma once
#if defined(__CUDACC_RTC__)
// NVRTC (the per-system specialised kernels of hc_jit.h) has no host headers: the math functions are built in
typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t; typedef unsigned long size_t;
#else
#include <cmath>
#include <cstdint>
#endif

#if defined(__CUDACC__)
#define HC_HD __host__ __device__ __forceinline__
#define HC_HDN __host__ __device__ __noinline__
#define HC_D __device__ __forceinline__
#else
#define HC_HD inline
#define HC_HDN inline
#define HC_D inline
#endif

namespace hc {

#define HC_EPS 2.220446049250313e-16
#if defined(__CUDA_ARCH__)
#define HC_INF (__longlong_as_double(0x7ff0000000000000LL))
#define HC_NAN (__longlong_as_double(0x7ff8000000000000LL))
#else
#define HC_INF (__builtin_inf())
#define HC_NAN (__builtin_nan(""))
#endif

// ---- complex fp64 (layout = double2 = Julia ComplexF64)
struct alignas(16) cx {
    double re, im;
};
HC_HD cx mk(double r, double i = 0.0) { cx z; z.re = r; z.im = i; return z; }
HC_HD cx operator+(cx a, cx b) { return mk(a.re + b.re, a.im + b.im); }
HC_HD cx operator-(cx a, cx b) { return mk(a.re - b.re, a.im - b.im); }
HC_HD cx operator-(cx a) { return mk(-a.re, -a.im); }
HC_HD cx operator*(cx a, cx b) { return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
HC_HD cx operator*(double a, cx b) { return mk(a * b.re, a * b.im); }
HC_HD cx operator*(cx b, double a) { return mk(a * b.re, a * b.im); }
HC_HD cx operator/(cx a, double b) { return mk(a.re / b, a.im / b); }
HC_HD cx conj(cx a) { return mk(a.re, -a.im); }
HC_HD double abs2(cx a) { return a.re * a.re + a.im * a.im; }
HC_HD double cabs(cx a) { return sqrt(abs2(a)); }  // reference fast_abs (src/utils.jl:207)
HC_HD bool cisnan(cx a) { return a.re != a.re || a.im != a.im; }
HC_HD bool ciszero(cx a) { return a.re == 0.0 && a.im == 0.0; }
// a * b + c
HC_HD cx cfma(cx a, cx b, cx c) {
    return mk(c.re + a.re * b.re - a.im * b.im, c.im + a.re * b.im + a.im * b.re);
}
// c - a * b
HC_HD cx cfnma(cx a, cx b, cx c) {
    return mk(c.re - a.re * b.re + a.im * b.im, c.im - a.re * b.im - a.im * b.re);
}
HC_HD cx cdiv(cx a, cx b) {  // div_fast: a * conj(b) / abs2(b)
    double d = 1.0 / abs2(b);
    return mk((a.re * b.re + a.im * b.im) * d, (a.im * b.re - a.re * b.im) * d);
}
HC_HD cx cinv(cx a) {
    double d = 1.0 / abs2(a);
    return mk(a.re * d, -a.im * d);
}
HC_HD cx csqr(cx z) { return mk((z.re + z.im) * (z.re - z.im), (z.re + z.re) * z.im); }
HC_HD cx ccb(cx z) {
    double a = (z.re + z.im) * (z.re - z.im), b = (z.re + z.re) * z.im;
    return mk(a * z.re - b * z.im, a * z.im + b * z.re);
}
HC_HD cx cpowi(cx x, int p) {  // Base.power_by_squaring; p < 0 -> inverse; p == 0 -> 1
    if (p == 0) return mk(1.0);
    int q = p < 0 ? -p : p;
    cx y = mk(1.0);
    bool have = false;
    cx b = x;
    while (q) {
        if (q & 1) { y = have ? y * b : b; have = true; }
        q >>= 1;
        if (q) b = csqr(b);
    }
    return p < 0 ? cinv(y) : y;
}

HC_HD double nanmin(double a, double b) { return a != a ? b : (b != b ? a : (a < b ? a : b)); }
HC_HD double nanmax(double a, double b) { return a != a ? b : (b != b ? a : (a > b ? a : b)); }
HC_HD double jmin(double a, double b) { return (a != a || b != b) ? HC_NAN : (a < b ? a : b); }
HC_HD double jmax(double a, double b) { return (a != a || b != b) ? HC_NAN : (a > b ? a : b); }
HC_HD double fmaxq(double a, double b) { return b > a ? b : a; }  // FastMath.max_fast
HC_HD double clampd(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }
HC_HD double nthroot(double x, int N) {  // src/utils.jl:408-422
    switch (N) {
        case 4: return sqrt(sqrt(x));
        case 2: return sqrt(x);
        case 3: return cbrt(x);
        case 1: return x;
        case 0: return 1.0;
        default: return pow(x, 1.0 / N);
    }
}
// cos(pi x), sin(pi x): CUDA built-ins; the g++ build of the device code takes glibc's cos / sin of pi x (|x| <= 2 here)
#if defined(__CUDACC__)
HC_HD double hc_cospi(double x) { return ::cospi(x); }
HC_HD double hc_sinpi(double x) { return ::sinpi(x); }
#else
inline double hc_cospi(double x) { return cos(3.141592653589793 * x); }
inline double hc_sinpi(double x) { return sin(3.141592653589793 * x); }
#endif
HC_HD double eps_of(double x) {  // Julia eps(x): ulp of |x|
    x = fabs(x);
    if (!(x < HC_INF)) return HC_NAN;
    if (x < 2.2250738585072014e-308) return 4.9406564584124654e-324;
    int e;
    frexp(x, &e);
    return ldexp(1.0, e - 53);
}

// ---- double-double
struct dd { double hi, lo; };
HC_HD dd mkdd(double h, double l = 0.0) { dd r; r.hi = h; r.lo = l; return r; }
HC_HD dd quick_two_sum(double a, double b) { double s = a + b; return mkdd(s, b - (s - a)); }
HC_HD dd two_sum(double a, double b) { double s = a + b, v = s - a; return mkdd(s, (a - (s - v)) + (b - v)); }
HC_HD dd two_diff(double a, double b) { double s = a - b, v = s - a; return mkdd(s, (a - (s - v)) - (b + v)); }
HC_HD dd two_prod(double a, double b) { double p = a * b; return mkdd(p, fma(a, b, -p)); }
HC_HD dd operator+(dd a, dd b) { dd s = two_sum(a.hi, b.hi); s.lo += (a.lo + b.lo); return quick_two_sum(s.hi, s.lo); }
HC_HD dd operator-(dd a, dd b) { dd s = two_diff(a.hi, b.hi); s.lo += a.lo; s.lo -= b.lo; return quick_two_sum(s.hi, s.lo); }
HC_HD dd operator-(dd a) { return mkdd(-a.hi, -a.lo); }
HC_HD dd operator*(dd a, dd b) { dd p = two_prod(a.hi, b.hi); p.lo += a.hi * b.lo + a.lo * b.hi; return quick_two_sum(p.hi, p.lo); }
HC_HD dd operator*(dd a, double b) { dd p = two_prod(a.hi, b); p.lo += a.lo * b; return quick_two_sum(p.hi, p.lo); }
HC_HD dd operator/(dd a, dd b) {
    double q1 = a.hi / b.hi;
    dd r = b * q1;
    dd s = two_diff(a.hi, r.hi);
    s.lo -= r.lo; s.lo += a.lo;
    double q2 = (s.hi + s.lo) / b.hi;
    return quick_two_sum(q1, q2);
}
HC_HD dd ddsqr(dd a) { dd p = two_prod(a.hi, a.hi); p.lo += 2.0 * a.hi * a.lo; p.lo += a.lo * a.lo; return quick_two_sum(p.hi, p.lo); }

// ---- complex double-double (layout: re.hi, re.lo, im.hi, im.lo = 32 B)
struct alignas(16) cdd { dd re, im; };
HC_HD cdd mkcdd(dd r, dd i) { cdd z; z.re = r; z.im = i; return z; }
HC_HD cdd tocdd(cx z) { return mkcdd(mkdd(z.re), mkdd(z.im)); }
HC_HD cx tocx(cdd z) { return mk(z.re.hi, z.im.hi); }
HC_HD cdd operator+(cdd a, cdd b) { return mkcdd(a.re + b.re, a.im + b.im); }
HC_HD cdd operator-(cdd a, cdd b) { return mkcdd(a.re - b.re, a.im - b.im); }
HC_HD cdd operator-(cdd a) { return mkcdd(-a.re, -a.im); }
HC_HD cdd operator*(cdd a, cdd b) { return mkcdd(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
HC_HD cdd mulc(cdd a, cx b) {  // cdd * ComplexF64 (exactly the promoted product)
    return mkcdd(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
HC_HD dd abs2(cdd a) { return a.re * a.re + a.im * a.im; }
HC_HD cdd cdiv(cdd a, cdd b) {
    dd d = abs2(b);
    return mkcdd((a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d);
}
HC_HD cdd cinv(cdd a) { dd d = abs2(a); return mkcdd(a.re / d, -(a.im / d)); }
HC_HD cdd csqr(cdd z) { return mkcdd((z.re + z.im) * (z.re - z.im), (z.re + z.re) * z.im); }
HC_HD cdd ccb(cdd z) {
    dd a = (z.re + z.im) * (z.re - z.im), b = (z.re + z.re) * z.im;
    return mkcdd(a * z.re - b * z.im, a * z.im + b * z.re);
}
HC_HD bool ciszero(cdd a) { return a.re.hi == 0.0 && a.re.lo == 0.0 && a.im.hi == 0.0 && a.im.lo == 0.0; }
HC_HD cdd cpowi(cdd x, int p) {
    if (p == 0) return tocdd(mk(1.0));
    int q = p < 0 ? -p : p;
    cdd y = x;
    bool have = false;
    cdd b = x;
    while (q) {
        if (q & 1) { y = have ? y * b : b; have = true; }
        q >>= 1;
        if (q) b = b * b;
    }
    return p < 0 ? cinv(y) : y;
}

}  // namespace hc
