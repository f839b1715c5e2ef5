// ORACLE (test infrastructure, NOT product code): CPU restatement of the scalar
// numerics of HomotopyContinuation.jl v2.20.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build or call this.
//
// Follows (reference file:line):
//   src/DoubleDouble.jl:13-66   error-free transforms
//   src/DoubleDouble.jl:185-191 (+), :223-231 (-), :260-266 (*), :311-327 (/),
//   src/DoubleDouble.jl:364-370 square, :382-409 power_by_squaring
//   src/model_kit/operations.jl:184-248 op_* kernels (complex specialisations)
//   src/utils.jl:207-212 fast_abs / nanmin / nanmax, :408-422 nthroot
//   Julia Base: Complex `*` (4 mul, 2 add, no contraction), FastMath.div_fast /
//   inv_fast (multiply by conjugate, divide by abs2), Base.power_by_squaring.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace orc {

constexpr double EPS = 2.220446049250313e-16;  // Julia eps(Float64)
constexpr double INF = std::numeric_limits<double>::infinity();
constexpr double NaN = std::numeric_limits<double>::quiet_NaN();

// ---------------------------------------------------------------- ComplexF64
struct cplx {
    double re, im;
    cplx() : re(0), im(0) {}
    cplx(double r) : re(r), im(0) {}
    cplx(double r, double i) : re(r), im(i) {}
};
inline cplx operator+(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
inline cplx operator-(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }
inline cplx operator-(cplx a) { return {-a.re, -a.im}; }
// Julia Base complex multiply: no FMA contraction (build with -ffp-contract=off)
inline cplx operator*(cplx a, cplx b) {
    return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
inline cplx operator*(double a, cplx b) { return {a * b.re, a * b.im}; }
inline cplx operator*(cplx b, double a) { return {a * b.re, a * b.im}; }
inline cplx operator/(cplx a, double b) { return {a.re / b, a.im / b}; }
inline cplx& operator+=(cplx& a, cplx b) { a = a + b; return a; }
inline cplx& operator-=(cplx& a, cplx b) { a = a - b; return a; }
inline cplx conj(cplx a) { return {a.re, -a.im}; }
inline double abs2(cplx a) { return a.re * a.re + a.im * a.im; }
inline double abs2(double a) { return a * a; }
// src/utils.jl:207 fast_abs(z) = sqrt(abs2(z))
inline double fast_abs(cplx a) { return std::sqrt(abs2(a)); }
inline double fast_abs(double a) { return std::fabs(a); }
// Base abs(::Complex) = hypot
inline double habs(cplx a) { return std::hypot(a.re, a.im); }
inline bool isnan(cplx a) { return std::isnan(a.re) || std::isnan(a.im); }
inline bool iszero(cplx a) { return a.re == 0.0 && a.im == 0.0; }
// Base.FastMath.div_fast(x::Complex, y::Complex) = x*conj(y) / abs2(y)
inline cplx div_fast(cplx a, cplx b) {
    double d = abs2(b);
    return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}
inline cplx inv_fast(cplx a) {
    double d = abs2(a);
    return {a.re / d, -a.im / d};
}
// Base `/` for Complex{Float64} (Smith-style robust division; used where the
// reference does NOT write @fastmath, e.g. `A[j,j] \ b[j]` linear_algebra.jl:288)
inline cplx div_robust(cplx a, cplx b) {
    if (std::fabs(b.re) >= std::fabs(b.im)) {
        double r = b.im / b.re, d = b.re + r * b.im;
        return {(a.re + a.im * r) / d, (a.im - a.re * r) / d};
    } else {
        double r = b.re / b.im, d = b.im + r * b.re;
        return {(a.re * r + a.im) / d, (a.im * r - a.re) / d};
    }
}
inline cplx cis(double th) { return {std::cos(th), std::sin(th)}; }

inline double nanmin(double a, double b) { return std::isnan(a) ? b : (std::isnan(b) ? a : std::fmin(a, b)); }
inline double nanmax(double a, double b) { return std::isnan(a) ? b : (std::isnan(b) ? a : std::fmax(a, b)); }
// Julia min/max propagate NaN
inline double jmin(double a, double b) { return (std::isnan(a) || std::isnan(b)) ? NaN : (a < b ? a : b); }
inline double jmax(double a, double b) { return (std::isnan(a) || std::isnan(b)) ? NaN : (a > b ? a : b); }
// Base.FastMath.max_fast(a,b) = ifelse(b > a, b, a)
inline double max_fast(double a, double b) { return b > a ? b : a; }
inline double min_fast(double a, double b) { return b < a ? b : a; }
inline double jclamp(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }

// src/utils.jl:408-422
inline double nthroot(double x, int N) {
    switch (N) {
        case 4: return std::sqrt(std::sqrt(x));
        case 2: return std::sqrt(x);
        case 3: return std::cbrt(x);
        case 1: return x;
        case 0: return 1.0;
        default: return std::pow(x, 1.0 / N);
    }
}
// eps(x::Float64): spacing of floats at x
inline double eps_of(double x) {
    x = std::fabs(x);
    if (!std::isfinite(x)) return NaN;
    return std::nextafter(x, INF) - x;
}

// ---------------------------------------------------------------- DoubleF64
struct dd {
    double hi, lo;
    dd() : hi(0), lo(0) {}
    dd(double h) : hi(h), lo(0) {}
    dd(double h, double l) : hi(h), lo(l) {}
};
inline void quick_two_sum(double a, double b, double& s, double& e) { s = a + b; e = b - (s - a); }
inline void two_sum(double a, double b, double& s, double& e) {
    s = a + b; double v = s - a; e = (a - (s - v)) + (b - v);
}
inline void two_diff(double a, double b, double& s, double& e) {
    s = a - b; double v = s - a; e = (a - (s - v)) - (b + v);
}
inline void two_prod(double a, double b, double& p, double& e) { p = a * b; e = std::fma(a, b, -p); }

inline dd operator+(dd a, dd b) {  // DoubleDouble.jl:185-191
    double hi, lo; two_sum(a.hi, b.hi, hi, lo);
    lo += (a.lo + b.lo);
    double h2, l2; quick_two_sum(hi, lo, h2, l2);
    return {h2, l2};
}
inline dd operator-(dd a, dd b) {  // :223-231
    double hi, lo; two_diff(a.hi, b.hi, hi, lo);
    lo += a.lo; lo -= b.lo;
    double h2, l2; quick_two_sum(hi, lo, h2, l2);
    return {h2, l2};
}
inline dd operator-(dd a) { return {-a.hi, -a.lo}; }
inline dd operator*(dd a, dd b) {  // :260-266 (sloppy)
    double p1, p2; two_prod(a.hi, b.hi, p1, p2);
    p2 += a.hi * b.lo + a.lo * b.hi;
    double h, l; quick_two_sum(p1, p2, h, l);
    return {h, l};
}
inline dd operator*(dd a, double b) {  // :247-253
    double p1, p2; two_prod(a.hi, b, p1, p2);
    p2 += a.lo * b;
    double h, l; quick_two_sum(p1, p2, h, l);
    return {h, l};
}
inline dd operator/(dd a, dd b) {  // :311-327
    double q1 = a.hi / b.hi;
    dd r = b * q1;
    double s1, s2; two_diff(a.hi, r.hi, s1, s2);
    s2 -= r.lo; s2 += a.lo;
    double q2 = (s1 + s2) / b.hi;
    double h, l; quick_two_sum(q1, q2, h, l);
    return {h, l};
}
inline dd square(dd a) {  // :364-370
    double p1, p2; two_prod(a.hi, a.hi, p1, p2);
    p2 += 2.0 * a.hi * a.lo;
    p2 += a.lo * a.lo;
    double h, l; quick_two_sum(p1, p2, h, l);
    return {h, l};
}

// ---------------------------------------------------------------- ComplexDF64
struct cdd {
    dd re, im;
    cdd() {}
    cdd(dd r, dd i) : re(r), im(i) {}
    explicit cdd(cplx z) : re(z.re), im(z.im) {}
};
inline cdd operator+(cdd a, cdd b) { return {a.re + b.re, a.im + b.im}; }
inline cdd operator-(cdd a, cdd b) { return {a.re - b.re, a.im - b.im}; }
inline cdd operator-(cdd a) { return {-a.re, -a.im}; }
inline cdd operator*(cdd a, cdd b) {  // Julia generic Complex{T} multiply
    return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
inline cdd& operator+=(cdd& a, cdd b) { a = a + b; return a; }
inline cdd& operator-=(cdd& a, cdd b) { a = a - b; return a; }
inline dd abs2(cdd a) { return a.re * a.re + a.im * a.im; }
inline cdd div_fast(cdd a, cdd b) {
    dd d = abs2(b);
    return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}
inline cdd inv_fast(cdd a) {
    dd d = abs2(a);
    return {a.re / d, -(a.im / d)};
}
inline bool iszero(cdd a) { return a.re.hi == 0.0 && a.re.lo == 0.0 && a.im.hi == 0.0 && a.im.lo == 0.0; }
inline cplx to_cplx(cdd a) { return {a.re.hi, a.im.hi}; }  // Float64(::DoubleF64) = hi
inline cplx to_cplx(cplx a) { return a; }

// ---------------------------------------------------------------- op kernels
// src/model_kit/operations.jl:184-248
template <class T> inline T from_c(cplx z);
template <> inline cplx from_c<cplx>(cplx z) { return z; }
template <> inline cdd from_c<cdd>(cplx z) { return cdd(z); }

inline cplx op_sqr(cplx z) { return {(z.re + z.im) * (z.re - z.im), (z.re + z.re) * z.im}; }
inline cdd op_sqr(cdd z) { return {(z.re + z.im) * (z.re - z.im), (z.re + z.re) * z.im}; }
inline cplx op_cb(cplx z) {
    double a = (z.re + z.im) * (z.re - z.im), b = (z.re + z.re) * z.im;
    return {a * z.re - b * z.im, a * z.im + b * z.re};
}
inline cdd op_cb(cdd z) {
    dd a = (z.re + z.im) * (z.re - z.im), b = (z.re + z.re) * z.im;
    return {a * z.re - b * z.im, a * z.im + b * z.re};
}
template <class T> inline T op_inv(T x) { return inv_fast(x); }
template <class T> inline T op_inv_not_zero(T x) { return iszero(x) ? x : inv_fast(x); }
template <class T> inline T op_invsqr(T x) { return op_sqr(inv_fast(x)); }
// Base.power_by_squaring(x, p) for p >= 1 (intfuncs.jl)
template <class T> inline T power_by_squaring(T x, int p) {
    if (p == 1) return x;
    if (p == 2) return x * x;
    int t = __builtin_ctz((unsigned)p) + 1;
    p >>= t;
    while (--t > 0) x = x * x;
    T y = x;
    while (p > 0) {
        t = __builtin_ctz((unsigned)p) + 1;
        p >>= t;
        while (--t >= 0) x = x * x;
        y = y * x;
    }
    return y;
}
template <class T> inline T op_pow_int(T x, int p) {
    if (p == 0) return from_c<T>(cplx(1.0));
    return p > 0 ? power_by_squaring(x, p) : inv_fast(power_by_squaring(x, -p));
}

}  // namespace orc
