// ORACLE (test infrastructure, NOT product code): per-path small dense linear algebra
// and norms (square systems only).
//
// Follows (reference file:line):
//   src/linear_algebra.jl:7-98     MatrixWorkspace, updated!
//   src/linear_algebra.jl:130-184  lu! (abs2 pivot, naive division, zero pivot skips scaling)
//   src/linear_algebra.jl:268-354  lu_ldiv!, adjoint solves, apply_ipiv!
//   src/linear_algebra.jl:389-408  ldiv!(x, WS, b)
//   src/linear_algebra.jl:432-497  skeel_row_scaling!, apply_row_scaling!
//   src/linear_algebra.jl:505-567  residual!, mixed/fixed precision iterative refinement
//   src/linear_algebra.jl:585-804  inverse_inf_norm_est, inf_norm, max_min_row, cond, egcond
//   src/linear_algebra.jl:809-885  Jacobian, ldiv!(x, J, b[, norm]), iterative_refinement!
//   src/norm.jl:36-40, 101-136, 150-234  WeightedNorm{InfNorm}, InfNorm
#pragma once
#include <vector>

#include "num.hpp"

namespace orc {

// ------------------------------------------------------------------ norms
struct WeightedNormOptions {  // norm.jl:36-40
    double scale_min = 1e-4, scale_abs_min = 1e-6, scale_max = 6.703903964971299e153;  // exp2(511)
};

inline double inf_norm(const cplx* x, int n) {  // norm.jl:194-213
    double dmax = abs2(x[0]);
    for (int i = 1; i < n; ++i) dmax = max_fast(dmax, abs2(x[i]));
    double d = std::sqrt(dmax);
    if (std::isinf(d)) {
        dmax = habs(x[0]);
        for (int i = 1; i < n; ++i) dmax = jmax(dmax, habs(x[i]));
        return dmax;
    }
    return d;
}
inline double inf_distance(const cplx* x, const cplx* y, int n) {  // norm.jl:150-167
    double dmax = abs2(x[0] - y[0]);
    for (int i = 1; i < n; ++i) dmax = max_fast(dmax, abs2(x[i] - y[i]));
    double d = std::sqrt(dmax);
    if (std::isinf(d)) {
        dmax = habs(x[0] - y[0]);
        for (int i = 1; i < n; ++i) dmax = jmax(dmax, abs2(x[i] - y[i]));  // sic (norm.jl:161)
        return dmax;
    }
    return d;
}

struct WeightedNorm {
    std::vector<double> w;
    WeightedNormOptions opt;
    void resize(int n) { w.assign(n, 1.0); }
    int size() const { return (int)w.size(); }
    double operator()(const cplx* x) const {  // norm.jl:214-234
        int n = size();
        double dmax = abs2(x[0] / w[0]);
        for (int i = 1; i < n; ++i) dmax = max_fast(dmax, abs2(x[i] / w[i]));
        double d = std::sqrt(dmax);
        if (std::isinf(d)) {
            dmax = habs(x[0] / w[0]);
            for (int i = 1; i < n; ++i) dmax = jmax(dmax, habs(x[i] / w[i]));
            return dmax;
        }
        return d;
    }
    double distance(const cplx* x, const cplx* y) const {  // norm.jl:168-192
        int n = size();
        double dmax = abs2((x[0] - y[0]) / w[0]);
        for (int i = 1; i < n; ++i) dmax = max_fast(dmax, abs2((x[i] - y[i]) / w[i]));
        double d = std::sqrt(dmax);
        if (std::isinf(d)) {
            dmax = habs((x[0] - y[0]) / w[0]);
            for (int i = 1; i < n; ++i) dmax = jmax(dmax, habs((x[i] - y[i]) / w[i]));
            return dmax;
        }
        return d;
    }
    void init(const cplx* x) {  // norm.jl:101-113 (point norm is the *unweighted* InfNorm)
        int n = size();
        double point_norm = inf_norm(x, n);
        for (int i = 0; i < n; ++i) {
            double wi = fast_abs(x[i]);
            if (wi < opt.scale_min * point_norm) wi = opt.scale_min * point_norm;
            else if (wi > opt.scale_max * point_norm) wi = opt.scale_max * point_norm;
            w[i] = jmax(wi, opt.scale_abs_min);
        }
    }
    void update(const cplx* x) {  // norm.jl:122-136
        int n = size();
        double norm_x = (*this)(x);
        for (int i = 0; i < n; ++i) {
            double wi = (fast_abs(x[i]) + w[i]) / 2;
            if (wi < opt.scale_min * norm_x) wi = opt.scale_min * norm_x;
            else if (wi > opt.scale_max * norm_x) wi = opt.scale_max * norm_x;
            if (std::isfinite(wi)) w[i] = jmax(wi, opt.scale_abs_min);
        }
    }
};

// ------------------------------------------------------------------ workspace
struct MatrixWorkspace {
    int n = 0;
    std::vector<cplx> A, lu;  // column-major n x n
    std::vector<int> ipiv;
    std::vector<double> row_scaling;
    bool factorized = false, scaled = false;
    std::vector<cdd> xbar, rbar;
    std::vector<cplx> r, dx, work;
    std::vector<double> rwork;
    long factorizations = 0, ldivs = 0;  // Jacobian stats, linear_algebra.jl:809-826

    void resize(int n_) {
        n = n_;
        A.assign((size_t)n * n, cplx()); lu = A;
        ipiv.assign(n, 0); row_scaling.assign(n, 1.0);
        xbar.assign(n, cdd()); rbar.assign(n, cdd());
        r.assign(n, cplx()); dx.assign(n, cplx()); work.assign(n, cplx()); rwork.assign(n, 0.0);
    }
    cplx& a(int i, int j) { return A[(size_t)j * n + i]; }
    void updated() {  // :88-98
        factorized = false; scaled = false;
        lu = A;
    }
};

inline void lu_factor(cplx* A, int* ipiv, int n) {  // :130-184
    for (int k = 0; k < n; ++k) {
        int kp = k;
        double amax = abs2(A[(size_t)k * n + k]);
        for (int i = k + 1; i < n; ++i) {
            double absi = abs2(A[(size_t)k * n + i]);
            if (absi > amax) { kp = i; amax = absi; }
        }
        ipiv[k] = kp;
        if (amax != 0.0) {  // !iszero(amax): NaN passes, as in Julia
            if (k != kp)
                for (int i = 0; i < n; ++i) std::swap(A[(size_t)i * n + k], A[(size_t)i * n + kp]);
            cplx Akk = A[(size_t)k * n + k];
            for (int i = k + 1; i < n; ++i) A[(size_t)k * n + i] = div_fast(A[(size_t)k * n + i], Akk);
        }
        for (int j = k + 1; j < n; ++j) {
            cplx Akj = A[(size_t)j * n + k];
            for (int i = k + 1; i < n; ++i) A[(size_t)j * n + i] -= A[(size_t)k * n + i] * Akj;
        }
    }
}
inline void lu_ldiv(cplx* x, const cplx* LU, const int* ipiv, int n, const cplx* b) {  // :310-316
    if (x != b) for (int i = 0; i < n; ++i) x[i] = b[i];
    for (int i = 0; i < n; ++i) if (i != ipiv[i]) std::swap(x[i], x[ipiv[i]]);
    for (int j = 0; j < n; ++j) {  // unit lower :295-308
        cplx xj = x[j];
        for (int i = j + 1; i < n; ++i) x[i] -= LU[(size_t)j * n + i] * xj;
    }
    for (int j = n - 1; j >= 0; --j) {  // upper :284-294
        cplx xj = x[j] = div_robust(x[j], LU[(size_t)j * n + j]);
        for (int i = 0; i < j; ++i) x[i] -= LU[(size_t)j * n + i] * xj;
    }
}
inline void lu_ldiv_adj(cplx* x, const cplx* LU, const int* ipiv, int n) {  // :318-354 (x in place)
    for (int j = 0; j < n; ++j) {  // adj upper
        cplx z = x[j];
        for (int i = 0; i < j; ++i) z -= conj(LU[(size_t)j * n + i]) * x[i];
        x[j] = div_robust(z, conj(LU[(size_t)j * n + j]));
    }
    for (int j = n - 1; j >= 0; --j) {  // adj unit lower
        cplx z = x[j];
        for (int i = n - 1; i >= j + 1; --i) z -= conj(LU[(size_t)j * n + i]) * x[i];
        x[j] = z;
    }
    for (int i = n - 1; i >= 0; --i) if (i != ipiv[i]) std::swap(x[i], x[ipiv[i]]);
}

inline void ws_factorize(MatrixWorkspace& W) { lu_factor(W.lu.data(), W.ipiv.data(), W.n); W.factorized = true; }

// LA.ldiv!(x, WS, b)  :389-408
inline void ws_ldiv(cplx* x, MatrixWorkspace& W, const cplx* b) {
    if (W.n == 1) { x[0] = div_robust(b[0], W.A[0]); return; }
    if (!W.factorized) ws_factorize(W);
    if (W.scaled) {
        for (int i = 0; i < W.n; ++i) x[i] = W.row_scaling[i] * b[i];
        lu_ldiv(x, W.lu.data(), W.ipiv.data(), W.n, x);
    } else {
        lu_ldiv(x, W.lu.data(), W.ipiv.data(), W.n, b);
    }
}

// skeel_row_scaling!(d, A, c; scaling_threshold)  :432-459
inline void skeel_row_scaling(double* d, const cplx* A, const double* c, int n, double scaling_threshold = -30.0) {
    for (int i = 0; i < n; ++i) d[i] = 0.0;
    for (int j = 0; j < n; ++j) {
        double cj = c[j];
        for (int i = 0; i < n; ++i) d[i] += fast_abs(A[(size_t)j * n + i]) * cj;
    }
    double m = d[0];
    for (int i = 1; i < n; ++i) m = jmax(m, d[i]);
    double s = scaling_threshold + m;  // sic: threshold added to the norm, not its exponent
    for (int i = 0; i < n; ++i) {
        int e = 0;
        if (d[i] != 0.0 && std::isfinite(d[i])) std::frexp(d[i], &e);
        if (e < s) d[i] = 1.0;
        else d[i] = std::exp2(-e);
    }
}
inline void apply_row_scaling(MatrixWorkspace& W) {  // :489-497
    int n = W.n;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) W.lu[(size_t)j * n + i] = W.lu[(size_t)j * n + i] * W.row_scaling[i];
    W.scaled = true;
}

// Jacobian-level ldiv!  :833-862
inline void jac_ldiv(cplx* x, MatrixWorkspace& W, const cplx* b, const WeightedNorm* norm = nullptr) {
    if (norm && !W.factorized) {
        skeel_row_scaling(W.row_scaling.data(), W.A.data(), norm->w.data(), W.n);
        apply_row_scaling(W);
    }
    W.factorizations += !W.factorized;
    W.ldivs += 1;
    ws_ldiv(x, W, b);
}

// residual!(r, A, x, b) = A x - b  :505-519
inline void residual(cplx* r, const cplx* A, const cplx* x, const cplx* b, int n) {
    for (int i = 0; i < n; ++i) r[i] = cplx();
    for (int j = 0; j < n; ++j) { cplx xj = x[j]; for (int i = 0; i < n; ++i) r[i] += A[(size_t)j * n + i] * xj; }
    for (int i = 0; i < n; ++i) r[i] -= b[i];
}
inline void residual_dd(cdd* r, const cplx* A, const cdd* x, const cplx* b, int n) {
    for (int i = 0; i < n; ++i) r[i] = cdd();
    for (int j = 0; j < n; ++j) { cdd xj = x[j]; for (int i = 0; i < n; ++i) r[i] += cdd(A[(size_t)j * n + i]) * xj; }
    for (int i = 0; i < n; ++i) r[i] -= cdd(b[i]);
}

// norm functor: weighted when wn != nullptr, else InfNorm
struct NormRef {
    const WeightedNorm* wn;
    int n;
    double operator()(const cplx* x) const { return wn ? (*wn)(x) : inf_norm(x, n); }
};

// :528-544
inline double mixed_precision_iterative_refinement(cplx* x, MatrixWorkspace& M, const cplx* b, NormRef norm) {
    int n = M.n;
    for (int i = 0; i < n; ++i) M.xbar[i] = cdd(x[i]);
    residual_dd(M.rbar.data(), M.A.data(), M.xbar.data(), b, n);
    for (int i = 0; i < n; ++i) M.r[i] = to_cplx(M.rbar[i]);
    ws_ldiv(M.dx.data(), M, M.r.data());
    for (int i = 0; i < n; ++i) x[i] -= M.dx[i];
    return norm(M.dx.data()) / norm(x);
}
// :553-567
inline double fixed_precision_iterative_refinement(cplx* x, MatrixWorkspace& M, const cplx* b, NormRef norm) {
    int n = M.n;
    residual(M.r.data(), M.A.data(), x, b, n);
    ws_ldiv(M.dx.data(), M, M.r.data());
    for (int i = 0; i < n; ++i) x[i] -= M.dx[i];
    return norm(M.dx.data()) / norm(x);
}
struct RefineResult { double accuracy; bool diverged; };
// iterative_refinement!(x, J, b, norm; max_iters, tol)  :864-885
inline RefineResult iterative_refinement(cplx* x, MatrixWorkspace& J, const cplx* b, NormRef norm, int max_iters, double tol) {
    J.ldivs += 1;
    double d = mixed_precision_iterative_refinement(x, J, b, norm);
    for (int i = 2; i <= max_iters; ++i) {
        J.ldivs += 1;
        double d2 = mixed_precision_iterative_refinement(x, J, b, norm);
        if (d2 < tol) return {d2, false};
        else if (d2 > 0.5 * d) return {d2, true};
        d = d2;
    }
    return {d, false};
}

// inverse_inf_norm_est  :585-682  (z = xi = y = work alias; x = rwork)
inline double inverse_inf_norm_est(MatrixWorkspace& W, const double* d_l, const double* d_r) {
    if (!W.factorized) ws_factorize(W);
    const double* rs = W.scaled ? W.row_scaling.data() : nullptr;
    int n = W.n;
    cplx* y = W.work.data();
    double* x = W.rwork.data();
    const cplx* LU = W.lu.data();
    const int* ipiv = W.ipiv.data();
    for (int i = 0; i < n; ++i) x[i] = 1.0 / n;
    if (d_r) for (int i = 0; i < n; ++i) x[i] /= d_r[i];
    for (int i = 0; i < n; ++i) y[i] = cplx(x[i]);
    lu_ldiv_adj(y, LU, ipiv, n);
    if (d_l) for (int i = 0; i < n; ++i) y[i] = y[i] / d_l[i];
    if (rs) for (int i = 0; i < n; ++i) y[i] = y[i] * rs[i];
    double gamma = 0; for (int i = 0; i < n; ++i) gamma += fast_abs(y[i]);
    for (int i = 0; i < n; ++i) y[i] = y[i] / fast_abs(y[i]);
    if (d_l) for (int i = 0; i < n; ++i) y[i] = y[i] / d_l[i];
    if (rs) for (int i = 0; i < n; ++i) y[i] = y[i] / rs[i];
    lu_ldiv(y, LU, ipiv, n, y);
    for (int i = 0; i < n; ++i) x[i] = d_r ? y[i].re / d_r[i] : y[i].re;
    int k = 2;
    while (true) {
        int j = 0; double maxx = std::fabs(x[0]);
        for (int i = 1; i < n; ++i) { double a = std::fabs(x[i]); if (a > maxx) { j = i; maxx = a; } }
        for (int i = 0; i < n; ++i) x[i] = 0.0;
        x[j] = 1.0;
        if (d_r) for (int i = 0; i < n; ++i) x[i] /= d_r[i];
        for (int i = 0; i < n; ++i) y[i] = cplx(x[i]);
        lu_ldiv_adj(y, LU, ipiv, n);
        if (d_l) for (int i = 0; i < n; ++i) y[i] = y[i] / d_l[i];
        if (rs) for (int i = 0; i < n; ++i) y[i] = y[i] * rs[i];
        double gbar = gamma;
        gamma = 0; for (int i = 0; i < n; ++i) gamma += fast_abs(y[i]);
        if (gamma <= gbar) { gamma = gbar; break; }
        for (int i = 0; i < n; ++i) y[i] = y[i] / fast_abs(y[i]);
        if (d_l) for (int i = 0; i < n; ++i) y[i] = y[i] / d_l[i];
        if (rs) for (int i = 0; i < n; ++i) y[i] = y[i] / rs[i];
        lu_ldiv(y, LU, ipiv, n, y);
        for (int i = 0; i < n; ++i) x[i] = d_r ? y[i].re / d_r[i] : y[i].re;
        k += 1;
        double ninf = 0; for (int i = 0; i < n; ++i) ninf = jmax(ninf, std::fabs(x[i]));
        if (x[j] == ninf || k > 2) break;
    }
    return nanmin(gamma, INF);
}
inline double ws_inf_norm(const MatrixWorkspace& W, const double* d_l, const double* d_r) {  // :684-707
    double norm = -INF; int n = W.n;
    for (int i = 0; i < n; ++i) {
        double ni = 0.0;
        for (int j = 0; j < n; ++j) ni += d_r ? fast_abs(W.A[(size_t)j * n + i]) * d_r[j] : fast_abs(W.A[(size_t)j * n + i]);
        if (d_l) ni *= d_l[i];
        norm = max_fast(norm, ni);
    }
    return norm;
}
inline double ws_cond(MatrixWorkspace& W, const double* d_l, const double* d_r) {  // :745-774
    if (W.n == 1) {
        double a = habs(W.A[0]);
        if (d_l) a *= d_l[0];
        if (d_r) a *= d_r[0];
        return 1.0 / a;
    }
    return inverse_inf_norm_est(W, d_l, d_r) * ws_inf_norm(W, d_l, d_r);
}

}  // namespace orc
