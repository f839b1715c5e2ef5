// ORACLE (test infrastructure, NOT product code): SegmentStepper, Predictor,
// NewtonCorrector and Tracker (Timme 2020, arXiv:1902.02968).
//
// Follows (reference file:line):
//   src/utils.jl:300-394            SegmentStepper
//   src/predictor.jl:72-371         Predictor (Pade(2,1) / cubic Hermite)
//   src/newton_corrector.jl:21-286  newton!, init_newton!, extended_prec_refinement_step!
//   src/tracker.jl:45-140, 164-178, 307-389   parameters / options / codes / state
//   src/tracker.jl:509-619          cond, step size control, check_terminated!
//   src/tracker.jl:639-844          init!, precision switching, refine_current_solution!
//   src/tracker.jl:851-996          step!, track!
#pragma once
#include "homotopy.hpp"
#include "linalg.hpp"

namespace orc {

// ------------------------------------------------------------ SegmentStepper
struct SegmentStepper {  // utils.jl:300-394
    cplx start, target;
    double abs_delta = 0;
    bool forward = true;
    double s = 0, s_prop = 0;
    void init(cplx st, cplx tg) {
        start = st; target = tg;
        abs_delta = habs(tg - st);
        forward = habs(st) < habs(tg);
        s = s_prop = forward ? 0.0 : abs_delta;
    }
    bool is_done() const { return forward ? s == abs_delta : s == 0.0; }
    void step_success() { s = s_prop; }
    void propose_step(double ds) {
        if (forward) s_prop = jmin(s + ds, abs_delta);
        else s_prop = jmax(s - ds, 0.0);
    }
    double dist_to_target() const { return forward ? abs_delta - s : s; }
    double ds() const { return forward ? s_prop - s : s - s_prop; }
    cplx t_helper(double sv) const {
        if (forward) {
            if (sv == 0.0) return start;
            if (sv == abs_delta) return target;
            return start + (sv / abs_delta) * (target - start);
        } else {
            if (sv == abs_delta) return start;
            if (sv == 0.0) return target;
            return target + (sv / abs_delta) * (start - target);
        }
    }
    cplx t() const { return t_helper(s); }
    cplx t_prop() const { return t_helper(s_prop); }
    cplx dt() const {
        if (forward) return ((s_prop - s) / abs_delta) * (target - start);
        return ((s - s_prop) / abs_delta) * (target - start);
    }
};

// ------------------------------------------------------------ options / codes
struct TrackerParameters {  // tracker.jl:45-62
    double a = 0.125, beta_a = 1.0, beta_omega_p = 3.0, beta_tau = 0.4, strict_beta_tau = 0.3;
    int min_newton_iters = 2;
};
struct TrackerOptions {  // tracker.jl:94-140
    int max_steps = 10000;
    double max_step_size = INF, max_initial_step_size = INF;
    bool extended_precision = true;
    double min_step_size = 1e-48, min_rel_step_size = 0.0;
    TrackerParameters parameters;
};
enum TrackerCode : int32_t {  // tracker.jl:166-176
    TC_tracking = 0, TC_success, TC_terminated_max_steps, TC_terminated_accuracy_limit,
    TC_terminated_ill_conditioned, TC_terminated_invalid_startvalue,
    TC_terminated_invalid_startvalue_singular_jacobian, TC_terminated_step_size_too_small,
    TC_terminated_unknown
};
enum NewtonCode : int32_t { NEWT_CONVERGED = 0, NEWT_TERMINATED, NEWT_MAX_ITERS, NEWT_SINGULARITY };
struct NewtonResult {  // newton_corrector.jl:21-29
    NewtonCode return_code; double accuracy; int iters; double omega, theta, mu_low, norm_dx0;
};
inline double _h(double a) { return 2 * a * (std::sqrt(4 * a * a + 1) - 2 * a); }  // tracker.jl:517

// ------------------------------------------------------------ Predictor
enum PredMethod { PM_Pade21, PM_Hermite };
inline cplx cpow_int(cplx z, int p) { return p == 0 ? cplx(1.0) : power_by_squaring(z, p); }

// predictor.jl:338-351 (branch = 0)
inline cplx t_to_s_plane(cplx t, int m) {
    double r = fast_abs(t);
    if (t.im == 0.0 && t.re > 0) return cplx(nthroot(r, m));
    double th = std::atan2(t.im, t.re);
    th = std::fmod(th, 2 * M_PI); if (th < 0) th += 2 * M_PI;  // mod2pi
    return nthroot(r, m) * cis(th / m);
}
// predictor.jl:354-371. tx rows: [0..n) value, [n..2n) derivative
inline void cubic_hermite(cplx* xh, const cplx* v0, const cplx* d0, cplx t0, const cplx* v1, const cplx* d1, cplx t1, cplx t, int n) {
    if (t0.im == 0 && t1.im == 0 && t.im == 0) {
        double T = t.re, T0 = t0.re, T1 = t1.re;
        double s = (T - T0) / (T1 - T0);
        double h00 = (1 + 2 * s) * ((1 - s) * (1 - s));
        double h10 = (T - T0) * ((1 - s) * (1 - s));
        double h01 = (s * s) * (3 - 2 * s);
        double h11 = (T - T0) * s * (s - 1);
        for (int i = 0; i < n; ++i) xh[i] = h00 * v0[i] + h10 * d0[i] + h01 * v1[i] + h11 * d1[i];
    } else {
        cplx one(1.0);
        cplx s = div_robust(t - t0, t1 - t0);
        cplx oms2 = (one - s) * (one - s);
        cplx h00 = (one + 2.0 * s) * oms2;
        cplx h10 = (t - t0) * oms2;
        cplx h01 = (s * s) * (cplx(3.0) - 2.0 * s);
        cplx h11 = (t - t0) * s * (s - one);
        for (int i = 0; i < n; ++i) xh[i] = h00 * v0[i] + h10 * d0[i] + h01 * v1[i] + h11 * d1[i];
    }
}

struct Predictor {  // predictor.jl:72-103
    int n = 0, m = 0;
    PredMethod method = PM_Pade21;
    int order = 4;
    double trust_region = INF, local_error = INF, cond_H_xdot = INF;
    std::vector<cplx> tx;       // 4 rows of n: x^0..x^3 (tx^0..tx^2 are views, predictor.jl:108-113)
    cplx t = cplx(NaN), prev_t = cplx(NaN);
    double tx_norm[4] = {0, 0, 0, 0};
    std::vector<cplx> xtemp, u;
    std::vector<cplx> prev_tx1; // 2 rows
    int winding_number = 1;
    cplx s = cplx(NaN), prev_s = cplx(NaN);
    std::vector<cplx> ty1, prev_ty1;

    void resize(int m_, int n_) {
        m = m_; n = n_;
        tx.assign((size_t)4 * n, cplx()); xtemp.assign(n, cplx()); u.assign(m, cplx());
        prev_tx1.assign((size_t)2 * n, cplx()); ty1.assign((size_t)2 * n, cplx()); prev_ty1.assign((size_t)2 * n, cplx());
    }
    void init() {  // :124-133
        cond_H_xdot = 1.0; winding_number = 1;
        t = prev_t = cplx(NaN); s = prev_s = cplx(NaN);
        trust_region = local_error = NaN;
    }
    cplx* x0() { return tx.data(); }
    cplx* x1() { return tx.data() + n; }
    cplx* x2() { return tx.data() + 2 * n; }
    cplx* x3() { return tx.data() + 3 * n; }

    // update!  :158-284.  xhat == nullptr <=> `nothing`
    void update(Homotopy& H, const cplx* x, cplx t_, MatrixWorkspace& J, const WeightedNorm& norm, const cplx* xhat) {
        int mw = winding_number;
        for (int i = 0; i < 2 * n; ++i) prev_tx1[i] = tx[i];
        prev_t = t; t = t_;
        if (mw > 1) { prev_s = s; s = t_to_s_plane(t_, mw); }
        if (!xhat) local_error = NaN;
        else {
            double ds = fast_abs(t_ - prev_t);
            local_error = norm.distance(xhat, x) / std::pow(ds, order);
        }
        for (int i = 0; i < n; ++i) x0()[i] = x[i];
        tx_norm[0] = norm(x);
        if (mw > 1) for (int i = 0; i < n; ++i) ty1[i] = x[i];

        NormRef wnorm{&norm, n}, inorm{nullptr, n};
        H.taylor(1, u.data(), x, t_);
        for (int i = 0; i < m; ++i) u[i] = -u[i];
        jac_ldiv(xtemp.data(), J, u.data());
        double delta = fixed_precision_iterative_refinement(xtemp.data(), J, u.data(), wnorm);
        cond_H_xdot = delta / EPS;
        const double tol1 = 1e-10;
        if (delta > tol1) iterative_refinement(xtemp.data(), J, u.data(), inorm, 5, tol1);
        tx_norm[1] = norm(xtemp.data());
        for (int i = 0; i < n; ++i) x1()[i] = xtemp[i];
        if (mw > 1) {
            cplx mu = mw == 2 ? 2.0 * s : (double)mw * cpow_int(s, mw - 1);
            for (int i = 0; i < n; ++i) ty1[n + i] = mu * xtemp[i];
            method = PM_Hermite; order = 4;
            trust_region = tx_norm[0] / tx_norm[1];
            if (std::isnan(local_error)) { double q = tx_norm[1] / tx_norm[0]; local_error = q * q * q; }
            return;
        }
        H.taylor(2, u.data(), tx.data(), t_);
        for (int i = 0; i < m; ++i) u[i] = -u[i];
        jac_ldiv(xtemp.data(), J, u.data());
        const double tol2 = 1e-10;
        if (delta > tol2) iterative_refinement(xtemp.data(), J, u.data(), wnorm, 4, tol2);
        tx_norm[2] = norm(xtemp.data());
        for (int i = 0; i < n; ++i) x2()[i] = xtemp[i];

        H.taylor(3, u.data(), tx.data(), t_);
        for (int i = 0; i < m; ++i) u[i] = -u[i];
        jac_ldiv(xtemp.data(), J, u.data());
        const double tol3 = 1e-4;
        if (delta > tol3) iterative_refinement(xtemp.data(), J, u.data(), wnorm, 3, tol3);
        tx_norm[3] = norm(xtemp.data());
        for (int i = 0; i < n; ++i) x3()[i] = xtemp[i];

        double tau = INF;
        for (int i = 0; i < n; ++i) {
            double c1 = fast_abs(x1()[i]), c2 = fast_abs(x2()[i]), c3 = fast_abs(x3()[i]);
            double lam = jmax(1e-6, c1);
            c1 /= lam; c2 /= lam * lam; c3 /= lam * lam * lam;
            double tol = 1e-14 * jmax(jmax(c1, c2), c3);
            if (!((c1 <= tol && c2 <= tol && c3 <= tol) || c2 <= tol)) {
                double ti = (c2 / c3) / lam;
                if (ti < tau) tau = ti;
            }
        }
        if (!std::isfinite(tau)) tau = tx_norm[2] / tx_norm[3];
        if (!std::isfinite(tau)) tau = tx_norm[0] / jmax(jmax(tx_norm[0], tx_norm[1]), jmax(tx_norm[2], tx_norm[3]));
        method = PM_Pade21; order = 4; trust_region = tau;
        if (std::isnan(local_error)) { double q = 1.0 / tau; local_error = (q * q) * (q * q); }
    }

    // predict!  :286-329
    void predict(cplx* xhat, cplx t_, cplx dt) {
        if (method == PM_Pade21) {
            double lam = trust_region, lam2 = lam * lam, lam3 = lam * lam * lam;
            const double tol = 1e-12;
            for (int i = 0; i < n; ++i) {
                cplx X = x0()[i], X1 = x1()[i], X2 = x2()[i], X3 = x3()[i];
                double c = fast_abs(X), c1 = fast_abs(X1), c2 = fast_abs(X2), c3 = fast_abs(X3);
                double a1 = c1 * lam, a2 = c2 * lam2, a3 = c3 * lam3;
                double tau = tol * std::sqrt(c * c + a1 * a1 + a2 * a2 + a3 * a3);
                if (c3 * lam3 <= tau || c2 * lam2 <= tau) {
                    xhat[i] = X + dt * (X1 + dt * X2);
                } else {
                    cplx d = cplx(1.0) - div_robust(dt * X3, X2);
                    xhat[i] = X + dt * (X1 + div_robust(dt * X2, d));
                }
            }
        } else {
            int mw = winding_number;
            cplx ps = t_to_s_plane(prev_t, mw), sc = t_to_s_plane(t_, mw), sp = t_to_s_plane(t_ + dt, mw);
            cplx prev_sm, sm;
            if (mw == 2) { prev_sm = 2.0 * ps; sm = 2.0 * sc; }
            else { prev_sm = (double)mw * cpow_int(ps, mw - 1); sm = (double)mw * cpow_int(sc, mw - 1); }
            for (int i = 0; i < n; ++i) { prev_ty1[i] = prev_tx1[i]; prev_ty1[n + i] = prev_sm * prev_tx1[n + i]; }
            for (int i = 0; i < n; ++i) { ty1[i] = tx[i]; ty1[n + i] = sm * tx[n + i]; }
            cubic_hermite(xhat, prev_ty1.data(), prev_ty1.data() + n, ps, ty1.data(), ty1.data() + n, sc, sp, n);
        }
    }
};

// ------------------------------------------------------------ Newton corrector
struct NewtonCorrector {  // newton_corrector.jl:35-53
    double a = 0.125, h_a = 0;
    std::vector<cplx> dx, r;
    std::vector<cdd> x_ext;
    void resize(double a_, int n, int m) {
        a = a_; h_a = 2 * a * (std::sqrt(4 * a * a + 1) - 2 * a);
        dx.assign(n, cplx()); r.assign(m, cplx()); x_ext.assign(n, cdd());
    }
};
inline void to_ext(std::vector<cdd>& xe, const cplx* x, int n) { for (int i = 0; i < n; ++i) xe[i] = cdd(x[i]); }

// extended_prec_refinement_step!  :55-78
inline double extended_prec_refinement_step(cplx* xbar, NewtonCorrector& NC, Homotopy& H, const cplx* x, cplx t,
                                            MatrixWorkspace& J, const WeightedNorm& norm, bool simple_newton_step = true) {
    int n = H.n;
    H.evaluate_and_jacobian(NC.r.data(), J.A.data(), x, t);
    to_ext(NC.x_ext, x, n);
    H.evaluate_dd(NC.r.data(), NC.x_ext.data(), t);
    J.updated();
    jac_ldiv(NC.dx.data(), J, NC.r.data(), &norm);
    iterative_refinement(NC.dx.data(), J, NC.r.data(), NormRef{&norm, n}, 3, 1e-8);
    for (int i = 0; i < n; ++i) xbar[i] = x[i] - NC.dx[i];
    if (simple_newton_step) {
        to_ext(NC.x_ext, xbar, n);
        H.evaluate_dd(NC.r.data(), NC.x_ext.data(), t);
        jac_ldiv(NC.dx.data(), J, NC.r.data(), &norm);
    }
    return norm(NC.dx.data());
}

// newton!  :80-205.  xbar may alias x0.
inline NewtonResult newton(cplx* xbar, NewtonCorrector& NC, Homotopy& H, const cplx* x0, cplx t, MatrixWorkspace& J,
                           const WeightedNorm& norm, double mu, double omega, bool extended_precision = false,
                           bool accurate_mu = false, bool first_correction = false) {
    const int n = H.n;
    const double a = NC.a, h_a = NC.h_a;
    cplx* dx = NC.dx.data();
    cplx* r = NC.r.data();
    NormRef wnorm{&norm, n};
    if (xbar != x0) for (int i = 0; i < n; ++i) xbar[i] = x0[i];
    cplx* xi = xbar;
    double mu_low = NaN, theta = NaN, norm_dxi = NaN, norm_dxim1 = NaN, norm_dx0 = NaN;
    double abar = a;
    for (int i = 0; i <= 10; ++i) {
        H.evaluate_and_jacobian(r, J.A.data(), xi, t);
        if (extended_precision) { to_ext(NC.x_ext, xi, n); H.evaluate_dd(r, NC.x_ext.data(), t); }
        J.updated();
        jac_ldiv(dx, J, r, &norm);
        if (extended_precision) iterative_refinement(dx, J, r, wnorm, 3, abar * abar);
        norm_dxi = norm(dx);
        if (std::isnan(norm_dxi)) return {NEWT_SINGULARITY, norm(dx), i + 1, omega, theta, mu_low, norm_dx0};
        for (int k = 0; k < n; ++k) xi[k] = xi[k] - dx[k];
        if (i == 0) norm_dx0 = norm_dxi;
        if (i == 1) omega = 2 * norm_dxi / (norm_dxim1 * norm_dxim1);
        if (i >= 1) theta = norm_dxi / norm_dxim1;
        if ((i >= 1 && theta > abar) || (i == 0 && !first_correction && 0.125 * norm_dx0 * omega > h_a)) {
            return {NEWT_TERMINATED, norm(dx), i + 1, omega, theta, mu_low, norm_dx0};
        } else if (omega * norm_dxi * norm_dxi < 2 * mu * std::sqrt(1 - 2 * h_a)) {
            H.evaluate_and_jacobian(r, J.A.data(), xi, t);
            J.updated();
            if (extended_precision) {
                jac_ldiv(dx, J, r);
                mu_low = norm(dx);
                to_ext(NC.x_ext, xi, n);
                H.evaluate_dd(r, NC.x_ext.data(), t);
            }
            jac_ldiv(dx, J, r);
            if (extended_precision) iterative_refinement(dx, J, r, wnorm, 3, abar * abar);
            for (int k = 0; k < n; ++k) xi[k] = xi[k] - dx[k];
            double norm_dxip1 = norm(dx);
            if (std::isnan(norm_dxip1)) return {NEWT_SINGULARITY, norm(dx), i + 1, omega, theta, mu_low, norm_dx0};
            else if (norm_dxip1 > std::sqrt(norm_dxi)) {
                theta = norm_dxip1 / norm_dxi;
                return {NEWT_TERMINATED, norm_dxip1, i + 2, omega, theta, mu_low, norm_dx0};
            }
            if (norm_dxip1 > 2 * mu && extended_precision) {
                to_ext(NC.x_ext, xi, n);
                H.evaluate_dd(r, NC.x_ext.data(), t);
                jac_ldiv(dx, J, r);
                norm_dxi = norm_dxip1;
                mu = norm_dxip1 = norm(dx);
            } else if (norm_dxip1 > 2 * mu || accurate_mu) {
                H.evaluate(r, xi, t);
                jac_ldiv(dx, J, r);
                mu = norm(dx);
            } else {
                mu = norm_dxip1;
            }
            if (i == 0) {
                double ob = 2 * norm_dxi / (norm_dxip1 * norm_dxip1);
                if (ob < omega) omega = ob; else omega *= 0.25;
            }
            return {NEWT_CONVERGED, mu, i + 2, omega, theta, mu_low, norm_dx0};
        }
        norm_dxim1 = norm_dxi;
        if (i >= 1) abar *= abar;
    }
    return {NEWT_MAX_ITERS, mu, 11, omega, theta, mu_low, norm_dx0};
}

struct InitNewtonResult { bool valid; double omega, mu; };
// init_newton!  :207-286
inline InitNewtonResult init_newton(cplx* xbar, NewtonCorrector& NC, Homotopy& H, const cplx* x0, cplx t, MatrixWorkspace& J,
                                    const WeightedNorm& norm, double /*a_kw*/, bool extended_precision) {
    const int n = H.n;
    const double a = NC.a;  // `@unpack a ... = NC` shadows the keyword (newton_corrector.jl:219)
    cplx* dx = NC.dx.data(); cplx* r = NC.r.data();
    H.evaluate_and_jacobian(r, J.A.data(), x0, t);
    if (extended_precision) { to_ext(NC.x_ext, x0, n); H.evaluate_dd(r, NC.x_ext.data(), t); }
    J.updated();
    jac_ldiv(dx, J, r, &norm);
    double v = norm(dx) + EPS;
    bool valid = false;
    double omega = NaN, mu = NaN;
    double eps_ = std::sqrt(v);
    for (int k = 1; k <= 3; ++k) {
        for (int i = 0; i < n; ++i) xbar[i] = x0[i] + cplx(eps_ * norm.w[i]);
        H.evaluate_and_jacobian(r, J.A.data(), xbar, t);
        if (extended_precision) { to_ext(NC.x_ext, xbar, n); H.evaluate_dd(r, NC.x_ext.data(), t); }
        J.updated();
        jac_ldiv(dx, J, r, &norm);
        for (int i = 0; i < n; ++i) xbar[i] = xbar[i] - dx[i];
        double norm_dx0 = norm(dx);
        if (extended_precision) { to_ext(NC.x_ext, xbar, n); H.evaluate_dd(r, NC.x_ext.data(), t); }
        else H.evaluate(r, xbar, t);
        jac_ldiv(dx, J, r, &norm);
        for (int i = 0; i < n; ++i) xbar[i] = xbar[i] - dx[i];
        double norm_dx1 = norm(dx) + EPS;
        if (norm_dx1 < a * norm_dx0) {
            omega = 2 * norm_dx1 / (norm_dx0 * norm_dx0);
            mu = norm_dx1;
            if (omega * mu > std::pow(a, 7)) {
                NewtonResult res = newton(xbar, NC, H, xbar, t, J, norm, std::pow(a, 7) / omega, omega, extended_precision, true);
                if (res.return_code == NEWT_CONVERGED) { valid = true; omega = res.omega; mu = res.accuracy; }
                else valid = false;
            } else { valid = true; break; }
        } else {
            eps_ *= std::sqrt(eps_);
        }
    }
    return {valid, omega, mu};
}

// ------------------------------------------------------------ Tracker
struct TrackerState {  // tracker.jl:307-338
    std::vector<cplx> x, xhat, xbar;
    SegmentStepper stepper;
    double ds_prev = 0, accuracy = 0, omega = 1, omega_prev = 1, mu = EPS, tau = INF, norm_dx0 = NaN;
    bool extended_prec = false, used_extended_prec = false, refined_extended_prec = false, keep_extended_prec = false;
    WeightedNorm norm;
    bool use_strict_beta_tau = false;
    MatrixWorkspace jacobian;
    TrackerCode code = TC_tracking;
    int accepted_steps = 0, rejected_steps = 0, last_steps_failed = 0, ext_accepted_steps = 0, ext_rejected_steps = 0;
    int steps() const { return accepted_steps + rejected_steps; }
    int ext_steps() const { return ext_accepted_steps + ext_rejected_steps; }
    cplx t() const { return stepper.t(); }
};

struct PathStats {  // event counters for the flop accounting of SURVEY.md 8(d)
    long evaljac = 0, eval = 0, eval_dd = 0, taylor[4] = {0, 0, 0, 0};
};

struct Tracker {
    Homotopy H;
    Predictor predictor;
    NewtonCorrector corrector;
    TrackerState state;
    TrackerOptions options;
    int m = 0, n = 0;

    void setup(const HomotopyDef* D, const TrackerOptions& opt, const WeightedNormOptions& nopt) {
        H.init(D); m = H.m; n = H.n; options = opt;
        state.x.assign(n, cplx()); state.xhat.assign(n, cplx()); state.xbar.assign(n, cplx());
        state.norm.resize(n); state.norm.opt = nopt;
        state.jacobian.resize(n);
        predictor.resize(m, n);
        corrector.resize(opt.parameters.a, n, m);
    }

    // LA.cond(tracker, x, t, d_l, d_r)  tracker.jl:509-514
    double cond(const cplx* x, cplx t, const double* d_l, const double* d_r) {
        H.evaluate_and_jacobian(corrector.r.data(), state.jacobian.A.data(), x, t);
        state.jacobian.updated();
        return ws_cond(state.jacobian, d_l, d_r);
    }

    double initial_step_size() {  // :520-539
        double a = options.parameters.beta_a * options.parameters.a;
        int p = predictor.order;
        double omega = state.omega, e = predictor.local_error;
        if (std::isinf(e)) e = 1e5;
        double tau = predictor.trust_region;
        double ds1 = nthroot((std::sqrt(1 + 2 * _h(a)) - 1) / (omega * e), p) / options.parameters.beta_omega_p;
        double ds2 = options.parameters.beta_tau * tau;
        double ds = nanmin(ds1, ds2);
        return jmin(jmin(ds, options.max_step_size), options.max_initial_step_size);
    }

    void update_stepsize(const NewtonResult& result) {  // :541-588
        TrackerState& st = state;
        const TrackerParameters& par = options.parameters;
        double a = par.beta_a * par.a;
        int p = predictor.order;
        double omega = jclamp(st.omega + 2 * (st.omega - st.omega_prev), st.omega, 8 * st.omega);
        double tau = st.tau;
        double ds;
        if (result.return_code == NEWT_CONVERGED) {
            double e = predictor.local_error;
            double ds1 = nthroot((std::sqrt(1 + 2 * _h(a)) - 1) / (omega * e), p) / par.beta_omega_p;
            double ds2 = par.beta_tau * tau;
            if (st.use_strict_beta_tau || st.stepper.dist_to_target() < ds2) ds2 = par.strict_beta_tau * tau;
            ds = jmin(nanmin(ds1, ds2), options.max_step_size);
            if (st.use_strict_beta_tau && st.stepper.dist_to_target() < ds) ds *= par.strict_beta_tau;
            ds = jmin(ds, 10 * st.ds_prev);
            if (st.last_steps_failed > 0) ds = jmin(ds, st.ds_prev);
        } else {
            int j = result.iters - 2;
            // 1 << j with j = -1 is 0 in Julia (shift by negative = right shift): nthroot(theta, 0) = 1
            int rootn = j >= 0 ? (1 << j) : 0;
            double Th = nthroot(result.theta, rootn);
            double h_Th = _h(Th), h_a = _h(0.5 * a);
            if (std::isnan(Th) || result.return_code == NEWT_SINGULARITY || std::isnan(result.accuracy) ||
                result.iters == 1 || h_Th < h_a) {
                ds = 0.25 * st.stepper.ds();
            } else {
                ds = nthroot((std::sqrt(1 + 2 * _h(0.5 * a)) - 1) / (std::sqrt(1 + 2 * _h(Th)) - 1), p) * st.stepper.ds();
            }
        }
        st.stepper.propose_step(ds);
    }

    void check_terminated() {  // :591-619
        TrackerState& st = state;
        double tol_acc;
        if (st.extended_prec || !options.extended_precision) {
            double a = options.parameters.a;
            tol_acc = std::pow(a, (1 << options.parameters.min_newton_iters) - 1) * _h(a);
        } else tol_acc = INF;
        cplx tp = st.stepper.t_prop(), t = st.stepper.t();
        if (st.stepper.is_done()) st.code = TC_success;
        else if (st.steps() >= options.max_steps) st.code = TC_terminated_max_steps;
        else if (st.omega * st.mu > tol_acc) st.code = TC_terminated_accuracy_limit;
        else if (st.stepper.ds() < options.min_step_size) st.code = TC_terminated_step_size_too_small;
        else if (fast_abs(tp - t) <= 2 * eps_of(fast_abs(t))) st.code = TC_terminated_step_size_too_small;
        else if (options.min_rel_step_size > 0 && !(tp.re == st.stepper.target.re && tp.im == st.stepper.target.im) &&
                 fast_abs(tp - t) < fast_abs(t) * options.min_rel_step_size)
            st.code = TC_terminated_step_size_too_small;
    }

    void update_predictor(const cplx* xhat = nullptr) {  // :621-624
        predictor.update(H, state.x.data(), state.t(), state.jacobian, state.norm, xhat);
    }

    // init!(tracker, x1, t1, t0; ...)  :639-754
    bool init(const cplx* x1, cplx t1, cplx t0, double omega = NaN, double mu = NaN, double tau = INF,
              double max_initial_step_size = INF, bool keep_steps = false, bool extended_precision = false) {
        TrackerState& st = state;
        for (int i = 0; i < n; ++i) st.x[i] = x1[i];
        st.stepper.init(t1, t0);
        st.ds_prev = 0.0; st.accuracy = EPS; st.omega = 1.0;
        st.keep_extended_prec = false; st.use_strict_beta_tau = false;
        st.norm.init(st.x.data());
        st.jacobian.factorizations = st.jacobian.ldivs = 0;
        st.code = TC_tracking;
        if (!keep_steps) st.accepted_steps = st.rejected_steps = st.ext_accepted_steps = st.ext_rejected_steps = 0;
        st.last_steps_failed = 0;
        cplx t = st.t();
        bool valid;
        if (std::isnan(omega) || std::isnan(mu)) {
            double a = options.parameters.a;
            InitNewtonResult r = init_newton(st.xbar.data(), corrector, H, st.x.data(), t, st.jacobian, st.norm, a, extended_precision);
            valid = r.valid; omega = r.omega; mu = r.mu;
            if (!valid && !extended_precision) {
                extended_precision = true;
                r = init_newton(st.xbar.data(), corrector, H, st.x.data(), t, st.jacobian, st.norm, a, true);
                valid = r.valid; omega = r.omega; mu = r.mu;
            }
        } else valid = true;
        st.used_extended_prec = st.extended_prec = extended_precision;
        if (!std::isnan(omega)) st.omega = omega;
        if (valid) {
            st.accuracy = mu;
            st.mu = jmax(mu, EPS);
        } else {
            H.evaluate_and_jacobian(corrector.r.data(), st.jacobian.A.data(), st.x.data(), t);
            bool anynan = false;
            for (auto& z : st.jacobian.A) anynan |= isnan(z);
            int corank = 0;
            if (anynan) st.code = TC_terminated_invalid_startvalue;  // overwritten below, as in the reference (:731-735)
            else corank = n - numerical_rank(st.jacobian.A, n, 1e-14);
            st.code = corank > 0 ? TC_terminated_invalid_startvalue_singular_jacobian : TC_terminated_invalid_startvalue;
            return false;
        }
        st.tau = tau;
        H.evaluate_and_jacobian(corrector.r.data(), st.jacobian.A.data(), st.x.data(), t);
        st.jacobian.updated();
        predictor.init();
        update_predictor();
        st.tau = predictor.trust_region;
        double ds = initial_step_size();
        ds = jmax(jmin(ds, max_initial_step_size), options.min_step_size);
        st.stepper.propose_step(ds);
        st.omega_prev = st.omega;
        return st.code == TC_tracking;
    }
    // init!(tracker, t0)  :756-766
    void init_continue(cplx t0, double max_initial_step_size = INF) {
        state.code = TC_tracking;
        state.stepper.init(state.t(), t0);
        double ds = initial_step_size();
        ds = jmin(ds, max_initial_step_size);
        state.stepper.propose_step(ds);
        state.ds_prev = 0.0;
    }

    // rank(J, rtol) via one-sided Jacobi SVD (LA.rank uses LAPACK SVD: tracker.jl:721,728)
    static int numerical_rank(const std::vector<cplx>& A, int n, double rtol) {
        std::vector<cplx> V(A);
        for (int sweep = 0; sweep < 60; ++sweep) {
            double off = 0;
            for (int p = 0; p < n; ++p)
                for (int q = p + 1; q < n; ++q) {
                    double app = 0, aqq = 0; cplx apq;
                    for (int i = 0; i < n; ++i) {
                        cplx vp = V[(size_t)p * n + i], vq = V[(size_t)q * n + i];
                        app += abs2(vp); aqq += abs2(vq); apq += conj(vp) * vq;
                    }
                    double g = habs(apq);
                    if (g <= 1e-300 || g <= 1e-17 * std::sqrt(app * aqq)) continue;
                    off = std::fmax(off, g / std::sqrt(app * aqq));
                    cplx ph = apq / g;
                    double zeta = (aqq - app) / (2 * g);
                    double tt = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
                    double c = 1 / std::sqrt(1 + tt * tt), s = c * tt;
                    for (int i = 0; i < n; ++i) {
                        cplx vp = V[(size_t)p * n + i], vq = V[(size_t)q * n + i] * conj(ph);
                        V[(size_t)p * n + i] = c * vp - s * vq;
                        V[(size_t)q * n + i] = (s * vp + c * vq) * ph;
                    }
                }
            if (off < 1e-15) break;
        }
        double smax = 0; std::vector<double> sv(n);
        for (int p = 0; p < n; ++p) { double s2 = 0; for (int i = 0; i < n; ++i) s2 += abs2(V[(size_t)p * n + i]); sv[p] = std::sqrt(s2); smax = std::fmax(smax, sv[p]); }
        int r = 0; for (int p = 0; p < n; ++p) if (sv[p] > rtol * smax) ++r;
        return r;
    }

    double use_extended_precision() {  // :788-813
        TrackerState& st = state;
        if (!options.extended_precision) return st.mu;
        if (st.extended_prec) return st.mu;
        st.extended_prec = true; st.used_extended_prec = true;
        double mu = st.mu;
        for (int i = 0; i < 2; ++i)
            mu = extended_prec_refinement_step(st.x.data(), corrector, H, st.x.data(), st.t(), st.jacobian, st.norm, false);
        st.mu = jmax(mu, EPS);
        return st.mu;
    }
    bool update_precision(double mu_low) {  // :768-786
        TrackerState& st = state;
        double a = options.parameters.a;
        if (!options.extended_precision) return false;
        if (st.extended_prec && !st.keep_extended_prec && !std::isnan(mu_low) && mu_low > st.mu) {
            if (mu_low * st.omega < std::pow(a, 7) * _h(a)) { st.extended_prec = false; st.mu = mu_low; }
        } else if (st.mu * st.omega > std::pow(a, 5) * _h(a)) {
            use_extended_precision();
        }
        return st.extended_prec;
    }
    double refine_current_solution(double min_tol = 4 * EPS, int nsteps = 3) {  // :815-844
        TrackerState& st = state;
        double mu = st.accuracy;
        double mub = extended_prec_refinement_step(st.xbar.data(), corrector, H, st.x.data(), st.t(), st.jacobian, st.norm, false);
        if (mub < mu) { st.x = st.xbar; mu = mub; }
        int k = 1;
        while (mu > min_tol && k <= nsteps) {
            mub = extended_prec_refinement_step(st.xbar.data(), corrector, H, st.x.data(), st.t(), st.jacobian, st.norm, true);
            if (mub < mu) { st.x = st.xbar; mu = mub; }
            k += 1;
        }
        return mu;
    }

    // step!  :851-926
    bool step() {
        TrackerState& st = state;
        cplx t = st.stepper.t(), dt = st.stepper.dt(), tp = st.stepper.t_prop();
        predictor.predict(st.xhat.data(), t, dt);
        st.norm.update(st.xhat.data());
        NewtonResult result = newton(st.xbar.data(), corrector, H, st.xhat.data(), tp, st.jacobian, st.norm, st.mu, st.omega,
                                     st.extended_prec, false, st.accepted_steps == 0);
        if (result.return_code == NEWT_CONVERGED) {
            st.x = st.xbar;
            st.ds_prev = st.stepper.ds();
            st.stepper.step_success();
            st.accuracy = result.accuracy;
            st.mu = jmax(result.accuracy, EPS);
            st.omega_prev = st.omega;
            st.omega = jmax(jmax(result.omega, 0.5 * st.omega), 0.1);
            update_precision(result.mu_low);
            if (st.stepper.is_done() && options.extended_precision && st.accuracy > 1e-14) {
                st.accuracy = refine_current_solution(1e-14);
                st.refined_extended_prec = true;
            }
            update_predictor(st.xhat.data());
            st.tau = predictor.trust_region;
            st.accepted_steps += 1;
            st.ext_accepted_steps += st.extended_prec;
            st.last_steps_failed = 0;
        } else {
            st.rejected_steps += 1;
            st.ext_rejected_steps += st.extended_prec;
            st.last_steps_failed += 1;
        }
        st.norm_dx0 = result.norm_dx0;
        update_stepsize(result);
        check_terminated();
        return !(st.last_steps_failed > 0);
    }

    // track!(tracker, x, t1, t0; ...)  :937-968
    TrackerCode track(const cplx* x, cplx t1, cplx t0, double omega = NaN, double mu = NaN, bool extended_precision = false,
                      double tau = INF, bool keep_steps = false, double max_initial_step_size = INF) {
        init(x, t1, t0, omega, mu, tau, max_initial_step_size, keep_steps, extended_precision);
        while (state.code == TC_tracking) step();
        return state.code;
    }
};

}  // namespace orc
