// ORACLE (test infrastructure, NOT product code): Valuation, EndgameTracker,
// PolyhedralTracker.track and the PathResult record.
//
// Follows (reference file:line):
//   src/valuation.jl:6-228            Valuation
//   src/endgame_tracker.jl:47-72      EndgameOptions
//   src/endgame_tracker.jl:97-141     EndgameTrackerCode (+ conversion from TrackerCode)
//   src/endgame_tracker.jl:177-222    EndgameTrackerState
//   src/endgame_tracker.jl:260-313    init!, track!
//   src/endgame_tracker.jl:329-530    step!, check_finite!, check_at_infinity!, switch_to_*!
//   src/endgame_tracker.jl:533-721    singular_endgame_step!, add_sample!, predict_endpoint!, tracking_stopped!
//   src/endgame_tracker.jl:847-914    PathResult(...), track
//   src/polyhedral.jl:414-530         track(::PolyhedralTracker, (cell, x))
//   src/homotopies/toric_homotopy.jl:66-112  update_weights!
//   src/path_result.jl:76-98          PathResult
#pragma once
#include <climits>

#include "tracker.hpp"

namespace orc {

// ------------------------------------------------------------ Valuation
struct Valuation {
    int n = 0;
    std::vector<double> val_x, val_tx, dval_x, dval_tx;
    std::vector<double> vx2, vx1, vd2, vd1, lx2, lx1, ld2, ld1;  // *_data tuples (2 = newest)
    double logt2 = NaN, logt1 = NaN;
    void resize(int n_) {
        n = n_;
        for (auto* v : {&val_x, &val_tx, &dval_x, &dval_tx, &vx2, &vx1, &vd2, &vd1, &lx2, &lx1, &ld2, &ld1}) v->assign(n, 0.0);
    }
    void init() {  // :39-50
        for (auto* v : {&val_x, &val_tx, &dval_x, &dval_tx, &vx2, &vx1, &vd2, &vd1, &lx2, &lx1, &ld2, &ld1}) v->assign(n, 0.0);
        logt2 = logt1 = NaN;
    }
};
inline double val_nu(cplx x, cplx xd, double t) {  // :53-61
    double mu = x.re * xd.re + x.im * xd.im;
    return t * (mu / abs2(x));
}
inline void val_nu_nu1(cplx x, cplx xd, cplx x2, double t, double& nu, double& nu1) {  // :63-77
    double xx = abs2(x);
    double mu = x.re * xd.re + x.im * xd.im;
    double l = mu / xx;
    double mu1 = x.re * x2.re + xd.re * xd.re + x.im * x2.im + xd.im * xd.im;
    double l1 = mu1 / xx - 2 * (l * l);
    nu = t * l; nu1 = t * l1 + l;
}
inline double finite_diff(double v, double s, double v2, double s2, double v1, double s1) {  // :138-141
    double D1 = s - s1, D2 = s - s2, D12 = s1 - s2;
    return (D2 * v1) / (D12 * D1) - ((D12 + D2) * v2) / (D12 * D2) - (D12 * v) / (D1 * D2);
}
inline void valuation_update(Valuation& val, Predictor& pred, double t) {  // :82-124
    double logt = std::log(t);
    bool diff = pred.winding_number > 1 && !std::isnan(val.logt2);
    int n = val.n;
    for (int i = 0; i < n; ++i) {
        cplx x = pred.x0()[i], xd = pred.x1()[i], x2 = pred.x2()[i], x3 = pred.x3()[i];
        double logx = std::log(fast_abs(x)), logxd = std::log(fast_abs(xd));
        if (diff) {
            double nu = val_nu(x, xd, t);
            double dnu = finite_diff(nu, logt, val.vx2[i], val.logt2, val.vx1[i], val.logt1);
            val.val_x[i] = nu; val.dval_x[i] = dnu;
            double vxd = finite_diff(logxd, logt, val.ld2[i], val.logt2, val.ld1[i], val.logt1);
            double dvxd = finite_diff(vxd, logt, val.vd2[i], val.logt2, val.vd1[i], val.logt1);
            val.val_tx[i] = vxd + 1; val.dval_tx[i] = dvxd;
        } else {
            double nu, nu1;
            val_nu_nu1(x, xd, 2.0 * x2, t, nu, nu1);
            val.val_x[i] = nu; val.dval_x[i] = t * nu1;
            double old2 = val.vx2[i];
            val.vx2[i] = nu; val.vx1[i] = old2;
            double vxd;
            val_nu_nu1(xd, 2.0 * x2, 6.0 * x3, t, vxd, nu1);
            val.val_tx[i] = vxd + 1; val.dval_tx[i] = t * nu1;
        }
        double o = val.lx2[i]; val.lx2[i] = logx; val.lx1[i] = o;
        o = val.ld2[i]; val.ld2[i] = logxd; val.ld1[i] = o;
    }
    val.logt1 = val.logt2; val.logt2 = logt;
}
inline double jmax3(double a, double b, double c) { return jmax(jmax(a, b), c); }
inline void at_infinity_tol(std::vector<double>& tols, const Valuation& val, double finite_tol, bool zero_is_finite) {  // :143-173
    for (int i = 0; i < val.n; ++i) {
        double vx = val.val_x[i];
        double e = jmax3(std::fabs(1.0 - val.val_tx[i] / vx), std::fabs(val.dval_x[i] / vx), std::fabs(val.dval_tx[i] / val.val_tx[i]));
        if (std::isnan(e)) { tols[i] = INF; continue; }
        if (vx + e < -finite_tol) tols[i] = e;
        else if (!zero_is_finite && vx - e > -finite_tol) tols[i] = e;
        else tols[i] = INF;
    }
}
inline bool val_is_finite(const Valuation& val, double finite_tol, bool zero_is_finite, int max_winding_number) {  // :175-205
    double delta = 1.0 / max_winding_number;
    for (int i = 0; i < val.n; ++i) {
        double vx = val.val_x[i];
        if (std::fabs(vx) < finite_tol) {
            if (!(std::fabs(val.dval_x[i]) < finite_tol) || val.val_tx[i] < 0.5 * delta) return false;
        } else if (zero_is_finite && vx > (delta - finite_tol)) {
            double e = jmax3(std::fabs(1.0 - val.val_tx[i] / vx), std::fabs(val.dval_x[i] / vx), std::fabs(val.dval_tx[i] / val.val_tx[i]));
            if (!(e < finite_tol)) return false;
        } else return false;
    }
    return true;
}
inline double jround(double x) { return std::nearbyint(x); }  // Julia round: ties to even
inline void estimate_winding_number(const Valuation& val, int max_winding_number, int& m, double& min_err) {  // :207-228
    m = 1; min_err = INF;
    for (int k = 1; k <= max_winding_number; ++k) {
        double err = 0.0;
        for (int i = 0; i < val.n; ++i) {
            double mv = k * val.val_tx[i];
            double e = std::fabs(jround(mv) - mv);
            err = e > err ? e : err;
        }
        if (err < min_err) { m = k; min_err = err; }
    }
}

// ------------------------------------------------------------ Endgame
struct EndgameOptions {  // endgame_tracker.jl:47-72
    double endgame_start = 0.1;
    int max_endgame_steps = 2000, max_endgame_extended_steps = 400;
    double min_cond = 1e6, min_cond_growth = 1e4, min_coord_growth = 100.0;
    bool zero_is_at_infinity = false, at_infinity_check = true, only_nonsingular = false;
    double singular_min_accuracy = 1e-6;
    int max_winding_number = 6;
    double val_finite_tol = 0.05, val_at_infinity_tol = 0.01;
    double sing_cond = 1e14, sing_accuracy = 1e-12, scaling_threshold = -30.0;
    int refine_steps = 3;
};
enum EGCode : int32_t {  // :100-117
    EG_tracking = 0, EG_success, EG_at_infinity, EG_at_zero, EG_terminated_accuracy_limit,
    EG_terminated_invalid_startvalue, EG_terminated_invalid_startvalue_singular_jacobian,
    EG_terminated_ill_conditioned, EG_terminated_max_steps, EG_terminated_max_extended_steps,
    EG_terminated_max_winding_number, EG_terminated_step_size_too_small, EG_terminated_unknown,
    EG_post_check_failed, EG_excess_solution, EG_polyhedral_failed
};
inline EGCode convert_code(TrackerCode c) {  // :119-139
    switch (c) {
        case TC_success: return EG_success;
        case TC_terminated_max_steps: return EG_terminated_max_steps;
        case TC_terminated_accuracy_limit: return EG_terminated_accuracy_limit;
        case TC_terminated_ill_conditioned: return EG_terminated_ill_conditioned;
        case TC_terminated_invalid_startvalue: return EG_terminated_invalid_startvalue;
        case TC_terminated_invalid_startvalue_singular_jacobian: return EG_terminated_invalid_startvalue_singular_jacobian;
        case TC_terminated_step_size_too_small: return EG_terminated_step_size_too_small;
        case TC_terminated_unknown: return EG_terminated_unknown;
        default: return EG_tracking;
    }
}

struct PathResult {  // path_result.jl:76-98
    int32_t return_code = EG_tracking;
    std::vector<cplx> solution;
    double t = NaN, accuracy = NaN, residual = NaN;
    bool singular = false;
    double condition_jacobian = NaN;
    int winding_number = 0;  // 0 = nothing
    bool extended_precision = false;
    std::vector<cplx> last_point; double last_t = NaN;
    bool has_valuation = false; std::vector<double> valuation;
    double omega = NaN, mu = NaN;
    int accepted_steps = 0, rejected_steps = 0, steps_eg = 0;
    bool extended_precision_used = false;
    long n_factorizations = 0, n_ldivs = 0;
};

struct EndgameTracker {
    Tracker tracker;
    EndgameOptions options;
    // state  :177-215
    EGCode code = EG_tracking;
    bool singular_endgame = false;
    Valuation val;
    int winding_number = 0;  // 0 = nothing
    std::vector<cplx> solution;
    double accuracy = NaN, cond = 1.0;
    bool singular = false;
    int steps_eg = 0;
    long ext_steps_eg_start = 0;
    bool jtz_prev = false, jtz_cur = false;  // jump_to_zero_failed
    std::vector<cplx> last_point; double last_t = NaN;
    std::vector<double> row_scaling, col_scaling;
    std::vector<double> at_infinity_starts, at_infinity_tols, at_infinity_abs_coords, at_infinity_conds;
    std::vector<cplx> samples[3];  // each 2 rows of n
    int sample_idx[3] = {0, 1, 2};
    double sample_times[3] = {0, 0, 0}, sample_conds[3] = {0, 0, 0};
    double singular_start = NaN;
    int singular_steps = 0;
    std::vector<cplx> prediction, prev_prediction;
    int n = 0, m = 0;

    void setup(const HomotopyDef* D, const TrackerOptions& topt, const EndgameOptions& eopt, const WeightedNormOptions& nopt) {
        tracker.setup(D, topt, nopt); options = eopt;
        n = tracker.n; m = tracker.m;
        val.resize(n);
        solution.assign(n, cplx()); last_point.assign(n, cplx());
        row_scaling.assign(m, 0.0); col_scaling.assign(n, 0.0);
        for (auto* v : {&at_infinity_starts, &at_infinity_tols, &at_infinity_abs_coords, &at_infinity_conds}) v->assign(n, 0.0);
        for (auto& s : samples) s.assign((size_t)2 * n, cplx());
        prediction.assign(n, cplx()); prev_prediction.assign(n, cplx());
    }

    void init(const cplx* x, double t1, double omega = NaN, double mu = NaN, bool extended_precision = false) {  // :260-294
        tracker.options.min_rel_step_size = 0.0;
        tracker.init(x, cplx(t1), cplx(0.0), omega, mu, INF, INF, false, extended_precision);
        code = convert_code(tracker.state.code);
        singular_endgame = false; jtz_prev = jtz_cur = false;
        val.init(); winding_number = 0;
        for (auto& z : solution) z = cplx(NaN, NaN);
        accuracy = NaN; cond = NaN; singular = false; steps_eg = 0; ext_steps_eg_start = LONG_MAX / 2;
        for (auto& d : row_scaling) d = 1; for (auto& d : col_scaling) d = 1;
        for (auto& d : at_infinity_starts) d = NaN;
        singular_steps = 0;
        sample_idx[0] = 0; sample_idx[1] = 1; sample_idx[2] = 2;
    }

    void row_scaling_update() {  // row_scaling!(d, WS, c, threshold) linear_algebra.jl:466-479
        skeel_row_scaling(row_scaling.data(), tracker.state.jacobian.A.data(), col_scaling.data(), n, options.scaling_threshold);
    }
    double jac_cond() { return ws_cond(tracker.state.jacobian, row_scaling.data(), col_scaling.data()); }

    bool check_finite() {  // :402-422
        if (!val_is_finite(val, options.val_finite_tol, !options.zero_is_at_infinity, options.max_winding_number)) return false;
        int mw; double merr;
        estimate_winding_number(val, options.max_winding_number, mw, merr);
        if (merr < options.val_finite_tol) {
            if (mw == 1 && !jtz_prev) return false;
            winding_number = mw;
            return true;
        }
        return false;
    }

    bool check_at_infinity() {  // :424-499
        if (!options.at_infinity_check) return false;
        at_infinity_tol(at_infinity_tols, val, options.val_finite_tol, !options.zero_is_at_infinity);
        double kappa = NaN;
        TrackerState& ts = tracker.state;
        double t = ts.t().re;
        for (int i = 0; i < n; ++i) {
            double tol = at_infinity_tols[i];
            if (tol < options.val_at_infinity_tol) {
                if (std::isnan(at_infinity_starts[i])) {
                    bool allnan = true;
                    for (double s : at_infinity_starts) allnan &= std::isnan(s);
                    if (allnan) {
                        col_scaling = ts.norm.w;
                        row_scaling_update();
                    }
                    kappa = jac_cond();
                    at_infinity_conds[i] = kappa;
                    at_infinity_abs_coords[i] = fast_abs(ts.x[i]);
                    at_infinity_starts[i] = t;
                } else {
                    if (std::isnan(kappa)) kappa = jac_cond();
                    double v = val.val_x[i];
                    bool at_zero = v > 0;
                    double cond_growth = kappa / at_infinity_conds[i];
                    double coord_growth = at_zero ? at_infinity_abs_coords[i] / fast_abs(ts.x[i])
                                                  : fast_abs(ts.x[i]) / at_infinity_abs_coords[i];
                    if (coord_growth > jclamp(std::pow(0.25, 4 * v), 20, options.min_coord_growth) &&
                        (cond_growth > options.min_cond_growth || kappa > jmax(1e8, options.min_cond))) {
                        code = at_zero ? EG_at_zero : EG_at_infinity;
                        return true;
                    }
                }
            } else if (!std::isnan(at_infinity_starts[i])) {
                at_infinity_starts[i] = NaN;
            }
        }
        return false;
    }

    void add_sample(double t) {  // :630-662
        const cplx* tx1 = tracker.predictor.tx.data();
        int mw = winding_number;
        double s = nthroot(t, mw);
        double mu = mw * std::pow(s, mw - 1);
        double kappa = jac_cond();
        int slot;
        if (singular_steps <= 2) slot = singular_steps;
        else {
            int first = sample_idx[0];
            sample_idx[0] = sample_idx[1]; sample_times[0] = sample_times[1]; sample_conds[0] = sample_conds[1];
            sample_idx[1] = sample_idx[2]; sample_times[1] = sample_times[2]; sample_conds[1] = sample_conds[2];
            sample_idx[2] = first;
            slot = 2;
        }
        std::vector<cplx>& ty = samples[sample_idx[slot]];
        for (int i = 0; i < n; ++i) { ty[i] = tx1[i]; ty[n + i] = mu * tx1[n + i]; }
        sample_times[slot] = s; sample_conds[slot] = kappa;
    }

    double predict_endpoint() {  // :664-693
        if (singular_steps < 2) return INF;
        auto S = [&](int k) -> std::vector<cplx>& { return samples[sample_idx[k]]; };
        if (singular_steps == 2)
            cubic_hermite(prediction.data(), S(0).data(), S(0).data() + n, cplx(sample_times[0]), S(1).data(), S(1).data() + n,
                          cplx(sample_times[1]), cplx(0.0), n);
        prev_prediction = prediction;
        cubic_hermite(prediction.data(), S(1).data(), S(1).data() + n, cplx(sample_times[1]), S(2).data(), S(2).data() + n,
                      cplx(sample_times[2]), cplx(0.0), n);
        double p = sample_times[2] / sample_times[1];
        double p2 = p * p;
        double err = inf_distance(prediction.data(), prev_prediction.data(), n) / std::fabs(p2 * p2 - 1);
        double norm_s = inf_norm(prediction.data(), n);
        if (norm_s > 1e-8) err /= norm_s;
        return err;
    }

    void switch_to_singular() {  // :501-524
        singular_endgame = true;
        double t = tracker.state.t().re;
        singular_start = t;
        bool allone = true;
        for (double d : row_scaling) allone &= (d == 1.0);
        if (allone) { col_scaling = tracker.state.norm.w; row_scaling_update(); }
        add_sample(t);
        singular_steps = 0;
        tracker.predictor.winding_number = winding_number;
        at_infinity_conds[0] = sample_conds[0];
        tracker.state.keep_extended_prec = true;
    }
    void switch_to_regular() {  // :525-530
        singular_endgame = false;
        tracker.predictor.winding_number = 1;
        tracker.init_continue(cplx(0.0));
    }

    void tracking_stopped() {  // :695-721
        accuracy = tracker.state.accuracy;
        if (code == EG_success && accuracy > 1e-14) tracker.refine_current_solution(1e-14, options.refine_steps);
        solution = tracker.state.x;
        winding_number = 0;
        if (code == EG_success) {
            col_scaling = tracker.state.norm.w;
            row_scaling_update();
            cond = tracker.cond(solution.data(), cplx(0.0), row_scaling.data(), col_scaling.data());
            singular = cond > options.sing_cond || accuracy > options.sing_accuracy;
        }
    }

    EGCode singular_endgame_step() {  // :533-628
        const double lambda = 0.25;
        double t = tracker.state.t().re;
        tracker.init_continue(cplx(lambda * t));
        bool max_steps = false;
        while (tracker.state.code == TC_tracking) {
            tracker.step();
            if ((steps_eg += 1) >= options.max_endgame_steps) { code = EG_terminated_max_steps; max_steps = true; break; }
            else if (tracker.state.ext_steps() - ext_steps_eg_start > options.max_endgame_extended_steps) {
                code = EG_terminated_max_extended_steps; max_steps = true; break;
            }
        }
        singular_steps += 1;
        if (!max_steps) {
            if (tracker.state.code != TC_success) {
                code = convert_code(tracker.state.code);
                tracking_stopped();
                return code;
            }
            valuation_update(val, tracker.predictor, lambda * t);
            int mh; double mh_err;
            estimate_winding_number(val, options.max_winding_number, mh, mh_err);
            if (mh != winding_number || mh_err > 0.1) { switch_to_regular(); return code; }
            add_sample(lambda * t);
            if (singular_steps < 2) return code;
            double acc = predict_endpoint();
            if (singular_steps == 2) { accuracy = acc; solution = prediction; return code; }
            else if (acc < accuracy && accuracy > 1e-12) { accuracy = acc; solution = prediction; return code; }
        }
        // @label prediction
        int mw = winding_number;
        double kappa = sample_conds[2];
        double zero_cond = 1.0 / (mw + 1);
        for (int i = 0; i < n; ++i) solution[i] = (val.val_x[i] < zero_cond ? 1.0 : 0.0) * prediction[i];
        double kappa0 = tracker.cond(solution.data(), cplx(0.0), row_scaling.data(), col_scaling.data());
        double J0_norm = ws_inf_norm(tracker.state.jacobian, row_scaling.data(), nullptr);
        if (accuracy < options.singular_min_accuracy &&
            (((mw > 1 && kappa > options.min_cond && nanmax(kappa0, 1.0 / J0_norm) > kappa) || (mw == 1 && kappa0 > 1e12)) ||
             max_steps || (n == 1 && 1.0 / J0_norm < options.min_cond))) {
            cond = jmax(kappa0, 1.0 / J0_norm);
            singular = true;
            return (code = EG_success);
        } else if (!max_steps) {
            switch_to_regular();
            return code;
        }
        return code;
    }

    EGCode step() {  // :329-399
        TrackerState& ts = tracker.state;
        if (steps_eg >= options.max_endgame_steps) return (code = EG_terminated_max_steps);
        else if (ts.ext_steps() - ext_steps_eg_start > options.max_endgame_extended_steps) {
            bool nonan = true;
            for (auto& z : solution) nonan &= !isnan(z);
            if (nonan && winding_number != 0 && accuracy < options.singular_min_accuracy) {
                cond = tracker.cond(solution.data(), cplx(0.0), row_scaling.data(), col_scaling.data());
                singular = true;
                return (code = EG_success);
            }
            return (code = EG_terminated_max_extended_steps);
        }
        last_point = ts.x;
        last_t = ts.t().re;
        if (singular_endgame) return singular_endgame_step();
        cplx tp = ts.stepper.t_prop();
        bool is_jump_to_zero = iszero(tp);
        bool step_success = tracker.step();
        code = convert_code(ts.code);
        if (code != EG_tracking) { tracking_stopped(); return code; }
        jtz_prev = jtz_cur; jtz_cur = is_jump_to_zero;
        double t = ts.t().re;
        if (!(t <= options.endgame_start)) return code;
        if (steps_eg == 0) ext_steps_eg_start = ts.ext_steps();
        steps_eg += 1;
        if (!step_success) return code;
        valuation_update(val, tracker.predictor, t);
        if (check_finite()) { switch_to_singular(); return code; }
        else if (check_at_infinity()) return code;
        return code;
    }

    EGCode track(const cplx* x, double t1, double omega = NaN, double mu = NaN, bool extended_precision = false) {  // :297-313
        init(x, t1, omega, mu, extended_precision);
        while (code == EG_tracking) step();
        return code;
    }

    PathResult path_result() {  // :847-888
        PathResult R;
        TrackerState& ts = tracker.state;
        Homotopy& H = tracker.H;
        std::vector<cplx>& r = tracker.corrector.r;
        if (code == EG_success) {
            R.t = 0.0; R.solution = solution;
            H.evaluate(r.data(), solution.data(), cplx(0.0));
        } else {
            R.t = ts.t().re; R.solution = ts.x;
            H.evaluate(r.data(), ts.x.data(), cplx(R.t));
        }
        R.residual = inf_norm(r.data(), m);
        R.return_code = code; R.singular = singular; R.accuracy = accuracy; R.condition_jacobian = cond;
        R.winding_number = winding_number;
        R.last_point = last_point; R.last_t = last_t;
        R.has_valuation = !(R.t > options.endgame_start);
        R.valuation = val.val_x;
        R.omega = ts.omega; R.mu = ts.mu; R.extended_precision = ts.extended_prec;
        R.accepted_steps = ts.accepted_steps; R.rejected_steps = ts.rejected_steps; R.steps_eg = steps_eg;
        R.extended_precision_used = ts.used_extended_prec;
        R.n_factorizations = ts.jacobian.factorizations; R.n_ldivs = ts.jacobian.ldivs;
        return R;
    }
};

// ------------------------------------------------------------ Polyhedral
// Host passes the *unscaled* per-cell weights s_ij = lifting_ij - beta_i + <a_ij, normal> with
// s = 0 on the two cell vertices of each support (toric_homotopy.jl:76-96); the min/max
// rescaling (:99-107) is done here because stage 1 may re-weight (polyhedral.jl:468-489).
struct PolyhedralTracker {
    Tracker toric;
    EndgameTracker generic;
    int P = 0;

    void setup(const HomotopyDef* toricD, const HomotopyDef* coeffD, const TrackerOptions& topt, const EndgameOptions& eopt,
               const WeightedNormOptions& nopt) {
        toric.setup(toricD, topt, nopt);
        generic.setup(coeffD, topt, eopt, nopt);
        P = toric.H.P;
    }
    void update_weights(const double* raw, bool use_min, double target, double& s_min, double& s_max) {
        s_max = 0.0; s_min = INF;
        std::vector<double>& w = toric.H.weights;
        for (int l = 0; l < P; ++l) {
            w[l] = raw[l];
            if (raw[l] != 0.0) { s_max = jmax(s_max, raw[l]); s_min = jmin(s_min, raw[l]); }
        }
        if (use_min) { double lam = s_min / target; for (auto& x : w) x /= lam; s_max = s_max / lam; s_min = target; }
        else { double lam = s_max / target; for (auto& x : w) x /= lam; s_min = s_min / lam; s_max = target; }
    }
    // vertex_mask[l] != 0 marks the cell's two vertices per support (weight exactly 0)
    PathResult track(const double* raw_weights, const cplx* x_inf) {  // polyhedral.jl:414-530
        double min_w, max_w;
        update_weights(raw_weights, true, 1.0, min_w, max_w);
        TrackerCode rc;
        double mu, omega;
        TrackerState& ts = toric.state;
        if (max_w < 10) {
            rc = toric.track(x_inf, cplx(0.0), cplx(1.0), 20.0, 1e-12, false, INF, false, 0.2);
            mu = ts.mu; omega = ts.omega;
        } else {
            double t0 = jclamp(std::pow(0.1, 10 / max_w), 0.9, 1 - 1e-6);
            rc = toric.track(x_inf, cplx(0.0), cplx(t0), 20.0, 1e-12, false, INF, false, 0.2);
            mu = ts.mu; omega = ts.omega;
            if (rc == TC_success) {
                update_weights(raw_weights, false, 10.0, min_w, max_w);
                double t_restart = std::pow(t0, 1 / min_w);
                double min_step = toric.options.min_step_size;
                toric.options.min_step_size = 0.0;
                std::vector<cplx> xcur = ts.x;
                rc = toric.track(xcur.data(), cplx(t_restart), cplx(1.0), omega, mu, false, 0.1 * t_restart, true);
                mu = ts.mu; omega = ts.omega;
                toric.options.min_step_size = min_step;
            }
        }
        if (rc != TC_success) {
            PathResult R;
            R.return_code = EG_polyhedral_failed; R.solution = ts.x; R.t = ts.t().re; R.accuracy = ts.accuracy;
            R.singular = false; R.condition_jacobian = NaN; R.residual = NaN; R.winding_number = 0;
            R.last_point = ts.x; R.last_t = ts.t().re; R.has_valuation = false; R.valuation.assign(toric.n, 0.0);
            R.omega = ts.omega; R.mu = ts.mu; R.extended_precision = ts.extended_prec;
            R.accepted_steps = ts.accepted_steps; R.rejected_steps = ts.rejected_steps;
            R.extended_precision_used = ts.used_extended_prec;
            return R;
        }
        // omega deliberately not passed (polyhedral.jl:515-521) => init_newton! runs again
        generic.track(ts.x.data(), 1.0, NaN, mu);
        PathResult R = generic.path_result();
        R.accepted_steps += ts.accepted_steps;
        R.rejected_steps += ts.rejected_steps;
        return R;
    }
};

}  // namespace orc
