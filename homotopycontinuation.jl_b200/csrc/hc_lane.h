// The path state machine of a lane group: one flat loop iteration = (at most) one tracker step,
// with the cheap phase logic of the endgame / polyhedral drivers around it.  A group whose path
// ends writes its PathResult and refills itself from the device-side path queue.
//
// The phases flatten the reference's nested loops (file:line):
//   PH_PLAIN    Tracker.track!                              src/tracker.jl:937-968
//   PH_EG       EndgameTracker.step! (regular)              src/endgame_tracker.jl:329-399
//   PH_SING     singular_endgame_step! inner tracking loop  src/endgame_tracker.jl:533-628
//   PH_TORIC_A/B  PolyhedralTracker stage 1 (toric 0 -> 1, optional re-weighted restart)
//                                                           src/polyhedral.jl:414-530
// plus Valuation (src/valuation.jl:39-228), check_finite!/check_at_infinity!/switch_to_*!
// (src/endgame_tracker.jl:402-530), add_sample!/predict_endpoint!/tracking_stopped!
// (:630-721) and the PathResult assembly (:847-888).
#pragma once
#include "hc_path.h"

namespace hc {

enum EGCode : int {  // src/endgame_tracker.jl:100-117
    EG_tracking = 0, EG_success, EG_at_infinity, EG_at_zero, EG_terminated_accuracy_limit,
    EG_terminated_invalid_startvalue, EG_terminated_invalid_startvalue_singular_jacobian,
    EG_terminated_ill_conditioned, EG_terminated_max_steps, EG_terminated_max_extended_steps,
    EG_terminated_max_winding_number, EG_terminated_step_size_too_small, EG_terminated_unknown,
    EG_post_check_failed, EG_excess_solution, EG_polyhedral_failed
};
HC_HD int convert_code(int c) {  // :119-139
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

enum Phase : int { PH_IDLE = 0, PH_PLAIN, PH_EG, PH_SING, PH_TORIC_A, PH_TORIC_B };
// Heavy, once-per-path work (tracker / endgame initialisation, the toric restart, the condition number and
// residual of a finished path) is not run where it arises: the lane parks with an event, and the kernel
// handles the events of a warp together (event_finish / event_begin below), each kind of work from ONE call
// site -- otherwise every lane would run its initialisations alone while the other 31 lanes of the warp wait.
enum Event : int { EV_NONE = 0, EV_START, EV_TORIC_NEXT, EV_FINISH_EG, EV_FINISH_PLAIN, EV_FINISH_POLY_FAILED, EV_HANDOFF };
// return_code of a path the first pass of a two-pass batch gave up on (never leaves the library): because it switched to
// extended precision (such paths end within a few hundred steps), or because of its step count (these can run to
// max_endgame_steps: the second pass starts them first)
enum { RC_HANDOFF = -1, RC_HANDOFF_STEPS = -2 };
enum InitKind : int { IK_NONE = 0, IK_TRACKER, IK_EG, IK_POLY_A, IK_POLY_B };
struct InitReq { int kind; cx t1, t0; double omega, mu, tau, max_init; bool keep_steps, ext; };
enum Mode : int { MODE_ENDGAME = 0, MODE_TRACKER = 1, MODE_POLYHEDRAL = 2 };

struct DevResults {  // caller's SoA, path-major (fields of src/path_result.jl:76-98)
    int* return_code; cx* solution; double* t; double* accuracy; double* residual; unsigned char* singular;
    double* condition_jacobian; int* winding_number; unsigned char* extended_precision; cx* last_point;
    double* last_t; double* valuation; unsigned char* has_valuation; double* omega; double* mu;
    int* accepted_steps; int* rejected_steps; int* steps_eg; unsigned char* extended_precision_used;
    long long* counters;  // 8 per path: factorizations, ldivs, evaljac, eval, eval_dd, taylor K=1, K=2, K=3
};

struct BatchIn {
    int mode;
    long long N;
    const cx* starts;          // path-major: starts[path * n + i]
    // start solutions produced on the device instead of streamed in:
    long long start_mod;       // > 0: path k starts from row k % start_mod of `starts` (many_solve: one start set for every parameter point)
    long long param_div;       // > 0: path k reads row k / param_div of the per-path parameters
    const cx* td_roots;        // total degree: roots of unity cis(2 pi j / d_i), j < d_i, variable after variable (host-made table)
    const int* td_degrees;     // total degree: d_1 .. d_n; path k <-> mixed-radix digits of td_first + k, first index fastest
    long long td_first;
    cx t1, t0;
    const double* omega_mu;    // optional 2 x N
    const int* cell_index;     // polyhedral: N (nullptr when the starts are made on the device: the cell follows from cell_first)
    const double* cell_weights;// polyhedral: ncells x P (unscaled s_ij, 0 on the cell's vertices)
    // binomial start solutions made on the device (BinomialSystemSolver, src/binomial_system.jl:55-106, 238-261):
    // per mixed cell the Hermite normal form H of its binomial system (n x n, row-major), the angles mu = U^T angle(b) / 2 pi
    // (mod 1) and the moduli r = exp(A^-T log|b|); path k of the batch = solution (k_first + k) - cell_first[cell] of its cell
    int ncells;
    const long long* cell_first;  // ncells + 1 prefix sums of the cell volumes
    const long long* bin_H;
    const double* bin_mu;
    const double* bin_r;
    long long k_first;
    // Two-pass batches (hc_api.cu: second_pass).  Pass 1, thread per path: a lane gives up a path (return_code =
    // RC_HANDOFF) that needs more than handoff_eg_steps endgame steps or handoff_steps tracker steps in total, or that
    // switches to extended precision -- one lane walks such a path 20 x slower than a CPU core while the rest of its CTA
    // waits.  The criteria depend on the path alone, so which pass tracks a path does not depend on timing.  Pass 2,
    // lane group per path: path k of the launch is path index_map[k] of the batch, tracked again from its start solution.
    int handoff_steps;            // 0 = off
    int handoff_eg_steps;
    int handoff_ext;
    const long long* index_map;   // nullptr = identity
};

template <int G, int S>
struct Lane : Path<G, S> {
    using B = Path<G, S>;
    using CV = typename B::CV; using RV = typename B::RV;
    using B::g; using B::H; using B::O; using B::M; using B::n; using B::pidx; using B::prow; using B::kind;
    using B::code; using B::accuracy; using B::omega; using B::mu; using B::tau; using B::winding;
    using B::extended_prec; using B::used_extended_prec; using B::refined_extended_prec; using B::keep_extended_prec;
    using B::accepted_steps; using B::rejected_steps; using B::factorized; using B::scaled;
    using B::min_step_size; using B::min_rel_step_size; using B::st_target;
    using B::n_fact; using B::n_ldiv; using B::n_evaljac; using B::n_eval; using B::n_evaldd; using B::n_tay1; using B::n_tay2; using B::n_tay3;
    using B::st_t; using B::tracker_init; using B::jac_cond; using B::cond_at; using B::skeel; using B::a_inf_norm;
    using B::refine_current_solution; using B::eval_f64; using B::vcopy;

    int phase, mode;
    int ev; bool stop_pending; InitReq req;  // parked event, deferred tracking_stopped!, pending tracker initialisation
    // ---- endgame state (src/endgame_tracker.jl:177-215)
    int eg_code; bool singular_endgame; int eg_winding;  // 0 = nothing
    double eg_accuracy, eg_cond; bool eg_singular;
    int steps_eg, ext_steps_eg_start;
    bool jtz_prev, jtz_cur, is_jump_to_zero;
    double last_t;
    int sidx0, sidx1, sidx2; double stime[3], scond[3];
    int singular_steps; double sing_t;
    double logt2, logt1;
    // ---- polyhedral
    int toric_acc, toric_rej; double poly_maxw, saved_min_step;
    int cell;  // mixed cell of the path
    int path_rounds, ho_steps, ho_eg_steps; bool ho_ext, ho_by_steps;  // tracker steps of the path so far (all stages); hand-off thresholds of the batch
    // ---- flop accounting totals of finished stages
    int c_fact, c_ldiv;

    // ================================================================ valuation
    HC_HD void val_init() { HC_COLD_N  HC_PAR(i, 12 * n) M.val[i] = 0.0; g.sync(); logt2 = logt1 = HC_NAN; }
    HC_HD static double fdiff(double v, double s, double v2, double s2, double v1, double s1) {
        double D1 = s - s1, D2 = s - s2, D12 = s1 - s2;
        return (D2 * v1) / (D12 * D1) - ((D12 + D2) * v2) / (D12 * D2) - (D12 * v) / (D1 * D2);
    }
    HC_HD static void nu_nu1(cx x, cx xd, cx x2, double t, double& nu, double& nu1) {
        double xx = abs2(x), mu_ = x.re * xd.re + x.im * xd.im, l = mu_ / xx;
        double mu1 = x.re * x2.re + xd.re * xd.re + x.im * x2.im + xd.im * xd.im;
        double l1 = mu1 / xx - 2 * (l * l);
        nu = t * l; nu1 = t * l1 + l;
    }
    HC_HDN void val_update(double t) {  // valuation.jl:82-124
        HC_COLD_N
        const int nn = n;
        const double logt = log(t);
        const bool diff = winding > 1 && logt2 == logt2;
        RV V = M.val;  // [val_x, val_tx, dval_x, dval_tx, vx2, vx1, vd2, vd1, lx2, lx1, ld2, ld1] x n
        HC_PAR(i, nn) {
            cx x = M.tx[i], xd = M.tx[nn + i], x2 = M.tx[2 * nn + i], x3 = M.tx[3 * nn + i];
            double logx = log(cabs(x)), logxd = log(cabs(xd));
            if (diff) {
                double nu = t * ((x.re * xd.re + x.im * xd.im) / abs2(x));
                double dnu = fdiff(nu, logt, V[4 * nn + i], logt2, V[5 * nn + i], logt1);
                V[i] = nu; V[2 * nn + i] = dnu;
                double vxd = fdiff(logxd, logt, V[10 * nn + i], logt2, V[11 * nn + i], logt1);
                double dvxd = fdiff(vxd, logt, V[6 * nn + i], logt2, V[7 * nn + i], logt1);
                V[nn + i] = vxd + 1; V[3 * nn + i] = dvxd;
            } else {
                double nu, nu1;
                nu_nu1(x, xd, 2.0 * x2, t, nu, nu1);
                V[i] = nu; V[2 * nn + i] = t * nu1;
                double o = V[4 * nn + i]; V[4 * nn + i] = nu; V[5 * nn + i] = o;
                double vxd;
                nu_nu1(xd, 2.0 * x2, 6.0 * x3, t, vxd, nu1);
                V[nn + i] = vxd + 1; V[3 * nn + i] = t * nu1;
            }
            double o = V[8 * nn + i]; V[8 * nn + i] = logx; V[9 * nn + i] = o;
            o = V[10 * nn + i]; V[10 * nn + i] = logxd; V[11 * nn + i] = o;
        }
        g.sync();
        logt1 = logt2; logt2 = logt;
    }
    HC_HD double eps_inf(int i) {
        HC_COLD_N
        double vx = M.val[i], vt = M.val[n + i];
        return jmax(jmax(fabs(1.0 - vt / vx), fabs(M.val[2 * n + i] / vx)), fabs(M.val[3 * n + i] / vt));
    }
    HC_HD bool val_is_finite() {  // valuation.jl:175-205
        HC_COLD_N
        const double ftol = O->val_finite_tol, delta = 1.0 / O->max_winding_number;
        const bool zero_is_finite = !O->zero_is_at_infinity;
        bool ok = true;
        HC_PAR(i, n) {
            double vx = M.val[i];
            if (fabs(vx) < ftol) {
                if (!(fabs(M.val[2 * n + i]) < ftol) || M.val[n + i] < 0.5 * delta) ok = false;
            } else if (zero_is_finite && vx > (delta - ftol)) {
                if (!(eps_inf(i) < ftol)) ok = false;
            } else ok = false;
        }
        return g.rall(ok);
    }
    HC_HD void estimate_winding(int& m, double& min_err) {  // valuation.jl:207-228
        HC_COLD_N
        m = 1; min_err = HC_INF;
        for (int k = 1; k <= O->max_winding_number; ++k) {
            double err = 0.0;
            HC_PAR(i, n) { double mv = k * M.val[n + i]; double e = fabs(rint(mv) - mv); err = e > err ? e : err; }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) { double w = g.xshfl(err, o); err = w > err ? w : err; }  // NaN never wins, as in the fold
            if (err < min_err) { m = k; min_err = err; }
        }
    }

    // ================================================================ endgame
    HC_HD double eg_jac_cond() { return jac_cond(&M.egrs, &M.egcs); }
    HC_HD void eg_scaling_update() {  // col_scaling .= weights(norm); row_scaling!(...)
        HC_COLD_N
        HC_PAR(i, n) M.egcs[i] = M.w[i];
        g.sync();
        skeel(M.egrs, M.egcs, O->scaling_threshold);
    }
    // init!(endgame_tracker, x, t1; omega, mu, extended_precision)  endgame_tracker.jl:260-294
    // (first half: min_rel_step_size = 0 and the tracker_init request, see eg_request; this is the rest)
    HC_HDN void eg_init_post() {
        HC_COLD_N
        eg_code = convert_code(code);
        singular_endgame = false; jtz_prev = jtz_cur = false;
        val_init(); eg_winding = 0;
        eg_accuracy = HC_NAN; eg_cond = HC_NAN; eg_singular = false; steps_eg = 0; ext_steps_eg_start = 0x3fffffff;
        HC_PAR(i, n) { M.sol[i] = mk(HC_NAN, HC_NAN); M.egrs[i] = 1.0; M.egcs[i] = 1.0; M.ais[i] = HC_NAN; M.lastp[i] = M.x[i]; }
        g.sync();
        singular_steps = 0; sidx0 = 0; sidx1 = 1; sidx2 = 2;
        last_t = HC_NAN;
    }
    HC_HDN void tracking_stopped() {  // :695-721
        HC_COLD_N
        eg_accuracy = accuracy;
        if (eg_code == EG_success && eg_accuracy > 1e-14) refine_current_solution(1e-14, O->refine_steps);
        vcopy(M.sol, M.x, n);
        eg_winding = 0;
        if (eg_code == EG_success) {
            eg_scaling_update();
            eg_cond = cond_at(M.sol, mk(0.0), &M.egrs, &M.egcs);
            eg_singular = eg_cond > O->sing_cond || eg_accuracy > O->sing_accuracy;
        }
    }
    HC_HD bool check_finite() {  // :402-422
        if (!val_is_finite()) return false;
        int mw; double merr;
        estimate_winding(mw, merr);
        if (merr < O->val_finite_tol) {
            if (mw == 1 && !jtz_prev) return false;
            eg_winding = mw;
            return true;
        }
        return false;
    }
    HC_HDN bool check_at_infinity() {  // :424-499
        HC_COLD_N
        if (!O->at_infinity_check) return false;
        const double ftol = O->val_finite_tol;
        const bool zero_is_finite = !O->zero_is_at_infinity;
        double kappa = HC_NAN;
        const double t = st_t().re;
        HC_PAR(i, n) {
            // at_infinity_tol!  valuation.jl:143-173
            double vx = M.val[i], e = eps_inf(i), tol;
            if (e != e) tol = HC_INF;
            else if (vx + e < -ftol) tol = e;
            else if (!zero_is_finite && vx - e > -ftol) tol = e;
            else tol = HC_INF;
            M.ait[i] = tol;
        }
        g.sync();
        // the coordinate loop carries state (kappa, first-flag scaling): every lane walks it with the same values
        for (int i = 0; i < n; ++i) {
            if (M.ait[i] < O->val_at_infinity_tol) {
                if (M.ais[i] != M.ais[i]) {
                    bool allnan = true;
                    for (int k = 0; k < n; ++k) allnan = allnan && (M.ais[k] != M.ais[k]);
                    if (allnan) eg_scaling_update();
                    kappa = eg_jac_cond();
                    double ax = cabs(M.x[i]);
                    g.sync();
                    if (g.lane == 0) { M.aic[i] = kappa; M.aia[i] = ax; M.ais[i] = t; }
                    g.sync();
                } else {
                    if (kappa != kappa) kappa = eg_jac_cond();
                    double v = M.val[i];
                    bool at_zero = v > 0;
                    double cond_growth = kappa / M.aic[i];
                    double coord_growth = at_zero ? M.aia[i] / cabs(M.x[i]) : cabs(M.x[i]) / M.aia[i];
                    if (coord_growth > clampd(pow(0.25, 4 * v), 20.0, O->min_coord_growth) &&
                        (cond_growth > O->min_cond_growth || kappa > jmax(1e8, O->min_cond))) {
                        eg_code = at_zero ? EG_at_zero : EG_at_infinity;
                        return true;
                    }
                }
            } else if (M.ais[i] == M.ais[i]) {
                g.sync();
                if (g.lane == 0) M.ais[i] = HC_NAN;
                g.sync();
            }
        }
        return false;
    }
    HC_HDN void add_sample(double t) {  // :630-662
        HC_COLD_N
        const int nn = n;
        const int mw = eg_winding;
        double s = nthroot(t, mw), mu_ = mw;
        for (int k = 0; k < mw - 1; ++k) mu_ *= s;
        double kappa = eg_jac_cond();
        int slot;
        if (singular_steps <= 2) slot = singular_steps;
        else {
            int first = sidx0;
            sidx0 = sidx1; stime[0] = stime[1]; scond[0] = scond[1];
            sidx1 = sidx2; stime[1] = stime[2]; scond[1] = scond[2];
            sidx2 = first; slot = 2;
        }
        int sid = slot == 0 ? sidx0 : (slot == 1 ? sidx1 : sidx2);
        CV ty = M.samp.at(sid * 2 * nn);
        HC_PAR(i, nn) { ty[i] = M.tx[i]; ty[nn + i] = mu_ * M.tx[nn + i]; }
        g.sync();
        if (slot == 0) { stime[0] = s; scond[0] = kappa; } else if (slot == 1) { stime[1] = s; scond[1] = kappa; } else { stime[2] = s; scond[2] = kappa; }
    }
    HC_HDN double predict_endpoint() {  // :664-693
        HC_COLD_N
        const int nn = n;
        if (singular_steps < 2) return HC_INF;
        CV S0 = M.samp.at(sidx0 * 2 * nn), S1 = M.samp.at(sidx1 * 2 * nn), S2 = M.samp.at(sidx2 * 2 * nn);
        if (singular_steps == 2) B::cubic_hermite(M.pred, S0, S0.at(nn), mk(stime[0]), S1, S1.at(nn), mk(stime[1]), mk(0.0));
        vcopy(M.ppred, M.pred, nn);
        B::cubic_hermite(M.pred, S1, S1.at(nn), mk(stime[1]), S2, S2.at(nn), mk(stime[2]), mk(0.0));
        double p = stime[2] / stime[1], p2 = p * p;
        double err = B::inf_dist(M.pred, M.ppred) / fabs(p2 * p2 - 1);
        double ns = B::inf_norm(M.pred);
        if (ns > 1e-8) err /= ns;
        return err;
    }
    HC_HDN void switch_to_singular() {  // :501-524
        HC_COLD_N
        singular_endgame = true;
        double t = st_t().re;
        bool allone = true;
        for (int i = 0; i < n; ++i) allone = allone && (M.egrs[i] == 1.0);
        if (allone) eg_scaling_update();
        add_sample(t);
        singular_steps = 0;
        winding = eg_winding;
        g.sync();
        if (g.lane == 0) M.aic[0] = scond[0];
        g.sync();
        keep_extended_prec = true;
    }
    HC_HD void switch_to_regular() {  // :525-530
        singular_endgame = false;
        winding = 1;
        B::tracker_init_continue(mk(0.0));
        phase = PH_EG;
    }
    // First half of step!(::EndgameTracker): returns false when the path ended without a tracker step.
    HC_HDN bool eg_pre() {  // :329-358
        HC_COLD_N
        if (steps_eg >= O->max_endgame_steps) { eg_code = EG_terminated_max_steps; return false; }
        if (B::ext_steps() - ext_steps_eg_start > O->max_endgame_extended_steps) {
            bool nonan = true;
            for (int i = 0; i < n; ++i) nonan = nonan && !cisnan(M.sol[i]);
            if (nonan && eg_winding != 0 && eg_accuracy < O->singular_min_accuracy) {
                eg_cond = cond_at(M.sol, mk(0.0), &M.egrs, &M.egcs);
                eg_singular = true; eg_code = EG_success;
            } else eg_code = EG_terminated_max_extended_steps;
            return false;
        }
        vcopy(M.lastp, M.x, n);
        last_t = st_t().re;
        if (singular_endgame) {  // begin singular_endgame_step!  :533-540
            sing_t = st_t().re;
            B::tracker_init_continue(mk(0.25 * sing_t));
            phase = PH_SING;
            return true;
        }
        is_jump_to_zero = ciszero(B::st_tp());
        return true;
    }
    // Second half of step!(::EndgameTracker) after a regular tracker step  :362-398
    HC_HDN void eg_post(bool step_success) {
        eg_code = convert_code(code);
        if (eg_code != EG_tracking) { stop_pending = true; return; }  // tracking_stopped! runs in the event round
        jtz_prev = jtz_cur; jtz_cur = is_jump_to_zero;
        double t = st_t().re;
        if (!(t <= O->endgame_start)) return;
        if (steps_eg == 0) ext_steps_eg_start = B::ext_steps();
        steps_eg += 1;
        if (!step_success) return;
        val_update(t);
        if (check_finite()) { switch_to_singular(); return; }
        check_at_infinity();
    }
    // After each tracker step of the singular endgame's inner loop  :541-628
    HC_HDN void sing_post() {
        HC_COLD_N
        bool max_steps = false;
        if ((steps_eg += 1) >= O->max_endgame_steps) { eg_code = EG_terminated_max_steps; max_steps = true; }
        else if (B::ext_steps() - ext_steps_eg_start > O->max_endgame_extended_steps) { eg_code = EG_terminated_max_extended_steps; max_steps = true; }
        if (!max_steps && code == TC_tracking) return;  // inner loop continues
        phase = PH_EG;
        singular_steps += 1;
        const double lt = 0.25 * sing_t;
        if (!max_steps) {
            if (code != TC_success) { eg_code = convert_code(code); stop_pending = true; return; }
            val_update(lt);
            int mh; double mh_err;
            estimate_winding(mh, mh_err);
            if (mh != eg_winding || mh_err > 0.1) { switch_to_regular(); return; }
            add_sample(lt);
            if (singular_steps < 2) return;
            double acc = predict_endpoint();
            if (singular_steps == 2 || (acc < eg_accuracy && eg_accuracy > 1e-12)) {
                eg_accuracy = acc;
                vcopy(M.sol, M.pred, n);
                return;
            }
        }
        const int mw = eg_winding;
        const double kappa = scond[2], zero_cond = 1.0 / (mw + 1);
        HC_PAR(i, n) M.sol[i] = (M.val[i] < zero_cond ? 1.0 : 0.0) * M.pred[i];
        g.sync();
        double kappa0 = cond_at(M.sol, mk(0.0), &M.egrs, &M.egcs);
        double J0 = a_inf_norm(&M.egrs, nullptr);
        if (eg_accuracy < O->singular_min_accuracy &&
            (((mw > 1 && kappa > O->min_cond && nanmax(kappa0, 1.0 / J0) > kappa) || (mw == 1 && kappa0 > 1e12)) || max_steps ||
             (n == 1 && 1.0 / J0 < O->min_cond))) {
            eg_cond = jmax(kappa0, 1.0 / J0);
            eg_singular = true;
            eg_code = EG_success;
        } else if (!max_steps) switch_to_regular();
    }

    // ================================================================ polyhedral stage 1
    // update_weights!(H, support, lifting, cell; min_weight | max_weight)  toric_homotopy.jl:66-112
    HC_HD void set_weights(const double* raw, bool use_min, double target, double& smin, double& smax) {
        const int P = H->P;
        smax = 0.0; smin = HC_INF;
        for (int l = 0; l < P; ++l) { double r = raw[l]; if (r != 0.0) { smax = jmax(smax, r); smin = jmin(smin, r); } }
        double lam = use_min ? smin / target : smax / target;
        B::tape_prog = B::tay_prog = nullptr; B::pv_kind = B::ps_kind = -1;  // new weights: cached toric parameters are stale
        g.sync();
        HC_PAR(l, P) M.tw[l] = raw[l] / lam;
        g.sync();
        if (use_min) { smax = smax / lam; smin = target; } else { smin = smin / lam; smax = target; }
    }

    // ================================================================ results
    HC_HDN void write_result(const DevResults& R, int rc, bool success_at_zero) {
        HC_COLD_N
        const long long k = pidx;
        const int nn = n;
        CV sol = success_at_zero ? M.sol : M.x;
        const double tt = success_at_zero ? 0.0 : st_t().re;
        HC_PAR(i, nn) { R.solution[k * nn + i] = sol[i]; R.last_point[k * nn + i] = M.lastp[i]; }
        if (g.lane == 0) {
            R.return_code[k] = rc;
            R.t[k] = tt;
            R.last_t[k] = last_t;
            R.omega[k] = omega; R.mu[k] = mu;
            R.extended_precision[k] = extended_prec; R.extended_precision_used[k] = used_extended_prec;
            R.accepted_steps[k] = accepted_steps; R.rejected_steps[k] = rejected_steps;
            if (R.counters) {
                long long* c = R.counters + 8 * k;
                c[0] = c_fact + n_fact; c[1] = c_ldiv + n_ldiv; c[2] = n_evaljac; c[3] = n_eval; c[4] = n_evaldd; c[5] = n_tay1; c[6] = n_tay2; c[7] = n_tay3;
            }
        }
    }
    HC_HDN void finish_eg(const DevResults& R) {  // PathResult(::EndgameTracker)  :847-888
        HC_COLD_N
        const long long k = pidx;
        const int nn = n;
        const bool ok = eg_code == EG_success;
        const double tt = ok ? 0.0 : st_t().re;
        eval_f64(M.r, nullptr, ok ? M.sol : M.x, mk(tt));
        const double res = B::inf_norm(M.r);
        write_result(R, eg_code, ok);
        HC_PAR(i, nn) R.valuation[k * nn + i] = M.val[i];
        if (g.lane == 0) {
            R.residual[k] = res;
            R.singular[k] = eg_singular; R.accuracy[k] = eg_accuracy; R.condition_jacobian[k] = eg_cond;
            R.winding_number[k] = eg_winding; R.steps_eg[k] = steps_eg;
            R.has_valuation[k] = !(tt > O->endgame_start);
            if (mode == MODE_POLYHEDRAL) { R.accepted_steps[k] += toric_acc; R.rejected_steps[k] += toric_rej; }
        }
        phase = PH_IDLE;
    }
    HC_HDN void finish_plain(const DevResults& R) {  // TrackerResult  tracker.jl:998-1012
        HC_COLD_N
        const long long k = pidx;
        const int nn = n;
        vcopy(M.lastp, M.x, nn);
        last_t = st_t().im;
        write_result(R, code, false);
        HC_PAR(i, nn) R.valuation[k * nn + i] = HC_NAN;
        if (g.lane == 0) {
            R.accuracy[k] = accuracy; R.residual[k] = HC_NAN; R.singular[k] = 0; R.condition_jacobian[k] = tau;
            R.winding_number[k] = 0; R.steps_eg[k] = 0; R.has_valuation[k] = 0;
            R.extended_precision[k] = extended_prec || refined_extended_prec;
            R.extended_precision_used[k] = used_extended_prec || refined_extended_prec;
        }
        phase = PH_IDLE;
    }
    HC_HDN void finish_poly_failed(const DevResults& R) {  // polyhedral.jl:491-513
        HC_COLD_N
        const long long k = pidx;
        const int nn = n;
        vcopy(M.lastp, M.x, nn);
        last_t = st_t().re;
        write_result(R, EG_polyhedral_failed, false);
        HC_PAR(i, nn) R.valuation[k * nn + i] = 0.0;
        if (g.lane == 0) {
            R.accuracy[k] = accuracy; R.residual[k] = HC_NAN; R.singular[k] = 0; R.condition_jacobian[k] = HC_NAN;
            R.winding_number[k] = 0; R.steps_eg[k] = 0; R.has_valuation[k] = 0;
        }
        phase = PH_IDLE;
    }

    // ================================================================ driver
    HC_HD void eg_request(double t1, double omega_, double mu_) {  // init!(endgame_tracker, ...) first half
        min_rel_step_size = 0.0;
        req.kind = IK_EG; req.t1 = mk(t1); req.t0 = mk(0.0); req.omega = omega_; req.mu = mu_; req.tau = HC_INF; req.max_init = HC_INF;
        req.keep_steps = false; req.ext = false;
    }
    // Start solution k of a polyhedral batch from its index (reference: PolyhedralStartSolutionsIterator +
    // BinomialSystemSolver, src/polyhedral.jl:104-144, src/binomial_system.jl:55-106, 238-261): the cell by binary search
    // in the volume prefix sums, the unit-root combination from the mixed-radix digits of the local index
    // (fill_unit_roots_combinations!: first coordinate slowest), the angles by the triangular solve over H in
    // double-double arithmetic reduced mod 2 as the reference does, x_j = r_j cis(2 pi alpha_j).
    HC_HDN void binomial_start(long long k, const BatchIn& Bt) {
        HC_COLD_N
        const int nn = n;
        const long long kk = (Bt.k_first + k) % Bt.cell_first[Bt.ncells];  // indices beyond the mixed volume wrap (replicated batches)
        int lo = 0, hi = Bt.ncells;
        while (hi - lo > 1) { const int mid = (lo + hi) / 2; if (Bt.cell_first[mid] <= kk) lo = mid; else hi = mid; }
        cell = lo;
        long long c = kk - Bt.cell_first[lo];
        const long long* Hm = Bt.bin_H + (size_t)lo * nn * nn;
        const double* mu = Bt.bin_mu + (size_t)lo * nn;
        const double* rr = Bt.bin_r + (size_t)lo * nn;
        // alpha_j lives in M.work[j] as a double-double (re = hi, im = lo); every lane of a group computes all of them
        long long tail = 1;  // product of the diagonal entries behind j
        for (int j = nn - 1; j >= 0; --j) {
            const long long dj = Hm[j * nn + j];
            const long long root = (c / tail) % dj;
            tail *= dj;
            dd a = (mkdd(mu[j]) + mkdd((double)root)) / mkdd((double)dj);
            a = a - mkdd(2.0 * rint(0.5 * a.hi));
            for (int q = nn - 1; q > j; --q) {
                const cx w = M.work[q];
                dd ak = (mkdd(w.re, w.im) * (double)Hm[q * nn + j]) / mkdd((double)dj);
                ak = ak - mkdd(2.0 * rint(0.5 * ak.hi));
                a = a - ak;
            }
            a = a - mkdd(2.0 * rint(0.5 * a.hi));  // rem(alpha, 2, RoundNearest)
            g.sync();
            if (g.lane == 0) M.work[j] = mk(a.hi, a.lo);
            g.sync();
        }
        HC_PAR(j, nn) {
            const double a = ((cx)M.work[j]).re;
            M.x[j] = mk(rr[j] * hc_cospi(2.0 * a), rr[j] * hc_sinpi(2.0 * a));
        }
    }
    // a new path: load the start, reset the per-path state, request the first tracker initialisation
    HC_HDN void start_pre(long long k, const BatchIn& Bt) {
        HC_COLD_N
        if (Bt.index_map) k = Bt.index_map[k];
        path_rounds = 0; ho_steps = Bt.handoff_steps; ho_eg_steps = Bt.handoff_eg_steps; ho_ext = Bt.handoff_ext != 0;
        pidx = k; mode = Bt.mode;
        const int nn = n;
        g.sync();
        prow = Bt.param_div > 0 ? k / Bt.param_div : k;
        if (Bt.td_roots) {  // TotalDegreeStartSolutionsIterator  total_degree.jl:235-262
            long long idx = Bt.td_first + k;
            int off = 0;
            for (int i = 0; i < nn; ++i) {
                const int d = Bt.td_degrees[i];
                const int j = (int)(idx % d);
                idx /= d;
                if (i % G == g.lane) M.x[i] = Bt.td_roots[off + j];
                off += d;
            }
        } else if (Bt.bin_H) binomial_start(k, Bt);
        else {
            const long long row = Bt.start_mod > 0 ? k % Bt.start_mod : k;
            HC_PAR(i, nn) M.x[i] = Bt.starts[row * nn + i];
        }
        if (Bt.mode == MODE_POLYHEDRAL && !Bt.bin_H) cell = Bt.cell_index[k];
        g.sync();
        refined_extended_prec = false; factorized = scaled = false; B::a_in_lu = B::rs_raw = false; stop_pending = false;
        min_step_size = O->min_step_size; min_rel_step_size = O->min_rel_step_size;
        B::tape_prog = B::tay_prog = nullptr; B::pv_kind = B::ps_kind = -1;
#if defined(HC_JIT_GEN)
        B::jit_bind_params();
#endif
        B::tol_acc_limit = pow(O->a, (double)((1 << O->min_newton_iters) - 1)) * hfun(O->a);
        n_fact = n_ldiv = n_evaljac = n_eval = n_evaldd = n_tay1 = n_tay2 = n_tay3 = 0; c_fact = c_ldiv = 0;
        toric_acc = toric_rej = 0;
        double om = HC_NAN, mu_ = HC_NAN;
        if (Bt.omega_mu) { om = Bt.omega_mu[2 * k]; mu_ = Bt.omega_mu[2 * k + 1]; }
        if (mode == MODE_TRACKER) {
            kind = H->kind;
            req.kind = IK_TRACKER; req.t1 = Bt.t1; req.t0 = Bt.t0; req.omega = om; req.mu = mu_; req.tau = HC_INF; req.max_init = HC_INF;
            req.keep_steps = false; req.ext = false;
        } else if (mode == MODE_ENDGAME) {
            kind = H->kind;
            eg_request(Bt.t1.re, om, mu_);
        } else {  // polyhedral.jl:414-465
            kind = H_TORIC;
            double smin, smax;
            const double* raw = Bt.cell_weights + (size_t)cell * H->P;
            set_weights(raw, true, 1.0, smin, smax);
            poly_maxw = smax;
            double tend = smax < 10 ? 1.0 : clampd(pow(0.1, 10 / smax), 0.9, 1 - 1e-6);
            req.kind = IK_POLY_A; req.t1 = mk(0.0); req.t0 = mk(tend); req.omega = 20.0; req.mu = 1e-12; req.tau = HC_INF; req.max_init = 0.2;
            req.keep_steps = false; req.ext = false;
        }
    }
    // the toric stage of a polyhedral path ended: re-weighted restart (:466-489) or hand-over to the
    // coefficient homotopy (:491-529)
    HC_HDN void toric_pre(const BatchIn& Bt) {
        if (phase == PH_TORIC_A && poly_maxw >= 10 && code == TC_success) {
            double smin, smax;
            const double* raw = Bt.cell_weights + (size_t)cell * H->P;
            double t0 = st_target.re;
            set_weights(raw, false, 10.0, smin, smax);
            double t_restart = pow(t0, 1 / smin);
            saved_min_step = min_step_size; min_step_size = 0.0;
            c_fact += n_fact; c_ldiv += n_ldiv;
            req.kind = IK_POLY_B; req.t1 = mk(t_restart); req.t0 = mk(1.0); req.omega = omega; req.mu = mu; req.tau = 0.1 * t_restart;
            req.max_init = HC_INF; req.keep_steps = true; req.ext = false;
            return;
        }
        if (phase == PH_TORIC_B) min_step_size = saved_min_step;
        if (code != TC_success) { ev = EV_FINISH_POLY_FAILED; return; }
        toric_acc = accepted_steps; toric_rej = rejected_steps;
        c_fact += n_fact; c_ldiv += n_ldiv;
        kind = H_COEFFICIENT;
        eg_request(1.0, HC_NAN, mu);  // omega deliberately not passed (:515-521)
    }
    // Event round, part 1: lanes whose path ended write their PathResult and become free (EV_START).
    HC_HD void event_finish(const DevResults& R) {
        if (ev == EV_FINISH_EG) {
            if (stop_pending) { tracking_stopped(); stop_pending = false; }
            finish_eg(R);
            ev = EV_START;
        } else if (ev == EV_FINISH_PLAIN) { finish_plain(R); ev = EV_START; }
        else if (ev == EV_FINISH_POLY_FAILED) { finish_poly_failed(R); ev = EV_START; }
        else if (ev == EV_HANDOFF) { if (g.lane == 0) R.return_code[pidx] = ho_by_steps ? RC_HANDOFF_STEPS : RC_HANDOFF; phase = PH_IDLE; ev = EV_START; }
    }
    // Event round, part 2: free lanes take path k (if any is left), toric lanes decide their next stage; all
    // requested tracker initialisations then run from one call site.
    HC_HD void event_begin(long long k, bool have_k, const BatchIn& Bt, const DevResults&) {
        req.kind = IK_NONE;
        if (ev == EV_START) {
            if (have_k) start_pre(k, Bt);
            else { ev = EV_NONE; phase = PH_IDLE; return; }  // queue drained: this lane is done
        } else if (ev == EV_TORIC_NEXT) toric_pre(Bt);
        if (req.kind == IK_NONE) return;  // (EV_FINISH_POLY_FAILED was set: handled in the next round)
        tracker_init(req.t1, req.t0, req.omega, req.mu, req.tau, req.max_init, req.keep_steps, req.ext);
        ev = EV_NONE;
        switch (req.kind) {
            case IK_TRACKER: phase = PH_PLAIN; if (code != TC_tracking) ev = EV_FINISH_PLAIN; break;
            case IK_EG: eg_init_post(); phase = PH_EG; if (eg_code != EG_tracking) ev = EV_FINISH_EG; break;
            case IK_POLY_A: phase = PH_TORIC_A; if (code != TC_tracking) ev = EV_TORIC_NEXT; break;
            default: phase = PH_TORIC_B; if (code != TC_tracking) ev = EV_TORIC_NEXT; break;  // IK_POLY_B
        }
    }
    // One flat iteration of an active lane (ev == EV_NONE, phase != PH_IDLE).
    // SYNC = true (lockstep kernels): called by every thread of the CTA each round, `act` = this lane is active.
    template <bool SYNC>
    HC_HD void iterate_t(bool act) {
        bool do_step = act;
        if (act && phase == PH_EG) do_step = eg_pre();
        const bool ok = B::template tracker_step_t<SYNC>(do_step);
        if (!act) return;
        switch (phase) {
            case PH_PLAIN: if (code != TC_tracking) ev = EV_FINISH_PLAIN; break;
            case PH_TORIC_A: case PH_TORIC_B: if (code != TC_tracking) ev = EV_TORIC_NEXT; break;
            case PH_SING: sing_post(); if (eg_code != EG_tracking) ev = EV_FINISH_EG; break;
            case PH_EG: if (do_step) eg_post(ok); if (eg_code != EG_tracking) ev = EV_FINISH_EG; break;
            default: break;
        }
        if (ho_steps > 0 && ev == EV_NONE) {
            const bool by_steps = ++path_rounds > ho_steps || ((phase == PH_EG || phase == PH_SING) && steps_eg > ho_eg_steps);
            if (by_steps || (ho_ext && extended_prec)) { ev = EV_HANDOFF; ho_by_steps = by_steps; }
        }
    }
    HC_HD void iterate(const BatchIn&, const DevResults&) { iterate_t<false>(true); }
};

}  // namespace hc
