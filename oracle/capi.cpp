// ORACLE (test infrastructure, NOT product code): C entry points + the threaded driver.
//
// The driver mirrors threaded_solve (reference src/solve.jl:628-709): one tracker copy per
// thread (:643-650), an atomic next-index counter (:641, 660-667), results written by
// path index (:637, 670).
#include "hc_oracle.h"

#include <atomic>
#include <map>
#include <mutex>
#include <memory>
#include <thread>

#include "endgame.hpp"

using namespace orc;

namespace {

Program make_program(const orc_program_desc* d) {
    Program p;
    p.instr.resize(d->n_instructions);
    for (int i = 0; i < d->n_instructions; ++i) {
        const int32_t* s = d->instructions + 6 * (size_t)i;
        Instr I; I.in[0] = s[0]; I.in[1] = s[1]; I.in[2] = s[2]; I.in[3] = s[3]; I.op = s[4]; I.out = s[5];
        p.instr[i] = I;
    }
    p.constants.resize(d->n_constants);
    for (int i = 0; i < d->n_constants; ++i) p.constants[i] = cplx(d->constants[2 * i], d->constants[2 * i + 1]);
    p.param_off = d->param_offset; p.P = d->n_params; p.t_index = d->t_index;
    p.var_off = d->var_offset; p.n = d->n_vars; p.out_dim = d->out_dim; p.tape_space = d->tape_space;
    for (int i = 0; i < d->n_u; ++i) p.u_assign.push_back({d->u_assign[2 * i], d->u_assign[2 * i + 1]});
    for (int i = 0; i < d->n_U; ++i) p.U_assign.push_back({d->U_assign[2 * i], d->U_assign[2 * i + 1]});
    finalize(p);
    return p;
}

std::vector<cplx> cvec(const double* p, int n) {
    std::vector<cplx> v(n);
    for (int i = 0; i < n; ++i) v[i] = cplx(p[2 * i], p[2 * i + 1]);
    return v;
}

void unpack_options(const orc_options* o, TrackerOptions& T, EndgameOptions& E, WeightedNormOptions& W) {
    T.max_steps = o->max_steps; T.max_step_size = o->max_step_size; T.max_initial_step_size = o->max_initial_step_size;
    T.extended_precision = o->extended_precision != 0; T.min_step_size = o->min_step_size; T.min_rel_step_size = o->min_rel_step_size;
    T.parameters.a = o->a; T.parameters.beta_a = o->beta_a; T.parameters.beta_omega_p = o->beta_omega_p;
    T.parameters.beta_tau = o->beta_tau; T.parameters.strict_beta_tau = o->strict_beta_tau;
    T.parameters.min_newton_iters = o->min_newton_iters;
    E.endgame_start = o->endgame_start; E.max_endgame_steps = o->max_endgame_steps;
    E.max_endgame_extended_steps = o->max_endgame_extended_steps; E.min_cond = o->min_cond;
    E.min_cond_growth = o->min_cond_growth; E.min_coord_growth = o->min_coord_growth;
    E.zero_is_at_infinity = o->zero_is_at_infinity != 0; E.at_infinity_check = o->at_infinity_check != 0;
    E.only_nonsingular = o->only_nonsingular != 0; E.singular_min_accuracy = o->singular_min_accuracy;
    E.max_winding_number = o->max_winding_number; E.val_finite_tol = o->val_finite_tol;
    E.val_at_infinity_tol = o->val_at_infinity_tol; E.sing_cond = o->sing_cond; E.sing_accuracy = o->sing_accuracy;
    E.scaling_threshold = o->scaling_threshold; E.refine_steps = o->refine_steps;
    W.scale_min = o->scale_min; W.scale_abs_min = o->scale_abs_min; W.scale_max = o->scale_max;
}

void store(const PathResult& R, orc_results* out, int64_t k, int n) {
    out->return_code[k] = R.return_code;
    for (int i = 0; i < n; ++i) {
        out->solution[2 * (k * n + i)] = R.solution[i].re; out->solution[2 * (k * n + i) + 1] = R.solution[i].im;
        out->last_point[2 * (k * n + i)] = R.last_point[i].re; out->last_point[2 * (k * n + i) + 1] = R.last_point[i].im;
        out->valuation[k * n + i] = R.valuation.empty() ? NaN : R.valuation[i];
    }
    out->t[k] = R.t; out->accuracy[k] = R.accuracy; out->residual[k] = R.residual; out->singular[k] = R.singular;
    out->condition_jacobian[k] = R.condition_jacobian; out->winding_number[k] = R.winding_number;
    out->extended_precision[k] = R.extended_precision; out->last_t[k] = R.last_t; out->has_valuation[k] = R.has_valuation;
    out->omega[k] = R.omega; out->mu[k] = R.mu; out->accepted_steps[k] = R.accepted_steps;
    out->rejected_steps[k] = R.rejected_steps; out->steps_eg[k] = R.steps_eg;
    out->extended_precision_used[k] = R.extended_precision_used;
    if (out->counters) {
        for (int c = 0; c < 8; ++c) out->counters[8 * k + c] = 0;
        out->counters[8 * k] = R.n_factorizations; out->counters[8 * k + 1] = R.n_ldivs;
    }
}

template <class Fn>
void parallel_for(int64_t N, int nthreads, Fn make_worker) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<int64_t> next{0};
    auto body = [&](int tid) {
        auto worker = make_worker(tid);
        while (true) {
            int64_t k = next.fetch_add(1);
            if (k >= N) break;
            worker(k);
        }
    };
    if (nthreads == 1) { body(0); return; }
    std::vector<std::thread> th;
    for (int i = 0; i < nthreads; ++i) th.emplace_back(body, i);
    for (auto& t : th) t.join();
}

static std::mutex g_hook_mutex;
static std::map<const void*, std::vector<double>> g_hook_weights;
}  // namespace

extern "C" {

void orc_options_default(orc_options* o) {
    TrackerOptions T; EndgameOptions E; WeightedNormOptions W;
    o->max_steps = T.max_steps; o->max_step_size = T.max_step_size; o->max_initial_step_size = T.max_initial_step_size;
    o->extended_precision = T.extended_precision; o->min_step_size = T.min_step_size; o->min_rel_step_size = T.min_rel_step_size;
    o->a = T.parameters.a; o->beta_a = T.parameters.beta_a; o->beta_omega_p = T.parameters.beta_omega_p;
    o->beta_tau = T.parameters.beta_tau; o->strict_beta_tau = T.parameters.strict_beta_tau;
    o->min_newton_iters = T.parameters.min_newton_iters;
    o->endgame_start = E.endgame_start; o->max_endgame_steps = E.max_endgame_steps;
    o->max_endgame_extended_steps = E.max_endgame_extended_steps; o->min_cond = E.min_cond;
    o->min_cond_growth = E.min_cond_growth; o->min_coord_growth = E.min_coord_growth;
    o->zero_is_at_infinity = E.zero_is_at_infinity; o->at_infinity_check = E.at_infinity_check;
    o->only_nonsingular = E.only_nonsingular; o->singular_min_accuracy = E.singular_min_accuracy;
    o->max_winding_number = E.max_winding_number; o->val_finite_tol = E.val_finite_tol;
    o->val_at_infinity_tol = E.val_at_infinity_tol; o->sing_cond = E.sing_cond; o->sing_accuracy = E.sing_accuracy;
    o->scaling_threshold = E.scaling_threshold; o->refine_steps = E.refine_steps;
    o->scale_min = W.scale_min; o->scale_abs_min = W.scale_abs_min; o->scale_max = W.scale_max;
}

void* orc_system_create(const orc_program_desc* eval, const orc_program_desc* jac) {
    System* s = new System();
    s->eval = make_program(eval);
    s->jac = make_program(jac);
    if (!s->eval.supported || !s->jac.supported) { delete s; return nullptr; }
    return s;
}
void orc_system_destroy(void* s) { delete (System*)s; }

void* orc_homotopy_create(const orc_homotopy_desc* d) {
    HomotopyDef* H = new HomotopyDef();
    H->kind = (HKind)d->kind;
    H->F = (const System*)d->F; H->G = (const System*)d->G;
    H->gamma = cplx(d->gamma[0], d->gamma[1]);
    if (d->G_params) H->G_params = cvec(d->G_params, d->n_G_params);
    if (d->F_params) H->F_params = cvec(d->F_params, d->n_F_params);
    if (d->kind == H_TORIC) H->system_coeffs = cvec(d->p, d->n_pq);
    else if (d->kind != H_STRAIGHT_LINE) { H->p = cvec(d->p, d->n_pq); H->q = cvec(d->q, d->n_pq); }
    return H;
}
void orc_homotopy_destroy(void* h) {
    { std::lock_guard<std::mutex> lock(g_hook_mutex); g_hook_weights.erase(h); }
    delete (HomotopyDef*)h;
}

int32_t orc_track_batch(void* Hv, const orc_options* o, int32_t mode, int64_t N, const double* starts, const double* t1,
                        const double* t0, const double* path_p, const double* path_q, const double* omega_mu,
                        orc_results* out, int32_t nthreads) {
    const HomotopyDef* D = (const HomotopyDef*)Hv;
    TrackerOptions T; EndgameOptions E; WeightedNormOptions W;
    unpack_options(o, T, E, W);
    const int n = D->n();
    const int P = D->F->eval.P;
    try {
        parallel_for(N, nthreads, [&](int) {
            auto eg = std::make_shared<EndgameTracker>();
            eg->setup(D, T, E, W);
            return [=](int64_t k) {
                Tracker& tr = eg->tracker;
                if (path_p) tr.H.p = cvec(path_p + 2 * k * P, P);
                if (path_q) tr.H.q = cvec(path_q + 2 * k * P, P);
                std::vector<cplx> x = cvec(starts + 2 * k * n, n);
                double om = omega_mu ? omega_mu[2 * k] : NaN, mu = omega_mu ? omega_mu[2 * k + 1] : NaN;
                if (mode == 0) {
                    eg->track(x.data(), t1[0], om, mu);
                    store(eg->path_result(), out, k, n);
                } else {
                    tr.state.refined_extended_prec = false;
                    tr.track(x.data(), cplx(t1[0], t1[1]), cplx(t0[0], t0[1]), om, mu);
                    PathResult R;  // TrackerResult, tracker.jl:998-1012
                    TrackerState& st = tr.state;
                    R.return_code = st.code; R.solution = st.x; R.t = st.t().re; R.accuracy = st.accuracy;
                    R.last_point = st.x; R.last_t = st.t().im; R.valuation.assign(n, NaN);
                    R.extended_precision = st.extended_prec || st.refined_extended_prec;
                    R.extended_precision_used = st.used_extended_prec || st.refined_extended_prec;
                    R.omega = st.omega; R.mu = st.mu; R.accepted_steps = st.accepted_steps; R.rejected_steps = st.rejected_steps;
                    R.condition_jacobian = st.tau;  // TrackerResult.tau travels in this slot
                    R.n_factorizations = st.jacobian.factorizations; R.n_ldivs = st.jacobian.ldivs;
                    store(R, out, k, n);
                }
            };
        });
    } catch (...) { return -1; }
    return 0;
}

int32_t orc_polyhedral_track_batch(void* Htoric, void* Hcoeff, const orc_options* o, int64_t N, const double* starts,
                                   const int32_t* cell_index, const double* cell_weights, int32_t ncells, orc_results* out,
                                   int32_t nthreads) {
    const HomotopyDef* DT = (const HomotopyDef*)Htoric;
    const HomotopyDef* DC = (const HomotopyDef*)Hcoeff;
    TrackerOptions T; EndgameOptions E; WeightedNormOptions W;
    unpack_options(o, T, E, W);
    const int n = DT->n();
    const int P = DT->F->eval.P;
    (void)ncells;
    try {
        parallel_for(N, nthreads, [&](int) {
            auto pt = std::make_shared<PolyhedralTracker>();
            pt->setup(DT, DC, T, E, W);
            return [=](int64_t k) {
                std::vector<cplx> x = cvec(starts + 2 * k * n, n);
                PathResult R = pt->track(cell_weights + (size_t)cell_index[k] * P, x.data());
                store(R, out, k, n);
            };
        });
    } catch (...) { return -1; }
    return 0;
}

// ------------------------------------------------------------------ test hooks
// A fresh Homotopy per hook call (no caching by handle address: handles get recycled).
struct HookH {
    Homotopy H;
    explicit HookH(void* Hv) {
        H.init((const HomotopyDef*)Hv);
        std::lock_guard<std::mutex> lock(g_hook_mutex);
        auto it = g_hook_weights.find(Hv);
        if (it != g_hook_weights.end()) for (int i = 0; i < H.P; ++i) H.weights[i] = it->second[i];
    }
};
static void put(double* dst, const std::vector<cplx>& v) { for (size_t i = 0; i < v.size(); ++i) { dst[2 * i] = v[i].re; dst[2 * i + 1] = v[i].im; } }

int32_t orc_toric_set_weights(void* Hv, const double* w) {
    const HomotopyDef* D = (const HomotopyDef*)Hv;
    std::lock_guard<std::mutex> lock(g_hook_mutex);
    g_hook_weights[Hv].assign(w, w + D->F->eval.P);
    return 0;
}
int32_t orc_evaluate(void* Hv, const double* x, const double* t, double* u) {
    HookH hh(Hv); Homotopy& H = hh.H;
    std::vector<cplx> xv = cvec(x, H.n), uv(H.m);
    H.evaluate(uv.data(), xv.data(), cplx(t[0], t[1]));
    put(u, uv); return 0;
}
int32_t orc_evaluate_dd(void* Hv, const double* x_hi, const double* x_lo, const double* t, double* u) {
    HookH hh(Hv); Homotopy& H = hh.H;
    std::vector<cdd> xv(H.n); std::vector<cplx> uv(H.m);
    for (int i = 0; i < H.n; ++i) xv[i] = cdd(dd(x_hi[2 * i], x_lo[2 * i]), dd(x_hi[2 * i + 1], x_lo[2 * i + 1]));
    H.evaluate_dd(uv.data(), xv.data(), cplx(t[0], t[1]));
    put(u, uv); return 0;
}
int32_t orc_evaluate_and_jacobian(void* Hv, const double* x, const double* t, double* u, double* U) {
    HookH hh(Hv); Homotopy& H = hh.H;
    std::vector<cplx> xv = cvec(x, H.n), uv(H.m), Uv((size_t)H.m * H.n);
    H.evaluate_and_jacobian(uv.data(), Uv.data(), xv.data(), cplx(t[0], t[1]));
    put(u, uv); put(U, Uv); return 0;
}
int32_t orc_taylor(void* Hv, int32_t K, const double* tx, const double* t, double* u) {
    HookH hh(Hv); Homotopy& H = hh.H;
    std::vector<cplx> xv = cvec(tx, K * H.n), uv(H.m);
    H.taylor(K, uv.data(), xv.data(), cplx(t[0], t[1]));
    put(u, uv); return 0;
}

void orc_la_solve(int32_t n, const double* A, const double* b, const double* weights, int32_t refine, double* x) {
    MatrixWorkspace W; W.resize(n);
    W.A = cvec(A, n * n); W.updated();
    std::vector<cplx> bv = cvec(b, n), xv(n);
    WeightedNorm wn; wn.resize(n);
    if (weights) for (int i = 0; i < n; ++i) wn.w[i] = weights[i];
    jac_ldiv(xv.data(), W, bv.data(), weights ? &wn : nullptr);
    if (refine == 1) fixed_precision_iterative_refinement(xv.data(), W, bv.data(), NormRef{nullptr, n});
    if (refine == 2) iterative_refinement(xv.data(), W, bv.data(), NormRef{nullptr, n}, 3, 1e-30);
    put(x, xv);
}
double orc_la_cond(int32_t n, const double* A, const double* d_l, const double* d_r) {
    MatrixWorkspace W; W.resize(n);
    W.A = cvec(A, n * n); W.updated();
    return ws_cond(W, d_l, d_r);
}
double orc_la_inverse_inf_norm_est(int32_t n, const double* A, const double* d_l, const double* d_r) {
    MatrixWorkspace W; W.resize(n);
    W.A = cvec(A, n * n); W.updated();
    return inverse_inf_norm_est(W, d_l, d_r);
}
void orc_dd_op(int32_t op, const double* a, const double* b, int32_t p, double* out) {
    dd x(a[0], a[1]), y(b ? b[0] : 0.0, b ? b[1] : 0.0), r;
    switch (op) {
        case 0: r = x + y; break;
        case 1: r = x - y; break;
        case 2: r = x * y; break;
        case 3: r = x / y; break;
        case 4: r = square(x); break;
        case 5: {  // DoubleDouble.jl:382-409 power_by_squaring
            if (p == 0) { r = dd(1.0); break; }
            int P = p < 0 ? -p : p;
            dd z = x;
            if (P == 1) r = z; else if (P == 2) r = square(z);
            else {
                int t = __builtin_ctz((unsigned)P) + 1; P >>= t;
                while (--t > 0) z = square(z);
                dd yv = z;
                while (P > 0) { t = __builtin_ctz((unsigned)P) + 1; P >>= t; while (--t >= 0) z = square(z); yv = yv * z; }
                r = yv;
            }
            if (p < 0) r = dd(1.0) / r;
        } break;
        default: r = dd(NaN);
    }
    out[0] = r.hi; out[1] = r.lo;
}
void orc_stepper_trace(const double* start, const double* target, const double* ds, int32_t k, double* out) {
    // out: per step 6 doubles: t.re t.im dt.re dt.im  is_done  dist_to_target, after propose(ds[i]) + step_success
    SegmentStepper S; S.init(cplx(start[0], start[1]), cplx(target[0], target[1]));
    for (int i = 0; i < k; ++i) {
        S.propose_step(ds[i]);
        cplx dt = S.dt(), tp = S.t_prop();
        S.step_success();
        cplx t = S.t();
        double* o = out + 8 * i;
        o[0] = t.re; o[1] = t.im; o[2] = dt.re; o[3] = dt.im; o[4] = S.is_done(); o[5] = S.dist_to_target(); o[6] = tp.re; o[7] = tp.im;
    }
}
void orc_norm_test(int32_t n, const double* x, const double* y, const double* w, double* out) {
    std::vector<cplx> xv = cvec(x, n), yv = cvec(y, n);
    WeightedNorm wn; wn.resize(n);
    for (int i = 0; i < n; ++i) wn.w[i] = w[i];
    out[0] = inf_norm(xv.data(), n); out[1] = inf_distance(xv.data(), yv.data(), n);
    out[2] = wn(xv.data()); out[3] = wn.distance(xv.data(), yv.data());
    wn.init(xv.data());
    for (int i = 0; i < n; ++i) out[4 + i] = wn.w[i];
    wn.update(yv.data());
    for (int i = 0; i < n; ++i) out[4 + n + i] = wn.w[i];
}
double orc_nthroot(double x, int32_t n) { return nthroot(x, n); }

void orc_taylor_op(int32_t op, int32_t K, const double* a, const double* b, const double* c, const double* d, int32_t p, double* out) {
    // a,b,c,d: 5 complex coefficients each; runs a 1-instruction tape
    Program P; P.n = 4; P.var_off = 0; P.out_dim = 1; P.tape_space = 5;
    Instr I; I.in[0] = 1; I.in[1] = (op == OP_POW_INT) ? p : 2; I.in[2] = 3; I.in[3] = 4; I.op = op; I.out = 5;
    Instr S; S.in[0] = S.in[1] = S.in[2] = S.in[3] = 5; S.op = OP_STOP; S.out = 5;
    P.instr = {I, S}; P.u_assign = {{1, 5}}; finalize(P);
    TaylorInterp T; T.init(&P);
    std::vector<cplx> x((size_t)5 * 4);
    const double* src[4] = {a, b, c, d};
    for (int v = 0; v < 4; ++v)
        for (int r = 0; r < 5; ++r) x[(size_t)r * 4 + v] = src[v] ? cplx(src[v][2 * r], src[v][2 * r + 1]) : cplx();
    std::vector<cplx> o((size_t)(K + 1));
    T.execute(K, o.data(), x.data(), 5, nullptr, nullptr, 0);
    for (int k = 0; k <= K; ++k) { out[2 * k] = o[k].re; out[2 * k + 1] = o[k].im; }
}
void orc_valuation_trace(int32_t n, int32_t steps, const double* tx, const double* t, double* out) {
    // tx: steps x 4 x n complex (x, x1, x2, x3 Taylor coefficients); out: steps x 4 x n (val_x, val_tx, dval_x, dval_tx)
    Valuation V; V.resize(n); V.init();
    Predictor P; P.resize(n, n);
    for (int s = 0; s < steps; ++s) {
        for (int i = 0; i < 4 * n; ++i) P.tx[i] = cplx(tx[2 * ((size_t)s * 4 * n + i)], tx[2 * ((size_t)s * 4 * n + i) + 1]);
        valuation_update(V, P, t[s]);
        double* o = out + (size_t)s * 4 * n;
        for (int i = 0; i < n; ++i) { o[i] = V.val_x[i]; o[n + i] = V.val_tx[i]; o[2 * n + i] = V.dval_x[i]; o[3 * n + i] = V.dval_tx[i]; }
    }
}

}  // extern "C"
