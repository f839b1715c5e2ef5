// One solution path per lane group (hc_coop.h): shared-memory views, homotopy evaluation, small
// dense linear algebra, norms, predictor, Newton corrector and the core tracker step.  Scalars are
// replicated in every lane of the group, vectors live in the path's shared-memory slab and are
// processed lane-strided; every routine is entered and left with the slab group-synchronised.
//
// Replaces (reference file:line):
//   src/homotopies/straight_line_homotopy.jl:81-154, parameter_homotopy.jl:66-101,
//   coefficient_homotopy.jl:89-141, toric_homotopy.jl:114-278         homotopy operator API
//   src/linear_algebra.jl:88-98, 130-184, 268-354, 389-408, 432-567, 585-774, 833-885
//   src/norm.jl:101-136, 150-234;  src/utils.jl:300-394 (SegmentStepper)
//   src/predictor.jl:158-371;  src/newton_corrector.jl:55-286
//   src/tracker.jl:509-619, 639-844, 851-926
#pragma once
#include "hc_tape.h"
#ifndef HC_JIT_PREFETCH
#define HC_JIT_PREFETCH 0
#endif

namespace hc {

// Lane-strided loops.  The trip count is the same for every lane of a group (the loop branch is warp-uniform) and
// the body is guarded instead: a lane group whose lanes leave `for (i = lane; i < N; i += G)` at different trips
// pays a divergence + reconvergence (BSSY / BSYNC, a refetch) per loop -- the second-pass capture showed BSYNC at 10 %
// of the instruction-fetch stalls.  For G == 1 both forms are the plain loop.
#define HC_PAR(i, N) for (int i##_b = 0, i = g.lane; i##_b < (N); i##_b += G, i += G) if (G == 1 || i < (N))
// load-only loops (reductions): unrolled so that the loads of four trips are in flight together
#define HC_PARU(i, N) _Pragma("unroll 4") for (int i##_b = 0, i = g.lane; i##_b < (N); i##_b += G, i += G) if (G == 1 || i < (N))

// CTA-wide barriers of the lockstep (SYNC) instantiations of the tracker step: every thread of the CTA passes each of
// them once per round whether or not its lane has work, so that the warps walk the same code together (a 32 KB
// instruction cache per SM cannot feed warps that stream different straight-line code).
#if defined(__CUDA_ARCH__)
HC_D bool cta_any(bool p) { return __syncthreads_or(p ? 1 : 0) != 0; }
HC_D void cta_sync() { __syncthreads(); }
#else
HC_HD bool cta_any(bool p) { return p; }
HC_HD void cta_sync() {}
#endif

enum HKind : int { H_STRAIGHT_LINE = 0, H_PARAMETER = 1, H_COEFFICIENT = 2, H_TORIC = 3 };

struct DevOptions {  // same field order as hc_options (include/hc_b200.h)
    int max_steps; double max_step_size, max_initial_step_size; int extended_precision;
    double min_step_size, min_rel_step_size;
    double a, beta_a, beta_omega_p, beta_tau, strict_beta_tau; int min_newton_iters;
    double endgame_start; int max_endgame_steps, max_endgame_extended_steps;
    double min_cond, min_cond_growth, min_coord_growth;
    int zero_is_at_infinity, at_infinity_check, only_nonsingular;
    double singular_min_accuracy; int max_winding_number;
    double val_finite_tol, val_at_infinity_tol, sing_cond, sing_accuracy, scaling_threshold;
    int refine_steps;
    double scale_min, scale_abs_min, scale_max;
};

struct DevHomotopy {
    int kind, n, P;            // square systems: m == n; P = #parameters of F
    DevProgram Fe, Fj, Ge, Gj; // eval / Jacobian programs of F (and of G for straight-line)
    cx gamma;
    const cx* G_params;        // straight-line: fixed parameters of G (scaling), length Ge.P
    const cx* F_params;        // straight-line: fixed parameters of F
    const cx* p;               // parameter/coefficient: start (t = 1); toric: system coefficients
    const cx* q;               // parameter/coefficient: target (t = 0)
    const cx* path_p;          // optional per-path start parameters, [path * P + i] (the caller's layout, no host transposition)
    const cx* path_q;          // optional per-path target parameters, [path * P + i]
    long long N;
    int tape_cx;               // per-path tape region, in cx units
};

// ---------------------------------------------------------------- per-path memory
// where the LU factors live: with the rest of the lane state, or (specialised thread-per-path kernels built with
// HC_JIT_LU_SMEM) thread-interleaved in shared memory
template <int S> struct LUView { using type = SV<cx, S>; };
#if defined(HC_JIT_LU_SMEM) && HC_JIT_LU_SMEM
template <> struct LUView<2> { using type = SV<cx, 3>; };
#endif
template <int S>
struct PathMem {
    using CV = SV<cx, S>; using RV = SV<double, S>; using IV = SV<int, S>;
    using LV = typename LUView<S>::type;
    CV x, xhat, xbar, tx, ptx1, ty1, pty1, xtemp, u, dx, r, A, wr, wdx, work;
    LV LU;
    CV sol, lastp, pred, ppred, samp, tape;
    CV pv;   // (unused)
    DV<S> rbd;
    RV w, rs, rwork, egrs, egcs, ais, ait, aia, aic, val, tw;
    RV ptw;  // (unused)
    CV wt;   // specialised kernels, toric stage: the pairs (w_i, t^w_i) at the cached t -- one 16-byte load gives both
    IV ipiv, perm;
};

// Group-per-path layout.  Carves the path's vectors out of two 16-byte aligned slabs: the hot one
// (everything a regular predictor-corrector step touches) lives in shared memory, the cold one
// (endgame samples, valuation history, at-infinity bookkeeping, Hermite-mode predictor data, DD
// accumulators) in a per-group scratch area in global memory.  With null bases it only counts.
struct SlabSizes { size_t hot, cold; };
// flags: bit 0 = specialised kernel (no fp64 / Taylor tape), bit 1 = the LU factors are not part of the slab (shared memory)
template <int S>
HC_HD SlabSizes carve(PathMem<S>& M, int n, int P, int tape_cx, unsigned char* hot, unsigned char* cold, int flags = 0) {
    const bool jit = (flags & 1) != 0, lu_out = (flags & 2) != 0;
    static_assert(S == 0 || S == 2, "flat layouts only");
    size_t off = 0;
    unsigned char* base = hot;
    auto C = [&](size_t k) { SV<cx, S> v = SV<cx, S>::make(base ? base + off : nullptr); off += 16 * k; return v; };
    auto R = [&](size_t k) { SV<double, S> v = SV<double, S>::make(base ? base + off : nullptr); off += 8 * k; return v; };
    M.x = C(n); M.xhat = C(n); M.xbar = C(n); M.tx = C(4 * n);
    M.xtemp = C(n); M.u = C(n); M.dx = C(n); M.r = C(n); M.A = C((size_t)n * n);
    if (!lu_out) { SV<cx, S> lu = C((size_t)n * n); M.LU.p = (decltype(M.LU.p))lu.p; }
    M.wr = C(n); M.wdx = C(n); M.work = C(n);
    // specialised kernels keep no fp64 / Taylor tape (slots are registers of the generated code): the parameter
    // values take its place in the hot slab, the tape of the DoubleDouble interpreter moves to the cold part
    if (!jit) M.tape = C((size_t)tape_cx);
    M.pv = M.tape;
    if (jit) M.wt = C(P > 0 ? P : 1); else M.wt = M.tape;
    M.w = R(n); M.rs = R(n); M.rwork = R(n); M.tw = R(P > 0 ? P : 1);
    M.ptw = M.tw;
    M.ipiv = SV<int, S>::make(base ? base + off : nullptr); off += 4 * (size_t)n;
    M.perm = SV<int, S>::make(base ? base + off : nullptr); off += 4 * (size_t)n;
    SlabSizes s;
    s.hot = (off + 15) & ~(size_t)15;
    off = 0; base = cold;
    M.ptx1 = C(2 * n); M.ty1 = C(2 * n); M.pty1 = C(2 * n);
    M.sol = C(n); M.lastp = C(n); M.pred = C(n); M.ppred = C(n); M.samp = C(6 * n);
    M.rbd.v = C(2 * n);
    if (jit) M.tape = C((size_t)tape_cx);
    M.egrs = R(n); M.egcs = R(n); M.ais = R(n); M.ait = R(n); M.aia = R(n); M.aic = R(n); M.val = R(12 * n);
    s.cold = (off + 15) & ~(size_t)15;
    return s;
}

HC_HD double nmax(double d, double v) { return (d != d || v != v) ? HC_NAN : (v > d ? v : d); }

// ---------------------------------------------------------------- the path
enum TrackerCode : int {  // src/tracker.jl:166-176
    TC_tracking = 0, TC_success, TC_terminated_max_steps, TC_terminated_accuracy_limit,
    TC_terminated_ill_conditioned, TC_terminated_invalid_startvalue,
    TC_terminated_invalid_startvalue_singular_jacobian, TC_terminated_step_size_too_small,
    TC_terminated_unknown
};
enum NewtonCode : int { NEWT_CONVERGED = 0, NEWT_TERMINATED, NEWT_MAX_ITERS, NEWT_SINGULARITY };
struct NewtonResult { int code; double accuracy; int iters; double omega, theta, mu_low, norm_dx0; };

HC_HD double hfun(double a) { return 2 * a * (sqrt(4 * a * a + 1) - 2 * a); }  // tracker.jl:517

static HC_HDN cx t_to_s_plane(cx t, int m) {  // predictor.jl:338-351 (cold: winding > 1 only)
    double r = cabs(t);
    if (t.im == 0.0 && t.re > 0) return mk(nthroot(r, m));
    double th = atan2(t.im, t.re);
    th = fmod(th, 6.283185307179586); if (th < 0) th += 6.283185307179586;
    double rr = nthroot(r, m);
    return mk(rr * cos(th / m), rr * sin(th / m));
}

template <int G, int S>
struct Path {
    using CV = SV<cx, S>; using RV = SV<double, S>; using IV = SV<int, S>; using LV = typename PathMem<S>::LV;
    Grp<G> g;
    // ---- immutable context
    const DevHomotopy* H;
    const DevOptions* O;
    PathMem<S> M;
#if defined(HC_JIT_N)
    static constexpr int n = HC_JIT_N;  // specialised kernel: every loop over the variables has a constant trip count
#else
    int n;
#endif
    long long pidx;
    long long prow;  // row of the per-path parameter arrays this path reads (pidx, or pidx / param_div for sweeps)
    int kind;  // current homotopy kind (polyhedral paths switch TORIC -> COEFFICIENT)
    // ---- tracker state (src/tracker.jl:307-338)
    cx st_start, st_target; double st_absd; bool st_forward; double st_s, st_sp;  // SegmentStepper
    double ds_prev, accuracy, omega, omega_prev, mu, tau;
    bool extended_prec, used_extended_prec, refined_extended_prec, keep_extended_prec, use_strict_beta_tau;
    bool factorized, scaled;  // MatrixWorkspace flags
    unsigned long long perm_bits, perm_bits2;  // row permutation of the register-blocked LU, 5 bits per row, rows 0-11 / 12-23
    bool a_in_lu, rs_raw;     // specialised kernels: the Jacobian sits in the LU buffer (factorize in place); M.rs holds raw Skeel row sums
    int code, accepted_steps, rejected_steps, last_steps_failed, ext_accepted_steps, ext_rejected_steps;
    const DevProgram* tape_prog; int tape_kind; cx tape_t;  // whose inputs (constants, parameters at t) the fp64 tape holds
    const DevProgram* tay_prog; int tay_kind; cx tay_t;     // ... and the series tape
    int pv_kind; cx pv_t;  // specialised kernels: homotopy kind and t the cached parameter values M.pv belong to (-1 = none)
    int ps_kind; cx ps_t;  // ... and the cached parameter series in M.tape
    const cx *jp, *jq;     // ... this path's start / target parameters
    double tol_acc_limit;  // accuracy-limit threshold of check_terminated (options only; cached per path: pow is 300 instructions)
    double min_step_size, min_rel_step_size;  // mutable copies (polyhedral.jl:474-488, endgame_tracker.jl:270)
    // ---- predictor (src/predictor.jl:72-103)
    int pm_hermite; double trust_region, local_error, cond_H;
    cx pt, pprev_t, ps, pprev_s; int winding;
    // ---- counters (src/linear_algebra.jl:809-826 + flop accounting of SURVEY.md 8(d))
    int n_fact, n_ldiv, n_evaljac, n_eval, n_evaldd, n_tay1, n_tay2, n_tay3;

    // Specialised kernels know n at compile time, which unrolls every loop over the variables.  That is what the
    // regular predictor-corrector step wants; the once-per-path and endgame-only code shadows it with the run-time
    // value so that its loops stay rolled (instruction footprint: the step should fit the instruction cache).
    HC_HD int nrt() const { return H->n; }
#define HC_COLD_N const int n = this->nrt(); (void)n;

    // ================================================================ norms (src/norm.jl)
    HC_HDN double inf_norm(CV x) {
        double d = -1.0;
        HC_PARU(i, n) d = nmax(d, abs2(x[i]));
        double r = sqrt(g.rmax(d));
        if (r == HC_INF) { d = 0; HC_PAR(i, n) { cx z = x[i]; d = nmax(d, hypot(z.re, z.im)); } r = g.rmax(d); }
        return r;
    }
    HC_HDN double inf_dist(CV x, CV y) {
        double d = -1.0;
        HC_PARU(i, n) d = nmax(d, abs2(x[i] - y[i]));
        double r = sqrt(g.rmax(d));
        if (r == HC_INF) { d = 0; HC_PAR(i, n) { cx z = x[i] - y[i]; d = nmax(d, hypot(z.re, z.im)); } r = g.rmax(d); }
        return r;
    }
    HC_HDN double wnorm(CV x) {
        double d = -1.0;
        HC_PARU(i, n) d = nmax(d, abs2(x[i] / M.w[i]));
        double r = sqrt(g.rmax(d));
        if (r == HC_INF) { d = 0; HC_PAR(i, n) { cx z = x[i] / M.w[i]; d = nmax(d, hypot(z.re, z.im)); } r = g.rmax(d); }
        return r;
    }
    HC_HDN double wdist(CV x, CV y) {
        double d = -1.0;
        HC_PARU(i, n) d = nmax(d, abs2((x[i] - y[i]) / M.w[i]));
        double r = sqrt(g.rmax(d));
        if (r == HC_INF) { d = 0; HC_PAR(i, n) { cx z = (x[i] - y[i]) / M.w[i]; d = nmax(d, hypot(z.re, z.im)); } r = g.rmax(d); }
        return r;
    }
    // Elementwise helpers.  With one thread per path (G == 1) the vectors sit in global memory, so the
    // loops are unrolled by four with all loads issued before the first store (memory-level parallelism).
    HC_HDN void vcopy(CV dst, CV src, int len) {
        if (G == 1) {
            int i = 0;
            for (; i + 3 < len; i += 4) { cx a0 = src[i], a1 = src[i + 1], a2 = src[i + 2], a3 = src[i + 3]; dst[i] = a0; dst[i + 1] = a1; dst[i + 2] = a2; dst[i + 3] = a3; }
            for (; i < len; ++i) dst[i] = src[i];
        } else { HC_PAR(i, len) dst[i] = src[i]; g.sync(); }
    }
    // y[i] -= a[i] * s for i in [lo, hi)
    template <class YV, class AV>
    HC_HD void col_fnma(YV y, AV a, cx s, int lo, int hi) {
        if (G == 1) {
            int i = lo;
            for (; i + 3 < hi; i += 4) {
                cx a0 = a[i], a1 = a[i + 1], a2 = a[i + 2], a3 = a[i + 3], y0 = y[i], y1 = y[i + 1], y2 = y[i + 2], y3 = y[i + 3];
                y[i] = cfnma(a0, s, y0); y[i + 1] = cfnma(a1, s, y1); y[i + 2] = cfnma(a2, s, y2); y[i + 3] = cfnma(a3, s, y3);
            }
            for (; i < hi; ++i) y[i] = cfnma(a[i], s, y[i]);
        } else for (int ib = lo, i = lo + g.lane; ib < hi; ib += G, i += G) { if (i < hi) y[i] = cfnma(a[i], s, y[i]); }
    }
    // outputs of a program run: MODE 0 dst = tape, 1 dst = s * tape, 2 dst += s * tape
    template <int MODE>
    HC_HDN void extract(CV dst, CV tape, const int2* asg, int cnt, cx s) {
        if (G == 1) {
            int k = 0;
            for (; k + 3 < cnt; k += 4) {
                int2 a0 = pld<S>(asg + k), a1 = pld<S>(asg + k + 1), a2 = pld<S>(asg + k + 2), a3 = pld<S>(asg + k + 3);
                cx v0 = tape[a0.y], v1 = tape[a1.y], v2 = tape[a2.y], v3 = tape[a3.y];
                if (MODE == 0) { dst[a0.x] = v0; dst[a1.x] = v1; dst[a2.x] = v2; dst[a3.x] = v3; }
                else if (MODE == 1) { dst[a0.x] = s * v0; dst[a1.x] = s * v1; dst[a2.x] = s * v2; dst[a3.x] = s * v3; }
                else {
                    cx d0 = dst[a0.x], d1 = dst[a1.x], d2 = dst[a2.x], d3 = dst[a3.x];
                    dst[a0.x] = cfma(s, v0, d0); dst[a1.x] = cfma(s, v1, d1); dst[a2.x] = cfma(s, v2, d2); dst[a3.x] = cfma(s, v3, d3);
                }
            }
            for (; k < cnt; ++k) {
                int2 a = pld<S>(asg + k);
                if (MODE == 0) dst[a.x] = tape[a.y]; else if (MODE == 1) dst[a.x] = s * tape[a.y]; else dst[a.x] = cfma(s, tape[a.y], dst[a.x]);
            }
        } else HC_PAR(k, cnt) {
            int2 a = pld<S>(asg + k);
            if (MODE == 0) dst[a.x] = tape[a.y]; else if (MODE == 1) dst[a.x] = s * tape[a.y]; else dst[a.x] = cfma(s, tape[a.y], dst[a.x]);
        }
    }

    // ================================================================ homotopy
    HC_HD cx param_p(int i) const { return H->path_p ? H->path_p[(size_t)prow * H->P + i] : pld<S>(H->p + i); }
    HC_HD cx param_q(int i) const { return H->path_q ? H->path_q[(size_t)prow * H->P + i] : pld<S>(H->q + i); }

    // log t and 1 / t of a toric homotopy: the same for every parameter (t^w = exp(w log t)), so the loops over the
    // parameters compute them once instead of once per parameter
    struct ToricT { double lt, ti; };
    HC_HD ToricT toric_t(cx t) const {
        ToricT r; r.lt = 0.0; r.ti = 0.0;
        if (kind == H_TORIC && t.re != 0.0) { r.lt = log(t.re); r.ti = 1.0 / t.re; }
        return r;
    }
    // Taylor coefficients c[0..4] of parameter i at t
    HC_HD void param_series(int i, cx t, const ToricT& tt, cx* c) const {
        c[1] = c[2] = c[3] = c[4] = mk(0.0);
        if (kind == H_TORIC) {  // toric_homotopy.jl:145-177, 220-264 (real t >= 0)
            cx u = pld<S>(H->p + i);
            double w = M.tw[i], tr = t.re;
            if (tr == 0.0) {
                c[0] = mk(0.0);
                if (w < 1e-12) c[0] = u;
                else if (fabs(w - 1.0) <= 1.4901161193847656e-08 * fmax(fabs(w), 1.0)) c[1] = u;
            } else {
                double tw = exp(w * tt.lt), ti = tt.ti;
                c[0] = u * tw;
                double tw1 = w * tw * ti; c[1] = u * tw1;
                double tw2 = 0.5 * (w - 1) * tw1 * ti; c[2] = u * tw2;
                double tw3 = (w - 2) * tw2 * ti / 3; c[3] = u * tw3;
                double tw4 = 0.25 * (w - 3) * tw3 * ti; c[4] = u * tw4;
            }
        } else {  // parameter_homotopy.jl:66-87, coefficient_homotopy.jl:89-107
            cx p = param_p(i), q = param_q(i);
            if (t.im == 0.0) c[0] = t.re * p + (1.0 - t.re) * q;
            else c[0] = t * p + (mk(1.0) - t) * q;
            c[1] = p - q;
        }
    }
    // value of parameter i at t (toric t == 0: weights that are exactly 0 survive, toric_homotopy.jl:160-165)
    HC_HD cx param_value(int i, cx t, const ToricT& tt) const {
        if (kind == H_TORIC) {
            cx u = pld<S>(H->p + i);
            double w = M.tw[i];
            if (t.re == 0.0) return w == 0.0 ? u : mk(0.0);
            return u * exp(w * tt.lt);
        }
        cx p = param_p(i), q = param_q(i);
        if (t.im == 0.0) return t.re * p + (1.0 - t.re) * q;
        return t * p + (mk(1.0) - t) * q;
    }

    HC_HD static void store_in(CV tape, int s, cx hi, cx, cx*) { tape[s] = hi; }
    HC_HD static void store_in(CV tape, int s, cx hi, cx lo, cdd*) { tstore(tape, s, mkcdd(mkdd(hi.re, lo.re), mkdd(hi.im, lo.im))); }

    template <class T>
    HC_HDN void load_inputs(const DevProgram& P, CV x, const CV* xlo, cx t, const cx* fixed) {
        T* tag = nullptr;
        CV tape = M.tape;
        // Ops never write into the input block, so constants, parameters and t survive an fp64 run: the
        // Newton iterations of one step (same program, same t) only refresh the variables.  That skips
        // P parameter evaluations (t^w = exp(w log t) each for a toric homotopy) per iteration.
        const bool keep = sizeof(T) == sizeof(cx) && tape_prog == &P && tape_kind == kind && tape_t.re == t.re && tape_t.im == t.im;
        if (!keep) {
            HC_PAR(i, P.C) store_in(tape, i, pld<S>(P.consts + i), mk(0.0), tag);
            const ToricT tt = toric_t(t);
            HC_PAR(i, P.P) store_in(tape, P.param_off + i, fixed ? pld<S>(fixed + i) : param_value(i, t, tt), mk(0.0), tag);
            if (P.t_slot >= 0 && g.lane == 0) store_in(tape, P.t_slot, t, mk(0.0), tag);
        }
        HC_PAR(i, P.n) store_in(tape, P.var_off + i, x[i], xlo ? (*xlo)[i] : mk(0.0), tag);
        g.sync();
        tape_prog = sizeof(T) == sizeof(cx) ? &P : nullptr; tape_kind = kind; tape_t = t;
        tay_prog = nullptr; ps_kind = -1;
    }

    // thread-per-path engines run the segmented interpreters, lane groups the levelised ones
    HC_HD void run_f64(const DevProgram& P) {
        if (G == 1) run_tape_seg(P.fops, P.segs, P.n_segs, M.tape); else run_tape<cx, G>(P, M.tape, g);
    }
    template <int K> HC_HD void run_taylor(const DevProgram& P) {
        if (G == 1) run_taylor_tape_seg<K>(P.fops, P.segs, P.n_segs, M.tape); else run_taylor_tape<K, G>(P, M.tape, g);
    }
    // u (and optionally the column-major Jacobian U) of H(x, t)
#if defined(HC_JIT_GEN)
    // ---- per-system generated code (hc_jitgen.h): straight-line evaluate / evaluate_and_jacobian / taylor with the
    // tape slots in registers.  The homotopy parameters are computed where the code uses them -- p_i, q_i come from
    // shared memory (or the path's row of the per-path arrays), parameter / coefficient homotopies need nothing else
    // (t p + (1 - t) q); the toric stage keeps t^w_i per lane (M.ptw, one exp per parameter and t) and, for the three
    // Taylor passes of a predictor update, the real factors of the coefficients 1..3 (M.tape viewed as doubles, idle
    // then: the DoubleDouble interpreter owns it only during eval_dd).  Complex t never reaches these kernels
    // (hc_api.cu keeps such batches on the interpreter).
    struct JPar { double tr, omt, g1, ti; bool toric, at0; const cx *pg, *qg; unsigned ps, qs; };
    HC_HDN void jit_refresh_ptw(cx t) {  // M.wt[i] = (w_i, t^w_i) (toric_homotopy.jl:145-177; t == 0: weights that are exactly 0 survive)
        if (pv_kind == kind && pv_t.re == t.re && pv_t.im == t.im) return;
        const int P = H->P;
        const ToricT tt = toric_t(t);
        const bool at0 = t.re == 0.0;
        int i = 0;
        for (; i + 3 < P; i += 4) {  // the four weight loads are in flight together
            const double w0 = M.tw[i], w1 = M.tw[i + 1], w2 = M.tw[i + 2], w3 = M.tw[i + 3];
            const double e0 = at0 ? (w0 == 0.0 ? 1.0 : 0.0) : exp(w0 * tt.lt), e1 = at0 ? (w1 == 0.0 ? 1.0 : 0.0) : exp(w1 * tt.lt);
            const double e2 = at0 ? (w2 == 0.0 ? 1.0 : 0.0) : exp(w2 * tt.lt), e3 = at0 ? (w3 == 0.0 ? 1.0 : 0.0) : exp(w3 * tt.lt);
            M.wt[i] = mk(w0, e0); M.wt[i + 1] = mk(w1, e1); M.wt[i + 2] = mk(w2, e2); M.wt[i + 3] = mk(w3, e3);
        }
        for (; i < P; ++i) { const double w = M.tw[i]; M.wt[i] = mk(w, at0 ? (w == 0.0 ? 1.0 : 0.0) : exp(w * tt.lt)); }
        pv_kind = kind; pv_t = t;
    }
    HC_HD JPar jit_par_ctx(cx t) {
        JPar c; c.toric = kind == H_TORIC; c.at0 = false; c.ti = 0.0;
        // where p and q are: generic pointers (per-path rows, host build) and, on the device, the shared-window address of
        // the staged arrays -- resolved once per function, not once per parameter
        c.pg = jp; c.qg = jq; c.ps = c.qs = 0u;
#if defined(__CUDA_ARCH__)
        if (S == 2) { c.ps = (unsigned)__cvta_generic_to_shared(H->p); c.qs = H->q ? (unsigned)__cvta_generic_to_shared(H->q) : 0u; }
#endif
        c.tr = t.re; c.omt = c.toric ? 0.0 : 1.0 - t.re; c.g1 = c.toric ? 0.0 : 1.0;
        if (c.toric) jit_refresh_ptw(t);
        return c;
    }
    HC_HD JPar jit_pser_ctx(cx t) {
        JPar c = jit_par_ctx(t);
        c.at0 = t.re == 0.0;
        c.ti = c.at0 ? 0.0 : 1.0 / t.re;
        return c;
    }
    // real factors of the Taylor coefficients 0..3 of a toric parameter: c_k = u_i F_k (same formulas as param_series)
    // (w_i, t^w_i) of a toric parameter; a separate statement of the generated code, issued well ahead of its use
    // (hc_jitgen.h hoists it), because this is a per-lane array that the L1 does not keep between uses
    template <int MODE> HC_HD cx jit_ldw(int i, const JPar& c) const { return (MODE == 4 || c.toric) ? (cx)M.wt[i] : mk(0.0); }
    HC_HD void jit_tfac(const cx wt, const JPar& c, double& f0, double& f1, double& f2, double& f3) const {
        const double w = wt.re, tw = wt.im;
        if (c.at0) {
            f0 = w < 1e-12 ? 1.0 : 0.0;
            f1 = (!(w < 1e-12) && fabs(w - 1.0) <= 1.4901161193847656e-08 * fmax(fabs(w), 1.0)) ? 1.0 : 0.0;
            f2 = f3 = 0.0;
        } else {
            f0 = tw;
            f1 = w * tw * c.ti;
            f2 = 0.5 * (w - 1) * f1 * c.ti;
            f3 = (w - 2) * f2 * c.ti / 3;
        }
    }
    // this path's start / target parameter arrays behind generic pointers (per-path rows in global memory)
    HC_HD void jit_bind_params() {
        jp = H->path_p ? H->path_p + (size_t)prow * H->P : H->p;
        jq = H->path_q ? H->path_q + (size_t)prow * H->P : H->q;
    }
    // PP: the batch carries per-path parameter rows; MODE 2: parameter / coefficient homotopy, 3: polyhedral driver
    // (toric or coefficient stage, decided per lane at run time), 4: toric homotopy
    HC_HD static cx jit_lds(unsigned a) {
        cx v = mk(0.0);
#if defined(__CUDA_ARCH__)
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.re), "=d"(v.im) : "r"(a));
#endif
        return v;
    }
    template <int PP> HC_HD cx jit_ldp(int i, const JPar& c) const { return (PP || S != 2) ? c.pg[i] : jit_lds(c.ps + 16u * (unsigned)i); }
    template <int PP> HC_HD cx jit_ldq(int i, const JPar& c) const { return (PP || S != 2) ? c.qg[i] : jit_lds(c.qs + 16u * (unsigned)i); }
    template <int MODE, int PP> HC_HD cx jit_par(int i, const JPar& c, const cx wt) const {
        const cx p = jit_ldp<PP>(i, c);
        if (MODE == 4) return p * wt.im;
        const cx q = jit_ldq<PP>(i, c);
        const double f = (MODE == 3 && c.toric) ? wt.im : c.tr;
        return mk(f * p.re + c.omt * q.re, f * p.im + c.omt * q.im);
    }
    template <int PP> HC_HD void jit_pser_lin(int i, const JPar& c, cx& c0, cx& c1) const {
        const cx p = jit_ldp<PP>(i, c), q = jit_ldq<PP>(i, c);
        c0 = mk(c.tr * p.re + c.omt * q.re, c.tr * p.im + c.omt * q.im);
        c1 = p - q;
    }
    template <int MODE, int PP> HC_HD void jit_pser(int i, const JPar& c, const cx wt, cx& c0, cx& c1, cx& c2, cx& c3) const {
        const cx p = jit_ldp<PP>(i, c);
        if (MODE == 4) {
            double f0, f1, f2, f3;
            jit_tfac(wt, c, f0, f1, f2, f3);
            c0 = p * f0; c1 = p * f1; c2 = p * f2; c3 = p * f3;
            return;
        }
        const cx q = jit_ldq<PP>(i, c);
        double f0 = c.tr, f1 = 1.0, f2 = 0.0, f3 = 0.0;
        if (c.toric) jit_tfac(wt, c, f0, f1, f2, f3);
        c0 = mk(f0 * p.re + c.omt * q.re, f0 * p.im + c.omt * q.im);
        c1 = mk(f1 * p.re - c.g1 * q.re, f1 * p.im - c.g1 * q.im);
        c2 = p * f2; c3 = p * f3;
    }
#include HC_JIT_GEN
    HC_HDN void eval_f64(CV u, const CV* U, CV x, cx t) {
        if (U) n_evaljac++; else n_eval++;
        if (U) { jit_evaljac(u, *U, *U, x, t, false, false); a_in_lu = false; rs_raw = false; } else jit_eval(u, x, t);
        g.sync();
    }
    // Evaluation of a corrector trip: the Jacobian goes straight into the LU buffer, which is then factorized in place
    // (no A -> LU copy, no second pass over A for the Skeel row sums: the generated code adds them up while the
    // entries are in registers).  keepA: a second copy lands in M.A -- the convergence trip's Jacobian is what the
    // predictor's refinement and the extended-precision refinement multiply with later.
    HC_HDN void eval_trip(CV u, CV x, cx t, bool keepA, bool rowsum) {
        n_evaljac++;
#if HC_JIT_PREFETCH & 2
        if (kind == H_TORIC && pv_kind == kind && pv_t.re == t.re) { const int P = H->P; for (int i = 0; i < P; ++i) M.wt.prefetch(i); }
#endif
        jit_evaljac(u, M.LU, M.A, x, t, keepA, rowsum);
        a_in_lu = true; rs_raw = rowsum;
    }
#else
    HC_HDN void eval_f64(CV u, const CV* U, CV x, cx t) {
        const int nn = n;
        const bool jac = U != nullptr;
        if (jac) n_evaljac++; else n_eval++;
        CV tape = M.tape;
        if (kind == H_STRAIGHT_LINE) {  // straight_line_homotopy.jl:96-124: u = (gamma t) G + (1 - t) F
            const cx ts = H->gamma * t, tt = mk(1.0) - t;
            const DevProgram& PG = jac ? H->Gj : H->Ge;
            const DevProgram& PF = jac ? H->Fj : H->Fe;
            // F first (plain stores), then the few entries of the start system are accumulated
            if (PF.nu != nn) HC_PAR(i, nn) u[i] = mk(0.0);
            if (jac && PF.nU != nn * nn) HC_PAR(i, nn * nn) (*U)[i] = mk(0.0);
            load_inputs<cx>(PF, x, nullptr, t, H->F_params);
            run_f64(PF);
            extract<1>(u, tape, PF.u_assign, PF.nu, tt);
            if (jac) extract<1>(*U, tape, PF.U_assign, PF.nU, tt);
            g.sync();
            load_inputs<cx>(PG, x, nullptr, t, H->G_params);
            run_f64(PG);
            extract<2>(u, tape, PG.u_assign, PG.nu, ts);
            if (jac) extract<2>(*U, tape, PG.U_assign, PG.nU, ts);
            g.sync();
        } else {
            const DevProgram& PF = jac ? H->Fj : H->Fe;
            if (PF.nu != nn) HC_PAR(i, nn) u[i] = mk(0.0);
            if (jac && PF.nU != nn * nn) HC_PAR(i, nn * nn) (*U)[i] = mk(0.0);
            load_inputs<cx>(PF, x, nullptr, t, nullptr);
            run_f64(PF);
            extract<0>(u, tape, PF.u_assign, PF.nu, mk(0.0));
            if (jac) extract<0>(*U, tape, PF.U_assign, PF.nU, mk(0.0));
            g.sync();
        }
    }
#endif  // HC_JIT_GEN
    // DoubleDouble re-evaluation of the residual, rounded to fp64 on store
    // (newton_corrector.jl:104-105; straight_line_homotopy.jl:81-94: combine in DD)
    HC_HDN void eval_dd(CV u, CV x, const CV* xlo, cx t) {
        const int nn = n;
        n_evaldd++;
        CV tape = M.tape;
        cdd* tag = nullptr;
        if (kind == H_STRAIGHT_LINE) {
            const cdd ts = tocdd(H->gamma * t), tt = tocdd(mk(1.0) - t);
            HC_PAR(i, nn) M.rbd.set(i, tocdd(mk(0.0)));
            load_inputs<cdd>(H->Ge, x, xlo, t, H->G_params);
            run_tape<cdd, G>(H->Ge, tape, g);
            HC_PAR(k, H->Ge.nu) { int2 a = H->Ge.u_assign[k]; M.rbd.set(a.x, ts * tload(tape, a.y, tag)); }
            g.sync();
            load_inputs<cdd>(H->Fe, x, xlo, t, H->F_params);
            run_tape<cdd, G>(H->Fe, tape, g);
            HC_PAR(k, H->Fe.nu) { int2 a = H->Fe.u_assign[k]; M.rbd.set(a.x, M.rbd.get(a.x) + tt * tload(tape, a.y, tag)); }
            g.sync();
            HC_PAR(i, nn) u[i] = tocx(M.rbd.get(i));
            g.sync();
        } else {
            const DevProgram& PF = H->Fe;
            if (PF.nu != nn) HC_PAR(i, nn) u[i] = mk(0.0);
            load_inputs<cdd>(PF, x, xlo, t, nullptr);
            run_tape<cdd, G>(PF, tape, g);
            HC_PAR(k, PF.nu) { int2 a = PF.u_assign[k]; u[a.x] = tocx(tload(tape, a.y, tag)); }
            g.sync();
        }
    }

    template <int K>
    HC_HDN void taylor_inputs(const DevProgram& P, CV tx, cx t, const cx* fixed) {
        constexpr int TS = HC_TS(K);
        CV tape = M.tape;
        tape_prog = nullptr;  // the series layout overwrites the fp64 input block
        // constants and the parameter series at t are written with all TS coefficients by the first pass of a
        // predictor update and serve the later passes (same program, same t): ops never write the input block
        const bool keep = K <= 3 && tay_prog == &P && tay_kind == kind && tay_t.re == t.re && tay_t.im == t.im;
        if (!keep) {
            HC_PAR(i, P.C) {
                tape[i * TS] = pld<S>(P.consts + i);
#pragma unroll
                for (int k = 1; k < TS; ++k) tape[i * TS + k] = mk(0.0);
            }
            const ToricT tt = toric_t(t);
            HC_PAR(i, P.P) {
                cx c[5];
                if (fixed) { c[0] = pld<S>(fixed + i); c[1] = c[2] = c[3] = c[4] = mk(0.0); }
                else param_series(i, t, tt, c);
                const int b = (P.param_off + i) * TS;
#pragma unroll
                for (int k = 0; k < TS; ++k) tape[b + k] = c[k];
            }
            if (P.t_slot >= 0 && g.lane == 0) {
                const int b = P.t_slot * TS;
                tape[b] = t; tape[b + 1] = mk(1.0);
#pragma unroll
                for (int k = 2; k < TS; ++k) tape[b + k] = mk(0.0);
            }
        }
        HC_PAR(i, P.n) {  // x series: rows 0..K-1 of tx, coefficient K zero padded
            const int b = (P.var_off + i) * TS;
#pragma unroll
            for (int k = 0; k < K; ++k) tape[b + k] = tx[k * n + i];
            tape[b + K] = mk(0.0);
        }
        g.sync();
        tay_prog = &P; tay_kind = kind; tay_t = t;
    }
    // u = K-th Taylor coefficient of lambda -> H(x(lambda), t + lambda); tx rows x^0..x^{K-1}
    template <int K>
    HC_HDN void taylor(CV u, CV tx, cx t) {
        if (K == 1) n_tay1++; else if (K == 2) n_tay2++; else n_tay3++;
#if defined(HC_JIT_GEN)
        static_assert(K <= 3, "specialised kernels generate the orders the tracker uses");
        if (K == 1) jit_taylor1(u, tx, t); else if (K == 2) jit_taylor2(u, tx, t); else jit_taylor3(u, tx, t);
        g.sync();
#else
        CV tape = M.tape;
        HC_PAR(i, n) u[i] = mk(0.0);
        if (kind == H_STRAIGHT_LINE) {  // straight_line_homotopy.jl:130-154
            taylor_inputs<K>(H->Ge, tx, t, H->G_params);
            run_taylor<K>(H->Ge);
            HC_PAR(k, H->Ge.nu) {
                int2 a = pld<S>(H->Ge.u_assign + k);
                u[a.x] = H->gamma * (tape[a.y * HC_TS(K) + K - 1] + t * tape[a.y * HC_TS(K) + K]);
            }
            g.sync();
            taylor_inputs<K>(H->Fe, tx, t, H->F_params);
            run_taylor<K>(H->Fe);
            HC_PAR(k, H->Fe.nu) {
                int2 a = pld<S>(H->Fe.u_assign + k);
                u[a.x] = u[a.x] + ((mk(1.0) - t) * tape[a.y * HC_TS(K) + K] - tape[a.y * HC_TS(K) + K - 1]);
            }
            g.sync();
        } else {  // parameter / coefficient / toric: parameters are series in lambda
            taylor_inputs<K>(H->Fe, tx, t, nullptr);
            run_taylor<K>(H->Fe);
            HC_PAR(k, H->Fe.nu) { int2 a = pld<S>(H->Fe.u_assign + k); u[a.x] = tape[a.y * HC_TS(K) + K]; }
            g.sync();
        }
#endif
    }

    // ================================================================ linear algebra
    // updated!(J) linear_algebra.jl:88-98: the copy A -> LU buffer is deferred to the factorization
    // (lu_prepare), where it is fused with the row scaling
    HC_HD void updated() { factorized = false; scaled = false; }
    HC_HDN void lu_prepare(bool scale) {
        const int nn = n;
        if (!scale) { for (int i = 0; i < nn * nn; i += G) if (i + g.lane < nn * nn) M.LU[i + g.lane] = M.A[i + g.lane]; g.sync(); return; }
        if (G == 1) {
            for (int j = 0; j < nn; ++j) {
                int i = 0;
                for (; i + 3 < nn; i += 4) {
                    cx a0 = M.A[j * nn + i], a1 = M.A[j * nn + i + 1], a2 = M.A[j * nn + i + 2], a3 = M.A[j * nn + i + 3];
                    double r0 = M.rs[i], r1 = M.rs[i + 1], r2 = M.rs[i + 2], r3 = M.rs[i + 3];
                    M.LU[j * nn + i] = a0 * r0; M.LU[j * nn + i + 1] = a1 * r1; M.LU[j * nn + i + 2] = a2 * r2; M.LU[j * nn + i + 3] = a3 * r3;
                }
                for (; i < nn; ++i) M.LU[j * nn + i] = M.A[j * nn + i] * M.rs[i];
            }
        } else {
            for (int j = 0; j < nn; ++j) HC_PAR(i, nn) M.LU[j * nn + i] = M.A[j * nn + i] * M.rs[i];
            g.sync();
        }
    }
    HC_HDN void lu_factor() {  // :130-184  right-looking, pivot = max abs2, first index wins ties
        const int nn = n;
        LV A = M.LU;
        for (int k = 0; k < nn; ++k) {
            double amax = -1.0; int kp = k;
            for (int ib = k, i = k + g.lane; ib < nn; ib += G, i += G) { if (i < nn) { double v = abs2(A[k * nn + i]); if (v > amax) { amax = v; kp = i; } } }
            g.argmax(amax, kp);
            if (g.lane == 0) M.ipiv[k] = kp;
            if (amax > 0.0) {
                if (kp != k) { HC_PAR(j, nn) { cx tmp = A[j * nn + k]; A[j * nn + k] = A[j * nn + kp]; A[j * nn + kp] = tmp; } g.sync(); }
                cx pinv = cinv(A[k * nn + k]);
                for (int ib = k + 1, i = k + 1 + g.lane; ib < nn; ib += G, i += G) { if (i < nn) A[k * nn + i] = A[k * nn + i] * pinv; }
                g.sync();
            }
            const int m = nn - k - 1;
            if (G == 1) {
                for (int j = k + 1; j < nn; ++j) col_fnma(A.at(j * nn), A.at(k * nn), A[j * nn + k], k + 1, nn);
            } else if (m > 0) {
                // trailing update, (column, row) pairs strided over the lanes
                int jc = g.lane / m, ir = g.lane - jc * m;
                const int dj = G / m, di = G - dj * m;
                for (int p0 = 0; p0 < m * m; p0 += G) {  // uniform trip count, guarded body (see HC_PAR)
                    if (jc < m) {
                        const int j = k + 1 + jc, i = k + 1 + ir;
                        A[j * nn + i] = cfnma(A[k * nn + i], A[j * nn + k], A[j * nn + i]);
                    }
                    jc += dj; ir += di;
                    if (ir >= m) { ir -= m; jc += 1; }
                }
            }
            g.sync();
        }
        factorized = true;
        n_fact++;
    }
    HC_HDN void lu_solve(CV x) {  // :310-316 (in place)
        const int nn = n;
        LV A = M.LU;
        if (g.lane == 0) for (int i = 0; i < nn; ++i) { int p = M.ipiv[i]; if (p != i) { cx tmp = x[i]; x[i] = x[p]; x[p] = tmp; } }
        g.sync();
        for (int j = 0; j < nn - 1; ++j) {
            cx xj = x[j];
            col_fnma(x, A.at(j * nn), xj, j + 1, nn);
            g.sync();
        }
        for (int j = nn - 1; j >= 0; --j) {
            cx xj = cdiv(x[j], A[j * nn + j]);
            g.sync();
            if (g.lane == 0) x[j] = xj;
            col_fnma(x, A.at(j * nn), xj, 0, j);
            g.sync();
        }
    }
    // ---- thread-per-path (G == 1) variants for n <= HC_REG_LU_MAX with the active column / the
    // right-hand side in REGISTERS (N is a compile-time size, every loop fully unrolled).
    // Left-looking (jki) order: column j is loaded once (fused with the A -> LU copy and the Skeel
    // row scaling of lu_prepare), receives the updates of the finished columns k < j from memory
    // -- loads that no longer sit on the dependency chain, so they overlap -- and is stored once:
    // n^3/3 loads + n^2 stores instead of 2n^3/3 loads + n^3/3 stores.  Every element still sees
    // its updates in increasing k with the same cfnma, so LU, ipiv and the solutions are bit-identical
    // to lu_prepare + lu_factor / lu_solve above (linear_algebra.jl:130-184, 310-316).
    // Fully unrolled variant (N^3 / 3 code, no selects: every index is a compile-time constant).  Lockstep kernels stream
    // their code once for all warps of the CTA, so for small N the shorter instruction count wins over the footprint
    // (HC_JIT_LU_UNROLL_MAX_N); same arithmetic, same pivots, bit-identical factors.
    template <int N, class AV>
    HC_HDN void lu_factor_reg_unrolled(bool scale, AV A) {
        LV LU = M.LU;
        int rowof[N];
#pragma unroll
        for (int i = 0; i < N; ++i) rowof[i] = i;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            cx col[N];
#pragma unroll
            for (int i = 0; i < N; ++i) col[i] = A[j * N + rowof[i]];
            if (scale) {
#pragma unroll
                for (int i = 0; i < N; ++i) col[i] = col[i] * M.rs[rowof[i]];
            }
#pragma unroll
            for (int k = 0; k < j; ++k) {
#pragma unroll
                for (int i = k + 1; i < N; ++i) col[i] = cfnma(LU[k * N + i], col[k], col[i]);
            }
            double amax = -1.0; int kp = j;
#pragma unroll
            for (int i = j; i < N; ++i) { double v = abs2(col[i]); if (v > amax) { amax = v; kp = i; } }
            M.ipiv[j] = kp;
            if (amax > 0.0) {
                if (kp != j) {
                    const cx cj = col[j]; cx ck = cj;
                    const int rj = rowof[j]; int rk = rj;
#pragma unroll
                    for (int i = j + 1; i < N; ++i) if (i == kp) { ck = col[i]; col[i] = cj; rk = rowof[i]; rowof[i] = rj; }
                    col[j] = ck; rowof[j] = rk;
#pragma unroll
                    for (int k = 0; k < j; ++k) { cx t0 = LU[k * N + j], t1 = LU[k * N + kp]; LU[k * N + j] = t1; LU[k * N + kp] = t0; }
                }
                const cx pinv = cinv(col[j]);
#pragma unroll
                for (int i = j + 1; i < N; ++i) col[i] = col[i] * pinv;
            }
#pragma unroll
            for (int i = 0; i < N; ++i) LU[j * N + i] = col[i];
        }
#pragma unroll
        { unsigned long long pk = 0, pk2 = 0; for (int i = 0; i < N; ++i) { if (i < 12) pk |= (unsigned long long)rowof[i] << (5 * i); else pk2 |= (unsigned long long)rowof[i] << (5 * (i - 12)); } perm_bits = pk; if (N > 12) perm_bits2 = pk2; }
        factorized = true;
        n_fact++;
    }
    // A: where the matrix sits -- M.A, or the LU buffer itself (in place: column j is read in full, in original row order,
    // before its permuted, eliminated form is written back; the columns to its right are still untouched)
    template <int N, class AV>
    HC_HDN void lu_factor_reg(bool scale, AV A) {
        LV LU = M.LU;
        int rowof[N];  // original row that currently sits in position i
#pragma unroll
        for (int i = 0; i < N; ++i) rowof[i] = i;
        // The loop over the columns stays rolled (j is uniform over the warp, so the `k < j` / `i >= j` guards below are
        // branches or predicates, not divergence): the instruction footprint is O(N^2) instead of O(N^3) -- a fully
        // unrolled N = 9 instance was 46 KB of SASS, more than the instruction cache of an SM holds.
#pragma unroll 1
        for (int j = 0; j < N; ++j) {
            cx col[N];
#pragma unroll
            for (int i = 0; i < N; ++i) col[i] = A[j * N + rowof[i]];
            if (scale) {
#pragma unroll
                for (int i = 0; i < N; ++i) col[i] = col[i] * M.rs[rowof[i]];
            }
#pragma unroll
            for (int k = 0; k < N - 1; ++k) {
                if (k >= j) break;
#pragma unroll
                for (int i = k + 1; i < N; ++i) col[i] = cfnma(LU[k * N + i], col[k], col[i]);
            }
            double amax = -1.0; int kp = j;
#pragma unroll
            for (int i = 0; i < N; ++i) { const double v = abs2(col[i]); if (i >= j && v > amax) { amax = v; kp = i; } }
            M.ipiv[j] = kp;
            cx pinv = mk(1.0);
            if (amax > 0.0) {
                if (kp != j) {  // swap positions j and kp: the column in registers, the finished columns in memory
                    cx cj = col[0], ck = col[0];
                    int rj = rowof[0], rk = rowof[0];
#pragma unroll
                    for (int i = 0; i < N; ++i) { if (i == j) { cj = col[i]; rj = rowof[i]; } if (i == kp) { ck = col[i]; rk = rowof[i]; } }
#pragma unroll
                    for (int i = 0; i < N; ++i) { if (i == j) { col[i] = ck; rowof[i] = rk; } else if (i == kp) { col[i] = cj; rowof[i] = rj; } }
                    for (int k = 0; k < j; ++k) { cx t0 = LU[k * N + j], t1 = LU[k * N + kp]; LU[k * N + j] = t1; LU[k * N + kp] = t0; }
                }
                cx pj = col[0];
#pragma unroll
                for (int i = 0; i < N; ++i) if (i == j) pj = col[i];
                pinv = cinv(pj);
#pragma unroll
                for (int i = 0; i < N; ++i) if (i > j) col[i] = col[i] * pinv;
            }
#pragma unroll
            for (int i = 0; i < N; ++i) LU[j * N + i] = col[i];
        }
#pragma unroll
        { unsigned long long pk = 0, pk2 = 0; for (int i = 0; i < N; ++i) { if (i < 12) pk |= (unsigned long long)rowof[i] << (5 * i); else pk2 |= (unsigned long long)rowof[i] << (5 * (i - 12)); } perm_bits = pk; if (N > 12) perm_bits2 = pk2; }
        factorized = true;
        n_fact++;
    }
    // x = (LU)^-1 P (scaled ? rs .* b : b); x may alias b
    template <int N>
    HC_HDN void lu_solve_reg(CV x, CV b, bool scale) {
        LV A = M.LU;
        cx xr[N];
        const unsigned long long pk = perm_bits, pk2 = N > 12 ? perm_bits2 : 0ull;  // scalars instead of N dependent index loads: the b[r] loads go out at once
#pragma unroll
        for (int i = 0; i < N; ++i) { const int r = (int)(((i < 12 ? pk >> (5 * i) : pk2 >> (5 * (i - 12)))) & 31u); xr[i] = b[r]; if (scale) xr[i] = M.rs[r] * xr[i]; }
#pragma unroll
        for (int j = 0; j < N - 1; ++j) {
#pragma unroll
            for (int i = j + 1; i < N; ++i) xr[i] = cfnma(A[j * N + i], xr[j], xr[i]);
        }
#pragma unroll
        for (int j = N - 1; j >= 0; --j) {
            xr[j] = cdiv(xr[j], A[j * N + j]);
#pragma unroll
            for (int i = 0; i < j; ++i) xr[i] = cfnma(A[j * N + i], xr[j], xr[i]);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = xr[i];
    }
#if defined(HC_JIT_N)
#define HC_REG_LU_MAX 20   // specialised kernels instantiate exactly their own N
#else
#define HC_REG_LU_MAX 12   // interpreter kernels: one instance per N of the dispatch below
#endif
#if defined(HC_JIT_N)
#define HC_REG_LU_DISPATCH(CALL) CALL((((HC_JIT_N) >= 2 && (HC_JIT_N) <= HC_REG_LU_MAX) ? (HC_JIT_N) : 2));
#else
#define HC_REG_LU_DISPATCH(CALL)                                                                    \
    switch (n) {                                                                                    \
        case 2: CALL(2); break; case 3: CALL(3); break; case 4: CALL(4); break; case 5: CALL(5); break; \
        case 6: CALL(6); break; case 7: CALL(7); break; case 8: CALL(8); break; case 9: CALL(9); break; \
        case 10: CALL(10); break; case 11: CALL(11); break; default: CALL(12); break;               \
    }
#endif
    HC_HD bool use_reg_lu() const { return G == 1 && S != 1 && n >= 2 && n <= HC_REG_LU_MAX; }
    HC_HD void factorize(bool scale) {
        if (use_reg_lu()) {
#if defined(HC_JIT_N) && defined(HC_JIT_LU_UNROLL_MAX_N)
            if (HC_JIT_N <= HC_JIT_LU_UNROLL_MAX_N) {
                if (a_in_lu) lu_factor_reg_unrolled<(HC_JIT_N >= 2 && HC_JIT_N <= HC_REG_LU_MAX) ? HC_JIT_N : 2>(scale, M.LU);
                else lu_factor_reg_unrolled<(HC_JIT_N >= 2 && HC_JIT_N <= HC_REG_LU_MAX) ? HC_JIT_N : 2>(scale, M.A);
                return;
            }
#endif
            if (a_in_lu) {
#define HC_CALL_(NN) lu_factor_reg<NN>(scale, M.LU)
                HC_REG_LU_DISPATCH(HC_CALL_)
#undef HC_CALL_
            } else {
#define HC_CALL_(NN) lu_factor_reg<NN>(scale, M.A)
                HC_REG_LU_DISPATCH(HC_CALL_)
#undef HC_CALL_
            }
        } else { lu_prepare(scale); lu_factor(); }
    }
    HC_HDN void lu_solve_adj(CV x) {  // :318-354 (in place)
        HC_COLD_N
        const int nn = n;
        LV A = M.LU;
        for (int j = 0; j < nn; ++j) {  // U^H z = x: column j of U gives a dot product
            double zr = 0.0, zi = 0.0;
            HC_PAR(i, j) { cx p = conj(A[j * nn + i]) * x[i]; zr += p.re; zi += p.im; }
            zr = g.rsum(zr); zi = g.rsum(zi);
            cx z = cdiv(x[j] - mk(zr, zi), conj(A[j * nn + j]));
            g.sync();
            if (g.lane == 0) x[j] = z;
            g.sync();
        }
        for (int j = nn - 1; j >= 0; --j) {  // L^H
            double zr = 0.0, zi = 0.0;
            for (int ib = j + 1, i = j + 1 + g.lane; ib < nn; ib += G, i += G) { if (i < nn) { cx p = conj(A[j * nn + i]) * x[i]; zr += p.re; zi += p.im; } }
            zr = g.rsum(zr); zi = g.rsum(zi);
            cx z = x[j] - mk(zr, zi);
            g.sync();
            if (g.lane == 0) x[j] = z;
            g.sync();
        }
        if (g.lane == 0) for (int i = nn - 1; i >= 0; --i) { int p = M.ipiv[i]; if (p != i) { cx tmp = x[i]; x[i] = x[p]; x[p] = tmp; } }
        g.sync();
    }
    // skeel_row_scaling!(d, A, c; threshold)  :432-459
    HC_HDN void skeel(RV d, RV c, double threshold) {
        HC_COLD_N
        const int nn = n;
        double m = -HC_INF;
        HC_PAR(i, nn) {
            double di = 0.0;
#pragma unroll 4
            for (int j = 0; j < nn; ++j) di += cabs(M.A[j * nn + i]) * c[j];
            d[i] = di;
            m = nmax(m, di);
        }
        m = g.rmax(m);
        double s = threshold + m;
        HC_PAR(i, nn) {
            int e = 0; double di = d[i];
            if (di != 0.0 && di == di && di < HC_INF) frexp(di, &e);
            d[i] = (e < s) ? 1.0 : ldexp(1.0, -e);
        }
        g.sync();
    }
    // second half of skeel for row sums that the generated Jacobian code left in d
    HC_HDN void skeel_finish(RV d, double threshold) {
        const int nn = n;
        double m = -HC_INF;
        HC_PAR(i, nn) m = nmax(m, (double)d[i]);
        m = g.rmax(m);
        const double s = threshold + m;
        HC_PAR(i, nn) {
            int e = 0; double di = d[i];
            if (di != 0.0 && di == di && di < HC_INF) frexp(di, &e);
            d[i] = (e < s) ? 1.0 : ldexp(1.0, -e);
        }
        g.sync();
    }
    // ldiv!(x, J, b[, norm])  :389-408, 833-862;  x may alias b
    HC_HDN void ldiv(CV x, CV b, bool with_norm) {
        const int nn = n;
        n_ldiv++;
        if (with_norm && !factorized) {
            if (rs_raw) { skeel_finish(M.rs, -30.0); rs_raw = false; } else skeel(M.rs, M.w, -30.0);
            scaled = true;
        }
        if (nn == 1) { cx v = cdiv(b[0], a_in_lu ? (cx)M.LU[0] : (cx)M.A[0]); g.sync(); if (g.lane == 0) x[0] = v; g.sync(); return; }
#if HC_JIT_PREFETCH & 1
        if (use_reg_lu()) { for (int i = 0; i < nn * nn; ++i) M.LU.prefetch(i); }
#endif
        if (!factorized) factorize(scaled);
        if (use_reg_lu()) {
            const bool sc = scaled;
#define HC_CALL_(NN) lu_solve_reg<NN>(x, b, sc)
            HC_REG_LU_DISPATCH(HC_CALL_)
#undef HC_CALL_
            return;
        }
        if (scaled) { HC_PAR(i, nn) x[i] = M.rs[i] * b[i]; g.sync(); }
        else if (x.p != b.p) vcopy(x, b, nn);
        lu_solve(x);
    }
    // one sweep of fixed precision refinement (:553-567); returns |dx| / |x| in the given norm
    HC_HDN double refine_fixed(CV x, CV b, bool weighted) {
        const int nn = n;
        HC_PAR(i, nn) {
            cx acc = -b[i];
#pragma unroll 4
            for (int j = 0; j < nn; ++j) acc = cfma(M.A[j * nn + i], x[j], acc);
            M.wr[i] = acc;
        }
        g.sync();
        n_ldiv--;  // workspace-level ldiv! is not counted by Jacobian.ldivs
        ldiv(M.wdx, M.wr, false);
        HC_PAR(i, nn) x[i] = x[i] - M.wdx[i];
        g.sync();
        return weighted ? wnorm(M.wdx) / wnorm(x) : inf_norm(M.wdx) / inf_norm(x);
    }
    // one sweep of mixed precision refinement (:528-544): residual A x - b accumulated in DD
    HC_HDN double refine_mixed(CV x, CV b, bool weighted) {
        HC_COLD_N
        const int nn = n;
        HC_PAR(i, nn) {
            cdd acc = tocdd(-b[i]);
            for (int j = 0; j < nn; ++j) {
                cx a = M.A[j * nn + i], xj = x[j];  // x converted exactly to DD, A stays fp64: products are exact two_prods
                dd rr = two_prod(a.re, xj.re) - two_prod(a.im, xj.im);
                dd ri = two_prod(a.re, xj.im) + two_prod(a.im, xj.re);
                acc = mkcdd(acc.re + rr, acc.im + ri);
            }
            M.wr[i] = tocx(acc);
        }
        g.sync();
        n_ldiv--;
        ldiv(M.wdx, M.wr, false);
        HC_PAR(i, nn) x[i] = x[i] - M.wdx[i];
        g.sync();
        return weighted ? wnorm(M.wdx) / wnorm(x) : inf_norm(M.wdx) / inf_norm(x);
    }
    // iterative_refinement!(x, J, b, norm; max_iters, tol)  :864-885
    HC_HDN void iterative_refinement(CV x, CV b, bool weighted, int max_iters, double tol) {
        n_ldiv++;
        double d = refine_mixed(x, b, weighted);
        for (int i = 2; i <= max_iters; ++i) {
            n_ldiv++;
            double d2 = refine_mixed(x, b, weighted);
            if (d2 < tol) return;
            if (d2 > 0.5 * d) return;
            d = d2;
        }
    }
    // Hager/Higham estimator of |diag(d_r)^-1 A^-1 diag(d_l)^-1|_inf  :585-682
    HC_HDN double inverse_inf_norm_est(const RV* dl, const RV* dr) {
        HC_COLD_N
        const int nn = n;
        if (!factorized) factorize(false);
        CV y = M.work; RV x = M.rwork;
        HC_PAR(i, nn) { double v = 1.0 / nn; if (dr) v /= (*dr)[i]; x[i] = v; y[i] = mk(v); }
        g.sync();
        lu_solve_adj(y);
        double gamma = 0;
        HC_PAR(i, nn) {
            cx v = y[i]; if (dl) v = v / (*dl)[i]; if (scaled) v = v * M.rs[i];
            double a = cabs(v); gamma += a;
            v = v / a; if (dl) v = v / (*dl)[i]; if (scaled) v = v / M.rs[i];
            y[i] = v;
        }
        gamma = g.rsum(gamma);
        g.sync();
        lu_solve(y);
        HC_PAR(i, nn) { cx yi = y[i]; x[i] = dr ? yi.re / (*dr)[i] : yi.re; }
        g.sync();
        int k = 2;
        while (true) {
            double mx = -1.0; int j = 0;
            HC_PAR(i, nn) { double a = fabs(x[i]); if (a != a) a = -0.5; if (a > mx) { mx = a; j = i; } }
            g.argmax(mx, j);
            g.sync();
            HC_PAR(i, nn) { double v = (i == j) ? 1.0 : 0.0; if (dr) v /= (*dr)[i]; x[i] = v; y[i] = mk(v); }
            g.sync();
            lu_solve_adj(y);
            double gbar = gamma; gamma = 0;
            HC_PAR(i, nn) {
                cx v = y[i]; if (dl) v = v / (*dl)[i]; if (scaled) v = v * M.rs[i];
                y[i] = v; gamma += cabs(v);
            }
            gamma = g.rsum(gamma);
            g.sync();
            if (gamma <= gbar) { gamma = gbar; break; }
            HC_PAR(i, nn) {
                cx v = y[i]; v = v / cabs(v); if (dl) v = v / (*dl)[i]; if (scaled) v = v / M.rs[i];
                y[i] = v;
            }
            g.sync();
            lu_solve(y);
            double ninf = 0;
            HC_PAR(i, nn) { cx yi = y[i]; double v = dr ? yi.re / (*dr)[i] : yi.re; x[i] = v; double a = fabs(v); ninf = a > ninf ? a : ninf; }
            ninf = g.rmax(ninf);
            g.sync();
            k += 1;
            if (x[j] == ninf || k > 2) break;
        }
        return nanmin(gamma, HC_INF);
    }
    HC_HDN double a_inf_norm(const RV* dl, const RV* dr) {  // :684-707
        HC_COLD_N
        const int nn = n;
        double nrm = -HC_INF;
        HC_PAR(i, nn) {
            double ni = 0.0;
            for (int j = 0; j < nn; ++j) ni += dr ? cabs(M.A[j * nn + i]) * (*dr)[j] : cabs(M.A[j * nn + i]);
            if (dl) ni *= (*dl)[i];
            nrm = ni > nrm ? ni : nrm;
        }
        return g.rmax(nrm);
    }
    HC_HDN double jac_cond(const RV* dl, const RV* dr) {  // :745-774
        HC_COLD_N
        if (n == 1) {
            cx a0 = M.A[0]; double a = hypot(a0.re, a0.im);
            if (dl) a *= (*dl)[0];
            if (dr) a *= (*dr)[0];
            return 1.0 / a;
        }
        double e = inverse_inf_norm_est(dl, dr);
        return e * a_inf_norm(dl, dr);
    }
    // LA.cond(tracker, x, t, d_l, d_r)  tracker.jl:509-514
    HC_HDN double cond_at(CV x, cx t, const RV* dl, const RV* dr) {
        eval_f64(M.r, &M.A, x, t);
        updated();
        return jac_cond(dl, dr);
    }

    // ================================================================ norm weights (norm.jl:101-136)
    HC_HDN void norm_init(CV x) {
        HC_COLD_N
        const double pn = inf_norm(x);
        HC_PAR(i, n) {
            double wi = cabs(x[i]);
            if (wi < O->scale_min * pn) wi = O->scale_min * pn;
            else if (wi > O->scale_max * pn) wi = O->scale_max * pn;
            M.w[i] = jmax(wi, O->scale_abs_min);
        }
        g.sync();
    }
    HC_HDN void norm_update(CV x) {
        const double nx = wnorm(x);
        HC_PAR(i, n) {
            double wi = (cabs(x[i]) + M.w[i]) / 2;
            if (wi < O->scale_min * nx) wi = O->scale_min * nx;
            else if (wi > O->scale_max * nx) wi = O->scale_max * nx;
            if (wi == wi && wi < HC_INF && wi > -HC_INF) M.w[i] = jmax(wi, O->scale_abs_min);
        }
        g.sync();
    }

    // ================================================================ stepper (utils.jl:300-394)
    HC_HD void st_init(cx a, cx b) {
        st_start = a; st_target = b;
        st_absd = hypot(b.re - a.re, b.im - a.im);
        st_forward = hypot(a.re, a.im) < hypot(b.re, b.im);
        st_s = st_sp = st_forward ? 0.0 : st_absd;
    }
    HC_HD bool st_done() const { return st_forward ? st_s == st_absd : st_s == 0.0; }
    HC_HD void st_propose(double ds) { st_sp = st_forward ? jmin(st_s + ds, st_absd) : jmax(st_s - ds, 0.0); }
    HC_HD double st_dist() const { return st_forward ? st_absd - st_s : st_s; }
    HC_HD double st_ds() const { return st_forward ? st_sp - st_s : st_s - st_sp; }
    HC_HD cx st_at(double s) const {
        if (st_forward) {
            if (s == 0.0) return st_start;
            if (s == st_absd) return st_target;
            return st_start + (s / st_absd) * (st_target - st_start);
        }
        if (s == st_absd) return st_start;
        if (s == 0.0) return st_target;
        return st_target + (s / st_absd) * (st_start - st_target);
    }
    HC_HD cx st_t() const { return st_at(st_s); }
    HC_HD cx st_tp() const { return st_at(st_sp); }
    HC_HD cx st_dt() const {
        double f = st_forward ? (st_sp - st_s) / st_absd : (st_s - st_sp) / st_absd;
        return f * (st_target - st_start);
    }

    // ================================================================ predictor
    HC_HD void pred_init() {  // predictor.jl:124-133
        cond_H = 1.0; winding = 1;
        pt = pprev_t = ps = pprev_s = mk(HC_NAN, 0.0);
        trust_region = local_error = HC_NAN;
        pm_hermite = 0;
    }
    // cold parts of pred_update (winding > 1, singular endgame only), out of line to keep the regular step small
    HC_HDN void pu_splane(cx t) { pprev_s = ps; ps = t_to_s_plane(t, winding); }
    HC_HDN void pu_hermite(double nrm0, double nrm1) {
        HC_COLD_N
        const int nn = n;
        CV x1 = M.tx.at(nn);
        cx mu_ = winding == 2 ? 2.0 * ps : (double)winding * cpowi(ps, winding - 1);
        HC_PAR(i, nn) { x1[i] = M.xtemp[i]; M.ty1[nn + i] = mu_ * M.xtemp[i]; }
        g.sync();
        pm_hermite = 1;
        trust_region = nrm0 / nrm1;
        if (local_error != local_error) { double q = nrm1 / nrm0; local_error = q * q * q; }
    }
    // update!(predictor, H, x, t, J, norm, xhat)  predictor.jl:158-284
    // SYNC = true (lockstep kernels): called by every thread of the CTA once per round; lanes without an accepted
    // step (`act` false) only keep the barriers company.  The warps re-align before each of the three Taylor passes.
    template <bool SYNC>
    HC_HDN void pred_update_t(bool act, cx t, bool have_xhat) {
        const int nn = n;
        CV x0 = M.tx, x1 = M.tx.at(nn), x2 = M.tx.at(2 * nn), x3 = M.tx.at(3 * nn);
        double nrm0 = 0, nrm1 = 0, nrm2 = 0, nrm3 = 0, delta = 0;
        if (SYNC) cta_sync();
        if (act) {
            // (x, x') of the previous update: read by the Hermite predictor only, i.e. once winding > 1; tx still holds
            // them here, so the copy into the cold slab waits until then
            if (winding > 1) { HC_PAR(i, 2 * nn) M.ptx1[i] = M.tx[i]; g.sync(); }
#if HC_JIT_PREFETCH & 1
            for (int i = 0; i < nn * nn; ++i) M.A.prefetch(i);  // refine_fixed multiplies with A after the first Taylor pass
#endif
#if HC_JIT_PREFETCH & 2
            if (kind == H_TORIC) { const int P = H->P; for (int i = 0; i < P; ++i) M.wt.prefetch(i); }
#endif
            pprev_t = pt; pt = t;
            if (winding > 1) pu_splane(t);
            if (!have_xhat) local_error = HC_NAN;
            else {
                double ds = cabs(t - pprev_t), d2 = ds * ds;
                local_error = wdist(M.xhat, M.x) / (d2 * d2);
            }
            HC_PAR(i, nn) { x0[i] = M.x[i]; if (winding > 1) M.ty1[i] = M.x[i]; }
            g.sync();
            nrm0 = wnorm(M.x);

            taylor<1>(M.u, M.tx, t);
            HC_PAR(i, nn) M.u[i] = -M.u[i];
            g.sync();
            ldiv(M.xtemp, M.u, false);
            delta = refine_fixed(M.xtemp, M.u, true);
            cond_H = delta / HC_EPS;
            if (delta > 1e-10) iterative_refinement(M.xtemp, M.u, false, 5, 1e-10);
            nrm1 = wnorm(M.xtemp);
            if (winding > 1) { pu_hermite(nrm0, nrm1); act = false; }
        }
        if (SYNC) cta_sync();
        if (act) {
            vcopy(x1, M.xtemp, nn);
            taylor<2>(M.u, M.tx, t);
            HC_PAR(i, nn) M.u[i] = -M.u[i];
            g.sync();
            ldiv(M.xtemp, M.u, false);
            if (delta > 1e-10) iterative_refinement(M.xtemp, M.u, true, 4, 1e-10);
            nrm2 = wnorm(M.xtemp);
            vcopy(x2, M.xtemp, nn);
        }
        if (SYNC) cta_sync();
        if (!act) return;
        taylor<3>(M.u, M.tx, t);
        HC_PAR(i, nn) M.u[i] = -M.u[i];
        g.sync();
        ldiv(M.xtemp, M.u, false);
        if (delta > 1e-4) iterative_refinement(M.xtemp, M.u, true, 3, 1e-4);
        nrm3 = wnorm(M.xtemp);
        vcopy(x3, M.xtemp, nn);

        double tau_ = HC_INF;
        HC_PAR(i, nn) {
            double c1 = cabs(x1[i]), c2 = cabs(x2[i]), c3 = cabs(x3[i]);
            double lam = jmax(1e-6, c1);
            c1 /= lam; c2 /= lam * lam; c3 /= lam * lam * lam;
            double tol = 1e-14 * jmax(jmax(c1, c2), c3);
            if (!((c1 <= tol && c2 <= tol && c3 <= tol) || c2 <= tol)) {
                double ti = (c2 / c3) / lam;
                if (ti < tau_) tau_ = ti;
            }
        }
        {   // min over the lanes; NaN candidates never win (as in the sequential `ti < tau`)
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) { double w = g.xshfl(tau_, o); tau_ = w < tau_ ? w : tau_; }
            tau_ = g.bcast(tau_, 0);
        }
        if (!(tau_ < HC_INF && tau_ > -HC_INF)) tau_ = nrm2 / nrm3;
        if (!(tau_ < HC_INF && tau_ > -HC_INF)) tau_ = nrm0 / jmax(jmax(nrm0, nrm1), jmax(nrm2, nrm3));
        pm_hermite = 0;
        trust_region = tau_;
        if (local_error != local_error) { double q = 1.0 / tau_; local_error = (q * q) * (q * q); }
    }
    HC_HD void pred_update(cx t, bool have_xhat) { pred_update_t<false>(true, t, have_xhat); }
    HC_HDN void cubic_hermite(CV xh, CV v0, CV d0, cx t0, CV v1, CV d1, cx t1, cx t) {  // predictor.jl:354-371
        HC_COLD_N
        const int nn = n;
        if (t0.im == 0 && t1.im == 0 && t.im == 0) {
            double T = t.re, T0 = t0.re, T1 = t1.re;
            double s = (T - T0) / (T1 - T0), oms2 = (1 - s) * (1 - s);
            double h00 = (1 + 2 * s) * oms2, h10 = (T - T0) * oms2, h01 = (s * s) * (3 - 2 * s), h11 = (T - T0) * s * (s - 1);
            HC_PAR(i, nn) xh[i] = h00 * v0[i] + h10 * d0[i] + h01 * v1[i] + h11 * d1[i];
        } else {
            cx one = mk(1.0), s = cdiv(t - t0, t1 - t0), oms2 = (one - s) * (one - s);
            cx h00 = (one + 2.0 * s) * oms2, h10 = (t - t0) * oms2, h01 = (s * s) * (mk(3.0) - 2.0 * s), h11 = (t - t0) * s * (s - one);
            HC_PAR(i, nn) xh[i] = h00 * v0[i] + h10 * d0[i] + h01 * v1[i] + h11 * d1[i];
        }
        g.sync();
    }
    HC_HDN void predict(cx t, cx dt) {  // predictor.jl:286-329
        const int nn = n;
        if (!pm_hermite) {
            double lam = trust_region, lam2 = lam * lam, lam3 = lam2 * lam;
            HC_PAR(i, nn) {
                cx X = M.tx[i], X1 = M.tx[nn + i], X2 = M.tx[2 * nn + i], X3 = M.tx[3 * nn + i];
                double c = cabs(X), a1 = cabs(X1) * lam, a2 = cabs(X2) * lam2, a3 = cabs(X3) * lam3;
                double tau_ = 1e-12 * sqrt(c * c + a1 * a1 + a2 * a2 + a3 * a3);
                if (a3 <= tau_ || a2 <= tau_) M.xhat[i] = X + dt * (X1 + dt * X2);
                else {
                    cx d = mk(1.0) - cdiv(dt * X3, X2);
                    M.xhat[i] = X + dt * (X1 + cdiv(dt * X2, d));
                }
            }
            g.sync();
        } else predict_hermite(t, dt);
    }
    // cold half of predict (winding > 1 only): kept out of line so that the regular step's code stays small
    HC_HDN void predict_hermite(cx t, cx dt) {
        HC_COLD_N
        const int nn = n;
        int mw = winding;
        cx s0 = t_to_s_plane(pprev_t, mw), s1 = t_to_s_plane(t, mw), sp = t_to_s_plane(t + dt, mw);
        cx psm, sm;
        if (mw == 2) { psm = 2.0 * s0; sm = 2.0 * s1; }
        else { psm = (double)mw * cpowi(s0, mw - 1); sm = (double)mw * cpowi(s1, mw - 1); }
        HC_PAR(i, nn) {
            M.pty1[i] = M.ptx1[i]; M.pty1[nn + i] = psm * M.ptx1[nn + i];
            M.ty1[i] = M.tx[i]; M.ty1[nn + i] = sm * M.tx[nn + i];
        }
        g.sync();
        cubic_hermite(M.xhat, M.pty1, M.pty1.at(nn), s0, M.ty1, M.ty1.at(nn), s1, sp);
    }

    // ================================================================ Newton corrector
    HC_HDN void axmy(CV out, CV a, CV b) {  // out = a - b
        if (G == 1) {
            int i = 0;
            for (; i + 3 < n; i += 4) {
                cx a0 = a[i], a1 = a[i + 1], a2 = a[i + 2], a3 = a[i + 3], b0 = b[i], b1 = b[i + 1], b2 = b[i + 2], b3 = b[i + 3];
                out[i] = a0 - b0; out[i + 1] = a1 - b1; out[i + 2] = a2 - b2; out[i + 3] = a3 - b3;
            }
            for (; i < n; ++i) out[i] = a[i] - b[i];
        } else { HC_PAR(i, n) out[i] = a[i] - b[i]; g.sync(); }
    }
    // extended_prec_refinement_step!  newton_corrector.jl:55-78 (xout may alias xin)
    HC_HDN double ext_refinement_step(CV xout, CV xin, cx t, bool simple_newton_step) {
        eval_f64(M.r, &M.A, xin, t);
        eval_dd(M.r, xin, nullptr, t);
        updated();
        ldiv(M.dx, M.r, true);
        iterative_refinement(M.dx, M.r, true, 3, 1e-8);
        axmy(xout, xin, M.dx);
        if (simple_newton_step) {
            eval_dd(M.r, xout, nullptr, t);
            ldiv(M.dx, M.r, true);
        }
        return wnorm(M.dx);
    }
    // newton!  newton_corrector.jl:80-205; iterates live in xbar (x0 may alias xbar).
    // The reference's loop body and its convergence branch (:142-197, one more Jacobian + solve)
    // are folded into ONE loop with a single evaluate/factorize/solve call site (`final` marks the
    // convergence pass), so that lanes of a warp that are in different Newton iterations still
    // execute the expensive primitives together.
    // SYNC = true (lockstep kernels): every thread of the CTA calls this once per round, `act` says whether its lane
    // has a corrector to run; the trip count is CTA-uniform (lanes that are done idle through the remaining trips) and
    // the warps re-align before the evaluation and before the factorization of every trip.
    struct NewtonState { double ndxi, ndxim1, abar, mu; bool final; int i; };
    // the part of a trip after evaluate + solve; true = the iteration ended and R is complete
    HC_HD bool newton_decide(NewtonResult& R, NewtonState& s, double nd, CV xi, cx t, bool ext, bool accurate_mu, bool first_correction, double h_a) {
        if (!s.final) {
            s.ndxi = nd;
            if (s.ndxi != s.ndxi) { R.code = NEWT_SINGULARITY; R.accuracy = s.ndxi; R.iters = s.i + 1; return true; }
            axmy(xi, xi, M.dx);
            if (s.i == 0) R.norm_dx0 = s.ndxi;
            if (s.i == 1) R.omega = 2 * s.ndxi / (s.ndxim1 * s.ndxim1);
            if (s.i >= 1) R.theta = s.ndxi / s.ndxim1;
            if ((s.i >= 1 && R.theta > s.abar) || (s.i == 0 && !first_correction && 0.125 * R.norm_dx0 * R.omega > h_a)) {
                R.code = NEWT_TERMINATED; R.accuracy = s.ndxi; R.iters = s.i + 1; return true;
            }
            if (R.omega * s.ndxi * s.ndxi < 2 * s.mu * sqrt(1 - 2 * h_a)) { s.final = true; return false; }
            if (s.i == 10) { R.code = NEWT_MAX_ITERS; R.accuracy = s.mu; R.iters = 11; return true; }
            s.ndxim1 = s.ndxi;
            if (s.i >= 1) s.abar *= s.abar;
            ++s.i;
            return false;
        }
        axmy(xi, xi, M.dx);
        double ndxip1 = nd;
        if (ndxip1 != ndxip1) { R.code = NEWT_SINGULARITY; R.accuracy = ndxip1; R.iters = s.i + 1; return true; }
        if (ndxip1 > sqrt(s.ndxi)) {
            R.theta = ndxip1 / s.ndxi; R.code = NEWT_TERMINATED; R.accuracy = ndxip1; R.iters = s.i + 2; return true;
        }
        if (ndxip1 > 2 * s.mu && ext) {
            eval_dd(M.r, xi, nullptr, t);
            ldiv(M.dx, M.r, false);
            s.ndxi = ndxip1;
            s.mu = ndxip1 = wnorm(M.dx);
        } else if (ndxip1 > 2 * s.mu || accurate_mu) {
            eval_f64(M.r, nullptr, xi, t);
            ldiv(M.dx, M.r, false);
            s.mu = wnorm(M.dx);
        } else s.mu = ndxip1;
        if (s.i == 0) {
            double ob = 2 * s.ndxi / (ndxip1 * ndxip1);
            if (ob < R.omega) R.omega = ob; else R.omega *= 0.25;
        }
        R.code = NEWT_CONVERGED; R.accuracy = s.mu; R.iters = s.i + 2; return true;
    }
    template <bool SYNC>
    HC_HDN NewtonResult newton_t(bool act, CV x0, cx t, double mu_, double omega_, bool ext, bool accurate_mu, bool first_correction) {
        const int nn = n;
        const double a = O->a, h_a = hfun(a);
        CV xi = M.xbar;
        if (act && xi.p != x0.p) vcopy(xi, x0, nn);
        NewtonResult R; R.mu_low = R.theta = R.norm_dx0 = HC_NAN; R.omega = omega_; R.code = NEWT_TERMINATED; R.accuracy = HC_NAN; R.iters = 0;
        NewtonState s; s.ndxi = s.ndxim1 = HC_NAN; s.abar = a; s.mu = mu_; s.final = false; s.i = 0;
        bool live = act;
        while (true) {
            if (SYNC) { if (!cta_any(live)) break; } else if (!live) break;
            if (live) {
#if defined(HC_JIT_GEN)
                if (use_reg_lu()) eval_trip(M.r, xi, t, s.final || ext, !s.final); else
#endif
                eval_f64(M.r, &M.A, xi, t);
                if (ext && !s.final) eval_dd(M.r, xi, nullptr, t);
                updated();
            }
            if (SYNC) cta_sync();
            if (live) {
                if (ext && s.final) {
                    ldiv(M.dx, M.r, false);
                    R.mu_low = wnorm(M.dx);
                    eval_dd(M.r, xi, nullptr, t);
                }
                ldiv(M.dx, M.r, !s.final);
                if (ext) iterative_refinement(M.dx, M.r, true, 3, s.abar * s.abar);
                const double nd = wnorm(M.dx);
                if (newton_decide(R, s, nd, xi, t, ext, accurate_mu, first_correction, h_a)) live = false;
            }
        }
        return R;
    }
    HC_HD NewtonResult newton(CV x0, cx t, double mu_, double omega_, bool ext, bool accurate_mu, bool first_correction) {
        return newton_t<false>(true, x0, t, mu_, omega_, ext, accurate_mu, first_correction);
    }
    // init_newton!  newton_corrector.jl:207-286
    HC_HDN bool init_newton(cx t, bool ext, double& omega_out, double& mu_out) {
        HC_COLD_N
        const int nn = n;
        const double a = O->a, a7 = a * a * a * a * a * a * a;
        eval_f64(M.r, &M.A, M.x, t);
        if (ext) eval_dd(M.r, M.x, nullptr, t);
        updated();
        ldiv(M.dx, M.r, true);
        double v = wnorm(M.dx) + HC_EPS;
        bool valid = false;
        omega_out = mu_out = HC_NAN;
        double e = sqrt(v);
        for (int k = 1; k <= 3; ++k) {
            HC_PAR(i, nn) M.xbar[i] = M.x[i] + mk(e * M.w[i]);
            g.sync();
            eval_f64(M.r, &M.A, M.xbar, t);
            if (ext) eval_dd(M.r, M.xbar, nullptr, t);
            updated();
            ldiv(M.dx, M.r, true);
            axmy(M.xbar, M.xbar, M.dx);
            double nd0 = wnorm(M.dx);
            if (ext) eval_dd(M.r, M.xbar, nullptr, t); else eval_f64(M.r, nullptr, M.xbar, t);
            ldiv(M.dx, M.r, true);
            axmy(M.xbar, M.xbar, M.dx);
            double nd1 = wnorm(M.dx) + HC_EPS;
            if (nd1 < a * nd0) {
                omega_out = 2 * nd1 / (nd0 * nd0);
                mu_out = nd1;
                if (omega_out * mu_out > a7) {
                    NewtonResult res = newton(M.xbar, t, a7 / omega_out, omega_out, ext, true, false);
                    if (res.code == NEWT_CONVERGED) { valid = true; omega_out = res.omega; mu_out = res.accuracy; }
                    else valid = false;
                } else { valid = true; break; }
            } else e *= sqrt(e);
        }
        return valid;
    }

    // ================================================================ tracker
    HC_HD int steps() const { return accepted_steps + rejected_steps; }
    HC_HD int ext_steps() const { return ext_accepted_steps + ext_rejected_steps; }

    HC_HD double initial_step_size() {  // tracker.jl:520-539
        double a = O->beta_a * O->a;
        double e = local_error;
        if (e == HC_INF || e == -HC_INF) e = 1e5;
        double ds1 = nthroot((sqrt(1 + 2 * hfun(a)) - 1) / (omega * e), 4) / O->beta_omega_p;
        double ds2 = O->beta_tau * trust_region;
        double ds = nanmin(ds1, ds2);
        return jmin(jmin(ds, O->max_step_size), O->max_initial_step_size);
    }
    HC_HDN double rejected_stepsize(const NewtonResult& R, double a) {  // tracker.jl:572-586 (cold: general n-th roots)
        int j = R.iters - 2;
        int rootn = j >= 0 ? (1 << j) : 0;
        double Th = rootn == 8 ? sqrt(sqrt(sqrt(R.theta))) : nthroot(R.theta, rootn);
        double hT = hfun(Th), ha = hfun(0.5 * a);
        if (Th != Th || R.code == NEWT_SINGULARITY || R.accuracy != R.accuracy || R.iters == 1 || hT < ha) return 0.25 * st_ds();
        return nthroot((sqrt(1 + 2 * ha) - 1) / (sqrt(1 + 2 * hT) - 1), 4) * st_ds();
    }
    HC_HDN void update_stepsize(const NewtonResult& R) {  // tracker.jl:541-588
        double a = O->beta_a * O->a;
        double om = clampd(omega + 2 * (omega - omega_prev), omega, 8 * omega);
        double ds;
        if (R.code == NEWT_CONVERGED) {
            double e = local_error;
            double ds1 = nthroot((sqrt(1 + 2 * hfun(a)) - 1) / (om * e), 4) / O->beta_omega_p;
            double ds2 = O->beta_tau * tau;
            if (use_strict_beta_tau || st_dist() < ds2) ds2 = O->strict_beta_tau * tau;
            ds = jmin(nanmin(ds1, ds2), O->max_step_size);
            if (use_strict_beta_tau && st_dist() < ds) ds *= O->strict_beta_tau;
            ds = jmin(ds, 10 * ds_prev);
            if (last_steps_failed > 0) ds = jmin(ds, ds_prev);
        } else ds = rejected_stepsize(R, a);
        st_propose(ds);
    }
    HC_HDN void check_terminated() {  // tracker.jl:591-619
        double tol_acc = HC_INF;
        if (extended_prec || !O->extended_precision) tol_acc = tol_acc_limit;  // a^(2^min_newton_iters - 1) h(a), set once per path
        cx tp = st_tp(), t = st_t();
        if (st_done()) code = TC_success;
        else if (steps() >= O->max_steps) code = TC_terminated_max_steps;
        else if (omega * mu > tol_acc) code = TC_terminated_accuracy_limit;
        else if (st_ds() < min_step_size) code = TC_terminated_step_size_too_small;
        else if (cabs(tp - t) <= 2 * eps_of(cabs(t))) code = TC_terminated_step_size_too_small;
        else if (min_rel_step_size > 0 && !(tp.re == st_target.re && tp.im == st_target.im) && cabs(tp - t) < cabs(t) * min_rel_step_size)
            code = TC_terminated_step_size_too_small;
    }
    // rank(J; rtol = 1e-14) < n ?  (tracker.jl:711-737).  One-sided Jacobi on the columns of LU (scratch);
    // only reached for invalid start values, so lane 0 does it alone.
    HC_HDN bool jacobian_rank_deficient() {
        HC_COLD_N
        const int nn = n;
        int deficient = 0;
        if (g.lane == 0) {
            LV V = M.LU;
            bool has_nan = false;
            for (int i = 0; i < nn * nn; ++i) { V[i] = M.A[i]; has_nan = has_nan || cisnan(M.A[i]); }
            if (!has_nan) {
                for (int sweep = 0; sweep < 60; ++sweep) {
                    double off = 0;
                    for (int p = 0; p < nn; ++p)
                        for (int q = p + 1; q < nn; ++q) {
                            double app = 0, aqq = 0; cx apq = mk(0.0);
                            for (int i = 0; i < nn; ++i) { cx vp = V[p * nn + i], vq = V[q * nn + i]; app += abs2(vp); aqq += abs2(vq); apq = cfma(conj(vp), vq, apq); }
                            double gg = hypot(apq.re, apq.im);
                            if (gg <= 1e-300 || gg <= 1e-17 * sqrt(app * aqq)) continue;
                            off = fmax(off, gg / sqrt(app * aqq));
                            cx ph = apq / gg;
                            double zeta = (aqq - app) / (2 * gg);
                            double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1 + zeta * zeta));
                            double c = 1 / sqrt(1 + tt * tt), s = c * tt;
                            for (int i = 0; i < nn; ++i) {
                                cx vp = V[p * nn + i], vq = V[q * nn + i] * conj(ph);
                                V[p * nn + i] = c * vp - s * vq;
                                V[q * nn + i] = (s * vp + c * vq) * ph;
                            }
                        }
                    if (off < 1e-15) break;
                }
                double smax = 0;
                for (int p = 0; p < nn; ++p) { double s2 = 0; for (int i = 0; i < nn; ++i) s2 += abs2(V[p * nn + i]); smax = fmax(smax, sqrt(s2)); }
                int r = 0;
                for (int p = 0; p < nn; ++p) { double s2 = 0; for (int i = 0; i < nn; ++i) s2 += abs2(V[p * nn + i]); if (sqrt(s2) > 1e-14 * smax) ++r; }
                deficient = r < nn;
            }
        }
        g.sync();
        return g.bcast(deficient, 0) != 0;
    }

    // init!(tracker, x1, t1, t0; omega, mu, tau, max_initial_step_size, keep_steps, extended_precision)
    // tracker.jl:639-754; M.x must already hold x1.
    HC_HDN bool tracker_init(cx t1, cx t0, double omega_, double mu_, double tau_, double max_init_step, bool keep_steps, bool ext) {
        st_init(t1, t0);
        ds_prev = 0.0; accuracy = HC_EPS; omega = 1.0;
        keep_extended_prec = false; use_strict_beta_tau = false;
        norm_init(M.x);
        n_fact = n_ldiv = 0;
        code = TC_tracking;
        if (!keep_steps) accepted_steps = rejected_steps = ext_accepted_steps = ext_rejected_steps = 0;
        last_steps_failed = 0;
        cx t = st_t();
        bool valid = true;
        if (omega_ != omega_ || mu_ != mu_) {
            valid = init_newton(t, ext, omega_, mu_);
            if (!valid && !ext) { ext = true; valid = init_newton(t, true, omega_, mu_); }
        }
        used_extended_prec = extended_prec = ext;
        if (omega_ == omega_) omega = omega_;
        if (valid) { accuracy = mu_; mu = jmax(mu_, HC_EPS); }
        else {
            eval_f64(M.r, &M.A, M.x, t);
            code = jacobian_rank_deficient() ? TC_terminated_invalid_startvalue_singular_jacobian : TC_terminated_invalid_startvalue;
            return false;
        }
        tau = tau_;
        eval_f64(M.r, &M.A, M.x, t);
        updated();
        pred_init();
        pred_update(t, false);
        tau = trust_region;
        double ds = initial_step_size();
        ds = jmax(jmin(ds, max_init_step), min_step_size);
        st_propose(ds);
        omega_prev = omega;
        return code == TC_tracking;
    }
    HC_HD void tracker_init_continue(cx t0) {  // init!(tracker, t0)  tracker.jl:756-766
        code = TC_tracking;
        st_init(st_t(), t0);
        st_propose(initial_step_size());
        ds_prev = 0.0;
    }
    HC_HDN void use_extended_precision() {  // tracker.jl:788-813
        if (!O->extended_precision || extended_prec) return;
        extended_prec = true; used_extended_prec = true;
        double m_ = mu;
        for (int i = 0; i < 2; ++i) m_ = ext_refinement_step(M.x, M.x, st_t(), false);
        mu = jmax(m_, HC_EPS);
    }
    HC_HD void update_precision(double mu_low) {  // tracker.jl:768-786
        if (!O->extended_precision) return;
        const double a = O->a, a5 = a * a * a * a * a, a7 = a5 * a * a;
        if (extended_prec && !keep_extended_prec && mu_low == mu_low && mu_low > mu) {
            if (mu_low * omega < a7 * hfun(a)) { extended_prec = false; mu = mu_low; }
        } else if (mu * omega > a5 * hfun(a)) use_extended_precision();
    }
    HC_HDN double refine_current_solution(double min_tol, int nsteps) {  // tracker.jl:815-844
        const int nn = n;
        double m_ = accuracy;
        double mb = ext_refinement_step(M.xbar, M.x, st_t(), false);
        if (mb < m_) { vcopy(M.x, M.xbar, nn); m_ = mb; }
        int k = 1;
        while (m_ > min_tol && k <= nsteps) {
            mb = ext_refinement_step(M.xbar, M.x, st_t(), true);
            if (mb < m_) { vcopy(M.x, M.xbar, nn); m_ = mb; }
            k += 1;
        }
        return m_;
    }
    // step!(tracker)  tracker.jl:851-926; returns true iff the step was accepted.
    // SYNC = true (lockstep kernels): every thread of the CTA calls this once per round, `act` = this lane steps.
    template <bool SYNC>
    HC_HDN bool tracker_step_t(bool act) {
        const int nn = n;
        cx t = mk(0.0), dt = mk(0.0), tp = mk(0.0);
        if (act) {
            t = st_t(); dt = st_dt(); tp = st_tp();
            predict(t, dt);
            norm_update(M.xhat);
        }
        NewtonResult R = newton_t<SYNC>(act, M.xhat, tp, mu, omega, extended_prec, false, accepted_steps == 0);
        const bool conv = act && R.code == NEWT_CONVERGED;
        if (conv) {
            vcopy(M.x, M.xbar, nn);
            ds_prev = st_ds();
            st_s = st_sp;
            accuracy = R.accuracy;
            mu = jmax(R.accuracy, HC_EPS);
            omega_prev = omega;
            omega = jmax(jmax(R.omega, 0.5 * omega), 0.1);
            update_precision(R.mu_low);
            if (st_done() && O->extended_precision && accuracy > 1e-14) {
                accuracy = refine_current_solution(1e-14, 3);
                refined_extended_prec = true;
            }
        }
        pred_update_t<SYNC>(conv, conv ? st_t() : t, true);
        if (conv) {
            tau = trust_region;
            accepted_steps += 1;
            ext_accepted_steps += extended_prec ? 1 : 0;
            last_steps_failed = 0;
        } else if (act) {
            rejected_steps += 1;
            ext_rejected_steps += extended_prec ? 1 : 0;
            last_steps_failed += 1;
        }
        if (!act) return false;
        update_stepsize(R);
        check_terminated();
        return last_steps_failed == 0;
    }
    HC_HD bool tracker_step() { return tracker_step_t<false>(true); }
};

}  // namespace hc
