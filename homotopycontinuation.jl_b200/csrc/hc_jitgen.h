// Host side: turns the reference's InstructionSequence of a system (the 24-byte, 1-based Instruction records the
// Julia host hands over, src/model_kit/instruction_sequence.jl:1-27, 145-254) into straight-line C++ member
// functions of hc::Path -- evaluate!, evaluate_and_jacobian! and taylor! (K = 1, 2, 3) of the homotopy -- that the
// specialised kernel of hc_jit.h compiles for sm_100a at run time (NVRTC).  This is the device analogue of the
// reference's default CompiledSystem / CompiledHomotopy (src/model_kit/compiled_system_homotopy.jl:178-243,
// generated functions per system): tape slots become registers, there is no dispatch loop, constants are literals,
// and Taylor coefficients that are known to be zero (constants, the padded top coefficient of x, parameters that are
// linear in t) never produce an instruction.
//
// Arithmetic per op follows the reference (file:line):
//   src/model_kit/operations.jl:184-248              op_* kernels (sqr as (x+y)(x-y), cb, invsqr, pow_int, ...)
//   src/model_kit/taylor.jl:607-878                  taylor_op_*: sqr 635-656, div 705-721, mul 723-737,
//                                                    pow_int 751-793, muladd 823-838, mulsub 840-862, submul 864-878
//   src/model_kit/instruction_interpreter.jl:252-442 which rule an op code selects
//   src/homotopies/straight_line_homotopy.jl:96-154  u = gamma t G + (1 - t) F and its Taylor coefficients
#pragma once
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hc_b200.h"
#include "hc_tape.h"

namespace hc {
namespace jit {

struct ProgCopy {  // deep copy of an hc_program_desc (the caller's arrays only live for the create call)
    std::vector<int32_t> instr, u_assign, U_assign;
    std::vector<double> consts;
    int param_offset = 0, n_params = 0, t_index = 0, var_offset = 0, n_vars = 0, out_dim = 0, tape_space = 0;
    bool valid = false;
    void assign(const hc_program_desc* d) {
        instr.assign(d->instructions, d->instructions + 6 * (size_t)d->n_instructions);
        consts.assign(d->constants, d->constants + 2 * (size_t)d->n_constants);
        u_assign.assign(d->u_assign, d->u_assign + 2 * (size_t)d->n_u);
        U_assign.assign(d->U_assign, d->U_assign + 2 * (size_t)d->n_U);
        param_offset = d->param_offset; n_params = d->n_params; t_index = d->t_index; var_offset = d->var_offset;
        n_vars = d->n_vars; out_dim = d->out_dim; tape_space = d->tape_space;
        valid = true;
    }
};

inline std::string lit(double d) {
    char b[64];
    if (d != d) return "HC_NAN";
    if (d > 1.7e308) return "HC_INF";
    if (d < -1.7e308) return "(-HC_INF)";
    snprintf(b, sizeof b, "%a", d);  // hexadecimal floating literal: exact
    return b;
}

// Straight-line code in SSA form with dead-code elimination.  A value is an id >= 0 (C++ name v<id>) or -1 = the
// exact zero, which every builder propagates symbolically.
class Emit {
public:
    struct Stmt { std::string text; std::vector<int> defs, deps; bool sink; bool hoist = false; };
    int hoist_distance = 0;  // statements by which loads marked `hoist` are moved ahead of their first use
    std::vector<Stmt> stmts;
    int nvals = 0;
    static std::string nm(int id) { return "v" + std::to_string(id); }
    int def(const std::string& expr, std::vector<int> deps) {
        const int id = nvals++;
        stmts.push_back({"const cx " + nm(id) + " = " + expr + ";", {id}, std::move(deps), false});
        return id;
    }
    // a statement that defines several values at once (`text` declares them)
    void group(const std::string& text, std::vector<int> defs, std::vector<int> deps) { stmts.push_back({text, std::move(defs), std::move(deps), false}); }
    int fresh() { return nvals++; }
    // a load with no inputs that should be in flight long before its value is needed
    int early_load(const std::string& expr) { const int id = def(expr, {}); stmts.back().hoist = true; return id; }
    void sink(const std::string& text, std::vector<int> deps) { stmts.push_back({text, {}, std::move(deps), true}); }
    // a sink that is emitted right after the statement that defines `id` (outputs leave their registers at once)
    std::vector<std::pair<int, std::string>> after;
    void sink_after(int id, const std::string& text) { after.push_back({id, text}); }
    // ---- complex arithmetic with symbolic zeros
    int mul(int a, int b) { return (a < 0 || b < 0) ? -1 : def(nm(a) + " * " + nm(b), {a, b}); }
    int add(int a, int b) { return a < 0 ? b : (b < 0 ? a : def(nm(a) + " + " + nm(b), {a, b})); }
    int neg(int a) { return a < 0 ? -1 : def("-" + nm(a), {a}); }
    int sub(int a, int b) { return b < 0 ? a : (a < 0 ? neg(b) : def(nm(a) + " - " + nm(b), {a, b})); }
    int fma(int a, int b, int c) {  // a * b + c
        if (a < 0 || b < 0) return c;
        if (c < 0) return mul(a, b);
        return def("cfma(" + nm(a) + ", " + nm(b) + ", " + nm(c) + ")", {a, b, c});
    }
    int fnma(int a, int b, int c) {  // c - a * b
        if (a < 0 || b < 0) return c;
        if (c < 0) return neg(mul(a, b));
        return def("cfnma(" + nm(a) + ", " + nm(b) + ", " + nm(c) + ")", {a, b, c});
    }
    int scale(double d, int a) { return a < 0 ? -1 : def(lit(d) + " * " + nm(a), {a}); }
    int un(const char* fn, int a) { return a < 0 ? -1 : def(std::string(fn) + "(" + nm(a) + ")", {a}); }
    int constant(double re, double im) { return (re == 0.0 && im == 0.0) ? def("mk(0.0)", {}) : def("mk(" + lit(re) + ", " + lit(im) + ")", {}); }

    std::string render(const char* indent = "        ") const {
        std::vector<char> live((size_t)nvals, 0), keep(stmts.size(), 0);
        for (const auto& a : after) live[a.first] = 1;
        for (size_t i = stmts.size(); i-- > 0;) {
            const Stmt& s = stmts[i];
            bool k = s.sink;
            for (int d : s.defs) k = k || live[d];
            if (!k) continue;
            keep[i] = 1;
            for (int d : s.deps) if (d >= 0) live[d] = 1;
        }
        // order of emission: kept statements in program order, hoisted loads moved `hoist_distance` statements earlier
        // (their mutual order is kept: they are volatile asm on the device)
        std::vector<size_t> order;
        for (size_t i = 0; i < stmts.size(); ++i) if (keep[i]) order.push_back(i);
        if (hoist_distance > 0) {
            std::vector<std::pair<long, size_t>> key;  // (position, statement); hoisted ones get an earlier position
            for (size_t k = 0; k < order.size(); ++k) key.push_back({stmts[order[k]].hoist ? (long)k * 2 - 2L * hoist_distance - 1 : (long)k * 2, order[k]});
            std::stable_sort(key.begin(), key.end(), [](const std::pair<long, size_t>& a, const std::pair<long, size_t>& b) { return a.first < b.first; });
            for (size_t k = 0; k < order.size(); ++k) order[k] = key[k].second;
        }
        std::string out;
        for (size_t i : order) {
            out += indent; out += stmts[i].text; out += "\n";
            for (int d : stmts[i].defs)
                for (const auto& a : after) if (a.first == d) { out += indent; out += a.second; out += "\n"; }
        }
        return out;
    }
};

typedef std::vector<int> Ser;  // coefficients 0..K of a truncated series (ids, -1 = zero)

// How a program reads its inputs.  mode: 0 = fixed parameters of F, 1 = fixed parameters of G (straight line),
// 2 = homotopy parameters, linear in t (parameter / coefficient), 3 = the polyhedral driver (toric stage, then
// coefficient stage: decided per lane at run time), 4 = toric homotopy (full series).
struct Leaves {
    Emit& E;
    const ProgCopy& pc;
    int mode, K;  // K < 0: scalar evaluation
    int pp;       // the batch carries per-path parameter rows
    std::vector<Ser> cst, par, var;
    Ser tser;
    Leaves(Emit& e, const ProgCopy& p, int mode_, int K_, int pp_ = 0) : E(e), pc(p), mode(mode_), K(K_), pp(pp_) {
        cst.resize((size_t)(pc.consts.size() / 2)); par.resize((size_t)pc.n_params); var.resize((size_t)pc.n_vars);
    }
    int width() const { return K < 0 ? 1 : K + 1; }
    Ser zeros() const { return Ser((size_t)width(), -1); }
    const Ser& constant(int i) {
        if (cst[i].empty()) { cst[i] = zeros(); cst[i][0] = E.constant(pc.consts[2 * i], pc.consts[2 * i + 1]); }
        return cst[i];
    }
    const Ser& param(int i) {
        if (!par[i].empty()) return par[i];
        Ser s = zeros();
        const std::string I = std::to_string(i);
        if (mode == 0) s[0] = E.def("pld<S>(H->F_params + " + I + ")", {});
        else if (mode == 1) s[0] = E.def("pld<S>(H->G_params + " + I + ")", {});
        else {
            const std::string T = "<" + std::to_string(mode) + ", " + std::to_string(pp) + ">";
            const int wt = mode >= 3 ? E.early_load("jit_ldw<" + std::to_string(mode) + ">(" + I + ", jc)") : -1;
            const std::string W = wt >= 0 ? ", " + Emit::nm(wt) : std::string(", mk(0.0)");
            const std::vector<int> wd = wt >= 0 ? std::vector<int>{wt} : std::vector<int>{};
            if (K < 1) s[0] = E.def("jit_par" + T + "(" + I + ", jc" + W + ")", wd);
            else if (mode == 2) {
                const int a = E.fresh(), b = E.fresh();
                E.group("cx " + Emit::nm(a) + ", " + Emit::nm(b) + "; jit_pser_lin<" + std::to_string(pp) + ">(" + I + ", jc, " + Emit::nm(a) + ", " + Emit::nm(b) + ");", {a, b}, {});
                s[0] = a; s[1] = b;
            } else {
                const int a = E.fresh(), b = E.fresh(), c = E.fresh(), d = E.fresh();
                E.group("cx " + Emit::nm(a) + ", " + Emit::nm(b) + ", " + Emit::nm(c) + ", " + Emit::nm(d) + "; jit_pser" + T + "(" + I + ", jc" + W + ", " + Emit::nm(a) +
                            ", " + Emit::nm(b) + ", " + Emit::nm(c) + ", " + Emit::nm(d) + ");", {a, b, c, d}, wd);
                s[0] = a; s[1] = b;
                if (K >= 2) s[2] = c;
                if (K >= 3) s[3] = d;
            }
        }
        return par[i] = s;
    }
    const Ser& variable(int i) {
        if (!var[i].empty()) return var[i];
        Ser s = zeros();
        const std::string I = std::to_string(i);
        if (K < 0) s[0] = E.early_load("x[" + I + "]");
        else for (int k = 0; k < K; ++k) s[k] = E.early_load("tx[" + std::to_string(k) + " * n + " + I + "]");  // coefficient K is the zero padding
        return var[i] = s;
    }
    const Ser& tvalue() {
        if (!tser.empty()) return tser;
        tser = zeros();
        tser[0] = E.def("t", {});
        if (K >= 1) tser[1] = E.constant(1.0, 0.0);
        return tser;
    }
};

struct SeriesOps {
    Emit& E;
    int K;  // series order; scalar evaluation: K = 0
    Ser t_add(const Ser& x, const Ser& y) { Ser r(x.size()); for (size_t k = 0; k < x.size(); ++k) r[k] = E.add(x[k], y[k]); return r; }
    Ser t_sub(const Ser& x, const Ser& y) { Ser r(x.size()); for (size_t k = 0; k < x.size(); ++k) r[k] = E.sub(x[k], y[k]); return r; }
    Ser t_neg(const Ser& x) { Ser r(x.size()); for (size_t k = 0; k < x.size(); ++k) r[k] = E.neg(x[k]); return r; }
    Ser t_mul(const Ser& x, const Ser& y) {  // taylor.jl:723-737
        Ser r(x.size());
        for (int k = 0; k <= K; ++k) {
            int c = E.mul(x[0], y[k]);
            for (int j = 1; j <= k; ++j) c = E.fma(x[j], y[k - j], c);
            r[k] = c;
        }
        return r;
    }
    Ser t_sqr(const Ser& x) {  // taylor.jl:635-656
        Ser r(x.size());
        r[0] = E.un("csqr", x[0]);
        for (int k = 1; k <= K; ++k) {
            int w = E.mul(x[0], x[k]);
            for (int j = 1; j <= (k - 1) / 2; ++j) w = E.fma(x[j], x[k - j], w);
            r[k] = (k % 2 == 0) ? E.add(E.scale(2.0, w), E.un("csqr", x[k / 2])) : E.add(w, w);
        }
        return r;
    }
    Ser t_div(const Ser& x, const Ser& y) {  // taylor.jl:705-721 (one reciprocal of y0, as the device interpreter does)
        Ser r(x.size());
        const int yinv = E.un("cinv", y[0]);
        for (int k = 0; k <= K; ++k) {
            int s = x[k];
            for (int j = 0; j < k; ++j) s = E.fnma(r[j], y[k - j], s);
            r[k] = E.mul(s, yinv);
        }
        return r;
    }
    Ser t_inv(const Ser& x) {  // taylor.jl:607-623
        Ser one(x.size(), -1);
        one[0] = E.constant(1.0, 0.0);
        return t_div(one, x);
    }
    Ser t_pow_int(const Ser& x, int r) {  // taylor.jl:751-793; a zero constant term yields the zero series
        Ser w(x.size(), -1);
        if (r == 0) { w[0] = E.constant(1.0, 0.0); return w; }
        if (K == 0) { w[0] = E.def("cpowi(" + Emit::nm(x[0]) + ", " + std::to_string(r) + ")", {x[0]}); return w; }
        w[0] = E.def("cpowi(" + Emit::nm(x[0]) + ", " + std::to_string(r) + ")", {x[0]});
        const int u0inv = E.un("cinv", x[0]);
        for (int k = 1; k <= K; ++k) {
            int s = -1, t = -1;
            for (int j = 1; j <= k; ++j) s = E.fma(w[k - j], E.scale((double)j, x[j]), s);
            s = E.scale((double)r, s);
            for (int j = 1; j <= k - 1; ++j) t = E.fma(x[k - j], E.scale((double)j, w[j]), t);
            const int d = E.sub(s, t);
            w[k] = d < 0 ? -1 : E.def("(" + Emit::nm(u0inv) + " * " + Emit::nm(d) + ") / " + lit((double)k), {u0inv, d});
        }
        for (int k = 0; k <= K; ++k)
            if (w[k] >= 0) w[k] = E.def("ciszero(" + Emit::nm(x[0]) + ") ? mk(0.0) : " + Emit::nm(w[k]), {x[0], w[k]});
        return w;
    }
    Ser t_muladd(const Ser& x, const Ser& y, const Ser& z) {  // :823-838
        Ser r(x.size());
        for (int k = 0; k <= K; ++k) { int c = z[k]; for (int j = 0; j <= k; ++j) c = E.fma(x[j], y[k - j], c); r[k] = c; }
        return r;
    }
    Ser t_mulsub(const Ser& x, const Ser& y, const Ser& z) {  // :840-862
        Ser r(x.size());
        for (int k = 0; k <= K; ++k) {
            int c = E.sub(E.mul(x[0], y[k]), z[k]);
            for (int j = 1; j <= k; ++j) c = E.fma(x[j], y[k - j], c);
            r[k] = c;
        }
        return r;
    }
    Ser t_submul(const Ser& x, const Ser& y, const Ser& z) {  // :864-878
        Ser r(x.size());
        for (int k = 0; k <= K; ++k) { int c = z[k]; for (int j = 0; j <= k; ++j) c = E.fnma(x[j], y[k - j], c); r[k] = c; }
        return r;
    }
};

// Runs the reference tape symbolically; returns the series (width 1 for scalar evaluation) of every tape slot.
inline std::vector<Ser> run_symbolic(Emit& E, const ProgCopy& pc, Leaves& L, int K) {
    const bool scalar = K < 0;
    SeriesOps T{E, scalar ? 0 : K};
    std::vector<Ser> cur((size_t)pc.tape_space + 1);
    const int C = (int)(pc.consts.size() / 2);
    // inputs are bound lazily (only what the tape reads is ever loaded)
    std::vector<int> in_kind((size_t)pc.tape_space + 1, 0), in_idx((size_t)pc.tape_space + 1, 0);
    for (int i = 0; i < C; ++i) { in_kind[i + 1] = 1; in_idx[i + 1] = i; }
    for (int i = 0; i < pc.n_params; ++i) { in_kind[pc.param_offset + 1 + i] = 2; in_idx[pc.param_offset + 1 + i] = i; }
    if (pc.t_index > 0) in_kind[pc.t_index] = 3;
    for (int i = 0; i < pc.n_vars; ++i) { in_kind[pc.var_offset + 1 + i] = 4; in_idx[pc.var_offset + 1 + i] = i; }
    auto rd = [&](int slot) -> const Ser& {
        if (slot < 1 || slot > pc.tape_space) throw std::string("tape index out of range");
        if (cur[slot].empty()) {
            switch (in_kind[slot]) {
                case 1: cur[slot] = L.constant(in_idx[slot]); break;
                case 2: cur[slot] = L.param(in_idx[slot]); break;
                case 3: cur[slot] = L.tvalue(); break;
                case 4: cur[slot] = L.variable(in_idx[slot]); break;
                default: throw std::string("tape reads a slot before it is written");
            }
        }
        return cur[slot];
    };
    const size_t ni = pc.instr.size() / 6;
    for (size_t i = 0; i < ni; ++i) {
        const int32_t* s = pc.instr.data() + 6 * i;
        const int op = s[4];
        if (op == OP_STOP) break;
        Ser r;
        if (scalar) {
            auto v = [&](int k) { return rd(s[k])[0]; };
            auto one = [&](int id) { return Ser(1, id); };
            switch (op) {
                case OP_CB: r = one(E.un("ccb", v(0))); break;
                case OP_INV: r = one(E.un("cinv", v(0))); break;
                case OP_INV_NOT_ZERO: { const int a = v(0); r = one(E.def("ciszero(" + Emit::nm(a) + ") ? " + Emit::nm(a) + " : cinv(" + Emit::nm(a) + ")", {a})); } break;
                case OP_INVSQR: r = one(E.un("csqr", E.un("cinv", v(0)))); break;
                case OP_NEG: r = one(E.neg(v(0))); break;
                case OP_SQR: r = one(E.un("csqr", v(0))); break;
                case OP_IDENTITY: r = one(v(0)); break;
                case OP_ADD: r = one(E.add(v(0), v(1))); break;
                case OP_SUB: r = one(E.sub(v(0), v(1))); break;
                case OP_MUL: r = one(E.mul(v(0), v(1))); break;
                case OP_DIV: { const int a = v(0), b = v(1); r = one(E.def("cdiv(" + Emit::nm(a) + ", " + Emit::nm(b) + ")", {a, b})); } break;
                case OP_POW_INT: r = T.t_pow_int(rd(s[0]), s[1]); break;
                case OP_ADD3: r = one(E.add(E.add(v(0), v(1)), v(2))); break;
                case OP_MUL3: r = one(E.mul(E.mul(v(0), v(1)), v(2))); break;
                case OP_MULADD: r = one(E.fma(v(0), v(1), v(2))); break;
                case OP_MULSUB: r = one(E.fma(v(0), v(1), E.neg(v(2)))); break;
                case OP_SUBMUL: r = one(E.fnma(v(0), v(1), v(2))); break;
                case OP_ADD4: r = one(E.add(E.add(E.add(v(0), v(1)), v(2)), v(3))); break;
                case OP_MUL4: r = one(E.mul(E.mul(E.mul(v(0), v(1)), v(2)), v(3))); break;
                case OP_MULMULADD: r = one(E.fma(v(2), v(3), E.mul(v(0), v(1)))); break;
                case OP_MULMULSUB: r = one(E.fnma(v(2), v(3), E.mul(v(0), v(1)))); break;
                default: throw std::string("unsupported op in tape: ") + std::to_string(op);
            }
        } else {
            auto v = [&](int k) -> const Ser& { return rd(s[k]); };
            switch (op) {
                case OP_CB: r = T.t_mul(T.t_sqr(v(0)), v(0)); break;  // taylor.jl:279-281
                case OP_INV: case OP_INV_NOT_ZERO: r = T.t_inv(v(0)); break;
                case OP_INVSQR: r = T.t_inv(T.t_sqr(v(0))); break;
                case OP_NEG: r = T.t_neg(v(0)); break;
                case OP_SQR: r = T.t_sqr(v(0)); break;
                case OP_IDENTITY: r = v(0); break;
                case OP_ADD: r = T.t_add(v(0), v(1)); break;
                case OP_SUB: r = T.t_sub(v(0), v(1)); break;
                case OP_MUL: r = T.t_mul(v(0), v(1)); break;
                case OP_DIV: r = T.t_div(v(0), v(1)); break;
                case OP_POW_INT: r = T.t_pow_int(v(0), s[1]); break;
                case OP_ADD3: r = T.t_add(T.t_add(v(0), v(1)), v(2)); break;
                case OP_MUL3: r = T.t_mul(T.t_mul(v(0), v(1)), v(2)); break;
                case OP_MULADD: r = T.t_muladd(v(0), v(1), v(2)); break;
                case OP_MULSUB: r = T.t_mulsub(v(0), v(1), v(2)); break;
                case OP_SUBMUL: r = T.t_submul(v(0), v(1), v(2)); break;
                case OP_ADD4: r = T.t_add(T.t_add(v(0), v(1)), T.t_add(v(2), v(3))); break;
                case OP_MUL4: r = T.t_mul(T.t_mul(v(0), v(1)), T.t_mul(v(2), v(3))); break;
                case OP_MULMULADD: r = T.t_add(T.t_mul(v(0), v(1)), T.t_mul(v(2), v(3))); break;
                case OP_MULMULSUB: r = T.t_sub(T.t_mul(v(0), v(1)), T.t_mul(v(2), v(3))); break;
                default: throw std::string("unsupported op in tape: ") + std::to_string(op);
            }
        }
        const int out = s[5];
        if (out < 1 || out > pc.tape_space) throw std::string("tape index out of range");
        if (in_kind[out]) throw std::string("instruction writes into the input block");
        cur[out] = r;
    }
    // make sure assigned output slots that are plain inputs are bound too
    for (size_t i = 0; i + 1 < pc.u_assign.size(); i += 2) rd(pc.u_assign[i + 1]);
    for (size_t i = 0; i + 1 < pc.U_assign.size(); i += 2) rd(pc.U_assign[i + 1]);
    return cur;
}

struct GenInput {
    int kind = 0;       // HKind of the homotopy handle
    bool poly = false;  // driven by the polyhedral tracker: toric stage first, coefficient stage second (kind switches at run time)
    bool path_params = false;  // the batch carries per-path parameter rows (hc_track_batch path_p / path_q, hc_track_sweep)
    int hoist = 1000;          // statements by which loads of lane state (x, the toric (w, t^w) pairs) run ahead of their use: all of them
                               // are issued at the top of the function, in flight together (measured: + 7 % on cyclic-7 polyhedral)
    int pmode() const { return poly ? 3 : (kind == H_TORIC ? 4 : 2); }
    const ProgCopy *Fe = nullptr, *Fj = nullptr, *Ge = nullptr, *Gj = nullptr;
    int n = 0;
};

// u / U outputs of one program: id per entry (-1 = structurally zero)
inline void outputs_of(const ProgCopy& pc, const std::vector<Ser>& cur, int coef, std::vector<int>& u, std::vector<int>* U) {
    u.assign((size_t)pc.out_dim, -1);
    for (size_t i = 0; i + 1 < pc.u_assign.size(); i += 2) u[pc.u_assign[i] - 1] = cur[pc.u_assign[i + 1]][coef];
    if (U) {
        U->assign((size_t)pc.out_dim * pc.n_vars, -1);
        for (size_t i = 0; i + 1 < pc.U_assign.size(); i += 2) (*U)[pc.U_assign[i] - 1] = cur[pc.U_assign[i + 1]][coef];
    }
}

inline void store_all(Emit& E, const char* dst, const std::vector<int>& v) {
    for (size_t i = 0; i < v.size(); ++i) {
        if (v[i] < 0) E.sink(std::string(dst) + "[" + std::to_string(i) + "] = mk(0.0);", {});
        else E.sink_after(v[i], std::string(dst) + "[" + std::to_string(i) + "] = " + Emit::nm(v[i]) + ";");
    }
}
// Jacobian entries (column-major, j = col * m + row): stored into U, into U2 as well if keepA, and -- if rowsum --
// added to the Skeel row sums d_i = sum_j |U_ij| w_j (linear_algebra.jl:432-459) while they are in registers
inline void store_jacobian(Emit& E, const std::vector<int>& U, int m, int n) {
    for (size_t j = 0; j < U.size(); ++j) {
        const std::string J = std::to_string(j), row = std::to_string(j % (size_t)m), col = std::to_string(j / (size_t)m);
        if (U[j] < 0) { E.sink("U[" + J + "] = mk(0.0); if (keepA) U2[" + J + "] = mk(0.0);", {}); continue; }
        const std::string v = Emit::nm(U[j]);
        E.sink_after(U[j], "U[" + J + "] = " + v + "; if (keepA) U2[" + J + "] = " + v + "; if (rowsum) jrs" + row + " += cabs(" + v + ") * jw" + col + ";");
    }
    (void)n;
}

// evaluate! (jac = false) / evaluate_and_jacobian! (jac = true) of the homotopy at (x, t)
inline std::string gen_scalar(const GenInput& in, bool jac) {
    Emit E;
    E.hoist_distance = in.hoist;
    std::vector<int> u, U;
    if (in.kind == H_STRAIGHT_LINE) {  // straight_line_homotopy.jl:96-124: u = (gamma t) G + (1 - t) F, entry by entry
        const ProgCopy& PF = jac ? *in.Fj : *in.Fe;
        const ProgCopy& PG = jac ? *in.Gj : *in.Ge;
        const int ts = E.def("H->gamma * t", {}), tt = E.def("mk(1.0) - t", {});
        Leaves LF(E, PF, 0, -1), LG(E, PG, 1, -1);
        std::vector<int> uf, Uf, ug, Ug;
        outputs_of(PF, run_symbolic(E, PF, LF, -1), 0, uf, jac ? &Uf : nullptr);
        outputs_of(PG, run_symbolic(E, PG, LG, -1), 0, ug, jac ? &Ug : nullptr);
        u.resize(uf.size());
        for (size_t i = 0; i < uf.size(); ++i) u[i] = E.fma(ts, ug[i], E.mul(tt, uf[i]));
        if (jac) { U.resize(Uf.size()); for (size_t i = 0; i < Uf.size(); ++i) U[i] = E.fma(ts, Ug[i], E.mul(tt, Uf[i])); }
    } else {
        const ProgCopy& PF = jac ? *in.Fj : *in.Fe;
        Leaves LF(E, PF, in.pmode(), -1, in.path_params);
        outputs_of(PF, run_symbolic(E, PF, LF, -1), 0, u, jac ? &U : nullptr);
    }
    store_all(E, "u", u);
    const int m = (int)u.size(), n = in.n;
    if (jac) store_jacobian(E, U, m, n);
    std::string s = jac ? "    template <class UV, class UV2>\n    HC_HDN void jit_evaljac(CV u, UV U, UV2 U2, CV x, cx t, bool keepA, bool rowsum) {\n" : "    HC_HDN void jit_eval(CV u, CV x, cx t) {\n";
    if (in.kind != H_STRAIGHT_LINE) s += "        const JPar jc = jit_par_ctx(t);\n";
    if (jac) {
        for (int i = 0; i < m; ++i) s += "        double jrs" + std::to_string(i) + " = 0.0;\n";
        for (int j = 0; j < n; ++j) s += "        double jw" + std::to_string(j) + " = 0.0;\n";
        s += "        if (rowsum) {";
        for (int j = 0; j < n; ++j) s += " jw" + std::to_string(j) + " = M.w[" + std::to_string(j) + "];";
        s += " }\n";
    }
    s += E.render();
    if (jac) {
        s += "        if (rowsum) {";
        for (int i = 0; i < m; ++i) s += " M.rs[" + std::to_string(i) + "] = jrs" + std::to_string(i) + ";";
        s += " }\n";
    }
    s += "    }\n";
    return s;
}

// taylor!(u, Val(K), H, tx, t): u = K-th Taylor coefficient of lambda -> H(x(lambda), t + lambda)
inline std::string gen_taylor(const GenInput& in, int K) {
    Emit E;
    E.hoist_distance = in.hoist;
    std::vector<int> u;
    if (in.kind == H_STRAIGHT_LINE) {  // straight_line_homotopy.jl:130-154
        Leaves LG(E, *in.Ge, 1, K), LF(E, *in.Fe, 0, K);
        const std::vector<Ser> cg = run_symbolic(E, *in.Ge, LG, K), cf = run_symbolic(E, *in.Fe, LF, K);
        std::vector<int> gK, gK1, fK, fK1;
        outputs_of(*in.Ge, cg, K, gK, nullptr); outputs_of(*in.Ge, cg, K - 1, gK1, nullptr);
        outputs_of(*in.Fe, cf, K, fK, nullptr); outputs_of(*in.Fe, cf, K - 1, fK1, nullptr);
        const int t = E.def("t", {}), gamma = E.def("H->gamma", {}), omt = E.def("mk(1.0) - t", {});
        u.resize(fK.size());
        for (size_t i = 0; i < u.size(); ++i) {
            const int g = E.mul(gamma, E.add(gK1[i], E.mul(t, gK[i])));
            u[i] = E.add(g, E.sub(E.mul(omt, fK[i]), fK1[i]));
        }
    } else {
        Leaves LF(E, *in.Fe, in.pmode(), K, in.path_params);
        outputs_of(*in.Fe, run_symbolic(E, *in.Fe, LF, K), K, u, nullptr);
    }
    store_all(E, "u", u);
    std::string s = "    HC_HDN void jit_taylor" + std::to_string(K) + "(CV u, CV tx, cx t) {\n";
    if (in.kind != H_STRAIGHT_LINE) s += "        const JPar jc = jit_pser_ctx(t);\n";
    s += E.render();
    s += "    }\n";
    return s;
}

// the generated member functions of hc::Path (included through HC_JIT_GEN inside the struct)
inline std::string generate_members(const GenInput& in) {
    std::string s = "// generated by hc_jitgen.h -- member functions of hc::Path<G, S>\n";
    s += gen_scalar(in, false);
    s += gen_scalar(in, true);
    for (int K = 1; K <= 3; ++K) s += gen_taylor(in, K);
    return s;
}

}  // namespace jit
}  // namespace hc
