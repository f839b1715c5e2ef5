// ORACLE (test infrastructure, NOT product code): ModelKit instruction tape,
// interpreter (ComplexF64 / ComplexDF64) and truncated-Taylor interpreter.
//
// Follows (reference file:line):
//   src/model_kit/instruction_sequence.jl:1-27    Instruction / InstructionSequence
//   src/model_kit/instruction_sequence.jl:145-254 tape layout
//       [constants | parameters | t | variables | registers | assignments], 1-based
//   src/model_kit/operations.jl:5-49              OpType enum (declaration order)
//   src/model_kit/instruction_interpreter.jl:136-192, 252-332  execute!/execute_instructions!
//   src/model_kit/instruction_interpreter.jl:335-491           execute_taylor!
//   src/model_kit/taylor.jl:607-878               taylor_op_* for the polynomial ops
//   src/model_kit/interpreted_system.jl:65-119    evaluate!/evaluate_and_jacobian!/taylor!
#pragma once
#include <cstring>
#include <stdexcept>
#include <utility>
#include <vector>

#include "num.hpp"

namespace orc {

enum Op : int32_t {
    OP_STOP = 0, OP_CB, OP_ACOS, OP_ASIN, OP_COS, OP_COSH, OP_EXP, OP_INV, OP_INV_NOT_ZERO,
    OP_INVSQR, OP_NEG, OP_SIN, OP_SINH, OP_SQR, OP_SQRT, OP_TAN, OP_TANH, OP_IDENTITY,
    OP_ADD, OP_DIV, OP_MUL, OP_SUB, OP_POW_INT, OP_POW,
    OP_ADD3, OP_MUL3, OP_MULADD, OP_MULSUB, OP_SUBMUL,
    OP_ADD4, OP_MUL4, OP_MULMULADD, OP_MULMULSUB
};

struct Instr {  // instruction_sequence.jl:1-5 (24 bytes, isbits)
    int32_t in[4];
    int32_t op;
    int32_t out;
};

struct Program {
    std::vector<Instr> instr;     // terminated by OP_STOP
    std::vector<cplx> constants;  // tape[1..C]
    int param_off = 0, P = 0;     // parameters at tape[param_off+1 .. param_off+P]
    int t_index = 0;              // 0 = none (System); else 1-based slot (Homotopy)
    int var_off = 0, n = 0;       // variables at tape[var_off+1 .. var_off+n]
    int out_dim = 0;              // m
    int tape_space = 0;
    std::vector<std::pair<int, int>> u_assign;  // (i, k): u[i] = tape[k], 1-based
    std::vector<std::pair<int, int>> U_assign;  // (j, k): U[j] = tape[k], col-major over (m, n)
    bool all_u = false, all_U = false;
    bool supported = true;        // false if a transcendental op occurs
};

inline void finalize(Program& p) {
    p.all_u = (int)p.u_assign.size() == p.out_dim;
    p.all_U = (int)p.U_assign.size() == p.out_dim * p.n;
    for (auto& I : p.instr) {
        switch (I.op) {
            case OP_ACOS: case OP_ASIN: case OP_COS: case OP_COSH: case OP_EXP: case OP_SIN:
            case OP_SINH: case OP_SQRT: case OP_TAN: case OP_TANH: case OP_POW:
                p.supported = false;
            default: break;
        }
    }
}

// ------------------------------------------------------------- F64 / DD tape
template <class T>
inline void run_tape(const Program& P, T* tape) {  // instruction_interpreter.jl:252-260
    const Instr* I = P.instr.data();
    for (;; ++I) {
        const int a1 = I->in[0], a2 = I->in[1], a3 = I->in[2], a4 = I->in[3];
        T& o = tape[I->out];
        switch (I->op) {
            case OP_STOP: return;
            case OP_CB: o = op_cb(tape[a1]); break;
            case OP_INV: o = op_inv(tape[a1]); break;
            case OP_INV_NOT_ZERO: o = op_inv_not_zero(tape[a1]); break;
            case OP_INVSQR: o = op_invsqr(tape[a1]); break;
            case OP_NEG: o = -tape[a1]; break;
            case OP_SQR: o = op_sqr(tape[a1]); break;
            case OP_IDENTITY: o = tape[a1]; break;
            case OP_ADD: o = tape[a1] + tape[a2]; break;
            case OP_DIV: o = div_fast(tape[a1], tape[a2]); break;
            case OP_MUL: o = tape[a1] * tape[a2]; break;
            case OP_SUB: o = tape[a1] - tape[a2]; break;
            case OP_POW_INT: o = op_pow_int(tape[a1], a2); break;
            case OP_ADD3: o = tape[a1] + tape[a2] + tape[a3]; break;
            case OP_MUL3: o = tape[a1] * tape[a2] * tape[a3]; break;
            case OP_MULADD: o = tape[a1] * tape[a2] + tape[a3]; break;
            case OP_MULSUB: o = tape[a1] * tape[a2] - tape[a3]; break;
            case OP_SUBMUL: o = tape[a3] - tape[a1] * tape[a2]; break;
            case OP_ADD4: o = tape[a1] + tape[a2] + tape[a3] + tape[a4]; break;
            case OP_MUL4: o = tape[a1] * tape[a2] * tape[a3] * tape[a4]; break;
            case OP_MULMULADD: o = tape[a1] * tape[a2] + tape[a3] * tape[a4]; break;
            case OP_MULMULSUB: o = tape[a1] * tape[a2] - tape[a3] * tape[a4]; break;
            default: throw std::runtime_error("oracle: unsupported op");
        }
    }
}

// An Interpreter{Vector{T}}: instruction_interpreter.jl:1-6, 76-86
template <class T>
struct Interp {
    const Program* P = nullptr;
    std::vector<T> tape;
    void init(const Program* p) {
        P = p;
        tape.assign(p->tape_space + 2, T());
        for (size_t i = 0; i < p->constants.size(); ++i) tape[i + 1] = from_c<T>(p->constants[i]);
    }
    // execute!(u, [U,] I, x, t, p)  instruction_interpreter.jl:273-332.
    // params are always ComplexF64 (parameter_homotopy.jl:89-92: parameters stay F64 in DD eval)
    void load_and_run(const T* x, const cplx* t, const cplx* params) {
        const Program& p = *P;
        for (int i = 0; i < p.P; ++i) tape[p.param_off + 1 + i] = from_c<T>(params[i]);
        if (p.t_index && t) tape[p.t_index] = from_c<T>(*t);
        for (int i = 0; i < p.n; ++i) tape[p.var_off + 1 + i] = x[i];
        run_tape(p, tape.data());
    }
    // u kept in the tape's own scalar type (needed by StraightLineHomotopy's DD combine,
    // straight_line_homotopy.jl:81-94)
    void execute_native(T* u, const T* x, const cplx* t, const cplx* params) {
        const Program& p = *P;
        load_and_run(x, t, params);
        for (int i = 0; i < p.out_dim; ++i) u[i] = T();
        for (auto& a : p.u_assign) u[a.first - 1] = tape[a.second];
    }
    void execute(cplx* u, cplx* U, const T* x, const cplx* t, const cplx* params) {
        const Program& p = *P;
        load_and_run(x, t, params);
        if (U) {
            if (!p.all_U) std::memset((void*)U, 0, sizeof(cplx) * p.out_dim * p.n);
            for (auto& a : p.U_assign) U[a.first - 1] = to_cplx(tape[a.second]);
        }
        if (u) {
            if (!p.all_u) for (int i = 0; i < p.out_dim; ++i) u[i] = cplx();
            for (auto& a : p.u_assign) u[a.first - 1] = to_cplx(tape[a.second]);
        }
    }
};

// ------------------------------------------------------------- Taylor tape
// TruncatedTaylorSeries{K+1}: taylor.jl:1-52.  We always carry 5 coefficients and
// only compute orders 0..K.
constexpr int TMAX = 5;
struct Series { cplx c[TMAX]; };

template <int K> inline Series t_add(const Series& x, const Series& y) {
    Series r; for (int k = 0; k <= K; ++k) r.c[k] = x.c[k] + y.c[k]; return r;
}
template <int K> inline Series t_sub(const Series& x, const Series& y) {
    Series r; for (int k = 0; k <= K; ++k) r.c[k] = x.c[k] - y.c[k]; return r;
}
template <int K> inline Series t_neg(const Series& x) {
    Series r; for (int k = 0; k <= K; ++k) r.c[k] = -x.c[k]; return r;
}
// taylor.jl:723-737
template <int K> inline Series t_mul(const Series& x, const Series& y) {
    Series r;
    for (int k = 0; k <= K; ++k) {
        cplx c = x.c[0] * y.c[k];
        for (int j = 1; j <= k; ++j) c = x.c[j] * y.c[k - j] + c;
        r.c[k] = c;
    }
    return r;
}
// taylor.jl:635-656
template <int K> inline Series t_sqr(const Series& x) {
    Series r;
    r.c[0] = op_sqr(x.c[0]);
    for (int k = 1; k <= K; ++k) {
        cplx w = x.c[0] * x.c[k];
        for (int j = 1; j <= (k - 1) / 2; ++j) w = x.c[j] * x.c[k - j] + w;
        if (k % 2 == 0) r.c[k] = 2.0 * w + op_sqr(x.c[k / 2]);
        else r.c[k] = w + w;
    }
    return r;
}
// taylor.jl:705-721
template <int K> inline Series t_div(const Series& x, const Series& y) {
    Series r;
    for (int k = 0; k <= K; ++k) {
        cplx s = x.c[k];
        for (int j = 0; j < k; ++j) s = s - r.c[j] * y.c[k - j];
        r.c[k] = div_fast(s, y.c[0]);
    }
    return r;
}
template <int K> inline Series t_inv(const Series& x) {  // taylor.jl:607-619
    Series one; one.c[0] = cplx(1.0);
    return t_div<K>(one, x);
}
// taylor.jl:751-793: w_k = u0^{-1} (r sum_j j u_j w_{k-j} - sum_j j w_j u_{k-j}) / k;
// a zero constant term yields the zero series (quirk of the reference, kept).
template <int K> inline Series t_pow_int(const Series& x, int r) {
    Series w;
    if (iszero(x.c[0])) return w;
    w.c[0] = op_pow_int(x.c[0], r);
    if (K == 0) return w;
    cplx u0inv = inv_fast(x.c[0]);
    for (int k = 1; k <= K; ++k) {
        cplx s;
        for (int j = 1; j <= k; ++j) s = w.c[k - j] * ((double)j * x.c[j]) + s;
        s = (double)r * s;
        cplx t;
        for (int j = 1; j <= k - 1; ++j) t = x.c[k - j] * ((double)j * w.c[j]) + t;
        w.c[k] = (u0inv * (s - t)) / (double)k;
    }
    return w;
}
template <int K> inline Series t_muladd(const Series& x, const Series& y, const Series& z) {  // :823-838
    Series r;
    for (int k = 0; k <= K; ++k) {
        cplx c = z.c[k];
        for (int j = 0; j <= k; ++j) c = x.c[j] * y.c[k - j] + c;
        r.c[k] = c;
    }
    return r;
}
template <int K> inline Series t_mulsub(const Series& x, const Series& y, const Series& z) {  // :840-862
    Series r;
    for (int k = 0; k <= K; ++k) {
        cplx c = x.c[0] * y.c[k] - z.c[k];
        for (int j = 1; j <= k; ++j) c = x.c[j] * y.c[k - j] + c;
        r.c[k] = c;
    }
    return r;
}
template <int K> inline Series t_submul(const Series& x, const Series& y, const Series& z) {  // :864-878
    Series r;
    for (int k = 0; k <= K; ++k) {
        cplx c = z.c[k];
        for (int j = 0; j <= k; ++j) c = c - x.c[j] * y.c[k - j];
        r.c[k] = c;
    }
    return r;
}

template <int K>
inline void run_taylor_tape(const Program& P, Series* tape) {  // instruction_interpreter.jl:335-442
    const Instr* I = P.instr.data();
    for (;; ++I) {
        const int a1 = I->in[0], a2 = I->in[1], a3 = I->in[2], a4 = I->in[3];
        Series& o = tape[I->out];
        switch (I->op) {
            case OP_STOP: return;
            case OP_CB: o = t_mul<K>(t_sqr<K>(tape[a1]), tape[a1]); break;  // taylor.jl:279-281
            case OP_INV: o = t_inv<K>(tape[a1]); break;
            case OP_INV_NOT_ZERO: o = t_inv<K>(tape[a1]); break;            // taylor.jl:621-623
            case OP_INVSQR: o = t_inv<K>(t_sqr<K>(tape[a1])); break;
            case OP_NEG: o = t_neg<K>(tape[a1]); break;
            case OP_SQR: o = t_sqr<K>(tape[a1]); break;
            case OP_IDENTITY: o = tape[a1]; break;
            case OP_ADD: o = t_add<K>(tape[a1], tape[a2]); break;
            case OP_DIV: o = t_div<K>(tape[a1], tape[a2]); break;
            case OP_MUL: o = t_mul<K>(tape[a1], tape[a2]); break;
            case OP_SUB: o = t_sub<K>(tape[a1], tape[a2]); break;
            case OP_POW_INT: o = t_pow_int<K>(tape[a1], a2); break;
            case OP_ADD3: o = t_add<K>(t_add<K>(tape[a1], tape[a2]), tape[a3]); break;
            case OP_MUL3: o = t_mul<K>(t_mul<K>(tape[a1], tape[a2]), tape[a3]); break;
            case OP_MULADD: o = t_muladd<K>(tape[a1], tape[a2], tape[a3]); break;
            case OP_MULSUB: o = t_mulsub<K>(tape[a1], tape[a2], tape[a3]); break;
            case OP_SUBMUL: o = t_submul<K>(tape[a1], tape[a2], tape[a3]); break;
            case OP_ADD4: o = t_add<K>(t_add<K>(tape[a1], tape[a2]), t_add<K>(tape[a3], tape[a4])); break;
            case OP_MUL4: o = t_mul<K>(t_mul<K>(tape[a1], tape[a2]), t_mul<K>(tape[a3], tape[a4])); break;
            case OP_MULMULADD: o = t_add<K>(t_mul<K>(tape[a1], tape[a2]), t_mul<K>(tape[a3], tape[a4])); break;
            case OP_MULMULSUB: o = t_sub<K>(t_mul<K>(tape[a1], tape[a2]), t_mul<K>(tape[a3], tape[a4])); break;
            default: throw std::runtime_error("oracle: unsupported taylor op");
        }
    }
}

struct TaylorInterp {
    const Program* P = nullptr;
    std::vector<Series> tape;
    void init(const Program* p) {
        P = p;
        tape.assign(p->tape_space + 2, Series());
        for (size_t i = 0; i < p->constants.size(); ++i) tape[i + 1].c[0] = p->constants[i];
    }
    // execute_taylor!(u, Val(K), I, x, t, p): instruction_interpreter.jl:444-491.
    //  x:  nx rows of n (row r = coefficient r of every variable), zero padded (taylor.jl:20-31)
    //  tp: np rows of P, zero padded;  t enters as (t, 1) (interpreted_homotopy.jl:89)
    //  out: (K+1) rows of m: out[k*m + i] = coefficient k of output i
    void execute(int K, cplx* out, const cplx* x, int nx, const cplx* t, const cplx* tp, int np) {
        const Program& p = *P;
        for (int i = 0; i < p.P; ++i) {
            Series s;
            for (int r = 0; r < np && r < TMAX; ++r) s.c[r] = tp[r * p.P + i];
            tape[p.param_off + 1 + i] = s;
        }
        if (p.t_index && t) {
            Series s; s.c[0] = *t; s.c[1] = cplx(1.0);
            tape[p.t_index] = s;
        }
        for (int i = 0; i < p.n; ++i) {
            Series s;
            for (int r = 0; r < nx && r < TMAX; ++r) s.c[r] = x[r * p.n + i];
            tape[p.var_off + 1 + i] = s;
        }
        switch (K) {
            case 1: run_taylor_tape<1>(p, tape.data()); break;
            case 2: run_taylor_tape<2>(p, tape.data()); break;
            case 3: run_taylor_tape<3>(p, tape.data()); break;
            case 4: run_taylor_tape<4>(p, tape.data()); break;
            default: throw std::runtime_error("oracle: taylor order must be 1..4");
        }
        for (int i = 0; i < (K + 1) * p.out_dim; ++i) out[i] = cplx();
        for (auto& a : p.u_assign)
            for (int k = 0; k <= K; ++k) out[k * p.out_dim + a.first - 1] = tape[a.second].c[k];
    }
};

// InterpretedSystem: interpreted_system.jl:23-51
struct System {
    Program eval, jac;  // eval tape; tape of [F; vec(dF/dx)]
    int m() const { return eval.out_dim; }
    int n() const { return eval.n; }
};

struct SystemWS {  // the mutable interpreters of one tracker copy (solve.jl:643-650 deepcopy per thread)
    const System* S = nullptr;
    Interp<cplx> ev, jc;
    Interp<cdd> evdd;
    TaylorInterp ty;
    void init(const System* s) {
        S = s; ev.init(&s->eval); jc.init(&s->jac); evdd.init(&s->eval); ty.init(&s->eval);
    }
    void evaluate(cplx* u, const cplx* x, const cplx* p, const cplx* t = nullptr) { ev.execute(u, nullptr, x, t, p); }
    void evaluate_dd(cplx* u, const cdd* x, const cplx* p, const cplx* t = nullptr) { evdd.execute(u, nullptr, x, t, p); }
    void evaluate_dd_native(cdd* u, const cdd* x, const cplx* p, const cplx* t = nullptr) { evdd.execute_native(u, x, t, p); }
    void evaluate_and_jacobian(cplx* u, cplx* U, const cplx* x, const cplx* p, const cplx* t = nullptr) {
        jc.execute(u, U, x, t, p);
    }
    void taylor(int K, cplx* out, const cplx* x, int nx, const cplx* tp, int np, const cplx* t = nullptr) {
        ty.execute(K, out, x, nx, t, tp, np);
    }
};

}  // namespace orc
