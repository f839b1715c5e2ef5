// Device-side ModelKit tape: packed instruction format and the three interpreters
// (ComplexF64, ComplexDF64, truncated Taylor series of order K <= 4), one path per thread.
//
// The tape is uniform across a batch, so the 32 lanes of a warp decode the same instruction
// (broadcast read) and execute it on 32 different paths in lock-step; each lane's working tape
// lives in a lane-interleaved slab (element i of lane l at base[i * stride + l]) so that every
// warp-wide tape access is one coalesced 512 B transaction.
//
// Replaces (reference file:line):
//   src/model_kit/instruction_interpreter.jl:136-192, 252-332  execute!/execute_instructions!
//   src/model_kit/instruction_interpreter.jl:335-491           execute_taylor!
//   src/model_kit/taylor.jl:607-878                            taylor_op_*
//   src/model_kit/operations.jl:5-49, 184-248                  OpType, op_*
#pragma once
#include "hc_common.h"

namespace hc {

enum Op : int {
    OP_STOP = 0, OP_CB, OP_ACOS, OP_ASIN, OP_COS, OP_COSH, OP_EXP, OP_INV, OP_INV_NOT_ZERO,
    OP_INVSQR, OP_NEG, OP_SIN, OP_SINH, OP_SQR, OP_SQRT, OP_TAN, OP_TANH, OP_IDENTITY,
    OP_ADD, OP_DIV, OP_MUL, OP_SUB, OP_POW_INT, OP_POW,
    OP_ADD3, OP_MUL3, OP_MULADD, OP_MULSUB, OP_SUBMUL,
    OP_ADD4, OP_MUL4, OP_MULMULADD, OP_MULMULSUB
};

// lane-interleaved view
template <class T>
struct SV {
    T* p;
    int s;
    // 32-bit index arithmetic (rows * lanes < 2^31) and a global-address-space hint: without them
    // every access costs a 64-bit multiply and a generic LD/ST (ncu r01: IMAD+LEA+R2UR = 47 % of issue)
    HC_HD T& operator[](int i) const {
#if defined(__CUDA_ARCH__)
        __builtin_assume(__isGlobal(p));
#endif
        return p[(unsigned)i * (unsigned)s];
    }
    HC_HD SV<T> at(int off) const { SV<T> r; r.p = p + (unsigned)off * (unsigned)s; r.s = s; return r; }
};
using CV = SV<cx>;
using RV = SV<double>;
using IV = SV<int>;

// Lane-interleaved vector of ComplexDF64: element i = cx rows 2i = (re.hi, re.lo) and 2i+1 =
// (im.hi, im.lo).  It MUST use the same 16-byte lane interleaving as CV: the tape region is viewed
// both ways, and lanes of a warp can be in different precisions at the same time.
struct DV {
    cx* p;
    int s;
    HC_HD cdd get(int i) const {
#if defined(__CUDA_ARCH__)
        __builtin_assume(__isGlobal(p));
#endif
        cx a = p[(unsigned)(2 * i) * (unsigned)s], b = p[(unsigned)(2 * i + 1) * (unsigned)s];
        return mkcdd(mkdd(a.re, a.im), mkdd(b.re, b.im));
    }
    HC_HD void set(int i, cdd v) const {
#if defined(__CUDA_ARCH__)
        __builtin_assume(__isGlobal(p));
#endif
        p[(unsigned)(2 * i) * (unsigned)s] = mk(v.re.hi, v.re.lo);
        p[(unsigned)(2 * i + 1) * (unsigned)s] = mk(v.im.hi, v.im.lo);
    }
};
HC_HD cx tget(CV t, int i) { return t[i]; }
HC_HD void tset(CV t, int i, cx v) { t[i] = v; }
HC_HD cdd tget(DV t, int i) { return t.get(i); }
HC_HD void tset(DV t, int i, cdd v) { t.set(i, v); }

#if !defined(__CUDACC__)
struct int2 { int x, y; };
#endif

struct alignas(16) PInstr {  // 16 B packed instruction, 0-based slots
    uint32_t w0;  // op | out << 8   (out < 2^24)
    uint32_t w1;  // in0 | in1 << 16 (slots < 2^16)
    uint32_t w2;  // in2 | in3 << 16
    int32_t lit;  // literal exponent of OP_POW_INT
};

struct DevProgram {
    const PInstr* instr;
    const cx* consts;   // C constants; slots [0, C) are read from here, never from the lane tape
    int C;
    int param_off, P;   // 0-based first parameter slot
    int t_slot;         // -1 = none
    int var_off, n;
    int out_dim, W;     // W = tape_space
    const int2* u_assign; int nu;  // (i, slot) 0-based
    const int2* U_assign; int nU;  // (j, slot) 0-based, j column-major over (out_dim, n)
};

// Tape metadata is read with plain (generic) loads: the kernel stages it into shared memory when
// it fits, otherwise it stays in global memory (L1-cached, warp-uniform broadcast either way).
HC_HD PInstr ld_instr(const PInstr* p) { return *p; }
HC_HD cx ld_const(const cx* p) { return *p; }
HC_HD double ld_real(const double* p) { return *p; }
HC_HD int2 ld_i2(const int2* p) { return *p; }

// ------------------------------------------------------------------ scalar tapes
// TapeT: CV (ComplexF64) or DV (ComplexDF64); slot s >= C lives at tape element s - C.
HC_HD cx fetch(const DevProgram& P, CV tape, int s) { return s < P.C ? ld_const(P.consts + s) : tape[s - P.C]; }
HC_HD cdd fetch(const DevProgram& P, DV tape, int s) { return s < P.C ? tocdd(ld_const(P.consts + s)) : tape.get(s - P.C); }

template <class T, class TapeT>
HC_HDN void run_tape(const DevProgram& P, TapeT tape) {
    const PInstr* ip = P.instr;
    for (;; ++ip) {
        PInstr I = ld_instr(ip);
        const int op = I.w0 & 0xff, out = (int)(I.w0 >> 8);
        const int a1 = I.w1 & 0xffff, a2 = I.w1 >> 16, a3 = I.w2 & 0xffff, a4 = I.w2 >> 16;
        if (op == OP_STOP) return;
        T r;
        T x = fetch(P, tape, a1);
        switch (op) {
            case OP_CB: r = ccb(x); break;
            case OP_INV: r = cinv(x); break;
            case OP_INV_NOT_ZERO: r = ciszero(x) ? x : cinv(x); break;
            case OP_INVSQR: r = csqr(cinv(x)); break;
            case OP_NEG: r = -x; break;
            case OP_SQR: r = csqr(x); break;
            case OP_IDENTITY: r = x; break;
            case OP_POW_INT: r = cpowi(x, I.lit); break;
            default: {
                T y = fetch(P, tape, a2);
                switch (op) {
                    case OP_ADD: r = x + y; break;
                    case OP_DIV: r = cdiv(x, y); break;
                    case OP_MUL: r = x * y; break;
                    case OP_SUB: r = x - y; break;
                    default: {
                        T z = fetch(P, tape, a3);
                        switch (op) {
                            case OP_ADD3: r = x + y + z; break;
                            case OP_MUL3: r = x * y * z; break;
                            case OP_MULADD: r = x * y + z; break;
                            case OP_MULSUB: r = x * y - z; break;
                            case OP_SUBMUL: r = z - x * y; break;
                            default: {
                                T w = fetch(P, tape, a4);
                                switch (op) {
                                    case OP_ADD4: r = x + y + z + w; break;
                                    case OP_MUL4: r = (x * y) * (z * w); break;
                                    case OP_MULMULADD: r = x * y + z * w; break;
                                    default: r = x * y - z * w; break;  // OP_MULMULSUB
                                }
                            }
                        }
                    }
                }
            }
        }
        tset(tape, out - P.C, r);
    }
}

// ------------------------------------------------------------------ Taylor tapes
template <int K>
struct Ser { cx c[K + 1]; };

template <int K>
HC_HD Ser<K> fetch_ser(const DevProgram& P, CV tape, int s) {
    Ser<K> r;
    if (s < P.C) {
        r.c[0] = ld_const(P.consts + s);
#pragma unroll
        for (int k = 1; k <= K; ++k) r.c[k] = mk(0.0);
    } else {
        const int b = (s - P.C) * (K + 1);
#pragma unroll
        for (int k = 0; k <= K; ++k) r.c[k] = tape[b + k];
    }
    return r;
}
template <int K> HC_HD Ser<K> t_add(const Ser<K>& x, const Ser<K>& y) { Ser<K> r;
#pragma unroll
    for (int k = 0; k <= K; ++k) r.c[k] = x.c[k] + y.c[k]; return r; }
template <int K> HC_HD Ser<K> t_sub(const Ser<K>& x, const Ser<K>& y) { Ser<K> r;
#pragma unroll
    for (int k = 0; k <= K; ++k) r.c[k] = x.c[k] - y.c[k]; return r; }
template <int K> HC_HD Ser<K> t_neg(const Ser<K>& x) { Ser<K> r;
#pragma unroll
    for (int k = 0; k <= K; ++k) r.c[k] = -x.c[k]; return r; }
template <int K> HC_HD Ser<K> t_mul(const Ser<K>& x, const Ser<K>& y) {
    Ser<K> r;
#pragma unroll
    for (int k = 0; k <= K; ++k) {
        cx c = x.c[0] * y.c[k];
#pragma unroll
        for (int j = 1; j <= k; ++j) c = cfma(x.c[j], y.c[k - j], c);
        r.c[k] = c;
    }
    return r;
}
template <int K> HC_HD Ser<K> t_muladd(const Ser<K>& x, const Ser<K>& y, const Ser<K>& z) {
    Ser<K> r;
#pragma unroll
    for (int k = 0; k <= K; ++k) {
        cx c = z.c[k];
#pragma unroll
        for (int j = 0; j <= k; ++j) c = cfma(x.c[j], y.c[k - j], c);
        r.c[k] = c;
    }
    return r;
}
template <int K> HC_HD Ser<K> t_submul(const Ser<K>& x, const Ser<K>& y, const Ser<K>& z) {
    Ser<K> r;
#pragma unroll
    for (int k = 0; k <= K; ++k) {
        cx c = z.c[k];
#pragma unroll
        for (int j = 0; j <= k; ++j) c = cfnma(x.c[j], y.c[k - j], c);
        r.c[k] = c;
    }
    return r;
}
template <int K> HC_HD Ser<K> t_sqr(const Ser<K>& x) {
    Ser<K> r;
    r.c[0] = csqr(x.c[0]);
#pragma unroll
    for (int k = 1; k <= K; ++k) {
        cx w = x.c[0] * x.c[k];
#pragma unroll
        for (int j = 1; j <= (k - 1) / 2; ++j) w = cfma(x.c[j], x.c[k - j], w);
        r.c[k] = (k % 2 == 0) ? (2.0 * w + csqr(x.c[k / 2])) : (w + w);
    }
    return r;
}
template <int K> HC_HD Ser<K> t_div(const Ser<K>& x, const Ser<K>& y) {
    Ser<K> r;
    cx yinv = cinv(y.c[0]);
#pragma unroll
    for (int k = 0; k <= K; ++k) {
        cx s = x.c[k];
#pragma unroll
        for (int j = 0; j < k; ++j) s = cfnma(r.c[j], y.c[k - j], s);
        r.c[k] = s * yinv;
    }
    return r;
}
template <int K> HC_HD Ser<K> t_one() { Ser<K> r;
#pragma unroll
    for (int k = 0; k <= K; ++k) r.c[k] = mk(k == 0 ? 1.0 : 0.0); return r; }
// reference taylor.jl:751-793 (a zero constant term gives the zero series, as in the reference)
template <int K> HC_HD Ser<K> t_powi(const Ser<K>& x, int rexp) {
    Ser<K> w;
#pragma unroll
    for (int k = 0; k <= K; ++k) w.c[k] = mk(0.0);
    if (ciszero(x.c[0])) return w;
    w.c[0] = cpowi(x.c[0], rexp);
    cx u0inv = cinv(x.c[0]);
#pragma unroll
    for (int k = 1; k <= K; ++k) {
        cx s = mk(0.0), t = mk(0.0);
#pragma unroll
        for (int j = 1; j <= k; ++j) s = cfma(w.c[k - j], (double)j * x.c[j], s);
        s = (double)rexp * s;
#pragma unroll
        for (int j = 1; j <= k - 1; ++j) t = cfma(x.c[k - j], (double)j * w.c[j], t);
        w.c[k] = (u0inv * (s - t)) / (double)k;
    }
    return w;
}

template <int K>
HC_HDN void run_taylor_tape(const DevProgram& P, CV tape) {
    const PInstr* ip = P.instr;
    for (;; ++ip) {
        PInstr I = ld_instr(ip);
        const int op = I.w0 & 0xff, out = (int)(I.w0 >> 8);
        const int a1 = I.w1 & 0xffff, a2 = I.w1 >> 16, a3 = I.w2 & 0xffff, a4 = I.w2 >> 16;
        if (op == OP_STOP) return;
        Ser<K> r;
        Ser<K> x = fetch_ser<K>(P, tape, a1);
        switch (op) {
            case OP_CB: r = t_mul<K>(t_sqr<K>(x), x); break;
            case OP_INV: case OP_INV_NOT_ZERO: r = t_div<K>(t_one<K>(), x); break;
            case OP_INVSQR: r = t_div<K>(t_one<K>(), t_sqr<K>(x)); break;
            case OP_NEG: r = t_neg<K>(x); break;
            case OP_SQR: r = t_sqr<K>(x); break;
            case OP_IDENTITY: r = x; break;
            case OP_POW_INT: r = t_powi<K>(x, I.lit); break;
            default: {
                Ser<K> y = fetch_ser<K>(P, tape, a2);
                switch (op) {
                    case OP_ADD: r = t_add<K>(x, y); break;
                    case OP_DIV: r = t_div<K>(x, y); break;
                    case OP_MUL: r = t_mul<K>(x, y); break;
                    case OP_SUB: r = t_sub<K>(x, y); break;
                    default: {
                        Ser<K> z = fetch_ser<K>(P, tape, a3);
                        switch (op) {
                            case OP_ADD3: r = t_add<K>(t_add<K>(x, y), z); break;
                            case OP_MUL3: r = t_mul<K>(t_mul<K>(x, y), z); break;
                            case OP_MULADD: r = t_muladd<K>(x, y, z); break;
                            case OP_MULSUB: r = t_sub<K>(t_mul<K>(x, y), z); break;
                            case OP_SUBMUL: r = t_submul<K>(x, y, z); break;
                            default: {
                                Ser<K> w = fetch_ser<K>(P, tape, a4);
                                switch (op) {
                                    case OP_ADD4: r = t_add<K>(t_add<K>(x, y), t_add<K>(z, w)); break;
                                    case OP_MUL4: r = t_mul<K>(t_mul<K>(x, y), t_mul<K>(z, w)); break;
                                    case OP_MULMULADD: r = t_muladd<K>(x, y, t_mul<K>(z, w)); break;
                                    default: r = t_sub<K>(t_mul<K>(x, y), t_mul<K>(z, w)); break;
                                }
                            }
                        }
                    }
                }
            }
        }
        const int b = (out - P.C) * (K + 1);
#pragma unroll
        for (int k = 0; k <= K; ++k) tape[b + k] = r.c[k];
    }
}

}  // namespace hc
