// Device-side ModelKit tape: levelised micro-op format and the three cooperative interpreters
// (ComplexF64, ComplexDF64, truncated Taylor series of order K <= 4).
//
// The host (hc_lower.h) lowers the reference's InstructionSequence into micro-ops of a handful of
// arithmetic classes, puts them into dependency levels and re-allocates the tape slots, so that
// the G lanes that track one path execute the independent instructions of a level together:
// lane l runs instructions l, l + G, ... of the level, then the group synchronises.  The working
// tape of the path (constants, parameters, t, variables, registers) lives in shared memory.
//
// Replaces (reference file:line):
//   src/model_kit/instruction_interpreter.jl:136-192, 252-332  execute!/execute_instructions!
//   src/model_kit/instruction_interpreter.jl:335-491           execute_taylor!
//   src/model_kit/taylor.jl:607-878                            taylor_op_*
//   src/model_kit/operations.jl:5-49, 184-248                  OpType, op_*
#pragma once
#include "hc_coop.h"

namespace hc {

enum Op : int {  // reference OpType, declaration order of src/model_kit/operations.jl:5-49
    OP_STOP = 0, OP_CB, OP_ACOS, OP_ASIN, OP_COS, OP_COSH, OP_EXP, OP_INV, OP_INV_NOT_ZERO,
    OP_INVSQR, OP_NEG, OP_SIN, OP_SINH, OP_SQR, OP_SQRT, OP_TAN, OP_TANH, OP_IDENTITY,
    OP_ADD, OP_DIV, OP_MUL, OP_SUB, OP_POW_INT, OP_POW,
    OP_ADD3, OP_MUL3, OP_MULADD, OP_MULSUB, OP_SUBMUL,
    OP_ADD4, OP_MUL4, OP_MULMULADD, OP_MULMULSUB
};

// micro-op classes: every reference op is a short sequence of these
enum MClass : int {
    MC_MM = 0,    // r = s1 a*b + s2 c*d
    MC_MA = 1,    // r = s1 a*b + s2 c
    MC_M = 2,     // r = s1 a*b
    MC_AA = 3,    // r = s1 a + s2 c
    MC_A = 4,     // r = s1 a
    MC_INV = 5,   // r = 1 / a            (inv_fast)
    MC_DIV = 6,   // r = a / b            (div_fast)
    MC_INVNZ = 7  // r = a == 0 ? a : 1/a (op_inv_not_zero)
};

struct alignas(16) MOp {  // 16 B, 0-based physical tape slots (< 2^16)
    uint32_t w0;  // out | cls << 16 | neg1 << 19 | neg2 << 20 | pair << 21 (independent of the next op)
    uint32_t w1;  // a | b << 16
    uint32_t w2;  // c | d << 16
    uint32_t pad;
};

#if !defined(__CUDACC__)
struct int2 { int x, y; };
#endif

// Thread-per-path fast format: 32-bit slot fields (no unpacking), the class + signs live in the
// segment table.  An MC_MM op occupies two entries (the second one carries d in .a).
struct alignas(16) FOp { uint32_t a, b, c, out; };

struct DevProgram {
    const FOp* fops;         // same order as ops
    const int2* segs;        // runs of ops of one level with the same (class, signs): (key, count); key = cls | n1 << 3 | n2 << 4
    int n_segs, n_fops;
    const MOp* ops;          // sorted by level
    const int* level_end;    // cumulative op count at the end of each level
    int n_levels, n_ops;
    const cx* consts; int C; // tape slots [0, C)
    int param_off, P;        // first parameter slot
    int t_slot;              // -1 = none
    int var_off, n;
    int out_dim, W;          // W = number of tape slots (inputs + registers)
    const int2* u_assign; int nu;  // (i, slot) 0-based
    const int2* U_assign; int nU;  // (j, slot) 0-based, j column-major over (out_dim, n)
};

// ------------------------------------------------------------------ vector views
// S = 0: contiguous behind a generic pointer (the lane group's shared-memory slab + its global
//        scratch, or the host in HC_HOST_SIM).
// S = 2: contiguous in the thread's LOCAL memory (thread-per-path engine, G = 1): the hardware
//        interleaves the lanes of a warp (a warp access to element i is one 512 B segment), L1 keeps
//        local lines write-back, and every access is an explicit ld.local / st.local on a 32-bit
//        window address -- a generic LD/ST costs a descriptor (2 x R2UR) and 64-bit address
//        arithmetic per access on sm_100a.  operator[] therefore returns a proxy, not a reference.
template <class T, int S> struct SV;
template <class T>
struct SV<T, 0> {
    T* p;
    HC_HD T& operator[](int i) const { return p[i]; }
    HC_HD SV at(int off) const { SV r; r.p = p + off; return r; }
    HC_HD void prefetch(int) const {}
    static HC_HD SV make(void* q) { SV r; r.p = (T*)q; return r; }
};

template <class T> struct LRef;  // element of a local-memory vector (device only)
template <>
struct LRef<cx> {
    unsigned a;
    HC_HD operator cx() const {
        cx v = mk(0.0);
#if defined(__CUDA_ARCH__)
        asm volatile("ld.local.v2.f64 {%0, %1}, [%2];" : "=d"(v.re), "=d"(v.im) : "r"(a));
#endif
        return v;
    }
    HC_HD const LRef& operator=(cx v) const {
#if defined(__CUDA_ARCH__)
        asm volatile("st.local.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.re), "d"(v.im));
#endif
        return *this;
    }
    HC_HD const LRef& operator=(const LRef& o) const { return *this = (cx)o; }
};
template <>
struct LRef<double> {
    unsigned a;
    HC_HD operator double() const {
        double v = 0.0;
#if defined(__CUDA_ARCH__)
        asm volatile("ld.local.f64 %0, [%1];" : "=d"(v) : "r"(a));
#endif
        return v;
    }
    HC_HD const LRef& operator=(double v) const {
#if defined(__CUDA_ARCH__)
        asm volatile("st.local.f64 [%0], %1;" ::"r"(a), "d"(v));
#endif
        return *this;
    }
    HC_HD const LRef& operator=(const LRef& o) const { return *this = (double)o; }
};
template <>
struct LRef<int> {
    unsigned a;
    HC_HD operator int() const {
        int v = 0;
#if defined(__CUDA_ARCH__)
        asm volatile("ld.local.s32 %0, [%1];" : "=r"(v) : "r"(a));
#endif
        return v;
    }
    HC_HD const LRef& operator=(int v) const {
#if defined(__CUDA_ARCH__)
        asm volatile("st.local.s32 [%0], %1;" ::"r"(a), "r"(v));
#endif
        return *this;
    }
    HC_HD const LRef& operator=(const LRef& o) const { return *this = (int)o; }
};
template <class T>
struct SV<T, 2> {
    unsigned p;  // byte address in the thread's local window
    HC_HD LRef<T> operator[](int i) const { LRef<T> r; r.a = p + (unsigned)i * (unsigned)sizeof(T); return r; }
    HC_HD SV at(int off) const { SV r; r.p = p + (unsigned)off * (unsigned)sizeof(T); return r; }
    // asks for the L1 line of element i (non-blocking): the lockstep kernels do this at the start of a phase for what the
    // phase is about to read, so that the loads of all warps overlap instead of each paying an L2 round trip in turn
    HC_HD void prefetch(int i) const {
#if defined(__CUDA_ARCH__)
        asm volatile("prefetch.local.L1 [%0];" ::"r"(p + (unsigned)i * (unsigned)sizeof(T)));
#endif
    }
    static HC_HD SV make(void* q) {
        SV r; r.p = 0u;
#if defined(__CUDA_ARCH__)
        if (q) r.p = (unsigned)__cvta_generic_to_local(q);
#endif
        return r;
    }
};

// S = 3: shared memory, the threads of the CTA interleaved (specialised kernels, HC_JIT_BLOCK threads): element i of this
// thread sits at p + i * HC_JIT_BLOCK * sizeof(T), so that a warp access to element i is one contiguous, conflict-free
// 512 B row -- the same picture as local memory, but on the SM by construction, not by cache luck.  Holds the LU factors
// (the most re-read piece of lane state: n^3 / 3 loads per factorization, n^2 per solve).
#if defined(HC_JIT_BLOCK)
template <class T> struct SRef;
template <>
struct SRef<cx> {
    unsigned a;
    HC_HD operator cx() const {
        cx v = mk(0.0);
#if defined(__CUDA_ARCH__)
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.re), "=d"(v.im) : "r"(a));
#endif
        return v;
    }
    HC_HD const SRef& operator=(cx v) const {
#if defined(__CUDA_ARCH__)
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.re), "d"(v.im));
#endif
        return *this;
    }
    HC_HD const SRef& operator=(const SRef& o) const { return *this = (cx)o; }
};
template <class T>
struct SV<T, 3> {
    unsigned p;  // shared-window byte address of this thread's element 0
    HC_HD SRef<T> operator[](int i) const { SRef<T> r; r.a = p + (unsigned)i * (unsigned)(HC_JIT_BLOCK * sizeof(T)); return r; }
    HC_HD SV at(int off) const { SV r; r.p = p + (unsigned)off * (unsigned)(HC_JIT_BLOCK * sizeof(T)); return r; }
    HC_HD void prefetch(int) const {}
};
#endif

template <int S>
struct DV {  // vector of ComplexDF64: element i = cx 2i (re.hi, re.lo) and 2i+1 (im.hi, im.lo)
    SV<cx, S> v;
    HC_HD cdd get(int i) const { cx a = v[2 * i], b = v[2 * i + 1]; return mkcdd(mkdd(a.re, a.im), mkdd(b.re, b.im)); }
    HC_HD void set(int i, cdd x) const { v[2 * i] = mk(x.re.hi, x.re.lo); v[2 * i + 1] = mk(x.im.hi, x.im.lo); }
};

// ------------------------------------------------------------------ scalar micro-ops
HC_HD cx mneg(cx a, bool s) { return s ? -a : a; }
HC_HD cdd mneg(cdd a, bool s) { return s ? -a : a; }
HC_HD cx mfma(cx a, cx b, cx c) { return cfma(a, b, c); }
HC_HD cdd mfma(cdd a, cdd b, cdd c) { return a * b + c; }

template <class TV> HC_HD cx tload(TV tape, int s, cx*) { return tape[s]; }
template <class TV> HC_HD cdd tload(TV tape, int s, cdd*) { cx a = tape[2 * s], b = tape[2 * s + 1]; return mkcdd(mkdd(a.re, a.im), mkdd(b.re, b.im)); }
template <class TV> HC_HD void tstore(TV tape, int s, cx v) { tape[s] = v; }
template <class TV> HC_HD void tstore(TV tape, int s, cdd v) { tape[2 * s] = mk(v.re.hi, v.re.lo); tape[2 * s + 1] = mk(v.im.hi, v.im.lo); }

template <class T, class TV>
HC_HD T eval_mop(const MOp I, TV tape) {
    const int cls = (I.w0 >> 16) & 7;
    const bool n1 = (I.w0 >> 19) & 1, n2 = (I.w0 >> 20) & 1;
    T* tag = nullptr;
    T a = mneg(tload(tape, I.w1 & 0xffff, tag), n1), r;
    switch (cls) {
        case MC_MM: {
            T b = tload(tape, I.w1 >> 16, tag), c = mneg(tload(tape, I.w2 & 0xffff, tag), n2), d = tload(tape, I.w2 >> 16, tag);
            r = mfma(c, d, a * b);
        } break;
        case MC_MA: { T b = tload(tape, I.w1 >> 16, tag), c = mneg(tload(tape, I.w2 & 0xffff, tag), n2); r = mfma(a, b, c); } break;
        case MC_M: { T b = tload(tape, I.w1 >> 16, tag); r = a * b; } break;
        case MC_AA: { T c = mneg(tload(tape, I.w2 & 0xffff, tag), n2); r = a + c; } break;
        case MC_A: r = a; break;
        case MC_INV: r = cinv(a); break;
        case MC_DIV: { T b = tload(tape, I.w1 >> 16, tag); r = cdiv(a, b); } break;
        default: r = ciszero(a) ? a : cinv(a); break;
    }
    return r;
}

// runs the levelised tape; the inputs must already be in the tape and visible (group-synced)
template <class T, int G, class TV>
HC_HDN void run_tape(const DevProgram& P, TV tape, const Grp<G>& g) {
    int beg = 0;
    for (int L = 0; L < P.n_levels; ++L) {
        const int end = P.level_end[L];
        if (G == 1) {  // thread-per-path: sequential program; flagged ops are independent of their successor
            for (int i = beg; i < end; ++i) {
                const MOp I = P.ops[i];
                if (sizeof(T) == sizeof(cx) && ((I.w0 >> 21) & 1)) {  // 2-way ILP: both ops' loads are in flight together
                    const MOp J = P.ops[++i];
                    T r0 = eval_mop<T>(I, tape), r1 = eval_mop<T>(J, tape);
                    tstore(tape, I.w0 & 0xffff, r0);
                    tstore(tape, J.w0 & 0xffff, r1);
                } else tstore(tape, I.w0 & 0xffff, eval_mop<T>(I, tape));
            }
        } else
        // two independent micro-ops per lane and trip: their loads overlap (ops of a level never alias)
        for (int i = beg + g.lane; i < end; i += 2 * G) {
            const MOp I0 = P.ops[i];
            if (i + G < end) {
                const MOp I1 = P.ops[i + G];
                T r0 = eval_mop<T>(I0, tape), r1 = eval_mop<T>(I1, tape);
                tstore(tape, I0.w0 & 0xffff, r0);
                tstore(tape, I1.w0 & 0xffff, r1);
            } else tstore(tape, I0.w0 & 0xffff, eval_mop<T>(I0, tape));
        }
        g.sync();
        beg = end;
    }
}

// ------------------------------------------------------------------ program cursors
// The thread-per-path kernel (S = 2) always runs from programs staged in shared memory and reads
// them with ld.shared on a 32-bit address; everything else walks a generic pointer.
template <class E, bool SHARED>
struct PCur {
    const E* p;
    HC_HD explicit PCur(const E* q) : p(q) {}
    HC_HD E get(int i) const { return p[i]; }
    HC_HD void adv(int k) { p += k; }
};
#if defined(__CUDA_ARCH__)
template <>
struct PCur<FOp, true> {
    unsigned a;
    HC_D explicit PCur(const FOp* q) : a((unsigned)__cvta_generic_to_shared(q)) {}
    HC_D FOp get(int i) const {
        FOp v;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.a), "=r"(v.b), "=r"(v.c), "=r"(v.out) : "r"(a + 16u * (unsigned)i));
        return v;
    }
    HC_D void adv(int k) { a += 16u * (unsigned)k; }
};
template <>
struct PCur<int2, true> {
    unsigned a;
    HC_D explicit PCur(const int2* q) : a((unsigned)__cvta_generic_to_shared(q)) {}
    HC_D int2 get(int i) const {
        int2 v;
        asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a + 8u * (unsigned)i));
        return v;
    }
    HC_D void adv(int k) { a += 8u * (unsigned)k; }
};
#endif
// Read-only program data (constants, output maps, homotopy parameters).  The thread-per-path kernel (S = 2)
// always runs with them staged in shared memory: ld.shared on a 32-bit address instead of a generic LD.
template <int S> HC_HD cx pld(const cx* p) {
#if defined(__CUDA_ARCH__)
    if (S == 2) {
        cx v;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.re), "=d"(v.im) : "r"((unsigned)__cvta_generic_to_shared(p)));
        return v;
    }
#endif
    return *p;
}
template <int S> HC_HD int2 pld(const int2* p) {
#if defined(__CUDA_ARCH__)
    if (S == 2) {
        int2 v;
        asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
        return v;
    }
#endif
    return *p;
}
template <class TV> struct ProgInShared { static constexpr bool value = false; };
template <> struct ProgInShared<SV<cx, 2>> { static constexpr bool value = true; };

// ------------------------------------------------------------------ segmented interpreter (G == 1)
// One dispatch per segment instead of one per micro-op: the loop of a segment knows class and
// signs at compile time (negations fold into the DFMA operand modifiers), reads the slot numbers
// straight out of the 16-byte FOp and handles four (then two) independent ops per trip with all loads issued
// before the first store (ops of a level never alias).  Results are bit-identical to eval_mop.
#ifndef HC_SEG_UNROLL4
#define HC_SEG_UNROLL4 0  // measured: the x4 trip costs more in instruction-cache misses than it gains in overlap (-4 %)
#endif
#define HC_KEY(cls, n1, n2) ((cls) | ((n1) << 3) | ((n2) << 4))

template <int CLS, bool N1, bool N2>
HC_HD cx fop_eval(cx a, cx b, cx c, cx d) {
    a = mneg(a, N1);
    if (CLS == MC_MM) return mfma(mneg(c, N2), d, a * b);
    if (CLS == MC_MA) return mfma(a, b, mneg(c, N2));
    if (CLS == MC_M) return a * b;
    if (CLS == MC_AA) return a + mneg(c, N2);
    if (CLS == MC_A) return a;
    if (CLS == MC_INV) return cinv(a);
    if (CLS == MC_DIV) return cdiv(a, b);
    return ciszero(a) ? a : cinv(a);
}
template <int CLS, bool N1, bool N2, class OC, class TV>
HC_HD void run_segment(OC& op, int cnt, TV tape) {
    constexpr bool useB = CLS == MC_MM || CLS == MC_MA || CLS == MC_M || CLS == MC_DIV;
    constexpr bool useC = CLS == MC_MM || CLS == MC_MA || CLS == MC_AA;
    constexpr int STEP = CLS == MC_MM ? 2 : 1;
    const cx z = mk(0.0);
#if HC_SEG_UNROLL4
    for (; cnt >= 4; cnt -= 4, op.adv(4 * STEP)) {
        const FOp I0 = op.get(0), I1 = op.get(STEP), I2 = op.get(2 * STEP), I3 = op.get(3 * STEP);
        const cx a0 = tape[I0.a], a1 = tape[I1.a], a2 = tape[I2.a], a3 = tape[I3.a];
        const cx b0 = useB ? tape[I0.b] : z, b1 = useB ? tape[I1.b] : z, b2 = useB ? tape[I2.b] : z, b3 = useB ? tape[I3.b] : z;
        const cx c0 = useC ? tape[I0.c] : z, c1 = useC ? tape[I1.c] : z, c2 = useC ? tape[I2.c] : z, c3 = useC ? tape[I3.c] : z;
        const cx d0 = CLS == MC_MM ? tape[op.get(1).a] : z, d1 = CLS == MC_MM ? tape[op.get(STEP + 1).a] : z;
        const cx d2 = CLS == MC_MM ? tape[op.get(2 * STEP + 1).a] : z, d3 = CLS == MC_MM ? tape[op.get(3 * STEP + 1).a] : z;
        const cx r0 = fop_eval<CLS, N1, N2>(a0, b0, c0, d0), r1 = fop_eval<CLS, N1, N2>(a1, b1, c1, d1);
        const cx r2 = fop_eval<CLS, N1, N2>(a2, b2, c2, d2), r3 = fop_eval<CLS, N1, N2>(a3, b3, c3, d3);
        tape[I0.out] = r0;
        tape[I1.out] = r1;
        tape[I2.out] = r2;
        tape[I3.out] = r3;
    }
    if (cnt >= 2) {
#else
    for (; cnt >= 2;) {
#endif
        const FOp I0 = op.get(0), I1 = op.get(STEP);
        const cx a0 = tape[I0.a], a1 = tape[I1.a];
        const cx b0 = useB ? tape[I0.b] : z, b1 = useB ? tape[I1.b] : z;
        const cx c0 = useC ? tape[I0.c] : z, c1 = useC ? tape[I1.c] : z;
        const cx d0 = CLS == MC_MM ? tape[op.get(1).a] : z, d1 = CLS == MC_MM ? tape[op.get(STEP + 1).a] : z;
        const cx r0 = fop_eval<CLS, N1, N2>(a0, b0, c0, d0), r1 = fop_eval<CLS, N1, N2>(a1, b1, c1, d1);
        tape[I0.out] = r0;
        tape[I1.out] = r1;
        cnt -= 2; op.adv(2 * STEP);
    }
    if (cnt) {
        const FOp I0 = op.get(0);
        const cx a0 = tape[I0.a];
        const cx b0 = useB ? tape[I0.b] : z, c0 = useC ? tape[I0.c] : z, d0 = CLS == MC_MM ? tape[op.get(1).a] : z;
        tape[I0.out] = fop_eval<CLS, N1, N2>(a0, b0, c0, d0);
        op.adv(STEP);
    }
}
#define HC_SEG_CASES(RUN)                                                                           \
    case HC_KEY(MC_MM, 0, 0): RUN(MC_MM, false, false); break;                                       \
    case HC_KEY(MC_MM, 0, 1): RUN(MC_MM, false, true); break;                                        \
    case HC_KEY(MC_MA, 0, 0): RUN(MC_MA, false, false); break;                                       \
    case HC_KEY(MC_MA, 0, 1): RUN(MC_MA, false, true); break;                                        \
    case HC_KEY(MC_MA, 1, 0): RUN(MC_MA, true, false); break;                                        \
    case HC_KEY(MC_M, 0, 0): RUN(MC_M, false, false); break;                                         \
    case HC_KEY(MC_AA, 0, 0): RUN(MC_AA, false, false); break;                                       \
    case HC_KEY(MC_AA, 0, 1): RUN(MC_AA, false, true); break;                                        \
    case HC_KEY(MC_A, 0, 0): RUN(MC_A, false, false); break;                                         \
    case HC_KEY(MC_A, 1, 0): RUN(MC_A, true, false); break;                                          \
    case HC_KEY(MC_INV, 0, 0): RUN(MC_INV, false, false); break;                                     \
    case HC_KEY(MC_DIV, 0, 0): RUN(MC_DIV, false, false); break;                                     \
    default: RUN(MC_INVNZ, false, false); break;  // the lowering emits no other (class, sign) pair

template <class TV>
HC_HDN void run_tape_seg(const FOp* fops, const int2* segs, int n_segs, TV tape) {
    PCur<FOp, ProgInShared<TV>::value> op(fops);
    PCur<int2, ProgInShared<TV>::value> sc(segs);
    for (int s = 0; s < n_segs; ++s) {
        const int2 sg = sc.get(s);
        switch (sg.x) {
#define HC_RUN_(CLS, N1, N2) run_segment<CLS, N1, N2>(op, sg.y, tape)
            HC_SEG_CASES(HC_RUN_)
#undef HC_RUN_
        }
    }
}

// ------------------------------------------------------------------ Taylor micro-ops
// Slot stride of a series tape: orders 1..3 (the tracker's three predictor passes) share the stride 4, so
// that the input block written by the first pass (constants, parameter series at t) serves all three.
#define HC_TS(K) ((K) <= 3 ? 4 : (K) + 1)

template <int K>
struct Ser { cx c[K + 1]; };

template <int K, class TV> HC_HD Ser<K> ser_load(TV tape, int s, bool neg) {
    Ser<K> r;
#pragma unroll
    for (int k = 0; k <= K; ++k) r.c[k] = mneg(tape[s * HC_TS(K) + k], neg);
    return r;
}
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
template <int K> HC_HD Ser<K> t_div(const Ser<K>& x, const Ser<K>& y) {  // taylor.jl:689-716
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
template <int K> HC_HD Ser<K> t_one() {
    Ser<K> r;
#pragma unroll
    for (int k = 0; k <= K; ++k) r.c[k] = mk(k == 0 ? 1.0 : 0.0);
    return r;
}

template <int K, class TV>
HC_HD void exec_mop_taylor(const MOp I, TV tape) {
    const int out = I.w0 & 0xffff, cls = (I.w0 >> 16) & 7;
    const bool n1 = (I.w0 >> 19) & 1, n2 = (I.w0 >> 20) & 1;
    Ser<K> a = ser_load<K>(tape, I.w1 & 0xffff, n1), r;
    switch (cls) {
        case MC_MM: {
            Ser<K> b = ser_load<K>(tape, I.w1 >> 16, false);
            r = t_mul<K>(a, b);
            Ser<K> c = ser_load<K>(tape, I.w2 & 0xffff, n2), d = ser_load<K>(tape, I.w2 >> 16, false);
            r = t_muladd<K>(c, d, r);
        } break;
        case MC_MA: {
            Ser<K> b = ser_load<K>(tape, I.w1 >> 16, false), c = ser_load<K>(tape, I.w2 & 0xffff, n2);
            r = t_muladd<K>(a, b, c);
        } break;
        case MC_M: { Ser<K> b = ser_load<K>(tape, I.w1 >> 16, false); r = t_mul<K>(a, b); } break;
        case MC_AA: {
            Ser<K> c = ser_load<K>(tape, I.w2 & 0xffff, n2);
#pragma unroll
            for (int k = 0; k <= K; ++k) r.c[k] = a.c[k] + c.c[k];
        } break;
        case MC_A: r = a; break;
        case MC_DIV: { Ser<K> b = ser_load<K>(tape, I.w1 >> 16, false); r = t_div<K>(a, b); } break;
        default: r = t_div<K>(t_one<K>(), a); break;  // MC_INV, MC_INVNZ (taylor.jl: inv and inv_not_zero share the rule)
    }
#pragma unroll
    for (int k = 0; k <= K; ++k) tape[out * HC_TS(K) + k] = r.c[k];
}

template <int K, int G, class TV>
HC_HDN void run_taylor_tape(const DevProgram& P, TV tape, const Grp<G>& g) {
    int beg = 0;
    for (int L = 0; L < P.n_levels; ++L) {
        const int end = P.level_end[L];
        for (int ib = beg, i = beg + g.lane; ib < end; ib += G, i += G) { if (i < end) exec_mop_taylor<K>(P.ops[i], tape); }  // uniform trips, guarded body
        g.sync();
        beg = end;
    }
}

// segmented Taylor interpreter (G == 1): class and signs are compile-time per segment
template <int K, int CLS, bool N1, bool N2, class OC, class TV>
HC_HD void run_taylor_segment(OC& op, int cnt, TV tape) {
    constexpr int STEP = CLS == MC_MM ? 2 : 1;
    for (; cnt > 0; --cnt, op.adv(STEP)) {
        const FOp I = op.get(0);
        Ser<K> a = ser_load<K>(tape, I.a, N1), r;
        if (CLS == MC_MM) {
            Ser<K> b = ser_load<K>(tape, I.b, false);
            r = t_mul<K>(a, b);
            Ser<K> c = ser_load<K>(tape, I.c, N2), d = ser_load<K>(tape, op.get(1).a, false);
            r = t_muladd<K>(c, d, r);
        } else if (CLS == MC_MA) {
            Ser<K> b = ser_load<K>(tape, I.b, false), c = ser_load<K>(tape, I.c, N2);
            r = t_muladd<K>(a, b, c);
        } else if (CLS == MC_M) {
            Ser<K> b = ser_load<K>(tape, I.b, false);
            r = t_mul<K>(a, b);
        } else if (CLS == MC_AA) {
            Ser<K> c = ser_load<K>(tape, I.c, N2);
#pragma unroll
            for (int k = 0; k <= K; ++k) r.c[k] = a.c[k] + c.c[k];
        } else if (CLS == MC_A) r = a;
        else if (CLS == MC_DIV) { Ser<K> b = ser_load<K>(tape, I.b, false); r = t_div<K>(a, b); }
        else r = t_div<K>(t_one<K>(), a);
#pragma unroll
        for (int k = 0; k <= K; ++k) tape[I.out * HC_TS(K) + k] = r.c[k];
    }
}
template <int K, class TV>
HC_HDN void run_taylor_tape_seg(const FOp* fops, const int2* segs, int n_segs, TV tape) {
    PCur<FOp, ProgInShared<TV>::value> op(fops);
    PCur<int2, ProgInShared<TV>::value> sc(segs);
    for (int s = 0; s < n_segs; ++s) {
        const int2 sg = sc.get(s);
        switch (sg.x) {
#define HC_RUN_(CLS, N1, N2) run_taylor_segment<K, CLS, N1, N2>(op, sg.y, tape)
            HC_SEG_CASES(HC_RUN_)
#undef HC_RUN_
        }
    }
}

}  // namespace hc
