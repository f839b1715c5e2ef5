// Host side: lowers a reference InstructionSequence (src/model_kit/instruction_sequence.jl:1-27,
// 145-254; 24-byte 1-based Instruction records handed over unchanged by the Julia host) into the
// levelised micro-op program the device interpreters run (hc_tape.h).
//
//   1. every reference op becomes 1-3 micro-ops in SSA form (the reference's compacted register
//      slots are renamed, so only true data dependencies remain),
//   2. micro-ops are put into dependency levels and sorted by arithmetic class inside a level
//      (lanes of a group that share a round then mostly share the code path),
//      (cap <= 0: the program stays sequential, for the thread-per-path engine)
//   3. tape slots are re-allocated by a linear scan over the levels: a slot is reused as soon as
//      its last reader sits in an earlier level, which keeps the per-path shared-memory tape as
//      small as the reference's own compaction pass does (instruction_sequence.jl:383-450).
#pragma once
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/hc_b200.h"
#include "hc_tape.h"

namespace hc {

struct LoweredProgram {
    std::vector<MOp> ops;
    std::vector<FOp> fops;        // thread-per-path fast format (hc_tape.h), same order as ops
    std::vector<int2> segs;
    std::vector<int> level_end;
    std::vector<cx> consts;
    int param_off = 0, P = 0, t_slot = -1, var_off = 0, n = 0, out_dim = 0, W = 0;
    std::vector<int2> u_assign, U_assign;
    int max_width = 0;
};

inline bool supported_op(int op) {
    switch (op) {
        case OP_STOP: case OP_CB: case OP_INV: case OP_INV_NOT_ZERO: case OP_INVSQR: case OP_NEG: case OP_SQR:
        case OP_IDENTITY: case OP_ADD: case OP_DIV: case OP_MUL: case OP_SUB: case OP_POW_INT: case OP_ADD3:
        case OP_MUL3: case OP_MULADD: case OP_MULSUB: case OP_SUBMUL: case OP_ADD4: case OP_MUL4:
        case OP_MULMULADD: case OP_MULMULSUB: return true;
        default: return false;
    }
}

inline LoweredProgram lower_program(const hc_program_desc* d, int cap, bool prio_height = true, bool pair = false, int seg_window = 0) {
    struct VOp { int cls, n1, n2, a, b, c, d, out, level; };
    LoweredProgram R;
    const int C0 = d->n_constants, P = d->n_params, n = d->n_vars;
    const bool has_t = d->t_index > 0;
    if (d->param_offset < C0 || d->var_offset < C0) throw std::string("inputs overlap the constants block");
    for (int i = 0; i < C0; ++i) R.consts.push_back(mk(d->constants[2 * i], d->constants[2 * i + 1]));
    int one_slot = -1;
    for (int i = 0; i < C0; ++i) if (R.consts[i].re == 1.0 && R.consts[i].im == 0.0) one_slot = i;
    // a literal x^0 needs the constant 1
    for (int i = 0; i < d->n_instructions; ++i) {
        const int32_t* s = d->instructions + 6 * (size_t)i;
        if (s[4] == OP_STOP) break;
        if (s[4] == OP_POW_INT && s[1] == 0 && one_slot < 0) { one_slot = (int)R.consts.size(); R.consts.push_back(mk(1.0)); }
    }
    const int C = (int)R.consts.size();
    R.param_off = C; R.P = P; R.t_slot = has_t ? C + P : -1; R.var_off = C + P + (has_t ? 1 : 0); R.n = n;
    R.out_dim = d->out_dim;
    const int Nin = R.var_off + n;

    // reference slot (0-based) -> current virtual id; inputs are their own physical slots
    std::vector<int> cur((size_t)d->tape_space, -1);
    for (int i = 0; i < C0; ++i) cur[i] = i;
    for (int i = 0; i < P; ++i) cur[d->param_offset + i] = R.param_off + i;
    if (has_t) cur[d->t_index - 1] = R.t_slot;
    for (int i = 0; i < n; ++i) cur[d->var_offset + i] = R.var_off + i;
    std::vector<char> is_input((size_t)d->tape_space, 0);
    for (int i = 0; i < (int)cur.size(); ++i) is_input[i] = cur[i] >= 0;

    std::vector<VOp> v;
    std::vector<int> level((size_t)Nin, 0);  // level of every virtual id
    auto emit = [&](int cls, int n1, int n2, int a, int b, int c, int dd) {
        VOp o{cls, n1, n2, a, b, c, dd, (int)level.size(), 0};
        int l = level[a];
        if (b >= 0) l = std::max(l, level[b]);
        if (c >= 0) l = std::max(l, level[c]);
        if (dd >= 0) l = std::max(l, level[dd]);
        o.level = l + 1;
        level.push_back(o.level);
        v.push_back(o);
        return o.out;
    };
    bool stopped = false;
    for (int i = 0; i < d->n_instructions && !stopped; ++i) {
        const int32_t* s = d->instructions + 6 * (size_t)i;
        const int op = s[4];
        if (!supported_op(op)) throw std::string("unsupported op in tape: ") + std::to_string(op);
        if (op == OP_STOP) { stopped = true; break; }
        auto rd = [&](int k) {
            int slot = s[k];
            if (slot < 1 || slot > d->tape_space) throw std::string("tape index out of range");
            if (cur[slot - 1] < 0) throw std::string("tape reads a slot before it is written");
            return cur[slot - 1];
        };
        const int x = rd(0);
        int r;
        switch (op) {
            case OP_CB: r = emit(MC_M, 0, 0, emit(MC_M, 0, 0, x, x, -1, -1), x, -1, -1); break;
            case OP_INV: r = emit(MC_INV, 0, 0, x, -1, -1, -1); break;
            case OP_INV_NOT_ZERO: r = emit(MC_INVNZ, 0, 0, x, -1, -1, -1); break;
            case OP_INVSQR: { int t = emit(MC_INV, 0, 0, x, -1, -1, -1); r = emit(MC_M, 0, 0, t, t, -1, -1); } break;
            case OP_NEG: r = emit(MC_A, 1, 0, x, -1, -1, -1); break;
            case OP_SQR: r = emit(MC_M, 0, 0, x, x, -1, -1); break;
            case OP_IDENTITY: r = emit(MC_A, 0, 0, x, -1, -1, -1); break;
            case OP_POW_INT: {  // Base.power_by_squaring order; the exponent is the literal input[2]
                const int p = s[1];
                if (p == 0) { r = emit(MC_A, 0, 0, one_slot, -1, -1, -1); break; }
                unsigned q = (unsigned)(p < 0 ? -(long long)p : p);
                int y = -1, b = x;
                while (q) {
                    if (q & 1u) y = y < 0 ? b : emit(MC_M, 0, 0, y, b, -1, -1);
                    q >>= 1;
                    if (q) b = emit(MC_M, 0, 0, b, b, -1, -1);
                }
                if (p < 0) y = emit(MC_INV, 0, 0, y, -1, -1, -1);
                r = (y == x) ? emit(MC_A, 0, 0, x, -1, -1, -1) : y;
            } break;
            case OP_ADD: r = emit(MC_AA, 0, 0, x, -1, rd(1), -1); break;
            case OP_SUB: r = emit(MC_AA, 0, 1, x, -1, rd(1), -1); break;
            case OP_DIV: r = emit(MC_DIV, 0, 0, x, rd(1), -1, -1); break;
            case OP_MUL: r = emit(MC_M, 0, 0, x, rd(1), -1, -1); break;
            case OP_ADD3: r = emit(MC_AA, 0, 0, emit(MC_AA, 0, 0, x, -1, rd(1), -1), -1, rd(2), -1); break;
            case OP_MUL3: r = emit(MC_M, 0, 0, emit(MC_M, 0, 0, x, rd(1), -1, -1), rd(2), -1, -1); break;
            case OP_MULADD: r = emit(MC_MA, 0, 0, x, rd(1), rd(2), -1); break;
            case OP_MULSUB: r = emit(MC_MA, 0, 1, x, rd(1), rd(2), -1); break;
            case OP_SUBMUL: r = emit(MC_MA, 1, 0, x, rd(1), rd(2), -1); break;
            // the reference's association (operations.jl: a + b + c + d and a * b * c * d fold from the left), so that
            // the interpreter rounds like the reference; the Taylor rules pair them ((a+b)+(c+d), taylor.jl) -- for a
            // sum the coefficients are independent and the pairing is exact per coefficient either way
            case OP_ADD4: {
                int t1 = emit(MC_AA, 0, 0, x, -1, rd(1), -1), t2 = emit(MC_AA, 0, 0, t1, -1, rd(2), -1);
                r = emit(MC_AA, 0, 0, t2, -1, rd(3), -1);
            } break;
            case OP_MUL4: {
                int t1 = emit(MC_M, 0, 0, x, rd(1), -1, -1), t2 = emit(MC_M, 0, 0, t1, rd(2), -1, -1);
                r = emit(MC_M, 0, 0, t2, rd(3), -1, -1);
            } break;
            case OP_MULMULADD: r = emit(MC_MM, 0, 0, x, rd(1), rd(2), rd(3)); break;
            default: r = emit(MC_MM, 0, 1, x, rd(1), rd(2), rd(3)); break;  // OP_MULMULSUB
        }
        const int out = s[5];
        if (out < 1 || out > d->tape_space) throw std::string("tape index out of range");
        if (is_input[out - 1]) throw std::string("instruction writes into the input block");
        cur[out - 1] = r;
    }
    if (!stopped) throw std::string("tape is not terminated by OP_STOP");

    // outputs
    std::vector<int> uv(d->n_u), Uv(d->n_U);
    for (int i = 0; i < d->n_u; ++i) {
        int idx = d->u_assign[2 * i] - 1, slot = d->u_assign[2 * i + 1] - 1;
        if (idx < 0 || idx >= d->out_dim || slot < 0 || slot >= d->tape_space || cur[slot] < 0) throw std::string("bad u assignment");
        uv[i] = cur[slot];
    }
    for (int i = 0; i < d->n_U; ++i) {
        int idx = d->U_assign[2 * i] - 1, slot = d->U_assign[2 * i + 1] - 1;
        if (idx < 0 || idx >= d->out_dim * n || slot < 0 || slot >= d->tape_space || cur[slot] < 0) throw std::string("bad U assignment");
        Uv[i] = cur[slot];
    }

    // List scheduling into rounds of at most `cap` independent micro-ops: an op is ready once all
    // its operands were produced in earlier rounds; among the ready ops the ones on the longest
    // remaining dependency chain go first (fewest rounds), ties in original tape order (the
    // reference's order keeps live ranges short).  One round = one "level" of the device loop.
    const int NV = (int)level.size();
    const int nops = (int)v.size();
    {
        std::vector<int> height((size_t)NV, 0);  // longest chain from the value to a sink
        for (int i = nops - 1; i >= 0; --i) {
            const VOp& o = v[i];
            for (int in : {o.a, o.b, o.c, o.d}) if (in >= 0) height[in] = std::max(height[in], height[o.out] + 1);
        }
        std::vector<int> done_round((size_t)NV, 0);     // round in which a value becomes available (inputs: 0)
        std::vector<char> scheduled((size_t)nops, 0);
        int remaining = nops, round = 0;
        std::vector<int> ready, newly;
        std::vector<int> wait_since((size_t)nops, 0);
        if (pair) cap = 2;  // thread-per-path engine with 2-way ILP: rounds of two independent ops, flattened below
        if (seg_window > 0) {
            // Segment scheduling (thread-per-path engines): a round = the ready ops of ONE (class, sign) key
            // among the next `seg_window` unscheduled ops in tape order -- long same-key runs for the
            // segment loops, while the window keeps the live ranges (tape slots) close to the reference's.
            std::vector<int> pending;
            for (int i = 0; i < nops; ++i) pending.push_back(i);
            while (!pending.empty()) {
                ++round;
                int cnt[32] = {0}, first[32];
                const int lim = std::min<int>((int)pending.size(), seg_window);
                std::vector<char> rdy((size_t)lim, 0);
                for (int p = 0; p < lim; ++p) {
                    const VOp& o = v[pending[p]];
                    bool ok = true;
                    for (int in : {o.a, o.b, o.c, o.d}) if (in >= Nin && done_round[in] == 0) ok = false;
                    if (!ok) continue;
                    rdy[p] = 1;
                    const int key = o.cls | (o.n1 << 3) | (o.n2 << 4);
                    if (cnt[key]++ == 0) first[key] = p;
                }
                int best = -1;
                for (int k = 0; k < 32; ++k) if (cnt[k] && (best < 0 || cnt[k] > cnt[best] || (cnt[k] == cnt[best] && first[k] < first[best]))) best = k;
                // the oldest pending op must not starve: take its key once it has waited for `seg_window` rounds
                {
                    const VOp& o = v[pending[0]];
                    const int key0 = o.cls | (o.n1 << 3) | (o.n2 << 4);
                    if (round - wait_since[pending[0]] >= 4) best = key0;
                }
                std::vector<int> rest;
                for (int p = 0; p < (int)pending.size(); ++p) {
                    const int i = pending[p];
                    const VOp& o = v[i];
                    if (p < lim && rdy[p] && (o.cls | (o.n1 << 3) | (o.n2 << 4)) == best) { v[i].level = round; level[o.out] = round; newly.push_back(o.out); }
                    else rest.push_back(i);
                }
                for (int id : newly) done_round[id] = round;  // results become visible to the NEXT round
                newly.clear();
                for (int p = 0; p < (int)rest.size() && p < seg_window; ++p) if (wait_since[rest[p]] == 0) wait_since[rest[p]] = round;
                pending.swap(rest);
            }
            remaining = 0;
        }
        if (cap <= 0 && seg_window <= 0) {  // sequential program (thread-per-path engine): original order, one op per allocation level
            for (int i = 0; i < nops; ++i) { v[i].level = i + 1; level[v[i].out] = i + 1; }
            remaining = 0;
        }
        while (remaining > 0) {
            ++round;
            ready.clear();
            for (int i = 0; i < nops; ++i) {
                if (scheduled[i]) continue;
                const VOp& o = v[i];
                bool ok = true;
                for (int in : {o.a, o.b, o.c, o.d}) if (in >= Nin && (done_round[in] == 0 || done_round[in] >= round)) ok = false;
                if (ok) ready.push_back(i);
            }
            std::stable_sort(ready.begin(), ready.end(), [&](int x, int y) {
                if (prio_height && height[v[x].out] != height[v[y].out]) return height[v[x].out] > height[v[y].out];
                return x < y;
            });
            if ((int)ready.size() > cap) ready.resize((size_t)cap);
            for (int i : ready) { scheduled[i] = 1; v[i].level = round; done_round[v[i].out] = round; level[v[i].out] = round; --remaining; }
        }
    }
    int n_levels = 0;
    for (const VOp& o : v) n_levels = std::max(n_levels, o.level);
    // last level that reads each virtual id (outputs stay live to the end)
    std::vector<int> last((size_t)NV, 0);
    for (const VOp& o : v) {
        last[o.out] = std::max(last[o.out], o.level);
        for (int in : {o.a, o.b, o.c, o.d}) if (in >= 0) last[in] = std::max(last[in], o.level);
    }
    for (int id : uv) last[id] = n_levels + 1;
    for (int id : Uv) last[id] = n_levels + 1;

    std::vector<int> order(v.size());
    for (size_t i = 0; i < v.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        const VOp &x = v[a], &y = v[b];
        if (x.level != y.level) return x.level < y.level;
        if (x.cls != y.cls) return x.cls < y.cls;
        return x.n1 * 2 + x.n2 < y.n1 * 2 + y.n2;
    });

    // linear-scan slot allocation over the levels
    std::vector<int> phys((size_t)NV, -1);
    for (int i = 0; i < Nin; ++i) phys[i] = i;
    std::vector<int> free_slots;
    std::vector<std::vector<int>> expire((size_t)n_levels + 3);
    int next_slot = Nin;
    size_t pos = 0;
    R.level_end.assign((size_t)n_levels, 0);
    for (int L = 1; L <= n_levels; ++L) {
        for (int id : expire[L - 1]) free_slots.push_back(phys[id]);  // last reader was in level L-1 or earlier
        int width = 0;
        while (pos < order.size() && v[order[pos]].level == L) {
            const VOp& o = v[order[pos]];
            int slot;
            if (!free_slots.empty()) { slot = free_slots.back(); free_slots.pop_back(); }
            else slot = next_slot++;
            phys[o.out] = slot;
            expire[std::min(last[o.out], n_levels + 1)].push_back(o.out);
            ++pos; ++width;
        }
        R.level_end[L - 1] = (int)pos;
        R.max_width = std::max(R.max_width, width);
    }
    std::vector<char> paired_first(v.size(), 0);
    if (pair) {  // mark the first op of every two-op round, then flatten to one sequential level
        int beg = 0;
        for (int e : R.level_end) { if (e - beg == 2) paired_first[beg] = 1; beg = e; }
    }
    if ((cap <= 0 && seg_window <= 0) || pair) { R.level_end.assign(1, nops); R.max_width = pair ? 2 : 1; }
    R.W = next_slot;
    if (R.W >= 65536) throw std::string("tape_space >= 65536 is not supported by the packed format");
    R.ops.resize(v.size());
    for (size_t i = 0; i < order.size(); ++i) {
        const VOp& o = v[order[i]];
        auto ph = [&](int id) { return (uint32_t)(id >= 0 ? phys[id] : 0); };
        MOp m;
        m.w0 = (uint32_t)phys[o.out] | ((uint32_t)o.cls << 16) | ((uint32_t)o.n1 << 19) | ((uint32_t)o.n2 << 20) |
               ((uint32_t)paired_first[i] << 21);
        m.w1 = ph(o.a) | (ph(o.b) << 16);
        m.w2 = ph(o.c) | (ph(o.d) << 16);
        m.pad = 0;
        R.ops[i] = m;
    }
    // fast format + segment table: a segment = maximal run of ops of one level with the same key
    {
        for (size_t i = 0; i < order.size(); ++i) {
            const VOp& o = v[order[i]];
            const int key = o.cls | (o.n1 << 3) | (o.n2 << 4);
            switch (key) {
                case HC_KEY(MC_MM, 0, 0): case HC_KEY(MC_MM, 0, 1): case HC_KEY(MC_MA, 0, 0): case HC_KEY(MC_MA, 0, 1):
                case HC_KEY(MC_MA, 1, 0): case HC_KEY(MC_M, 0, 0): case HC_KEY(MC_AA, 0, 0): case HC_KEY(MC_AA, 0, 1):
                case HC_KEY(MC_A, 0, 0): case HC_KEY(MC_A, 1, 0): case HC_KEY(MC_INV, 0, 0): case HC_KEY(MC_DIV, 0, 0):
                case HC_KEY(MC_INVNZ, 0, 0): break;
                default: throw std::string("internal: micro-op (class, sign) pair without a segment loop");
            }
            auto ph = [&](int id) { return (uint32_t)(id >= 0 ? phys[id] : 0); };
            FOp f; f.a = ph(o.a); f.b = ph(o.b); f.c = ph(o.c); f.out = (uint32_t)phys[o.out];
            R.fops.push_back(f);
            if (o.cls == MC_MM) { FOp g; g.a = ph(o.d); g.b = g.c = g.out = 0; R.fops.push_back(g); }
            const bool new_level = i == 0 || v[order[i - 1]].level != o.level;  // (sequential programs: one op per level)
            if (new_level || R.segs.back().x != key) { int2 sg; sg.x = key; sg.y = 0; R.segs.push_back(sg); }
            R.segs.back().y += 1;
        }
    }
    R.u_assign.resize(d->n_u); R.U_assign.resize(d->n_U);
    for (int i = 0; i < d->n_u; ++i) { R.u_assign[i].x = d->u_assign[2 * i] - 1; R.u_assign[i].y = phys[uv[i]]; }
    for (int i = 0; i < d->n_U; ++i) { R.U_assign[i].x = d->U_assign[2 * i] - 1; R.U_assign[i].y = phys[Uv[i]]; }
    return R;
}

}  // namespace hc
