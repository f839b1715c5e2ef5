"""Host-side tape tooling: a small expression DAG -> ModelKit ``InstructionSequence`` compiler.

In production the Julia host builds the tapes with ``ModelKit.instruction_sequence(F)`` and
``jacobian_instruction_sequence(F)`` (reference src/model_kit/instruction_sequence.jl:133-143,
256-272) and hands them to the C ABI unchanged.  Julia is absent from this image, so the tapes
of the benchmark/test systems are produced here instead.  The output follows the reference
tape format exactly (src/model_kit/instruction_sequence.jl:1-27, 145-254):

* ``Instruction`` = ``(input::NTuple{4,Int32}, op::Int32, output::Int32)``, 24 bytes, 1-based,
* tape layout ``[constants | parameters | t | variables | registers | assignments]``,
* unused inputs repeat the previous index, ``OP_POW_INT`` keeps the literal exponent in
  ``input[2]``, the last instruction is ``OP_STOP``,
* the Jacobian tape is the tape of ``[F; vec(dF/dx)]`` with ``output_dim = length(F)``
  (src/model_kit/instruction_interpreter.jl:94-110).

The instruction *order* and CSE differ from SymEngine's (any semantically equal tape is
acceptable: the reference tests pin values to rtol 1e-12, not tape contents).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field

import numpy as np

# OpType values: declaration order of src/model_kit/operations.jl:5-49
(OP_STOP, OP_CB, OP_ACOS, OP_ASIN, OP_COS, OP_COSH, OP_EXP, OP_INV, OP_INV_NOT_ZERO, OP_INVSQR,
 OP_NEG, OP_SIN, OP_SINH, OP_SQR, OP_SQRT, OP_TAN, OP_TANH, OP_IDENTITY, OP_ADD, OP_DIV, OP_MUL,
 OP_SUB, OP_POW_INT, OP_POW, OP_ADD3, OP_MUL3, OP_MULADD, OP_MULSUB, OP_SUBMUL, OP_ADD4, OP_MUL4,
 OP_MULMULADD, OP_MULMULSUB) = range(33)

OP_NAMES = ["STOP", "CB", "ACOS", "ASIN", "COS", "COSH", "EXP", "INV", "INV_NOT_ZERO", "INVSQR",
            "NEG", "SIN", "SINH", "SQR", "SQRT", "TAN", "TANH", "IDENTITY", "ADD", "DIV", "MUL",
            "SUB", "POW_INT", "POW", "ADD3", "MUL3", "MULADD", "MULSUB", "SUBMUL", "ADD4", "MUL4",
            "MULMULADD", "MULMULSUB"]


# --------------------------------------------------------------------------- expression DAG
class Graph:
    """Hash-consed expression DAG.  Node = tuple; id = index into ``nodes``."""

    def __init__(self):
        self.nodes: list[tuple] = []
        self.index: dict[tuple, int] = {}
        self._diff: dict[tuple[int, int], int] = {}

    def _mk(self, node: tuple) -> int:
        i = self.index.get(node)
        if i is None:
            i = len(self.nodes)
            self.nodes.append(node)
            self.index[node] = i
        return i

    # leaves
    def const(self, c) -> int:
        c = complex(c)
        return self._mk(("c", c.real + 0.0, c.imag + 0.0))

    def var(self, i: int) -> int:
        return self._mk(("x", i))

    def param(self, i: int) -> int:
        return self._mk(("p", i))

    def t(self) -> int:
        return self._mk(("t",))

    def is_const(self, a: int):
        nd = self.nodes[a]
        return complex(nd[1], nd[2]) if nd[0] == "c" else None

    # arithmetic with light simplification
    def add(self, a: int, b: int) -> int:
        ca, cb = self.is_const(a), self.is_const(b)
        if ca is not None and cb is not None:
            return self.const(ca + cb)
        if ca == 0:
            return b
        if cb == 0:
            return a
        if self.nodes[b][0] == "neg":
            return self.sub(a, self.nodes[b][1])
        if self.nodes[a][0] == "neg":
            return self.sub(b, self.nodes[a][1])
        if a > b:
            a, b = b, a
        return self._mk(("add", a, b))

    def sub(self, a: int, b: int) -> int:
        ca, cb = self.is_const(a), self.is_const(b)
        if ca is not None and cb is not None:
            return self.const(ca - cb)
        if cb == 0:
            return a
        if ca == 0:
            return self.neg(b)
        if a == b:
            return self.const(0)
        if self.nodes[b][0] == "neg":
            return self.add(a, self.nodes[b][1])
        return self._mk(("sub", a, b))

    def neg(self, a: int) -> int:
        ca = self.is_const(a)
        if ca is not None:
            return self.const(-ca)
        if self.nodes[a][0] == "neg":
            return self.nodes[a][1]
        return self._mk(("neg", a))

    def mul(self, a: int, b: int) -> int:
        ca, cb = self.is_const(a), self.is_const(b)
        if ca is not None and cb is not None:
            return self.const(ca * cb)
        if ca == 0 or cb == 0:
            return self.const(0)
        if ca == 1:
            return b
        if cb == 1:
            return a
        if ca == -1:
            return self.neg(b)
        if cb == -1:
            return self.neg(a)
        if self.nodes[a][0] == "neg":
            return self.neg(self.mul(self.nodes[a][1], b))
        if self.nodes[b][0] == "neg":
            return self.neg(self.mul(a, self.nodes[b][1]))
        if a == b:
            return self.pow(a, 2)
        # const * (const * y) -> (const*const) * y
        if ca is not None and self.nodes[b][0] == "mul":
            c2 = self.is_const(self.nodes[b][1])
            if c2 is not None:
                return self.mul(self.const(ca * c2), self.nodes[b][2])
        if cb is not None and self.nodes[a][0] == "mul":
            c2 = self.is_const(self.nodes[a][1])
            if c2 is not None:
                return self.mul(self.const(cb * c2), self.nodes[a][2])
        if a > b:
            a, b = b, a
        return self._mk(("mul", a, b))

    def pow(self, a: int, k: int) -> int:
        k = int(k)
        if k == 0:
            return self.const(1)
        if k == 1:
            return a
        ca = self.is_const(a)
        if ca is not None:
            return self.const(ca ** k)
        if self.nodes[a][0] == "pow" and self.nodes[a][2] > 0 and k > 0:
            return self.pow(self.nodes[a][1], self.nodes[a][2] * k)
        return self._mk(("pow", a, k))

    def div(self, a: int, b: int) -> int:
        cb = self.is_const(b)
        if cb is not None:
            return self.mul(a, self.const(1 / cb))
        return self._mk(("div", a, b))

    # symbolic differentiation w.r.t. variable j
    def diff(self, a: int, j: int) -> int:
        key = (a, j)
        r = self._diff.get(key)
        if r is not None:
            return r
        nd = self.nodes[a]
        k = nd[0]
        if k == "x":
            r = self.const(1 if nd[1] == j else 0)
        elif k in ("c", "p", "t"):
            r = self.const(0)
        elif k == "add":
            r = self.add(self.diff(nd[1], j), self.diff(nd[2], j))
        elif k == "sub":
            r = self.sub(self.diff(nd[1], j), self.diff(nd[2], j))
        elif k == "neg":
            r = self.neg(self.diff(nd[1], j))
        elif k == "mul":
            r = self.add(self.mul(self.diff(nd[1], j), nd[2]), self.mul(nd[1], self.diff(nd[2], j)))
        elif k == "pow":
            d = self.diff(nd[1], j)
            r = self.mul(self.mul(self.const(nd[2]), self.pow(nd[1], nd[2] - 1)), d)
        elif k == "div":
            da, db = self.diff(nd[1], j), self.diff(nd[2], j)
            r = self.div(self.sub(self.mul(da, nd[2]), self.mul(nd[1], db)), self.pow(nd[2], 2))
        else:
            raise ValueError(k)
        self._diff[key] = r
        return r

    # numeric evaluation (python complex or mpmath.mpc via ``ctx``)
    def evaluate(self, outs, x, p=(), t=None, ctx=None):
        conv = (lambda z: complex(z)) if ctx is None else (lambda z: ctx.mpc(z.real, z.imag) if isinstance(z, complex) else ctx.mpc(z))
        val: dict[int, object] = {}
        order = self.topo(outs)
        for a in order:
            nd = self.nodes[a]
            k = nd[0]
            if k == "c":
                v = conv(complex(nd[1], nd[2]))
            elif k == "x":
                v = x[nd[1]]
            elif k == "p":
                v = p[nd[1]]
            elif k == "t":
                v = t
            elif k == "add":
                v = val[nd[1]] + val[nd[2]]
            elif k == "sub":
                v = val[nd[1]] - val[nd[2]]
            elif k == "neg":
                v = -val[nd[1]]
            elif k == "mul":
                v = val[nd[1]] * val[nd[2]]
            elif k == "pow":
                v = val[nd[1]] ** nd[2]
            elif k == "div":
                v = val[nd[1]] / val[nd[2]]
            val[a] = v
        return [val[o] for o in outs]

    def topo(self, outs):
        """Children-first order of all nodes reachable from ``outs`` (iterative DFS)."""
        seen, order = set(), []
        for o in outs:
            if o in seen:
                continue
            stack = [(o, False)]
            while stack:
                a, done = stack.pop()
                if done:
                    order.append(a)
                    continue
                if a in seen:
                    continue
                seen.add(a)
                stack.append((a, True))
                nd = self.nodes[a]
                if nd[0] in ("add", "sub", "mul", "div"):
                    stack.append((nd[2], False))
                    stack.append((nd[1], False))
                elif nd[0] in ("neg", "pow"):
                    stack.append((nd[1], False))
        return order

    # polynomial expansion in the variables: {exponent tuple: coefficient}, parameters/t numeric
    def expand(self, a: int, nvars: int, p=(), t=None, _memo=None):
        memo = {} if _memo is None else _memo
        zero = (0,) * nvars

        def pmul(A, B):
            out = {}
            for ea, ca in A.items():
                for eb, cb in B.items():
                    e = tuple(i + j for i, j in zip(ea, eb))
                    out[e] = out.get(e, 0) + ca * cb
            return out

        for b in self.topo([a]):
            if b in memo:
                continue
            nd = self.nodes[b]
            k = nd[0]
            if k == "c":
                r = {zero: complex(nd[1], nd[2])}
            elif k == "x":
                e = [0] * nvars
                e[nd[1]] = 1
                r = {tuple(e): 1.0 + 0j}
            elif k == "p":
                r = {zero: complex(p[nd[1]])}
            elif k == "t":
                r = {zero: complex(t)}
            elif k == "add":
                r = dict(memo[nd[1]])
                for e, c in memo[nd[2]].items():
                    r[e] = r.get(e, 0) + c
            elif k == "sub":
                r = dict(memo[nd[1]])
                for e, c in memo[nd[2]].items():
                    r[e] = r.get(e, 0) - c
            elif k == "neg":
                r = {e: -c for e, c in memo[nd[1]].items()}
            elif k == "mul":
                r = pmul(memo[nd[1]], memo[nd[2]])
            elif k == "pow":
                if nd[2] < 0:
                    raise ValueError("not a polynomial")
                r = {zero: 1.0 + 0j}
                for _ in range(nd[2]):
                    r = pmul(r, memo[nd[1]])
            else:
                raise ValueError("not a polynomial")
            memo[b] = r
        return {e: c for e, c in memo[a].items() if c != 0}


class Expr:
    """Operator-overloading handle on a Graph node."""
    __slots__ = ("g", "i")
    __array_priority__ = 1000

    def __init__(self, g: Graph, i: int):
        self.g, self.i = g, i

    def _c(self, o):
        return o.i if isinstance(o, Expr) else self.g.const(o)

    def __add__(self, o): return Expr(self.g, self.g.add(self.i, self._c(o)))
    def __radd__(self, o): return Expr(self.g, self.g.add(self._c(o), self.i))
    def __sub__(self, o): return Expr(self.g, self.g.sub(self.i, self._c(o)))
    def __rsub__(self, o): return Expr(self.g, self.g.sub(self._c(o), self.i))
    def __mul__(self, o): return Expr(self.g, self.g.mul(self.i, self._c(o)))
    def __rmul__(self, o): return Expr(self.g, self.g.mul(self._c(o), self.i))
    def __truediv__(self, o): return Expr(self.g, self.g.div(self.i, self._c(o)))
    def __neg__(self): return Expr(self.g, self.g.neg(self.i))
    def __pow__(self, k): return Expr(self.g, self.g.pow(self.i, k))
    def diff(self, j: int): return Expr(self.g, self.g.diff(self.i, j))
    def __repr__(self): return f"Expr({self.g.nodes[self.i]})"


# --------------------------------------------------------------------------- tape programs
@dataclass
class Program:
    """Mirror of InstructionSequence (src/model_kit/instruction_sequence.jl:7-27)."""
    instructions: np.ndarray   # (L, 6) int32: in[4], op, out; 1-based; last row OP_STOP
    constants: np.ndarray      # (C,) complex128
    param_offset: int
    n_params: int
    t_index: int               # 0 = none
    var_offset: int
    n_vars: int
    u_assign: np.ndarray       # (nu, 2) int32 (i, k)
    U_assign: np.ndarray       # (nU, 2) int32 (j, k), j column-major over (out_dim, n_vars)
    out_dim: int
    tape_space: int

    @property
    def n_instructions(self):
        return int(self.instructions.shape[0])


def _lower(g: Graph, outs: list[int], nvars: int, nparams: int, has_t: bool, out_dim: int) -> Program:
    order = g.topo(outs)
    # use counts (for fusion decisions)
    uses: dict[int, int] = {}
    for a in order:
        nd = g.nodes[a]
        if nd[0] in ("add", "sub", "mul", "div"):
            uses[nd[1]] = uses.get(nd[1], 0) + 1
            uses[nd[2]] = uses.get(nd[2], 0) + 1
        elif nd[0] in ("neg", "pow"):
            uses[nd[1]] = uses.get(nd[1], 0) + 1
    for o in outs:
        uses[o] = uses.get(o, 0) + 1

    consts = [a for a in order if g.nodes[a][0] == "c"]
    # a constant that is only ever the exponent-free operand still needs a slot
    const_slot = {a: i + 1 for i, a in enumerate(consts)}
    C = len(consts)
    param_offset = C
    t_index = C + nparams + 1 if has_t else 0
    var_offset = C + nparams + (1 if has_t else 0)
    input_block = var_offset + nvars

    def leaf_slot(a):
        nd = g.nodes[a]
        if nd[0] == "c":
            return const_slot[a]
        if nd[0] == "p":
            return param_offset + 1 + nd[1]
        if nd[0] == "t":
            return t_index
        if nd[0] == "x":
            return var_offset + 1 + nd[1]
        return None

    def is_mul(a):
        return g.nodes[a][0] == "mul" and uses.get(a, 0) == 1 and a not in fused_skip

    # virtual instructions: (op, [operand node ids or ('lit', k)], result node)
    virt = []
    fused_skip: set[int] = set()
    # decide fusions top-down: walk in reverse topological order so that parents claim children
    fuse: dict[int, tuple] = {}
    outset = set(outs)
    for a in reversed(order):
        nd = g.nodes[a]
        if a in fused_skip:
            continue
        if nd[0] == "add":
            l, r = nd[1], nd[2]
            lm = g.nodes[l][0] == "mul" and uses.get(l, 0) == 1 and l not in outset
            rm = g.nodes[r][0] == "mul" and uses.get(r, 0) == 1 and r not in outset
            if lm and rm:
                fuse[a] = (OP_MULMULADD, g.nodes[l][1], g.nodes[l][2], g.nodes[r][1], g.nodes[r][2])
                fused_skip.update((l, r))
            elif lm:
                fuse[a] = (OP_MULADD, g.nodes[l][1], g.nodes[l][2], r)
                fused_skip.add(l)
            elif rm:
                fuse[a] = (OP_MULADD, g.nodes[r][1], g.nodes[r][2], l)
                fused_skip.add(r)
        elif nd[0] == "sub":
            l, r = nd[1], nd[2]
            lm = g.nodes[l][0] == "mul" and uses.get(l, 0) == 1 and l not in outset
            rm = g.nodes[r][0] == "mul" and uses.get(r, 0) == 1 and r not in outset
            if lm and rm:
                fuse[a] = (OP_MULMULSUB, g.nodes[l][1], g.nodes[l][2], g.nodes[r][1], g.nodes[r][2])
                fused_skip.update((l, r))
            elif lm:
                fuse[a] = (OP_MULSUB, g.nodes[l][1], g.nodes[l][2], r)
                fused_skip.add(l)
            elif rm:
                fuse[a] = (OP_SUBMUL, g.nodes[r][1], g.nodes[r][2], l)
                fused_skip.add(r)
    for a in order:
        nd = g.nodes[a]
        k = nd[0]
        if k in ("c", "x", "p", "t") or a in fused_skip:
            continue
        if a in fuse:
            f = fuse[a]
            virt.append((f[0], list(f[1:]), a))
        elif k == "add":
            virt.append((OP_ADD, [nd[1], nd[2]], a))
        elif k == "sub":
            virt.append((OP_SUB, [nd[1], nd[2]], a))
        elif k == "mul":
            virt.append((OP_MUL, [nd[1], nd[2]], a))
        elif k == "div":
            virt.append((OP_DIV, [nd[1], nd[2]], a))
        elif k == "neg":
            virt.append((OP_NEG, [nd[1]], a))
        elif k == "pow":
            if nd[2] == 2:
                virt.append((OP_SQR, [nd[1]], a))
            elif nd[2] == 3:
                virt.append((OP_CB, [nd[1]], a))
            elif nd[2] == -1:
                virt.append((OP_INV, [nd[1]], a))
            else:
                virt.append((OP_POW_INT, [nd[1], ("lit", nd[2])], a))

    # register allocation with liveness (reference: reduce_space, instruction_sequence.jl:354-450)
    last_use: dict[int, int] = {}
    for idx, (_, args, _) in enumerate(virt):
        for b in args:
            if not isinstance(b, tuple):
                last_use[b] = idx
    n_assign_slots = sum(1 for o in outs if leaf_slot(o) is None and g.is_const(o) != 0)
    slot: dict[int, int] = {}
    free: list[int] = []
    next_reg = input_block + 1
    max_reg = input_block
    out_nodes = {o for o in outs if leaf_slot(o) is None}
    rows = []
    for idx, (op, args, res) in enumerate(virt):
        ins = []
        for b in args:
            if isinstance(b, tuple):
                ins.append(int(b[1]))
            else:
                s = leaf_slot(b)
                ins.append(s if s is not None else slot[b])
        # free registers whose last use is this instruction (outputs are never freed)
        for b in args:
            if isinstance(b, tuple):
                continue
            if leaf_slot(b) is None and last_use.get(b) == idx and b not in out_nodes and b in slot:
                free.append(slot[b])
                last_use[b] = -1
        if res in out_nodes:
            slot[res] = -1  # patched below
        else:
            if free:
                slot[res] = free.pop()
            else:
                slot[res] = next_reg
                next_reg += 1
            max_reg = max(max_reg, slot[res])
        while len(ins) < 4:
            ins.append(ins[-1])
        rows.append([ins[0], ins[1], ins[2], ins[3], op, res])
    # assignment slots after the registers
    assign_slot: dict[int, int] = {}
    nxt = max_reg + 1
    for o in outs:
        if o in out_nodes and o not in assign_slot:
            assign_slot[o] = nxt
            nxt += 1
    tape_space = max(nxt - 1, input_block, 1)
    # patch: rows reference nodes for outputs
    node_slot = dict(slot)
    node_slot.update(assign_slot)
    final = []
    for idx, (op, args, res) in enumerate(virt):
        ins = []
        for b in args:
            if isinstance(b, tuple):
                ins.append(int(b[1]))
            else:
                s = leaf_slot(b)
                ins.append(s if s is not None else node_slot[b])
        while len(ins) < 4:
            ins.append(ins[-1])
        final.append([ins[0], ins[1], ins[2], ins[3], op, node_slot[res]])
    final.append([tape_space] * 4 + [OP_STOP, tape_space])

    u_assign, U_assign = [], []
    for k, o in enumerate(outs):
        if g.is_const(o) == 0:
            continue  # zero outputs are not assigned (instruction_sequence.jl:441-445)
        s = leaf_slot(o)
        if s is None:
            s = node_slot[o]
        if k < out_dim:
            u_assign.append((k + 1, s))
        else:
            U_assign.append((k + 1 - out_dim, s))
    cvals = np.array([complex(g.nodes[a][1], g.nodes[a][2]) for a in consts], dtype=np.complex128)
    return Program(
        instructions=np.array(final, dtype=np.int32).reshape(-1, 6),
        constants=cvals,
        param_offset=param_offset, n_params=nparams, t_index=t_index,
        var_offset=var_offset, n_vars=nvars,
        u_assign=np.array(u_assign, dtype=np.int32).reshape(-1, 2),
        U_assign=np.array(U_assign, dtype=np.int32).reshape(-1, 2),
        out_dim=out_dim, tape_space=tape_space,
    )


@dataclass
class System:
    """A polynomial system F(x; p) with its eval and Jacobian tapes (InterpretedSystem)."""
    graph: Graph
    exprs: list[int]
    n_vars: int
    n_params: int
    eval_program: Program = field(init=False)
    jac_program: Program = field(init=False)

    def __post_init__(self):
        g = self.graph
        m = len(self.exprs)
        self.eval_program = _lower(g, list(self.exprs), self.n_vars, self.n_params, False, m)
        jac = [g.diff(e, j) for j in range(self.n_vars) for e in self.exprs]  # vec(), column-major
        self.jac_program = _lower(g, list(self.exprs) + jac, self.n_vars, self.n_params, False, m)

    @property
    def n_eqs(self):
        return len(self.exprs)

    def evaluate(self, x, p=(), ctx=None):
        return self.graph.evaluate(self.exprs, x, p, None, ctx)

    def jacobian(self, x, p=(), ctx=None):
        g = self.graph
        m, n = self.n_eqs, self.n_vars
        jac = [g.diff(e, j) for j in range(n) for e in self.exprs]
        v = g.evaluate(jac, x, p, None, ctx)
        return [[v[j * m + i] for j in range(n)] for i in range(m)]

    def support_coefficients(self, p=()):
        """(supports, coeffs) per equation: reference ModelKit.support_coefficients."""
        out_s, out_c = [], []
        memo = {}
        for e in self.exprs:
            poly = self.graph.expand(e, self.n_vars, p, None, memo)
            exps = sorted(poly.keys(), reverse=True)
            out_s.append(np.array(exps, dtype=np.int64).reshape(-1, self.n_vars).T)
            out_c.append(np.array([poly[k] for k in exps], dtype=np.complex128))
        return out_s, out_c


def make_system(builder, n_vars: int, n_params: int = 0) -> System:
    """``builder(x, p) -> list[Expr]`` with x, p lists of Expr."""
    g = Graph()
    x = [Expr(g, g.var(i)) for i in range(n_vars)]
    p = [Expr(g, g.param(i)) for i in range(n_params)]
    exprs = builder(x, p)
    ids = [e.i if isinstance(e, Expr) else g.const(e) for e in exprs]
    return System(g, ids, n_vars, n_params)


def system_from_support(supports, n_vars: int) -> System:
    """polyhedral_system(support): F_i = sum_j c_ij x^{a_ij}, coefficients are parameters
    (reference src/polyhedral.jl:166-177)."""
    g = Graph()
    x = [Expr(g, g.var(i)) for i in range(n_vars)]
    exprs, k = [], 0
    for A in supports:
        A = np.asarray(A)
        s = None
        for j in range(A.shape[1]):
            c = Expr(g, g.param(k)); k += 1
            mon = None
            for i in range(n_vars):
                if A[i, j] > 0:
                    f = x[i] ** int(A[i, j])
                    mon = f if mon is None else mon * f
            term = c if mon is None else c * mon
            s = term if s is None else s + term
        exprs.append(s)
    return System(g, [e.i for e in exprs], n_vars, k)


def system_from_terms(supports, coeffs, n_vars: int) -> System:
    """F_i = sum_j coeffs[i][j] x^{supports[i][:, j]} with constant coefficients."""
    g = Graph()
    x = [Expr(g, g.var(i)) for i in range(n_vars)]
    exprs = []
    for A, cs in zip(supports, coeffs):
        A = np.asarray(A)
        s = None
        for j in range(A.shape[1]):
            mon = None
            for i in range(n_vars):
                if A[i, j] > 0:
                    f = x[i] ** int(A[i, j])
                    mon = f if mon is None else mon * f
            term = Expr(g, g.const(cs[j])) if mon is None else complex(cs[j]) * mon
            s = term if s is None else s + term
        exprs.append(s)
    return System(g, [e.i for e in exprs], n_vars, 0)
