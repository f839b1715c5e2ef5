"""Algorithmic FLOP accounting of the path-tracking hot path.

Unit = one path.  FLOPs(path) = sum over events count(event) * cost(event); the counts come from
the per-path counters the kernel maintains (hc_results.counters), the costs are fixed by the tapes
and n.  Real-flop cost of the complex ops: add/sub/neg 2, mul 6, sqr 5, cb 11, muladd/mulsub/submul
8, mulmuladd/mulmulsub 14, add3 4, add4 6, mul3 12, mul4 18, inv 8, div 14,
pow_int(p) 6 (floor(log2 p) + popcount(p) - 1) (+8 if p < 0)."""
from __future__ import annotations

import numpy as np

from . import modelkit as mk

_COST = {mk.OP_ADD: 2, mk.OP_SUB: 2, mk.OP_NEG: 2, mk.OP_MUL: 6, mk.OP_SQR: 5, mk.OP_CB: 11, mk.OP_MULADD: 8,
         mk.OP_MULSUB: 8, mk.OP_SUBMUL: 8, mk.OP_MULMULADD: 14, mk.OP_MULMULSUB: 14, mk.OP_ADD3: 4, mk.OP_ADD4: 6,
         mk.OP_MUL3: 12, mk.OP_MUL4: 18, mk.OP_INV: 8, mk.OP_INV_NOT_ZERO: 8, mk.OP_INVSQR: 13, mk.OP_DIV: 14,
         mk.OP_IDENTITY: 0}
# (real adds, real muls) per op, for the DoubleDouble variant: add -> 11 flops, mul -> 9 flops
_ADDMUL = {mk.OP_ADD: (2, 0), mk.OP_SUB: (2, 0), mk.OP_NEG: (0, 0), mk.OP_MUL: (2, 4), mk.OP_SQR: (3, 2), mk.OP_CB: (5, 6),
           mk.OP_MULADD: (4, 4), mk.OP_MULSUB: (4, 4), mk.OP_SUBMUL: (4, 4), mk.OP_MULMULADD: (6, 8),
           mk.OP_MULMULSUB: (6, 8), mk.OP_ADD3: (4, 0), mk.OP_ADD4: (6, 0), mk.OP_MUL3: (4, 8), mk.OP_MUL4: (6, 12),
           mk.OP_INV: (1, 4), mk.OP_INV_NOT_ZERO: (1, 4), mk.OP_INVSQR: (4, 6), mk.OP_DIV: (3, 8), mk.OP_IDENTITY: (0, 0)}
_MULTYPE = {mk.OP_MUL: 1, mk.OP_SQR: 1, mk.OP_CB: 2, mk.OP_MULADD: 1, mk.OP_MULSUB: 1, mk.OP_SUBMUL: 1, mk.OP_MULMULADD: 2,
            mk.OP_MULMULSUB: 2, mk.OP_MUL3: 2, mk.OP_MUL4: 3, mk.OP_INV: 1, mk.OP_INV_NOT_ZERO: 1, mk.OP_INVSQR: 2,
            mk.OP_DIV: 1}
_ADDTYPE = {mk.OP_ADD: 1, mk.OP_SUB: 1, mk.OP_NEG: 1, mk.OP_ADD3: 2, mk.OP_ADD4: 3, mk.OP_MULADD: 0, mk.OP_MULMULADD: 1,
            mk.OP_MULMULSUB: 1, mk.OP_MULSUB: 0}


def _pow_cost(p: int) -> int:
    q = abs(int(p))
    c = 6 * (q.bit_length() - 1 + bin(q).count("1") - 1) if q > 0 else 0
    return c + (8 if p < 0 else 0)


def tape_costs(prog: mk.Program) -> dict:
    f64 = dd = 0
    tay = {1: 0, 2: 0, 3: 0}
    for row in prog.instructions:
        op = int(row[4])
        if op == mk.OP_STOP:
            break
        if op == mk.OP_POW_INT:
            c = _pow_cost(int(row[1]))
            f64 += c
            nm = c // 6
            dd += nm * (4 * 9 + 2 * 11)
            for K in tay:
                tay[K] += 8 * (K + 1) * (K + 2) // 2 * max(nm, 1)
            continue
        f64 += _COST[op]
        a, m = _ADDMUL[op]
        dd += 11 * a + 9 * m
        for K in tay:
            tay[K] += 8 * ((K + 1) * (K + 2) // 2) * _MULTYPE.get(op, 0) + 2 * (K + 1) * _ADDTYPE.get(op, 0)
    return {"f64": f64, "dd": dd, "taylor": tay}


def homotopy_costs(F: mk.System, G: mk.System | None = None) -> dict:
    """Cost of one evaluate!, evaluate_and_jacobian!, DD evaluate! and taylor!(K) of the homotopy."""
    n = F.n_vars
    e, j = tape_costs(F.eval_program), tape_costs(F.jac_program)
    c = {"eval": e["f64"], "evaljac": j["f64"], "eval_dd": e["dd"], "taylor": dict(e["taylor"]), "n": n}
    if G is not None:  # straight line: two tapes + the combine u = ts*G + tt*F (8 flops per entry)
        ge, gj = tape_costs(G.eval_program), tape_costs(G.jac_program)
        c["eval"] += ge["f64"] + 8 * n
        c["evaljac"] += gj["f64"] + 8 * (n + n * n)
        c["eval_dd"] += ge["dd"] + 80 * n
        for K in c["taylor"]:
            c["taylor"][K] += ge["taylor"][K] + 16 * n
    else:          # parameter / coefficient homotopy: p(t) = t p + (1 - t) q
        c["eval"] += 8 * F.n_params
        c["evaljac"] += 8 * F.n_params
    c["lu"] = 8 * n ** 3 / 3 + 6 * n ** 2
    c["solve"] = 8 * n ** 2
    c["step_misc"] = 60 * n
    return c


def batch_flops(costs: dict, counters: np.ndarray, accepted: np.ndarray, rejected: np.ndarray) -> float:
    """Total algorithmic FLOPs of a batch from the (N, 8) event counters of hc_results."""
    cs = counters.sum(axis=0).astype(np.float64)
    fact, ldiv, evaljac, ev, evdd, t1, t2, t3 = cs
    steps = float(accepted.sum() + rejected.sum())
    return (evaljac * costs["evaljac"] + ev * costs["eval"] + evdd * costs["eval_dd"] + t1 * costs["taylor"][1]
            + t2 * costs["taylor"][2] + t3 * costs["taylor"][3] + fact * costs["lu"] + ldiv * costs["solve"]
            + steps * costs["step_misc"])


def path_bytes(n: int, P_per_path: int = 0, polyhedral: bool = False) -> int:
    """Algorithmic HBM bytes per path: start in + result out."""
    b = 16 * n + 40 * n + 130 + 16 * P_per_path
    if polyhedral:
        b += 8 * P_per_path + 32 * n
    return b
