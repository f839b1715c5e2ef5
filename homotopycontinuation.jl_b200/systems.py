"""Definitions of the benchmark / test systems (host side).

Sources (reference file:line):
  katsura(n)     test/model_kit/slp_test.jl:78-85
  cyclic(n)      test/test_systems.jl:72-77
  tritangents    test/test_systems.jl:31-54   (the 20 cubic coefficients are parameters)
  cyclooctane    benchmarks/cyclooctane.jl:4-20, test/test_systems.jl:57-69
  bio-chem 1     benchmarks/bio-chemical-rection-networks.jl:19-26
"""
from __future__ import annotations

import itertools
import math

import numpy as np

from .modelkit import Expr, System, make_system


def katsura(n: int) -> System:
    def build(x, p):
        eqs = []
        for m in range(n):
            s = None
            for l in range(-n, n + 1):
                if abs(m - l) <= n:
                    term = x[abs(l)] * x[abs(m - l)]
                    s = term if s is None else s + term
            eqs.append(s - x[m])
        lin = x[0]
        for i in range(1, n + 1):
            lin = lin + 2 * x[i]
        eqs.append(lin - 1)
        return eqs
    return make_system(build, n + 1)


def cyclic(n: int) -> System:
    def build(z, p):
        eqs = []
        for m in range(0, n - 1):
            s = None
            for j in range(1, n + 1):
                prod = None
                for k in range(j, j + m + 1):
                    v = z[(k - 1) % n]
                    prod = v if prod is None else prod * v
                s = prod if s is None else s + prod
            eqs.append(s)
        prod = z[0]
        for i in range(1, n):
            prod = prod * z[i]
        eqs.append(prod - 1)
        return eqs
    return make_system(build, n)


def _dense_poly_exponents(nvars: int, d: int):
    """Exponents of all monomials of degree <= d (order fixed here; any fixed order works since
    the coefficients are generic)."""
    out = []
    for deg in range(d, -1, -1):
        for e in itertools.product(range(deg + 1), repeat=nvars):
            if sum(e) == deg:
                out.append(e)
    return out


def tritangents() -> System:
    """12 equations in (h, x, y, z) with the 20 coefficients c of a dense cubic as parameters."""
    exps = _dense_poly_exponents(3, 3)
    assert len(exps) == 20

    def det3(M):
        return (M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1])
                - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0])
                + M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]))

    def build(v, c):
        h = v[0:3]
        eqs = []
        for blk in range(3):
            x = v[3 + 3 * blk: 6 + 3 * blk]
            Q = x[2] - x[0] * x[1]
            Cc = None
            for k, e in enumerate(exps):
                mon = None
                for i in range(3):
                    if e[i]:
                        f = x[i] ** e[i]
                        mon = f if mon is None else mon * f
                term = c[k] if mon is None else c[k] * mon
                Cc = term if Cc is None else Cc + term
            idx = [3 + 3 * blk + i for i in range(3)]
            dQ = [Q.diff(j) for j in idx]
            dC = [Cc.diff(j) for j in idx]
            M = [[h[i], dQ[i], dC[i]] for i in range(3)]
            eqs += [h[0] * x[0] + h[1] * x[1] + h[2] * x[2] - 1, Q, Cc, det3(M)]
        return eqs
    return make_system(build, 12, 20)


def cyclooctane() -> System:
    """15 distance quadrics + 2 linear slices A z = b; parameters = [vec(A) (2x17, col-major); b]."""
    c2 = 2.0

    def build(zv, p):
        def Z(col):
            if col == 0:
                return [0.0, 0.0, 0.0]
            if 1 <= col <= 5:
                return [zv[3 * (col - 1) + r] for r in range(3)]
            if col == 6:
                return [zv[15], zv[16], 0.0]
            return [math.sqrt(c2), 0.0, 0.0]

        def dist2(a, b):
            s = None
            for r in range(3):
                d = Z(a)[r] - Z(b)[r] if isinstance(Z(a)[r], Expr) else (-(Z(b)[r] - Z(a)[r]) if isinstance(Z(b)[r], Expr) else Z(a)[r] - Z(b)[r])
                if isinstance(d, Expr):
                    t = d * d
                else:
                    t = d * d
                    if t == 0.0:
                        continue
                s = t if s is None else s + t
            return s
        eqs = [dist2(i, i + 1) - c2 for i in range(7)]
        eqs += [dist2(i, i + 2) - 8 * c2 / 3 for i in range(6)]
        eqs.append(dist2(6, 0) - 8 * c2 / 3)
        eqs.append(dist2(7, 1) - 8 * c2 / 3)
        for r in range(2):
            s = None
            for j in range(17):
                t = p[2 * j + r] * zv[j]
                s = t if s is None else s + t
            eqs.append(s - p[34 + r])
        return eqs
    return make_system(build, 17, 36)


BIOCHEM1_PVALS = np.array([0.04, 0.04, 1.0, 1.0, 10.0, 0.0, 0.04, 35.0, 0.1, 0.04])


def biochem1() -> System:
    def build(x, p):
        return [
            -x[0] * x[2] * p[2] - x[0] * p[1] + p[0],
            x[0] * x[1] * x[2] * p[7] * p[8] + x[0] * x[2] * p[6] * p[7] * p[8]
            - x[1] * x[2] * p[4] * p[5] - x[1] * p[4] * p[5] * p[9] - x[1] * x[2] * p[3]
            - x[1] * p[3] * p[9],
            x[1] + x[2] - 1.0,
        ]
    return make_system(build, 3, 10)


# bio-chemical reaction networks 2 - 4 of the reference benchmark (benchmarks/bio-chemical-rection-networks.jl:46-61, 106,
# 128-131; networks 2 and 3 are the same polynomials at two parameter points)
BIOCHEM2_PVALS = np.array([0.005, 0.1, 2.8, 10, 100, 0.1, 0.01, 0.0])
BIOCHEM3_PVALS = np.array([0.005, 0.1, 2.8, 10, 100, 0.1, 0.01, 1.0])
BIOCHEM4_PVALS = np.array([1.0, 0.2, 1.0])


def biochem2() -> System:
    def build(x, p):
        p34 = p[2] ** 4
        return [
            -x[0] ** 5 * x[1] * p[4] + x[0] ** 4 * x[2] * p[5] * p[7] - x[0] * x[1] * p34 * p[4] + x[2] * p34 * p[5] * p[7]
            - x[0] ** 5 * p[6] + x[0] ** 4 * x[2] * p[3] - x[0] * p34 * p[6] + x[2] * p34 * p[3] + x[0] ** 4 * p[0] + x[0] ** 4 * p[1] + p[0] * p34,
            -x[0] ** 5 * x[1] * p[4] - x[0] * x[1] * p34 * p[4] - x[0] ** 4 * x[1] * p[6] + x[0] ** 4 * x[2] * p[3] - x[1] * p34 * p[6]
            + x[2] * p34 * p[3] + x[0] ** 4 * p[0] + x[0] ** 4 * p[1] + p[0] * p34,
            x[0] * x[1] * p[4] - x[2] * p[5] * p[7] - x[2] * p[3] - x[2] * p[6],
        ]
    return make_system(build, 3, 8)


def biochem4() -> System:
    return make_system(lambda x, p: [-x[0] ** 3 * p[2] - x[0] * p[1] ** 2 * p[2] + x[0] ** 2 * p[0]], 1, 3)
