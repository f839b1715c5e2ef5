"""Systems of the reference's golden test cases, restated for the tape compiler."""
import json
import os

import numpy as np

import hcb200
from hcb200.modelkit import make_system

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    d = json.load(open(os.path.join(GOLDEN, name)))
    return {k: np.array([complex(a, b) for a, b in v]) for k, v in d.items()}


def steiner_higher_prec():
    """reference test/test_cases/steiner_higher_prec.jl:106-139: variables [a; vec(y)], parameters vec(v)."""
    def build(X, V):
        a = X[0:5]
        eqs = []
        for i in range(5):
            x = [X[5 + 2 * i], X[5 + 2 * i + 1]]
            c = V[6 * i: 6 * i + 6]
            f = a[0] * x[0] ** 2 + a[1] * x[0] * x[1] + a[2] * x[1] ** 2 + a[3] * x[0] + a[4] * x[1] + 1
            g = c[0] * x[0] ** 2 + c[1] * x[0] * x[1] + c[2] * x[1] ** 2 + c[3] * x[0] + c[4] * x[1] + c[5]
            df = [2 * a[0] * x[0] + a[1] * x[1] + a[3], a[1] * x[0] + 2 * a[2] * x[1] + a[4]]
            dg = [2 * c[0] * x[0] + c[1] * x[1] + c[3], c[1] * x[0] + 2 * c[2] * x[1] + c[4]]
            eqs += [f, g, df[0] * dg[1] - df[1] * dg[0]]
        return eqs
    return make_system(build, 15, 30), _load("steiner_higher_prec.json")


def four_bar():
    """reference test/test_cases/four_bar.jl:2-25: variables [x,a,y,b,x^,a^,y^,b^,gamma(8),gamma^(8)],
    parameters [delta(8); delta^(8)]."""
    def build(X, P):
        x, a, y, b, xh, ah, yh, bh = X[0:8]
        gam, gamh = X[8:16], X[16:24]
        dl, dlh = P[0:8], P[8:16]
        D1 = [(ah * x - dlh[i] * x) * gam[i] + (a * xh - dl[i] * xh) * gamh[i] + (ah - xh) * dl[i]
              + (a - x) * dlh[i] - dl[i] * dlh[i] for i in range(8)]
        D2 = [(bh * y - dlh[i] * y) * gam[i] + (b * yh - dl[i] * yh) * gamh[i] + (bh - yh) * dl[i]
              + (b - y) * dlh[i] - dl[i] * dlh[i] for i in range(8)]
        D3 = [gam[i] * gamh[i] + gam[i] + gamh[i] for i in range(8)]
        return D1 + D2 + D3
    return make_system(build, 24, 16), _load("four_bar.json")


# ---- systems of reference test/endgame_test.jl and test/tracker_test.jl (known answers, no fixtures needed)
def hyperbolic_6_6():
    """test/endgame_test.jl:7-24 (y = 1): two roots of multiplicity 6, winding number 3 on all 12 paths."""
    return make_system(lambda v, p: [0.75 * v[0] ** 4 + 1.5 * v[0] ** 2 - 2.5 * v[0] ** 2 * v[1] ** 2 + 0.75
                                     - 2.5 * v[1] ** 2 + 0.75 * v[1] ** 4,
                                     10 * v[0] ** 2 * v[1] + 10 * v[1] - 6 * v[1] ** 3], 2)


def singular_1():
    """test/endgame_test.jl:26-38 (z = 1): one singular solution of multiplicity 3 and one nonsingular."""
    return make_system(lambda v, p: [v[0] ** 2 + 2 * v[1] ** 2 + 2j * v[1],
                                     (18 + 3j) * v[0] * v[1] + 7j * v[1] ** 2 - (3 - 18j) * v[0] - 14 * v[1] - 7j], 2)


def wilkinson(d: int):
    """test/endgame_test.jl:40-47: expand(prod(x - i for i = 1:d)); the coefficients are exact in fp64 for d = 12."""
    c = [1]
    for i in range(1, d + 1):
        c = [(c[k - 1] if k > 0 else 0) - i * (c[k] if k < len(c) else 0) for k in range(len(c) + 1)]
    return make_system(lambda v, p: [sum(float(c[k]) * v[0] ** k for k in range(1, d + 1)) + float(c[0])], 1)


def winding_number_family(d: int):
    """test/endgame_test.jl:63-70."""
    a = [0.257, -0.139, -1.73, -0.199, 1.79, -1.32]
    return make_system(lambda v, p: [(a[0] * v[0] ** d + a[1] * v[1]) * (a[2] * v[0] + a[3] * v[1]) + 1,
                                     (a[0] * v[0] ** d + a[1] * v[1]) * (a[4] * v[0] + a[5] * v[1]) + 1], 2)


def mohab():
    """test/endgame_test.jl:77-110, variable order [x, z, y]: 693 nonsingular solutions out of 900 paths."""
    def f(v, p):
        x, z, y = v
        return [-9091098778555951517 * x ** 3 * y ** 4 * z ** 2 + 5958442613080401626 * y ** 2 * z ** 7
                + 17596733865548170996 * x ** 2 * z ** 6 - 17979170986378486474 * x * y * z ** 6
                - 2382961149475678300 * x ** 4 * y ** 3 - 15412758154771986214 * x * y ** 3 * z ** 3 + 133,
                -10798198881812549632 * x ** 6 * y ** 3 * z - 11318272225454111450 * x * y ** 9
                - 14291416869306766841 * y ** 9 * z - 5851790090514210599 * y ** 2 * z ** 8
                + 15067068695242799727 * x ** 2 * y ** 3 * z ** 4 + 7716112995720175148 * x ** 3 * y * z ** 3 + 171,
                13005416239846485183 * x ** 7 * y ** 3 + 4144861898662531651 * x ** 5 * z ** 4
                - 8026818640767362673 * x ** 6 - 6882178109031199747 * x ** 2 * y ** 4
                + 7240929562177127812 * x ** 2 * y ** 3 * z + 5384944853425480296 * x * y * z ** 4 + 88]
    return make_system(f, 3)


MOHAB_GAMMA = -0.9132549847010242 + 0.4073884300256109j   # "path jumping happened with too loose default config"


def pinned_framework():
    """test/tracker_test.jl:137-219 (issue 454): a planar bar framework with three pinned vertices, sliced by
    sum of the free y-coordinates = b.  The initial configuration p0 is a singular point of the system.
    Returns (system, start, b0)."""
    p0 = np.array([[0, 0], [1, 0], [1, 1], [2, 1], [1.5, 0.5], [3, 1], [4, 1], [5, 1], [4.5, 0.5], [5, 0], [6, 0]], float)
    E = [(1, 2), (2, 3), (2, 5), (3, 4), (3, 5), (4, 5), (4, 6), (5, 9), (6, 7), (7, 8), (7, 9), (8, 9), (8, 10), (9, 10), (10, 11)]
    pinned, free = [1, 6, 11], [2, 3, 4, 5, 7, 8, 9, 10]
    idx = {(i, k): 2 * j + k for j, i in enumerate(free) for k in range(2)}

    def f(v, p):
        X = lambda i, k: float(p0[i - 1, k]) if i in pinned else v[idx[(i, k)]]
        eqs = [sum((X(i, k) - X(j, k)) ** 2 for k in range(2)) - float(sum((p0[i - 1, k] - p0[j - 1, k]) ** 2 for k in range(2)))
               for i, j in E]
        eqs.append(sum(v[idx[(i, 1)]] for i in free) - p[0])
        return eqs
    start = np.array([p0[i - 1, k] for i in free for k in range(2)], complex)
    return make_system(f, 16, 1), start, float(sum(p0[i - 1, 1] for i in free))
