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
