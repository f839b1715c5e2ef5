"""Host-side start systems (north star item 4: generated on the host and streamed in).

total_degree: reference src/total_degree.jl:32-117 (gamma :35, scaling :62, G = s.*(x.^D .- 1) :102,
              StraightLineHomotopy(G, F; gamma) :106), start solutions :235-266.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass

import numpy as np

from .modelkit import Expr, System, make_system


@dataclass
class TotalDegreeStart:
    F: System
    G: System
    degrees: np.ndarray
    scaling: np.ndarray        # parameters of G
    gamma: complex
    target_parameters: np.ndarray | None
    chart: np.ndarray | None = None   # homogeneous input: F and G carry the row v'x - 1 of this affine chart
    original: System | None = None    # overdetermined input: the system before squaring up (F = [I A] original)
    A: np.ndarray | None = None       # ... and the randomisation matrix

    def n_paths(self) -> int:
        return int(np.prod(self.degrees))

    def start_solutions(self) -> np.ndarray:
        """All prod(D) starts x_i = cis(2 pi k_i / d_i); first index fastest
        (Iterators.product order, total_degree.jl:241, 258-262)."""
        D = self.degrees
        N = self.n_paths()
        out = np.empty((N, len(D)), dtype=np.complex128)
        idx = np.arange(N)
        for i, d in enumerate(D):
            k = idx % d
            idx = idx // d
            out[:, i] = np.exp(2j * np.pi * k / d)
            # exact values at the axes, as cis() gives in Julia
        if self.chart is not None:   # homogeneous: [x_1 .. x_{n-1}, 1] put on the chart v'x = 1 (total_degree.jl:262, on_chart!)
            out = np.concatenate([out, np.ones((N, 1), dtype=np.complex128)], axis=1)
            out /= (out @ self.chart)[:, None]
        return out


def total_degree(F: System, gamma: complex, target_parameters=None, chart_seed: int = 11) -> TotalDegreeStart:
    p = () if target_parameters is None else list(np.asarray(target_parameters, dtype=np.complex128))
    supports, coeffs = F.support_coefficients(p)
    # src/total_degree.jl:48-61, 94-108: a system whose every polynomial has all its monomials of one degree is homogeneous;
    # the reference then tracks on a random affine chart v'x = 1 (AffineChartHomotopy, src/homotopies/affine_chart_homotopy.jl
    # :42-98) with G = s .* (x[1:n-1].^D .- x[n].^D).  The chart is built into the systems here: F and G both get the row
    # v'x - 1 (in the straight-line homotopy that row is (gamma t + 1 - t)(v'x - 1): the same zero set; the reference keeps
    # it unscaled) -- a host-side construction, the device tracks an ordinary square system (first half).
    if all(len(set(int(d) for d in A.sum(axis=0))) == 1 for A in supports):
        return _total_degree_on_chart(F, gamma, target_parameters, supports, coeffs, chart_seed)
    if F.n_eqs > F.n_vars:
        return _total_degree_squared_up(F, gamma, target_parameters, supports, coeffs, chart_seed)
    if F.n_eqs != F.n_vars:
        raise NotImplementedError("underdetermined system")
    D = np.array([int(A.sum(axis=0).max()) for A in supports], dtype=np.int64)
    scaling = np.array([np.abs(c).max() for c in coeffs], dtype=np.float64)
    n = F.n_vars
    G = make_system(lambda x, s: [s[i] * (x[i] ** int(D[i]) - 1) for i in range(n)], n, n)
    tp = None if target_parameters is None else np.asarray(target_parameters, dtype=np.complex128)
    return TotalDegreeStart(F, G, D, scaling, complex(gamma), tp)


def _total_degree_on_chart(F: System, gamma, target_parameters, supports, coeffs, chart_seed: int) -> TotalDegreeStart:
    n = F.n_vars
    if F.n_eqs != n - 1:
        raise NotImplementedError("homogeneous system: n - 1 equations in n variables expected (overdetermined systems are squared up by the "
                                  "reference, src/total_degree.jl:72-88; not built here)")
    rng = np.random.default_rng(chart_seed)
    v = (rng.normal(size=n) + 1j * rng.normal(size=n)) / np.sqrt(2)   # randn(ComplexF64, n), affine_chart_homotopy.jl:35
    D = np.array([int(A.sum(axis=0).max()) for A in supports], dtype=np.int64)
    scaling = np.array([np.abs(c).max() for c in coeffs], dtype=np.float64)
    g = F.graph
    x = [Expr(g, g.var(i)) for i in range(n)]
    row = x[0] * complex(v[0])
    for i in range(1, n):
        row = row + x[i] * complex(v[i])
    Fc = System(g, list(F.exprs) + [(row - 1.0).i], n, F.n_params)

    def start(xs, s):
        r = xs[0] * complex(v[0])
        for i in range(1, n):
            r = r + xs[i] * complex(v[i])
        return [s[i] * (xs[i] ** int(D[i]) - xs[n - 1] ** int(D[i])) for i in range(n - 1)] + [r - 1.0]
    G = make_system(start, n, n - 1)
    tp = None if target_parameters is None else np.asarray(target_parameters, dtype=np.complex128)
    return TotalDegreeStart(Fc, G, D, scaling, complex(gamma), tp, chart=v)


def square_up(F: System, A: np.ndarray) -> System:
    """RandomizedSystem(F, A) with the identity block: [I A] F, the first n equations plus A times the other m - n
    (reference src/systems/randomized_system.jl:8-12, 40-47; built symbolically on the host, the device tracks a square system)."""
    n, m = F.n_vars, F.n_eqs
    g = F.graph
    E = [Expr(g, e) for e in F.exprs]
    rows = []
    for i in range(n):
        r = E[i]
        for j in range(m - n):
            r = r + E[n + j] * complex(A[i, j])
        rows.append(r.i)
    return System(g, rows, n, F.n_params)


def _total_degree_squared_up(F: System, gamma, target_parameters, supports, coeffs, seed: int) -> TotalDegreeStart:
    """Overdetermined affine system (reference src/total_degree.jl:66-92): equations sorted by descending degree, squared up
    with a random A, D = the n largest degrees, scaling = [I A] scaling.  The excess solutions the randomisation adds are
    marked afterwards (excess_solution_check)."""
    n, m = F.n_vars, F.n_eqs
    D = np.array([int(A.sum(axis=0).max()) for A in supports], dtype=np.int64)
    scaling = np.array([np.abs(c).max() for c in coeffs], dtype=np.float64)
    perm = np.argsort(-D, kind="stable")
    Fs = System(F.graph, [F.exprs[k] for k in perm], n, F.n_params)
    D, scaling = D[perm], scaling[perm]
    rng = np.random.default_rng(seed)
    A = (rng.normal(size=(n, m - n)) + 1j * rng.normal(size=(n, m - n))) / np.sqrt(2)
    Fr = square_up(Fs, A)
    sc = scaling[:n].astype(np.complex128) + A @ scaling[n:]
    Dn = D[:n]
    G = make_system(lambda x, s: [s[i] * (x[i] ** int(Dn[i]) - 1) for i in range(n)], n, n)
    tp = None if target_parameters is None else np.asarray(target_parameters, dtype=np.complex128)
    return TotalDegreeStart(Fr, G, Dn, sc, complex(gamma), tp, original=Fs, A=A)


def excess_solution_check(td: TotalDegreeStart, res, tol_factor: float = 100.0, tol_min: float = 1e-8):
    """Marks the endpoints that solve the squared-up system but not the original one with return code 14
    (EndgameTrackerCode.excess_solution).  The reference (src/overdetermined.jl:28-64) decides nonsingular endpoints with a
    Newton iteration on the overdetermined system and singular ones by their residual; this is the residual rule for
    both: ||F(x)||_inf > max(100 residual(path), 1e-8).  Host side (it runs once per solve on the successes only)."""
    if td.original is None:
        return res
    p = () if td.target_parameters is None else list(td.target_parameters)
    for k in np.flatnonzero(res.return_code == 1):
        r = np.abs(np.array(td.original.evaluate(list(res.solution[k]), p), dtype=np.complex128)).max()
        if r > max(tol_factor * float(res.residual[k]), tol_min):
            res.return_code[k] = 14
            res.residual[k] = r
    return res
