"""Host-side polyhedral start system (north star item 4: start solutions are generated on the host
and streamed in).

What this restates (reference file:line):
  polyhedral(support, start_coeffs, target_coeffs)   src/polyhedral.jl:358-412
      origin added to every support lacking it (:375-394), F = polyhedral_system(support) with one
      parameter per term (:166-177), ToricHomotopy(F, start_coeffs), CoefficientHomotopy(F, p, q)
  PolyhedralStartSolutionsIterator                   src/polyhedral.jl:9-160
      one binomial system per mixed cell, `volume(cell)` start solutions each
  BinomialSystemSolver                               src/binomial_system.jl:45-106, 160-261
      x^A = b through the Hermite normal form A U = H: angular part by a triangular solve over the
      d^ = prod(H_ii) combinations of unit roots, modulus by a real linear solve
  update_weights!                                    src/homotopies/toric_homotopy.jl:66-97
      s_ij = w_ij - beta_i + <a_ij, normal>, 0 on the two cell vertices (the min/max rescaling is
      applied per path on the device, hc_lane.h set_weights)

The mixed cells themselves come from MixedSubdivisions.jl (`= "1"`, Project.toml:40), a dependency
that is not vendored under /root/reference: `fine_mixed_cells(support)` draws a random integer
lifting and enumerates the fine mixed cells of the induced regular subdivision.  Its published
contract is restated here: a mixed cell is one edge (a_i, b_i) per support together with the inner
normal `normal` (lifted normal (normal, 1)) such that
    <a_i, normal> + w_i[a_i] = <b_i, normal> + w_i[b_i] = beta_i < <c, normal> + w_i[c]   for all other c,
and volume(cell) = |det [a_i - b_i]_i|.  The enumeration below is a depth-first search over the
supports with linear-programming feasibility pruning.  No reference fixture pins cells or liftings
("parity unpinned" at this boundary); the checks are sum(volumes) == mixed
volume (70 / 924, test/polyhedral_test.jl:38-46) and the final solution counts.  In production the
Julia host passes MixedSubdivisions' cells through the C ABI instead.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from fractions import Fraction

import numpy as np

from .modelkit import System, system_from_support


# --------------------------------------------------------------------------- mixed cells
@dataclass
class MixedCell:
    indices: list          # [(a_i, b_i)] 0-based columns of support i
    normal: np.ndarray     # inner normal, float64 (n,)
    beta: np.ndarray       # (n,) min_j <A_i[:, j], normal> + w_i[j]
    volume: int


def _lp_feasible(A_eq, b_eq, A_ub, b_ub, n):
    from scipy.optimize import linprog
    if isinstance(A_eq, list) and len(A_eq):
        A_eq, b_eq = np.vstack(A_eq), np.concatenate(b_eq)
    r = linprog(np.zeros(n), A_ub=A_ub if len(A_ub) else None, b_ub=b_ub if len(b_ub) else None,
                A_eq=A_eq if len(A_eq) else None, b_eq=b_eq if len(b_eq) else None,
                bounds=[(None, None)] * n, method="highs")
    return r.status == 0


def mixed_cells(support, lifting) -> list[MixedCell]:
    """Fine mixed cells of the subdivision induced by `lifting` (see module docstring)."""
    n = len(support)
    S = [np.asarray(A, dtype=np.int64) for A in support]
    W = [np.asarray(w, dtype=np.int64) for w in lifting]
    assert all(A.shape[0] == n for A in S)
    order = sorted(range(n), key=lambda i: S[i].shape[1])  # fewest points first
    cells: list[MixedCell] = []

    def minimal_rows(i, a):
        """rows (c - a) . alpha >= w_a - w_c  as  -(c - a) . alpha <= w_c - w_a"""
        A, w = S[i], W[i]
        others = [c for c in range(A.shape[1]) if c != a]
        return -(A[:, others] - A[:, [a]]).T.astype(float), (w[others] - w[a]).astype(float)

    def leaf(chosen):
        idx = [None] * n
        for i, a, b in chosen:
            idx[i] = (a, b)
        E = np.array([S[i][:, idx[i][0]] - S[i][:, idx[i][1]] for i in range(n)], dtype=np.int64)
        rhs = np.array([W[i][idx[i][1]] - W[i][idx[i][0]] for i in range(n)], dtype=float)
        det = _int_det(E)
        if det == 0:
            return
        alpha = np.linalg.solve(E.astype(float), rhs)
        beta = np.empty(n)
        for i in range(n):
            vals = S[i].T.astype(float) @ alpha + W[i]
            a, b = idx[i]
            beta[i] = 0.5 * (vals[a] + vals[b])
            rest = np.delete(vals, [a, b])
            if rest.size and rest.min() <= beta[i] + 1e-9 * max(1.0, abs(beta[i])):
                return
        cells.append(MixedCell(idx, alpha, beta, abs(det)))

    def dfs(level, chosen, A_eq, b_eq, A_ub, b_ub):
        if level == n:
            leaf(chosen)
            return
        i = order[level]
        m = S[i].shape[1]
        feas = []
        if m == 2:
            feas = [0, 1]
        else:
            for a in range(m):
                Ra, ra = minimal_rows(i, a)
                if _lp_feasible(A_eq, b_eq, np.vstack(A_ub + [Ra]), np.concatenate(b_ub + [ra]), n):
                    feas.append(a)
        for x in range(len(feas)):
            for y in range(x + 1, len(feas)):
                a, b = feas[x], feas[y]
                Ra, ra = minimal_rows(i, a)
                e = (S[i][:, a] - S[i][:, b]).astype(float)[None, :]
                eb = np.array([float(W[i][b] - W[i][a])])
                Aeq2, beq2 = np.vstack(A_eq + [e]), np.concatenate(b_eq + [eb])
                if np.linalg.matrix_rank(Aeq2) < Aeq2.shape[0]:
                    continue
                Aub2, bub2 = A_ub + [Ra], b_ub + [ra]
                if level + 1 < n and not _lp_feasible(Aeq2, beq2, np.vstack(Aub2), np.concatenate(bub2), n):
                    continue
                dfs(level + 1, chosen + [(i, a, b)], A_eq + [e], b_eq + [eb], Aub2, bub2)

    dfs(0, [], [], [], [], [])
    cells.sort(key=lambda c: c.indices)
    return cells


def _int_det(M) -> int:
    """Exact determinant of an integer matrix (Bareiss)."""
    A = [[int(v) for v in row] for row in np.asarray(M)]
    n = len(A)
    sign, prev = 1, 1
    for k in range(n - 1):
        if A[k][k] == 0:
            sw = next((r for r in range(k + 1, n) if A[r][k] != 0), None)
            if sw is None:
                return 0
            A[k], A[sw] = A[sw], A[k]
            sign = -sign
        for i in range(k + 1, n):
            for j in range(k + 1, n):
                A[i][j] = (A[i][j] * A[k][k] - A[i][k] * A[k][j]) // prev
        prev = A[k][k]
    return sign * A[n - 1][n - 1]


def mixed_volume(support, seed: int = 0) -> int:
    rng = np.random.default_rng(seed)
    lifting = [rng.integers(-2 ** 12, 2 ** 12, size=np.asarray(A).shape[1]) for A in support]
    return sum(c.volume for c in mixed_cells(support, lifting))


# --------------------------------------------------------------------------- binomial systems
def hnf(A):
    """Column-style Hermite normal form A U = H, H lower triangular with positive diagonal and
    0 <= H[i, j] < H[i, i] for j < i (reference src/binomial_system.jl:267-373; exact Python
    integers, so the reference's overflow fallback to BigInt is not needed)."""
    n = len(A)
    H = [[int(v) for v in row] for row in np.asarray(A)]
    U = [[int(i == j) for j in range(n)] for i in range(n)]

    def colop(M, j, k, p, q, r, s):  # (col_j, col_k) <- (p col_j + q col_k, r col_j + s col_k)
        for row in M:
            a, b = row[j], row[k]
            row[j], row[k] = p * a + q * b, r * a + s * b

    for i in range(n):
        for k in range(i + 1, n):  # zero H[i, k] against the pivot column i
            if H[i][k] != 0:
                g, p, q = _gcdx(H[i][i], H[i][k])
                r, s = -H[i][k] // g, H[i][i] // g
                colop(H, i, k, p, q, r, s)
                colop(U, i, k, p, q, r, s)
        if H[i][i] < 0:
            for M in (H, U):
                for row in M:
                    row[i] = -row[i]
        if H[i][i] == 0:
            raise ZeroDivisionError("singular binomial system")
        for j in range(i):  # reduce the entries left of the diagonal
            f = H[i][j] // H[i][i]
            if f:
                for M in (H, U):
                    for row in M:
                        row[j] -= f * row[i]
    return np.array(H, dtype=object), np.array(U, dtype=object)


def _gcdx(a, b):
    """g = p a + q b with g = gcd(a, b) >= 0"""
    x0, y0, x1, y1 = 1, 0, 0, 1
    while b:
        q = a // b
        a, b = b, a - q * b
        x0, x1 = x1, x0 - q * x1
        y0, y1 = y1, y0 - q * y1
    if a < 0:
        a, x0, y0 = -a, -x0, -y0
    return a, x0, y0


def solve_binomial_system(A, b) -> np.ndarray:
    """All |det A| solutions x in (C*)^n of  prod_i x_i^{A[i, j]} = b_j  (columns of A = binomials).
    Returns an (n, d^) complex array like BinomialSystemSolver.X (src/binomial_system.jl:238-261)."""
    A = np.asarray(A)
    n = A.shape[0]
    b = np.asarray(b, dtype=np.complex128)
    H, U = hnf(A)
    diag = [int(H[i][i]) for i in range(n)]
    dhat = 1
    for d in diag:
        dhat *= d
    # angular part: x = e^{2 pi i alpha};  A^T alpha = gamma (mod 1)  <=>  H^T (U^-1 alpha) ... solve with
    # y = U^T-transformed angles exactly in rationals: (A U)^T alpha = U^T gamma (mod 1)
    gamma = [Fraction(float(np.angle(z)) / (2 * np.pi)) for z in b]
    mu = [sum(int(U[i][j]) * gamma[i] for i in range(n)) for j in range(n)]
    mu = [m - 2 * round(m / 2) for m in mu]   # rem(mu, 2, RoundNearest) as the reference (:88): fixes the order inside a cell
    X = np.empty((n, dhat), dtype=np.complex128)
    # unit-root combinations in the reference's order (fill_unit_roots_combinations!, :55-70)
    table = np.zeros((n, dhat), dtype=np.int64)
    d, e = dhat, 1
    for i in range(n):
        d //= diag[i]
        k = 0
        for _ in range(e):
            for j in range(diag[i]):
                table[i, k:k + d] = j
                k += d
        e *= diag[i]
    for c in range(dhat):
        alpha = [Fraction(0)] * n
        for j in range(n - 1, -1, -1):  # H^T is upper triangular: sum_{k >= j} H[k, j] alpha_k = mu_j + root_j
            s = mu[j] + int(table[j, c])
            for k in range(j + 1, n):
                s -= int(H[k][j]) * alpha[k]
            a = s / diag[j]
            alpha[j] = a - 2 * round(a / 2)   # the reference keeps alpha in [-1, 1] (rem(alpha, 2, RoundNearest), :101);
            # the representative matters: alpha_k enters alpha_j through H[k, j] / H[j, j], so it fixes the order inside a cell
        for j in range(n):
            a = float(alpha[j])
            X[j, c] = complex(np.cos(2 * np.pi * a), np.sin(2 * np.pi * a))
    # modulus: A^T log|x| = log|b|
    r = np.exp(np.linalg.solve(A.T.astype(float), np.log(np.abs(b))))
    return X * r[:, None]


# --------------------------------------------------------------------------- the start system
@dataclass
class PolyhedralStart:
    F: System                      # polyhedral_system(support): one parameter per term
    support: list                  # n x m_i integer matrices (origin added)
    start_coeffs: np.ndarray       # flat, P
    target_coeffs: np.ndarray      # flat, P
    lifting: list
    cells: list
    offsets: np.ndarray = field(default=None)

    @property
    def n(self):
        return len(self.support)

    def n_paths(self) -> int:
        return sum(c.volume for c in self.cells)

    def start_solutions(self):
        """(starts (N, n) complex, cell_index (N,) int32): the order of PolyhedralStartSolutionsIterator."""
        xs, ci = [], []
        for k, cell in enumerate(self.cells):
            A = np.stack([self.support[i][:, a] - self.support[i][:, b] for i, (a, b) in enumerate(cell.indices)], axis=1)
            bb = np.array([-self.start_coeffs[self.offsets[i] + b] / self.start_coeffs[self.offsets[i] + a]
                           for i, (a, b) in enumerate(cell.indices)])
            X = solve_binomial_system(A, bb)
            assert X.shape[1] == cell.volume
            xs.append(X.T)
            ci += [k] * cell.volume
        return np.concatenate(xs, axis=0), np.array(ci, dtype=np.int32)

    def binomial_systems(self):
        """(A, b) of every mixed cell: columns of A = support differences, b_i = -c_i[b] / c_i[a] (binomial_system.jl:45-53)."""
        for cell in self.cells:
            A = np.stack([self.support[i][:, a] - self.support[i][:, b] for i, (a, b) in enumerate(cell.indices)], axis=1)
            bb = np.array([-self.start_coeffs[self.offsets[i] + b] / self.start_coeffs[self.offsets[i] + a]
                           for i, (a, b) in enumerate(cell.indices)])
            yield A, bb

    def binomial_data(self) -> dict:
        """Per-cell inputs of hc_polyhedral_track_cells (start solutions made on the device): volume, Hermite normal
        form H (A U = H), mu = rem(U^T angle(b) / 2 pi, 2) and r = exp(A^-T log|b|) -- the host half of
        BinomialSystemSolver (hnf!, the coordinate change of compute_angular_part!, the modulus solve of solve!,
        src/binomial_system.jl:72-90, 238-261); the device does the d^ triangular solves."""
        vol, Hs, mus, rs = [], [], [], []
        for (A, bb), cell in zip(self.binomial_systems(), self.cells):
            H, U = hnf(A)
            n = A.shape[0]
            gamma = [Fraction(float(np.angle(z)) / (2 * np.pi)) for z in bb]
            mu = []
            for j in range(n):
                m = sum(int(U[i][j]) * gamma[i] for i in range(n))
                m -= 2 * round(m / 2)   # rem(mu, 2, RoundNearest)
                mu.append(float(m))
            vol.append(int(cell.volume)); Hs.append(np.array(H.tolist(), dtype=np.int64)); mus.append(mu)
            rs.append(np.exp(np.linalg.solve(A.T.astype(float), np.log(np.abs(bb)))))
        return {"volume": np.array(vol, dtype=np.int64), "H": np.array(Hs, dtype=np.int64), "mu": np.array(mus), "r": np.array(rs)}

    def cell_weights(self) -> np.ndarray:
        """ncells x P raw weights s_ij (0 on each cell's two vertices), toric_homotopy.jl:76-97."""
        P = len(self.start_coeffs)
        out = np.zeros((len(self.cells), P))
        for k, cell in enumerate(self.cells):
            for i, A in enumerate(self.support):
                s = np.asarray(self.lifting[i], dtype=float) - cell.beta[i] + A.T.astype(float) @ cell.normal
                a, b = cell.indices[i]
                s[a] = s[b] = 0.0
                out[k, self.offsets[i]:self.offsets[i] + A.shape[1]] = s
        return out


def polyhedral(F: System, target_parameters=None, seed_coeffs: int = 7, seed_origin: int = 8, seed_lifting: int = 9,
               cache: str | None = None) -> PolyhedralStart:
    """polyhedral(F) with every random input made explicit (SURVEY.md 8c "RNG"): start coefficients
    (0.9 + 0.2u) cis(2 pi v) |c|_inf (src/polyhedral.jl:344, 354), CN(0,1) start coefficients of the
    added origins (:386), integer lifting.  `cache`: JSON file holding lifting + cells (the
    enumeration is the slow part and is deterministic in the seeds)."""
    p = () if target_parameters is None else list(np.asarray(target_parameters, dtype=np.complex128))
    supports, tcoeffs = F.support_coefficients(p)
    n = F.n_vars
    if F.n_eqs != n:
        raise NotImplementedError("only square systems")
    r1, r2 = np.random.default_rng(seed_coeffs), np.random.default_rng(seed_origin)
    supp, sc, tc = [], [], []
    for A, c in zip(supports, tcoeffs):
        m = A.shape[1]
        s = (0.9 + 0.2 * r1.random(m)) * np.exp(2j * np.pi * r1.random(m)) * np.abs(c).max()
        if not (A == 0).all(axis=0).any():
            A = np.concatenate([A, np.zeros((n, 1), dtype=A.dtype)], axis=1)
            s = np.append(s, (r2.normal() + 1j * r2.normal()) / np.sqrt(2))
            c = np.append(c, 0.0)
        supp.append(A); sc.append(s); tc.append(c)
    offsets = np.concatenate([[0], np.cumsum([A.shape[1] for A in supp])]).astype(int)
    lifting = cells = None
    if cache and os.path.exists(cache):
        d = json.load(open(cache))
        if d["support"] == [A.tolist() for A in supp]:
            lifting = [np.array(w) for w in d["lifting"]]
            cells = [MixedCell([tuple(e) for e in c["indices"]], np.array(c["normal"]), np.array(c["beta"]), c["volume"]) for c in d["cells"]]
    if cells is None:
        rl = np.random.default_rng(seed_lifting)
        lifting = [rl.integers(-2 ** 12, 2 ** 12, size=A.shape[1]) for A in supp]
        cells = mixed_cells(supp, lifting)
        if cache:
            json.dump({"support": [A.tolist() for A in supp], "lifting": [w.tolist() for w in lifting],
                       "cells": [{"indices": [list(map(int, e)) for e in c.indices], "normal": c.normal.tolist(),
                                  "beta": c.beta.tolist(), "volume": int(c.volume)} for c in cells]}, open(cache, "w"))
    return PolyhedralStart(system_from_support(supp, n), supp, np.concatenate(sc), np.concatenate(tc).astype(np.complex128),
                           lifting, cells, offsets)
