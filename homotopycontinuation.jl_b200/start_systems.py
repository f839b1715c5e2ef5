"""Host-side start systems (north star item 4: generated on the host and streamed in).

total_degree: reference src/total_degree.jl:32-117 (gamma :35, scaling :62, G = s.*(x.^D .- 1) :102,
              StraightLineHomotopy(G, F; gamma) :106), start solutions :235-266.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass

import numpy as np

from .modelkit import System, make_system


@dataclass
class TotalDegreeStart:
    F: System
    G: System
    degrees: np.ndarray
    scaling: np.ndarray        # parameters of G
    gamma: complex
    target_parameters: np.ndarray | None

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
        return out


def total_degree(F: System, gamma: complex, target_parameters=None) -> TotalDegreeStart:
    p = () if target_parameters is None else list(np.asarray(target_parameters, dtype=np.complex128))
    supports, coeffs = F.support_coefficients(p)
    if F.n_eqs != F.n_vars:
        raise NotImplementedError("only square affine systems (SURVEY.md section 8: all five configs)")
    # src/total_degree.jl:48-61: a system whose every polynomial has all its monomials of one degree is homogeneous; the
    # reference then tracks on an affine chart with G = s .* (x[1:n-1].^D .- x[n].^D).  Charts are outside this path
    # (SURVEY.md section 8f-3), so say so instead of silently building the affine start system.
    if all(len(set(int(d) for d in A.sum(axis=0))) == 1 for A in supports):
        raise NotImplementedError("homogeneous system: the reference puts it on an affine chart (src/total_degree.jl:94-108), "
                                  "which this path does not build")
    D = np.array([int(A.sum(axis=0).max()) for A in supports], dtype=np.int64)
    scaling = np.array([np.abs(c).max() for c in coeffs], dtype=np.float64)
    n = F.n_vars
    G = make_system(lambda x, s: [s[i] * (x[i] ** int(D[i]) - 1) for i in range(n)], n, n)
    tp = None if target_parameters is None else np.asarray(target_parameters, dtype=np.complex128)
    return TotalDegreeStart(F, G, D, scaling, complex(gamma), tp)
