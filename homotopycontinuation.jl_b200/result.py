"""Host-side mirror of the reference's `Result` post-processing: multiplicity clustering of the singular
endpoints and the `ResultStatistics` counts ("solution classes") that BASELINE.json asks to be identical.

Follows (reference file:line):
  src/unique_points.jl:247-285, 314-346   UniquePoints.add! / multiplicities: a point joins the first stored point
                                          within max(atol, rtol * norm(v)) (Euclidean norm, atol 1e-14, rtol 1e-8)
  src/result.jl:46-84, 102-110            MultiplicityInfo over the *singular* path results only; every member of a
                                          cluster but its first is a "multiple result" and is skipped when counting
  src/result.jl:146-212                   ResultStatistics
  src/path_result.jl:223-299              is_at_infinity / is_failed / is_singular / is_real
The clustering stays on the host in the reference too (it needs all endpoints at once); the GPU path only
delivers the PathResult fields it reads.
"""
from __future__ import annotations

from dataclasses import asdict, dataclass

import numpy as np

# EndgameTrackerCode values (src/endgame_tracker.jl:100-117)
SUCCESS, AT_INFINITY, AT_ZERO, EXCESS_SOLUTION = 1, 2, 3, 14


def multiplicities(V: np.ndarray, atol: float = 1e-14, rtol: float = 1e-8) -> list[list[int]]:
    """Indices of points that coincide, grouped (reference `multiplicities(solution, results)`)."""
    V = np.asarray(V)
    reps: list[int] = []          # index of the first point of every cluster
    clusters: dict[int, list[int]] = {}
    for i in range(len(V)):
        v = V[i]
        rad = max(atol, rtol * float(np.linalg.norm(v)))
        found = -1
        if reps:
            d = np.linalg.norm(V[reps] - v[None, :], axis=1)
            k = int(np.argmin(d))
            if d[k] <= rad:
                found = reps[k]
        if found < 0:
            reps.append(i)
        else:
            clusters.setdefault(found, [found]).append(i)
    return list(clusters.values())


@dataclass
class ResultStatistics:
    total: int
    nonsingular: int
    singular: int
    singular_with_multiplicity: int
    real: int
    real_nonsingular: int
    real_singular: int
    real_singular_with_multiplicity: int
    at_infinity: int
    excess_solution: int
    failed: int

    def asdict(self) -> dict:
        return asdict(self)


def is_real(solution: np.ndarray, atol: float = 1e-6, rtol: float = 0.0) -> np.ndarray:
    m = np.abs(solution.imag).max(axis=1)
    out = m < atol
    if rtol != 0.0:
        out |= m < rtol * np.abs(solution).sum(axis=1)
    return out


def statistics(res, real_atol: float = 1e-6, real_rtol: float = 0.0) -> ResultStatistics:
    """ResultStatistics(Result(path_results)) of a batch in EndgameTrackerCode order (mode 0 / polyhedral)."""
    rc = np.asarray(res.return_code)
    success = rc == SUCCESS
    at_inf = (rc == AT_INFINITY) | (rc == AT_ZERO)
    excess = rc == EXCESS_SOLUTION
    failed = ~(success | at_inf | excess)
    singular = success & (np.asarray(res.singular) != 0)
    sing_idx = np.flatnonzero(singular)
    mult = np.ones(len(rc), dtype=np.int64)
    multiple = np.zeros(len(rc), dtype=bool)
    wn = np.asarray(res.winding_number)
    for cluster in multiplicities(res.solution[sing_idx]):
        members = sing_idx[cluster]
        multiple[members[1:]] = True
        for m in members:  # assign_multiplicities!: max(k, winding_number)
            mult[m] = max(len(cluster), int(wn[m]) if wn[m] > 0 else 1)
    real = is_real(np.asarray(res.solution), real_atol, real_rtol)
    keep = ~multiple
    s = singular & keep
    ns = success & ~singular & keep
    return ResultStatistics(
        total=int(len(rc)), nonsingular=int(ns.sum()), singular=int(s.sum()), singular_with_multiplicity=int(mult[s].sum()),
        real=int((real & (s | ns)).sum()), real_nonsingular=int((real & ns).sum()), real_singular=int((real & s).sum()),
        real_singular_with_multiplicity=int(mult[real & s].sum()), at_infinity=int((at_inf & keep).sum()),
        excess_solution=int((excess & keep).sum()), failed=int((failed & keep).sum()))
