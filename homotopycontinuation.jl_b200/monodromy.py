"""Batched monodromy loops: the host driver of `monodromy_solve` restated over the batch calls.

The reference walks one solution at a time around a parameter loop p -> p1 -> p2 -> p (reference
src/monodromy.jl:1509-1583 `track(egtracker, r, loop, ...)`: two plain `Tracker` segments that hand omega / mu on,
then the `EndgameTracker` back to p), adds the endpoint if it is a new nonsingular solution (`add_tracked_result!`,
:1176-1200, a `UniquePoints` lookup), and queues every new solution on the same loop again (:1188-1196).  Paths are
independent, so here a whole queue is ONE batch per segment -- three `hc_track_batch` calls per round -- and the
lookup of the endpoints against the known solutions is one call of the device-side duplicate filter
(`hc_unique_points_filter`, csrc/hc_api.cu) followed by a short host pass over the survivors.

Stopping rules of the reference (MonodromyOptions, src/monodromy.jl:26-120): `target_solutions_count`,
`max_loops_no_progress` (5), optional `single_loop_per_start_solution`.  Loops: p1, p2 ~ CN(0, 1)^m (the
reference's default `independent_normal` sampler, :224-232).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import capi


@dataclass
class MonodromyResult:
    returncode: str                 # "success" (target count reached) | "heuristic_stop" (no progress) | "invalid_startvalue"
    solutions: np.ndarray           # (count, n)
    parameters: np.ndarray
    loops: int = 0
    tracked_loops: int = 0          # loops tracked successfully, over all solutions (MonodromyStatistics.tracked_loops)
    tracked_loops_after_last_success: int = 0
    batches: list = field(default_factory=list)   # paths per round (what one segment call carried)


def unique_filter(api: capi.CApi, known: np.ndarray, cand: np.ndarray, atol: float = 1e-14, rtol: float = 1e-8) -> np.ndarray:
    """For every candidate the index of the first `known` point within max(atol, rtol * norm(candidate)) (Euclidean
    norm; UniquePoints' default metric, reference src/unique_points.jl:247-285), or -1.  Runs on the device when the
    library has `hc_unique_points_filter` (libhc_b200), in numpy otherwise (the oracle has no such entry point)."""
    known = np.ascontiguousarray(known, dtype=np.complex128)
    cand = np.ascontiguousarray(cand, dtype=np.complex128)
    N = cand.shape[0]
    out = np.full(N, -1, dtype=np.int64)
    if N == 0 or known.shape[0] == 0:
        return out
    fn = getattr(api, "_unique_points_filter", None)
    if fn is not None:
        rc = fn(known.shape[1], known.shape[0], known.view(np.float64).ctypes.data_as(capi.c_double_p), N,
                cand.view(np.float64).ctypes.data_as(capi.c_double_p), atol, rtol, out.ctypes.data_as(capi.c_int64_p))
        if rc:
            raise RuntimeError("unique_points_filter failed: " + capi._last_error(api))
        return out
    for i in range(N):
        d = np.linalg.norm(known - cand[i][None, :], axis=1)
        hit = np.flatnonzero(d <= max(atol, rtol * float(np.linalg.norm(cand[i]))))
        if len(hit):
            out[i] = hit[0]
    return out


def monodromy_solve(api: capi.CApi, F, start_solutions, p, target_solutions_count: int | None = None,
                    max_loops_no_progress: int = 5, single_loop_per_start_solution: bool = False, seed: int = 0x42,
                    options: capi.Options | None = None, max_loops: int = 1000) -> MonodromyResult:
    """All solutions of F(x; p) = 0 reachable from `start_solutions` by monodromy (reference `monodromy_solve(F, sols, p)`)."""
    p = np.asarray(p, dtype=np.complex128).reshape(-1)
    S = np.asarray(start_solutions, dtype=np.complex128).reshape(-1, F.n_vars)
    rng = np.random.default_rng(seed)
    hF = api.system(F)
    H = api.homotopy(capi.H_PARAMETER, hF, p=p, q=p)
    can_set = getattr(api, "_homotopy_set_parameters", None) is not None

    def segment(a, b):
        """the homotopy of one loop edge: parameters!(tracker, a, b) -- or a new handle where the library has no setter (oracle)"""
        if can_set:
            H.set_parameters(p=a, q=b)
            return H
        return api.homotopy(capi.H_PARAMETER, hF, p=a, q=b)
    opts = options if options is not None else api.default_options()
    # the start solutions must be solutions: one Newton-refining endgame track p -> p (t from 1 to 0 at fixed parameters)
    r0 = H.track_batch(S, options=opts)
    ok = (r0.return_code == 1) & (r0.singular == 0)
    if not ok.any():
        return MonodromyResult("invalid_startvalue", np.zeros((0, F.n_vars), np.complex128), p)
    sols = np.zeros((0, F.n_vars), dtype=np.complex128)
    omega_mu = np.zeros((0, 2))

    def add_new(X, om):
        """adds the rows of X that are neither known nor repeated inside X; returns their indices in X"""
        nonlocal sols, omega_mu
        hit = unique_filter(api, sols, X)
        fresh = []
        for i in np.flatnonzero(hit < 0):     # survivors: compare among themselves in order (add! is sequential)
            v = X[i]
            rad = max(1e-14, 1e-8 * float(np.linalg.norm(v)))
            if fresh and np.linalg.norm(X[fresh] - v[None, :], axis=1).min() <= rad:
                continue
            fresh.append(int(i))
        if fresh:
            sols = np.concatenate([sols, X[fresh]])
            omega_mu = np.concatenate([omega_mu, om[fresh]])
        return fresh

    add_new(r0.solution[ok], np.stack([r0.omega[ok], r0.mu[ok]], axis=1))
    res = MonodromyResult("heuristic_stop", sols, p)
    loops_no_progress = 0
    done = lambda: target_solutions_count is not None and len(sols) >= target_solutions_count
    while not done() and loops_no_progress < max_loops_no_progress and res.loops < max_loops:
        if res.loops > 0 and single_loop_per_start_solution:
            break
        m = len(p)
        p1 = (rng.normal(size=m) + 1j * rng.normal(size=m)) / np.sqrt(2)
        p2 = (rng.normal(size=m) + 1j * rng.normal(size=m)) / np.sqrt(2)
        res.loops += 1
        queue = np.arange(len(sols))
        progress = False
        while len(queue) and not done():
            X, om = sols[queue], omega_mu[queue]
            res.batches.append(len(queue))
            alive = np.arange(len(queue))
            for a, b, mode in ((p, p1, 1), (p1, p2, 1), (p2, p, 0)):
                r = segment(a, b).track_batch(X, options=opts, mode=mode, omega_mu=om)
                good = r.return_code == 1
                if mode == 0:
                    good &= r.singular == 0
                alive, X, om = alive[good], r.solution[good], np.stack([r.omega[good], r.mu[good]], axis=1)
                if len(alive) == 0:
                    break
            res.tracked_loops += len(alive)
            res.tracked_loops_after_last_success += len(alive)
            if len(alive) == 0:
                break
            first_new = len(sols)
            fresh = add_new(X, om)
            if fresh:
                progress = True
                res.tracked_loops_after_last_success = 0
                queue = np.arange(first_new, len(sols)) if not single_loop_per_start_solution else np.zeros(0, dtype=int)
            else:
                queue = np.zeros(0, dtype=int)
        loops_no_progress = 0 if progress else loops_no_progress + 1
    segment(p, p)
    res.solutions = sols
    res.returncode = "success" if done() else "heuristic_stop"
    return res
