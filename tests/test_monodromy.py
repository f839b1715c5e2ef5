"""Batched monodromy loops: the loop driver of hcb200.monodromy over the batch calls, and the
duplicate filter hc_unique_points_filter.  Known answer of the reference's own test (test/monodromy_test.jl:4-31): the
Euclidean-distance-degree system of the toric variety of A = [3 2 1 0; 0 1 2 3] has 21 solutions for generic
parameters; `monodromy_solve` finds them from one start pair, stops with `success` when told the count and with a
heuristic stop otherwise."""
import numpy as np
import pytest

from hcb200 import monodromy
from hcb200.modelkit import make_system


def toric_ed():
    """F(t, y; u) = [phi(t) + y - u; Dphi(t)' y], phi = (t1^3, t1^2 t2, t1 t2^2, t2^3)  (monodromy_test.jl:4-12)"""
    def build(v, p):
        t1, t2, y = v[0], v[1], v[2:6]
        phi = [t1 ** 3, t1 ** 2 * t2, t1 * t2 ** 2, t2 ** 3]
        d1 = [3 * t1 ** 2, 2 * t1 * t2, t2 ** 2, 0 * t1]
        d2 = [0 * t1, t1 ** 2, 2 * t1 * t2, 3 * t2 ** 2]
        return [phi[j] + y[j] - p[j] for j in range(4)] + [sum(d1[j] * y[j] for j in range(4)), sum(d2[j] * y[j] for j in range(4))]
    return make_system(build, 6, n_params=4)


def start_pair(seed=1):
    """a solution and its parameters: t random, y in the null space of Dphi(t)', u = phi(t) + y (find_start_pair)"""
    rng = np.random.default_rng(seed)
    t1, t2 = rng.normal(size=2) + 1j * rng.normal(size=2)
    phi = np.array([t1 ** 3, t1 ** 2 * t2, t1 * t2 ** 2, t2 ** 3])
    D = np.array([[3 * t1 ** 2, 0], [2 * t1 * t2, t1 ** 2], [t2 ** 2, 2 * t1 * t2], [0, 3 * t2 ** 2]])
    _, _, Vh = np.linalg.svd(D.T)           # D' y = 0 (plain transpose, as in the system)
    null = Vh[2:].conj().T
    y = null @ (rng.normal(size=2) + 1j * rng.normal(size=2))
    assert np.abs(D.T @ y).max() < 1e-12
    return np.concatenate([[t1, t2], y]), phi + y


def check(api):
    F = toric_ed()
    x0, p0 = start_pair()
    r = monodromy.monodromy_solve(api, F, [x0], p0, target_solutions_count=21, max_loops_no_progress=50)
    assert r.returncode == "success" and len(r.solutions) == 21
    # all different, all solutions of F(x; p0) = 0
    d = np.abs(r.solutions[:, None, :] - r.solutions[None, :, :]).max(axis=2) + np.eye(21)
    assert d.min() > 1e-6
    H = api.homotopy(1, api.system(F), p=p0, q=p0)
    for x in r.solutions:
        assert np.abs(H.evaluate(x, 0.0)).max() < 1e-10
    # same seed, same run (monodromy_test.jl:33-42)
    r2 = monodromy.monodromy_solve(api, F, [x0], p0, target_solutions_count=21, max_loops_no_progress=50)
    assert r2.tracked_loops == r.tracked_loops and r2.loops == r.loops
    # without the count: heuristic stop after 5 loops without a new solution, still 21
    r3 = monodromy.monodromy_solve(api, F, [x0], p0)
    assert r3.returncode == "heuristic_stop" and len(r3.solutions) == 21
    # a point that is no solution
    bad = monodromy.monodromy_solve(api, F, [x0 + 0.3], p0, target_solutions_count=21)
    assert bad.returncode == "invalid_startvalue"
    return r


def test_monodromy_on_the_oracle(oracle):
    check(oracle)


def test_monodromy_on_the_host_compiled_device_code(sim):
    r = check(sim)
    assert max(r.batches) > 1   # later rounds carry many paths per call


def test_duplicate_filter_host_build(sim):
    _filter_case(sim)


def _filter_case(api):
    rng = np.random.default_rng(4)
    known = rng.normal(size=(300, 5)) + 1j * rng.normal(size=(300, 5))
    cand = rng.normal(size=(200, 5)) + 1j * rng.normal(size=(200, 5))
    pick = rng.integers(0, 300, size=80)
    cand[:80] = known[pick] * (1 + 1e-10 * rng.normal(size=(80, 1)))     # inside the radius 1e-8 ||v||
    cand[80:100] = known[pick[:20]] * (1 + 1e-6)                          # outside
    known[7] = known[3]                                                   # a repeated known point: the FIRST one matches
    cand[100] = known[7]
    got = monodromy.unique_filter(api, known, cand)
    assert api._unique_points_filter is not None
    want = np.full(200, -1)
    for i in range(200):
        d = np.linalg.norm(known - cand[i], axis=1)
        hit = np.flatnonzero(d <= max(1e-14, 1e-8 * np.linalg.norm(cand[i])))
        want[i] = hit[0] if len(hit) else -1
    assert (got == want).all() and (got[:80] >= 0).all() and (got[80:100] < 0).all() and got[100] == 3


@pytest.mark.gpu
def test_monodromy_on_the_gpu(gpu):
    check(gpu)
    _filter_case(gpu)
