"""Known answers of the reference's own endgame / tracker tests (test/endgame_test.jl, test/tracker_test.jl:137-219),
checked on every backend behind the C ABI: the CPU oracle (pins the oracle), the device code compiled for the host
(kernel logic without a GPU) and, marked `gpu`, libhc_b200.so on a B200."""
import numpy as np
import pytest

import ref_systems
from helpers import straight_line
from hcb200 import capi

GAMMA = 0.4 + 1.3j   # reference test/endgame_tracker_test.jl:5


@pytest.fixture(params=["oracle", "sim", pytest.param("gpu", marks=pytest.mark.gpu)])
def api(request):
    return request.getfixturevalue(request.param)


def clusters(points, tol=1e-5, rtol=None):
    """Multiplicity clustering of endpoints (host side in the reference: src/result.jl:146-212)."""
    reps, counts = [], []
    for p in points:
        for i, q in enumerate(reps):
            if np.linalg.norm(p - q) < (tol if rtol is None else rtol * np.linalg.norm(q)):
                counts[i] += 1
                break
        else:
            reps.append(p); counts.append(1)
    return counts


def test_hyperbolic_6_6(api):  # test/endgame_test.jl:7-24
    td, H = straight_line(api, ref_systems.hyperbolic_6_6(), GAMMA)
    r = H.track_batch(td.start_solutions())
    assert td.n_paths() == 12
    assert (r.return_code == 1).all() and (r.winding_number == 3).all() and r.singular.all()
    assert sorted(clusters(r.solution)) == [6, 6]                      # 2 results of multiplicity 6
    assert np.allclose(np.abs(r.solution[:, 0]), 1, atol=1e-5) and np.allclose(r.solution[:, 1], 0, atol=1e-5)


def test_singular_1(api):  # test/endgame_test.jl:26-38
    td, H = straight_line(api, ref_systems.singular_1(), GAMMA)
    r = H.track_batch(td.start_solutions())
    assert (r.return_code == 1).all()
    assert int(r.singular.sum()) == 3 and sorted(clusters(r.solution[r.singular == 1])) == [3]   # 1 singular solution, multiplicity 3
    assert int((r.singular == 0).sum()) == 1                                                     # 1 nonsingular
    assert np.allclose(r.solution[r.singular == 1], [0, -1j], atol=1e-5)


def test_wilkinson_12(api):  # test/endgame_test.jl:40-47, endgame_options = (only_nonsingular = true,)
    td, H = straight_line(api, ref_systems.wilkinson(12), GAMMA)
    r = H.track_batch(td.start_solutions(), options=api.default_options(only_nonsingular=1))
    assert (r.return_code == 1).all()
    x = r.solution[:, 0]
    assert list(np.round(np.sort(x.real)).astype(int)) == list(range(1, 13))
    assert np.abs(x.imag).max() < 1e-4


@pytest.mark.parametrize("d", [2, 6])
def test_x_minus_10_to_the_d(api, d):  # test/endgame_test.jl:49-54
    from hcb200.modelkit import make_system
    td, H = straight_line(api, make_system(lambda v, p: [(v[0] - 10) ** d], 1), GAMMA)
    r = H.track_batch(td.start_solutions())
    assert int((r.winding_number == d).sum()) == d


@pytest.mark.parametrize("d", [2, 4, 6])
def test_winding_number_family(api, d):  # test/endgame_test.jl:63-70
    td, H = straight_line(api, ref_systems.winding_number_family(d), GAMMA)
    r = H.track_batch(td.start_solutions())
    assert td.n_paths() == (d + 1) ** 2
    assert int((r.return_code == 1).sum()) == d + 1


@pytest.mark.parametrize("gamma", [GAMMA, ref_systems.MOHAB_GAMMA])
def test_mohab_693(api, gamma):  # test/endgame_test.jl:77-110
    td, H = straight_line(api, ref_systems.mohab(), gamma)
    r = H.track_batch(td.start_solutions(), nthreads=8)
    assert td.n_paths() == 900 and list(td.degrees) == [9, 10, 10]
    ok = r.return_code == 1
    assert int((ok & (r.singular == 0)).sum()) == 693
    # all endpoints distinct, i.e. no path jumping (the closest pair of true solutions is 6e-9 apart, relatively)
    assert len(clusters(r.solution[ok], rtol=1e-10)) == 693


@pytest.mark.parametrize("db", [1.3e-3, -0.7e-3])
def test_invalid_startvalue_singular_jacobian(api, db):  # test/tracker_test.jl:137-219 (issue 454)
    F, start, b0 = ref_systems.pinned_framework()
    H = api.homotopy(capi.H_PARAMETER, api.system(F), p=[b0], q=[b0 + db])
    r = H.track_batch([start], mode=1)
    assert capi.TRACKER_CODES[r.return_code[0]] == "terminated_invalid_startvalue_singular_jacobian"


# ---- polyhedral start systems (mixed cells from hcb200.polyhedral with explicit seeds, SURVEY.md 8c "RNG")
def track_polyhedral(api, F):
    from hcb200 import polyhedral as ph
    ps = ph.polyhedral(F)
    S, ci = ps.start_solutions()
    h = api.system(ps.F)
    Ht = api.homotopy(capi.H_TORIC, h, p=ps.start_coeffs)
    Hc = api.homotopy(capi.H_COEFFICIENT, h, p=ps.start_coeffs, q=ps.target_coeffs)
    return ps, capi.polyhedral_track_batch(api, Ht, Hc, S, ci, ps.cell_weights())


def test_polyhedral_affine_and_torus_solutions(api):  # test/polyhedral_test.jl:2-9 (only_torus = false)
    from hcb200.modelkit import make_system
    f = make_system(lambda v, p: [2 * v[1] + 3 * v[1] ** 2 - v[0] * v[1] ** 3, v[0] + 4 * v[0] ** 2 - 2 * v[0] ** 3 * v[1]], 2)
    ps, r = track_polyhedral(api, f)
    assert ps.n_paths() == 8
    assert int((r.return_code == 1).sum()) == 6


def test_polyhedral_affine_square(api):  # test/solve_test.jl:94-101
    from helpers import system_2x2
    ps, r = track_polyhedral(api, system_2x2())
    assert int((r.return_code == 1).sum()) == 2
    assert sorted(capi.ENDGAME_CODES[c] for c in r.return_code) == ["at_infinity", "at_infinity", "success", "success"]


def test_many_parameters_solver(api):  # test/solve_test.jl:437-537: circle x line, 100 parameter points, 2 solutions each
    from hcb200.modelkit import make_system
    F = make_system(lambda v, p: [v[0] ** 2 + v[1] ** 2 - 1, p[0] * v[0] + p[1] * v[1] + p[2]], 2, 3)
    rng = np.random.default_rng(2024)
    p0 = (rng.normal(size=3) + 1j * rng.normal(size=3)) / np.sqrt(2)
    td, H0 = straight_line(api, F, GAMMA, p0)
    r0 = H0.track_batch(td.start_solutions())
    S0 = r0.solution[r0.return_code == 1]
    assert len(S0) == 2
    params = rng.random((100, 3)).astype(np.complex128)
    H = api.homotopy(capi.H_PARAMETER, api.system(F), p=p0, q=params[0])
    if getattr(api, "_track_sweep", None) is not None:
        r = capi.track_sweep(H, S0, params)                                     # many_solve entry point
    else:
        r = H.track_batch(np.tile(S0, (100, 1)), path_q=np.repeat(params, 2, axis=0))
    assert (r.return_code == 1).all() and not r.singular.any()
    sol = r.solution.reshape(100, 2, 2)
    assert (np.abs(sol[:, 0] - sol[:, 1]).max(axis=1) > 1e-8).all()              # 2 distinct solutions per point
    x, y = r.solution[:, 0], r.solution[:, 1]
    q = np.repeat(params, 2, axis=0)
    assert np.abs(x ** 2 + y ** 2 - 1).max() < 1e-10 and np.abs(q[:, 0] * x + q[:, 1] * y + q[:, 2]).max() < 1e-10


# ---- reference test/homotopies_test.jl: evaluate!, evaluate_and_jacobian!, taylor! K = 1..4 at a random complex t against
#      the symbolic homotopy (here: 60-digit arithmetic on the expression DAG), rtol 1e-12
@pytest.fixture(params=["oracle", "sim"])
def cpu_api(request):
    return request.getfixturevalue(request.param)


def _check_against_symbolic(H, h, n, rng):
    import mpmath as mp
    mp.mp.prec = 220
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    t = complex(rng.normal(), rng.normal())
    mx, mt = [mp.mpc(v) for v in x], mp.mpc(t)
    ref_u = np.array([complex(v) for v in h(mx, mt)])
    u, U = H.evaluate_and_jacobian(x, t)
    assert np.abs(H.evaluate(x, t) - ref_u).max() <= 1e-12 * np.abs(ref_u).max()
    assert np.abs(u - ref_u).max() <= 1e-12 * np.abs(ref_u).max()
    ref_U = np.array([[complex(mp.diff(lambda z: h(mx[:j] + [z] + mx[j + 1:], mt)[i], mx[j])) for j in range(n)] for i in range(n)])
    assert np.abs(U.reshape(n, n, order="F") - ref_U).max() <= 1e-11 * np.abs(ref_U).max()
    X = rng.normal(size=(4, n)) + 1j * rng.normal(size=(4, n))
    for K in (1, 2, 3, 4):
        def g(l, i):
            xl = [sum(mp.mpc(X[k][j]) * l ** k for k in range(K)) for j in range(n)]
            return h(xl, mt + l)[i]
        ref = np.array([complex(mp.taylor(lambda l: g(l, i), 0, K)[K]) for i in range(n)])
        got = H.taylor(K, X[:K], t)
        assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max(), K


def test_parameter_homotopy_operators(cpu_api):  # test/homotopies_test.jl:59-78
    from hcb200.modelkit import make_system
    f = lambda v, p: [(2 * v[0] ** 2 + p[1] ** 2 * v[1] ** 3 + 2 * p[0] * v[0] * v[1]) ** 3, (p[0] + p[2]) ** 4 * v[0] + v[1] ** 2]
    F = make_system(f, 2, 3)
    p, q = np.array([5.2, -1.3, 9.3]), np.array([2.6, 3.3, 2.3])
    H = cpu_api.homotopy(capi.H_PARAMETER, cpu_api.system(F), p=p, q=q)
    _check_against_symbolic(H, lambda x, t: f(x, [t * a + (1 - t) * b for a, b in zip(p, q)]), 2, np.random.default_rng(3))


def test_straight_line_homotopy_operators(cpu_api):  # test/homotopies_test.jl:80-93: H = t F + (1 - t) G (gamma = 1)
    from hcb200.modelkit import make_system
    a, b, c = 0.31, 0.77, 0.52
    f = lambda v, p: [(2 * v[0] ** 2 + b ** 2 * v[1] ** 3 + 2 * a * v[0] * v[1]) ** 3, (a + c) ** 4 * v[0] + v[1] ** 2]
    g = lambda v, p: [(2 * v[1] ** 2 + b ** 2 * v[0] ** 3 + 2 * a * v[0] * v[1]) ** 2, (a - c) ** 3 * v[1] + v[0] ** 2]
    start, target = cpu_api.system(make_system(f, 2)), cpu_api.system(make_system(g, 2))
    H = cpu_api.homotopy(capi.H_STRAIGHT_LINE, target, start, gamma=1.0, G_params=[], F_params=[])
    _check_against_symbolic(H, lambda x, t: [t * u + (1 - t) * w for u, w in zip(f(x, None), g(x, None))], 2, np.random.default_rng(4))
