"""Pins the CPU oracle against the golden vectors / known answers of the reference's own tests."""
import ctypes as C

import mpmath as mp
import numpy as np
import pytest

import ref_systems
from helpers import straight_line, system_2x2
from hcb200 import capi, systems
from hcb200.modelkit import make_system

DP = C.POINTER(C.c_double)


def _dp(a):
    return a.ctypes.data_as(DP)


def test_nthroot(oracle):  # reference test/utils_test.jl:19-24
    f = oracle.raw.orc_nthroot
    f.restype, f.argtypes = C.c_double, [C.c_double, C.c_int32]
    assert f(42.2, 0) == 1.0 and f(42.2, 1) == 42.2
    assert f(42.1, 2) == np.sqrt(42.1) and abs(f(42.1, 3) - np.cbrt(42.1)) <= 1e-15
    assert f(42.1, 4) == np.sqrt(np.sqrt(42.1)) and f(42.1, 5) == 42.1 ** (1 / 5)


def _stepper(oracle, start, target, ds):
    ds = np.asarray(ds, dtype=np.float64)
    out = np.zeros((len(ds), 8))
    s, t = np.array([start.real, start.imag]), np.array([target.real, target.imag])
    oracle.raw.orc_stepper_trace(_dp(s), _dp(t), _dp(ds), C.c_int32(len(ds)), _dp(out))
    return out


def test_segment_stepper(oracle):  # reference test/utils_test.jl:25-83 (exact values)
    e = np.exp2
    o = _stepper(oracle, 0j, 1 + 0j, [e(-60), e(-50), 4])
    assert o[0, 0] == e(-60) and o[0, 2] == e(-60) and o[0, 4] == 0
    assert o[1, 6] == e(-60) + e(-50) and o[1, 0] == e(-60) + e(-50)
    assert o[2, 6] == 1.0 and o[2, 0] == 1.0 and o[2, 4] == 1
    o = _stepper(oracle, 1j, 0j, [e(-20), e(-10), 2])
    assert o[0, 7] == 1 - e(-20) and o[0, 6] == 0 and o[0, 3] == -e(-20)
    assert o[1, 7] == 1 - (e(-20) + e(-10))
    assert o[2, 6] == 0.0 and o[2, 7] == 0.0 and o[2, 4] == 1


def test_norms(oracle):  # reference test/norm_test.jl:1-36 incl. the 2^700 overflow fallback
    x = np.array([2j, 3 - 1j, 5 + 2j]); y = np.array([-2j, 3 - 1j, 5 + 2j])
    for scale in (1.0, np.exp2(700)):
        w = np.array([4.0, 2.0, 2.0]); out = np.zeros(4 + 6)
        xs, ys = np.ascontiguousarray(x * scale), np.ascontiguousarray(y * scale)
        oracle.raw.orc_norm_test(C.c_int32(3), _dp(xs.view(np.float64)), _dp(ys.view(np.float64)), _dp(w), _dp(out))
        assert np.isclose(out[0], scale * abs(5 + 2j)) and np.isclose(out[1], scale * 4)
        assert np.isclose(out[2], scale * 0.5 * abs(5 + 2j)) and np.isclose(out[3], scale * 0.25 * 4)


def test_double_double(oracle):  # reference test/double_double_test.jl:67-75
    mp.mp.prec = 300
    rng = np.random.default_rng(0)
    f = oracle.raw.orc_dd_op
    def dd(v):
        hi = float(v); return np.array([hi, float(v - mp.mpf(hi))])
    for _ in range(10):
        X = mp.mpf(rng.random()) * 20 - 10 + mp.mpf(rng.random()) * mp.mpf(2) ** -60
        Y = mp.mpf(rng.random()) * 20 - 10 + mp.mpf(rng.random()) * mp.mpf(2) ** -60
        a, b = dd(X), dd(Y)
        X, Y = mp.mpf(a[0]) + mp.mpf(a[1]), mp.mpf(b[0]) + mp.mpf(b[1])
        for op, ref, atol in ((0, X + Y, 1e-30), (1, X - Y, 1e-30), (2, X * Y, 1e-29), (3, X / Y, 1e-26)):
            out = np.zeros(2)
            f(C.c_int32(op), _dp(a), _dp(b), C.c_int32(0), _dp(out))
            assert abs(mp.mpf(out[0]) + mp.mpf(out[1]) - ref) < atol * max(1, abs(ref))
        out = np.zeros(2)
        f(C.c_int32(5), _dp(a), None, C.c_int32(5), _dp(out))
        assert abs(mp.mpf(out[0]) + mp.mpf(out[1]) - X ** 5) < 1e-25 * max(1, abs(X ** 5))


@pytest.mark.parametrize("n", [3, 13, 31])
def test_ldiv(oracle, n):  # reference test/linear_algebra_test.jl:51-64
    rng = np.random.default_rng(n)
    A = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    b = rng.normal(size=n) + 1j * rng.normal(size=n)
    x = np.zeros(n, np.complex128)
    Af = np.ascontiguousarray(A.T)  # column-major
    oracle.raw.orc_la_solve(C.c_int32(n), _dp(Af.view(np.float64)), _dp(b.view(np.float64)), None, C.c_int32(0), _dp(x.view(np.float64)))
    assert np.allclose(x, np.linalg.solve(A, b), rtol=1e-10)
    w = np.abs(rng.normal(size=n)) + 0.1
    oracle.raw.orc_la_solve(C.c_int32(n), _dp(Af.view(np.float64)), _dp(b.view(np.float64)), _dp(w), C.c_int32(2), _dp(x.view(np.float64)))
    assert np.allclose(x, np.linalg.solve(A, b), rtol=1e-13)


def test_cond_estimator(oracle):  # reference test/linear_algebra_test.jl:66-109
    rng = np.random.default_rng(5)
    est = oracle.raw.orc_la_inverse_inf_norm_est; est.restype = C.c_double
    cond = oracle.raw.orc_la_cond; cond.restype = C.c_double
    opn = lambda M: np.abs(M).sum(axis=1).max()
    d_r = rng.random() * 10.0 ** np.linspace(-6, 6, 6)
    d_l = rng.random() * 10.0 ** np.linspace(6, -6, 6)
    A = rng.normal(size=(6, 6))
    def call(fn, M, dl, dr):
        Mf = np.ascontiguousarray(M.T.astype(np.complex128))
        return fn(C.c_int32(6), _dp(Mf.view(np.float64)), _dp(dl) if dl is not None else None, _dp(dr) if dr is not None else None)
    B = A @ np.diag(1 / d_r)
    assert 0.1 <= opn(np.linalg.inv(B)) / call(est, B, None, None) <= 10
    assert 0.1 <= opn(np.linalg.inv(B @ np.diag(d_r))) / call(est, B, None, d_r) <= 10
    assert 0.1 <= opn(np.linalg.inv(np.diag(d_r) @ B)) / call(est, B, d_r, None) <= 10
    assert 0.1 <= np.linalg.cond(B @ np.diag(d_r), np.inf) / call(cond, B, None, d_r) <= 10
    D = np.diag(1 / d_l) @ A @ np.diag(1 / d_r)
    assert 0.1 <= opn(np.linalg.inv(np.diag(d_l) @ D @ np.diag(d_r))) / call(est, D, d_l, d_r) <= 10
    assert call(est, D, d_r, d_l) > 100 * call(est, D, d_l, d_r)
    assert 0.1 <= np.linalg.cond(np.diag(d_l) @ D @ np.diag(d_r), np.inf) / call(cond, D, d_l, d_r) <= 10


OPS = {"CB": 1, "INV": 7, "INVSQR": 9, "NEG": 10, "SQR": 13, "ADD": 18, "DIV": 19, "MUL": 20, "SUB": 21, "POW_INT": 22,
       "ADD3": 24, "MUL3": 25, "MULADD": 26, "MULSUB": 27, "SUBMUL": 28, "ADD4": 29, "MUL4": 30, "MULMULADD": 31, "MULMULSUB": 32}


@pytest.mark.parametrize("K", [1, 2, 3, 4])
def test_taylor_ops(oracle, K):  # reference test/model_kit/operations_test.jl:1-59
    mp.mp.prec = 200
    rng = np.random.default_rng(K)
    ser = [rng.normal(size=5) + 1j * rng.normal(size=5) for _ in range(4)]
    def S(v):
        return lambda l: sum(mp.mpc(c) * l ** k for k, c in enumerate(ser[v]))
    a, b, c, d = (S(v) for v in range(4))
    fns = {"CB": lambda l: a(l) ** 3, "INV": lambda l: 1 / a(l), "INVSQR": lambda l: 1 / a(l) ** 2, "NEG": lambda l: -a(l),
           "SQR": lambda l: a(l) ** 2, "ADD": lambda l: a(l) + b(l), "DIV": lambda l: a(l) / b(l), "MUL": lambda l: a(l) * b(l),
           "SUB": lambda l: a(l) - b(l), "POW_INT": lambda l: a(l) ** 5, "ADD3": lambda l: a(l) + b(l) + c(l),
           "MUL3": lambda l: a(l) * b(l) * c(l), "MULADD": lambda l: a(l) * b(l) + c(l), "MULSUB": lambda l: a(l) * b(l) - c(l),
           "SUBMUL": lambda l: c(l) - a(l) * b(l), "ADD4": lambda l: a(l) + b(l) + c(l) + d(l),
           "MUL4": lambda l: a(l) * b(l) * c(l) * d(l), "MULMULADD": lambda l: a(l) * b(l) + c(l) * d(l),
           "MULMULSUB": lambda l: a(l) * b(l) - c(l) * d(l)}
    flat = [np.ascontiguousarray(s).view(np.float64) for s in ser]
    for name, fn in fns.items():
        out = np.zeros(2 * (K + 1))
        oracle.raw.orc_taylor_op(C.c_int32(OPS[name]), C.c_int32(K), _dp(flat[0]), _dp(flat[1]), _dp(flat[2]), _dp(flat[3]),
                                 C.c_int32(5), _dp(out))
        ref = mp.taylor(fn, 0, K)
        got = out.view(np.complex128)
        for k in range(K + 1):
            assert abs(complex(ref[k]) - got[k]) <= 1e-12 * max(1.0, abs(complex(ref[k]))), (name, k)


SYSTEMS = {
    "katsura5": lambda: (systems.katsura(5), None), "cyclic5": lambda: (systems.cyclic(5), None),
    "cyclic7": lambda: (systems.cyclic(7), None), "tritangents": lambda: (systems.tritangents(), 20),
    "cyclooctane": lambda: (systems.cyclooctane(), 36), "biochem1": lambda: (systems.biochem1(), 10),
    "steiner": lambda: (ref_systems.steiner_higher_prec()[0], 30), "four_bar": lambda: (ref_systems.four_bar()[0], 16),
}


@pytest.mark.parametrize("name", sorted(SYSTEMS))
def test_tape_evaluation(oracle, name):
    """evaluate!, evaluate_and_jacobian!, DD evaluate!, taylor! K=1..3 against 60-digit arithmetic on the
    expression DAG (reference test/model_kit/e2e_test.jl:42-91: rtol 1e-12)."""
    mp.mp.prec = 200
    F, P = SYSTEMS[name]()
    rng = np.random.default_rng(7)
    n = F.n_vars
    p = rng.normal(size=P or 0) + 1j * rng.normal(size=P or 0)
    q = rng.normal(size=P or 0) + 1j * rng.normal(size=P or 0)
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    t = 0.37
    H = oracle.homotopy(capi.H_PARAMETER, oracle.system(F), p=p, q=q)
    pt = t * p + (1 - t) * q
    mpx = [mp.mpc(v) for v in x]; mpp = [mp.mpc(v) for v in pt]
    ref_u = np.array([complex(v) for v in F.evaluate(mpx, mpp, ctx=mp)])
    ref_U = np.array([[complex(v) for v in row] for row in F.jacobian(mpx, mpp, ctx=mp)])
    scale = max(1.0, np.abs(ref_u).max())
    u, U = H.evaluate_and_jacobian(x, t)
    assert np.abs(u - ref_u).max() <= 1e-12 * scale
    assert np.abs(U - ref_U).max() <= 1e-12 * max(1.0, np.abs(ref_U).max())
    assert np.abs(H.evaluate(x, t) - ref_u).max() <= 1e-12 * scale
    # DD evaluation: hi + lo input, result correct to ~1e-28 relative before rounding to fp64
    xlo = x * 2.0 ** -55
    mpx2 = [mp.mpc(a) + mp.mpc(b) for a, b in zip(x, xlo)]
    ref_dd = np.array([complex(v) for v in F.evaluate(mpx2, mpp, ctx=mp)])
    assert np.abs(H.evaluate_dd(x, xlo, t) - ref_dd).max() <= 4e-16 * scale
    # Taylor coefficients of lambda -> F(x(lambda); p(t + lambda))
    xs = [x, rng.normal(size=n) + 1j * rng.normal(size=n), rng.normal(size=n) + 1j * rng.normal(size=n)]
    for K in (1, 2, 3):
        def g(l, i):
            xl = [sum(mp.mpc(xs[k][j]) * l ** k for k in range(K)) for j in range(n)]
            pl = [mp.mpc(a) + l * (mp.mpc(b) - mp.mpc(c)) for a, b, c in zip(pt, p, q)]
            return F.evaluate(xl, pl, ctx=mp)[i]
        ref = np.array([complex(mp.taylor(lambda l: g(l, i), 0, K)[K]) for i in range(F.n_eqs)])
        got = H.taylor(K, np.stack(xs[:K]), t)
        assert np.abs(got - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max()), K


def test_tracker_parameter_homotopy(oracle):  # reference test/tracker_test.jl:2-27
    F = make_system(lambda v, p: [v[0] ** 2 - p[0], v[0] * v[1] - p[0] + p[1]], 2, 2)
    H = oracle.homotopy(capi.H_PARAMETER, oracle.system(F), p=[1, 0], q=[2, 4])
    r = H.track_batch([[1, 1]], mode=1)
    assert capi.TRACKER_CODES[r.return_code[0]] == "success"
    assert r.accepted_steps[0] + r.rejected_steps[0] <= 5
    assert np.allclose(r.solution[0], [np.sqrt(2), -np.sqrt(2)])
    back = H.track_batch([r.solution[0]], mode=1, t1=0.0, t0=1.0, omega_mu=[r.omega[0], r.mu[0]])
    assert capi.TRACKER_CODES[back.return_code[0]] == "success"


def test_invalid_start_value(oracle):  # reference test/tracker_test.jl:93-105
    F = make_system(lambda v, p: [v[0] ** 2 + v[1] ** 2 - 3, 2 * v[0] ** 2 + 0.5 * v[0] * v[1] + 3 * v[1] ** 2 - 2], 2)
    td, H = straight_line(oracle, F, 1j)
    r = H.track_batch([[100, -100]], mode=1)
    assert capi.TRACKER_CODES[r.return_code[0]] == "terminated_invalid_startvalue"


def test_endgame_tracker_path_results(oracle):  # reference test/endgame_tracker_test.jl:4-45
    td, H = straight_line(oracle, system_2x2(), 0.4 + 1.3j)
    r = H.track_batch(td.start_solutions())
    codes = [capi.ENDGAME_CODES[c] for c in r.return_code]
    assert codes == ["success", "success", "at_infinity", "at_infinity"]
    assert r.accuracy[0] < 1e-12 and r.residual[0] < 1e-12
    assert r.accepted_steps[0] + r.rejected_steps[0] < 20 and r.rejected_steps[0] == 0
    assert r.winding_number[0] == 0 and not r.singular[1] and r.condition_jacobian[1] < 1e3
    assert np.allclose(r.valuation[2], [-1, -1], rtol=1e-3)
    td, H = straight_line(oracle, make_system(lambda v, p: [(v[0] - 10) ** 2], 1), np.exp(2j * np.pi * 0.37))
    r = H.track_batch(td.start_solutions())
    assert r.winding_number[0] == 2 and 0 < r.last_t[0] < 0.1
    assert (np.abs(r.solution.imag) < 1e-6).all() and r.singular.all()
    assert np.allclose(r.solution.real, 10, atol=1e-6)


def test_steiner_higher_precision(oracle):  # reference test/test_cases/steiner_higher_prec.jl:140-152
    F, g = ref_systems.steiner_higher_prec()
    H = oracle.homotopy(capi.H_PARAMETER, oracle.system(F), p=g["p"], q=g["q"])
    r = H.track_batch([g["s_p"]], mode=1)
    assert capi.TRACKER_CODES[r.return_code[0]] == "success"
    assert np.allclose(r.solution[0], g["s_q"], rtol=1.5e-8, atol=0)   # Julia `≈`: rtol = sqrt(eps)
    assert r.extended_precision_used[0]
    back = H.track_batch([r.solution[0]], mode=1, t1=0.0, t0=1.0, omega_mu=[r.omega[0], r.mu[0]])
    assert capi.TRACKER_CODES[back.return_code[0]] == "success"
    assert np.allclose(back.solution[0], g["s_p"], rtol=1.5e-8, atol=0)


def test_four_bar(oracle):  # reference test/test_cases/four_bar.jl:90-97
    F, g = ref_systems.four_bar()
    H = oracle.homotopy(capi.H_PARAMETER, oracle.system(F), p=g["p"], q=g["q"])
    r = H.track_batch([g["s"]], mode=1)
    assert capi.TRACKER_CODES[r.return_code[0]] == "success"


def test_winding_numbers(oracle):  # reference test/endgame_test.jl:47-52: (x-10)^d -> d paths, winding number d
    for d in (2, 3, 4, 5):
        td, H = straight_line(oracle, make_system(lambda v, p: [(v[0] - 10) ** d], 1), np.exp(2j * np.pi * 0.123))
        r = H.track_batch(td.start_solutions())
        assert (r.return_code == 1).all() and (r.winding_number == d).all(), (d, r.return_code, r.winding_number)


def test_cyclic7_total_degree_count(oracle):  # reference test/endgame_test.jl:2-5: 924 solutions
    td, H = straight_line(oracle, systems.cyclic(7), np.exp(2j * np.pi * 0.7133))
    r = H.track_batch(td.start_solutions(), nthreads=8)
    assert td.n_paths() == 5040
    assert int((r.return_code == 1).sum()) == 924
