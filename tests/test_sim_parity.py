"""Kernel logic without a GPU: the device headers compiled for the host (tests/host_sim) must agree
with the oracle.  This is a unit test of the code that runs on the B200, not a product path."""
import numpy as np
import pytest

import ref_systems
from helpers import assert_batches_match, straight_line, system_2x2
from hcb200 import capi, systems
from hcb200.modelkit import make_system


def both(oracle, sim, build):
    return [build(api) for api in (oracle, sim)]


def test_operator_api(oracle, sim):
    rng = np.random.default_rng(1)
    F = systems.katsura(4)
    x = rng.normal(size=5) + 1j * rng.normal(size=5)
    tx = np.stack([x, 0.3 * x + 0.1j, 0.01 * x * x])
    out = []
    for api in (oracle, sim):
        td, H = straight_line(api, F, 0.4 + 1.3j)
        u, U = H.evaluate_and_jacobian(x, 0.37)
        out.append([u, U, H.evaluate(x, 0.37), H.evaluate_dd(x, x * 2.0 ** -54, 0.37)] + [H.taylor(K, tx[:K], 0.37) for K in (1, 2, 3)])
    for a, b in zip(*out):
        assert np.abs(a - b).max() <= 1e-13 * max(1.0, np.abs(a).max())


def test_operator_api_batched(oracle, sim):
    """hc_evaluate_batch: evaluate! / evaluate_and_jacobian! at many points in one call == the single-point calls"""
    rng = np.random.default_rng(2)
    td, H = straight_line(sim, systems.katsura(4), 0.4 + 1.3j)
    _, Ho = straight_line(oracle, systems.katsura(4), 0.4 + 1.3j)
    X = rng.normal(size=(37, 5)) + 1j * rng.normal(size=(37, 5))
    u, U = H.evaluate_batch(X, 0.37, jacobian=True)
    u2 = H.evaluate_batch(X, 0.37)
    for k in (0, 5, 36):
        a, A = Ho.evaluate_and_jacobian(X[k], 0.37)
        assert np.abs(u[k] - a).max() <= 1e-13 * max(1.0, np.abs(a).max()) and np.abs(U[k] - A).max() <= 1e-13 * max(1.0, np.abs(A).max())
        assert np.abs(u2[k] - a).max() <= 1e-13 * max(1.0, np.abs(a).max())
    assert H.evaluate_batch(X[:0], 0.0).shape == (0, 5)


def test_endgame_paths(oracle, sim):
    ro, rs = both(oracle, sim, lambda api: (lambda td_H: td_H[1].track_batch(td_H[0].start_solutions()))(straight_line(api, system_2x2(), 0.4 + 1.3j)))
    assert_batches_match(ro, rs)
    assert np.allclose(ro.valuation[2:], rs.valuation[2:], atol=1e-6)
    F1 = make_system(lambda v, p: [(v[0] - 10) ** 3], 1)
    ro, rs = both(oracle, sim, lambda api: (lambda td_H: td_H[1].track_batch(td_H[0].start_solutions()))(straight_line(api, F1, np.exp(0.77j))))
    assert_batches_match(ro, rs)
    assert (rs.winding_number == 3).all() and rs.singular.all()


def test_katsura6(oracle, sim):
    ro, rs = both(oracle, sim, lambda api: (lambda td_H: td_H[1].track_batch(td_H[0].start_solutions()))(straight_line(api, systems.katsura(6), 0.4 + 1.3j)))
    assert_batches_match(ro, rs)
    assert (rs.return_code == 1).sum() == 64
    assert abs(int(ro.accepted_steps.sum()) - int(rs.accepted_steps.sum())) <= 0.02 * ro.accepted_steps.sum()


def test_steiner_double_double(oracle, sim):
    F, g = ref_systems.steiner_higher_prec()
    rs = sim.homotopy(capi.H_PARAMETER, sim.system(F), p=g["p"], q=g["q"]).track_batch([g["s_p"]], mode=1)
    assert capi.TRACKER_CODES[rs.return_code[0]] == "success" and rs.extended_precision_used[0]
    assert np.allclose(rs.solution[0], g["s_q"], rtol=1.5e-8, atol=0)


def test_parameter_sweep_per_path_targets(oracle, sim):
    F = systems.biochem1()
    rng = np.random.default_rng(5)
    p1 = rng.normal(size=10) + 1j * rng.normal(size=10)
    # start solutions of the generic instance via total degree
    td, H0 = straight_line(oracle, F, 0.4 + 1.3j, p1)
    r0 = H0.track_batch(td.start_solutions())
    starts = r0.solution[(r0.return_code == 1)]
    assert len(starts) >= 1
    K = 6
    q = systems.BIOCHEM1_PVALS[None, :] * np.exp(0.5 * rng.normal(size=(K, 10)))
    S = np.repeat(starts[None], K, axis=0).reshape(-1, 3)
    Q = np.repeat(q[:, None, :], len(starts), axis=1).reshape(-1, 10).astype(np.complex128)
    res = [api.homotopy(capi.H_PARAMETER, api.system(F), p=p1, q=q[0]).track_batch(S, path_q=Q) for api in (oracle, sim)]
    assert_batches_match(*res)
    # per-path targets really are used: path block k equals a homogeneous run with q[k]
    one = sim.homotopy(capi.H_PARAMETER, sim.system(F), p=p1, q=q[3]).track_batch(starts)
    blk = slice(3 * len(starts), 4 * len(starts))
    assert (one.return_code == res[1].return_code[blk]).all()
    assert np.allclose(one.solution, res[1].solution[blk], rtol=1e-12, atol=1e-14)


def _same(a, b):
    for x, y in zip(a.arrays(), b.arrays()):
        assert np.array_equal(x, y, equal_nan=x.dtype.kind in "fc")


def test_total_degree_starts_made_by_the_tracker(oracle, sim):
    """hc_track_total_degree: index -> start solution inside the tracker gives bit-identical results
    to the streamed-in iterator values, for the whole range and for an index sub-range (a multi-GPU shard)."""
    td, H = straight_line(sim, systems.katsura(4), 0.4 + 1.3j)
    S = td.start_solutions()
    ref = H.track_batch(S)
    got = capi.track_total_degree(H, td.degrees)
    assert got.N == len(S) == 16
    _same(ref, got)
    part = capi.track_total_degree(H, td.degrees, first=5, count=7)
    _same(H.track_batch(S[5:12]), part)
    # mixed degrees: the first index runs fastest (Iterators.product order, total_degree.jl:241)
    F = make_system(lambda v, p: [v[0] ** 3 - 2 * v[1] + 1, v[0] * v[1] + v[1] ** 2 - 3, v[2] - v[0] - 1], 3)
    td, H = straight_line(sim, F, 0.3 - 0.9j)
    assert list(td.degrees) == [3, 2, 1]
    _same(H.track_batch(td.start_solutions()), capi.track_total_degree(H, td.degrees))
    _, Ho = straight_line(oracle, F, 0.3 - 0.9j)
    assert_batches_match(Ho.track_batch(td.start_solutions()), capi.track_total_degree(H, td.degrees))
    with pytest.raises(RuntimeError, match="exceeds"):
        capi.track_total_degree(H, td.degrees, first=4, count=3)
    with pytest.raises(RuntimeError):
        capi.track_total_degree(Ho, td.degrees)   # the oracle has no such entry point


def test_sweep_entry_point_matches_replicated_batch(oracle, sim):
    """hc_track_sweep (many_solve, src/solve.jl:815-881): S starts x M parameter points with one parameter column per
    point == hc_track_batch with starts and parameters replicated per path."""
    F = systems.biochem1()
    rng = np.random.default_rng(11)
    p1 = rng.normal(size=10) + 1j * rng.normal(size=10)
    td, H0 = straight_line(oracle, F, 0.4 + 1.3j, p1)
    r0 = H0.track_batch(td.start_solutions())
    starts = r0.solution[(r0.return_code == 1)]
    k, M = len(starts), 5
    q = (systems.BIOCHEM1_PVALS[None, :] * np.exp(0.5 * rng.normal(size=(M, 10)))).astype(np.complex128)
    H = sim.homotopy(capi.H_PARAMETER, sim.system(F), p=p1, q=q[0])
    S = np.repeat(starts[None], M, axis=0).reshape(-1, 3)
    Q = np.repeat(q[:, None, :], k, axis=1).reshape(-1, 10)
    _same(H.track_batch(S, path_q=Q), capi.track_sweep(H, starts, q))
    Ho = oracle.homotopy(capi.H_PARAMETER, oracle.system(F), p=p1, q=q[0])
    assert_batches_match(Ho.track_batch(S, path_q=Q), capi.track_sweep(H, starts, q))
    _, Hs = straight_line(sim, F, 0.4 + 1.3j, p1)   # a straight-line homotopy is not a sweep target
    with pytest.raises(RuntimeError, match="parameter homotopy"):
        capi.track_sweep(Hs, starts, q)


def test_polyhedral_cyclic5(oracle, sim):
    """PolyhedralTracker two-stage track (reference src/polyhedral.jl:414-530): 70 mixed-volume paths, 70 solutions
    (reference test/polyhedral_test.jl:38-46)."""
    from hcb200 import polyhedral as ph
    ps = ph.polyhedral(systems.cyclic(5))
    S, ci = ps.start_solutions()
    cw = ps.cell_weights()
    res = []
    for api in (oracle, sim):
        h = api.system(ps.F)
        Ht = api.homotopy(capi.H_TORIC, h, p=ps.start_coeffs)
        Hc = api.homotopy(capi.H_COEFFICIENT, h, p=ps.start_coeffs, q=ps.target_coeffs)
        res.append(capi.polyhedral_track_batch(api, Ht, Hc, S, ci, cw))
    assert_batches_match(*res)
    assert (res[1].return_code == 1).sum() == 70
    assert len(np.unique(np.round(res[1].solution, 6), axis=0)) == 70


def test_polyhedral_starts_made_on_the_device(oracle, sim):
    """hc_polyhedral_track_cells: the start solutions of every mixed cell are made from per-cell data
    (H, mu, r) by the device code -- BinomialSystemSolver's unit-root combinations and double-double triangular solve,
    reference src/binomial_system.jl:55-106.  They solve their binomial systems (test/binomial_system_test.jl:107-120
    checks residuals the same way) and the tracked batch equals the oracle's run from the host-made start solutions;
    an index range (first, count) gives the slice, indices beyond the mixed volume wrap."""
    from hcb200 import polyhedral as ph
    for F in (systems.cyclic(5), systems.katsura(4)):
        ps = ph.polyhedral(F)
        S, ci = ps.start_solutions()
        cw, cells = ps.cell_weights(), ps.binomial_data()
        assert int(cells["volume"].sum()) == len(S)
        def handles(api):
            h = api.system(ps.F)
            return (api.homotopy(capi.H_TORIC, h, p=ps.start_coeffs),
                    api.homotopy(capi.H_COEFFICIENT, h, p=ps.start_coeffs, q=ps.target_coeffs))
        ref = capi.polyhedral_track_batch(oracle, *handles(oracle), S, ci, cw)
        Ht, Hc = handles(sim)
        dev = capi.polyhedral_track_cells(sim, Ht, Hc, cells, cw)
        assert_batches_match(ref, dev)
        # against the same code started from the host-made solutions: the starts agree to rounding, so do the step counts
        host = capi.polyhedral_track_batch(sim, Ht, Hc, S, ci, cw)
        assert (host.return_code == dev.return_code).all()
        assert np.abs(host.accepted_steps - dev.accepted_steps).max() <= 3
        assert np.abs(host.solution - dev.solution).max() <= 1e-12 * np.abs(host.solution).max()
        lo, cnt = len(S) // 3, len(S) // 2
        part = capi.polyhedral_track_cells(sim, Ht, Hc, cells, cw, first=lo, count=cnt)
        assert (part.return_code == dev.return_code[lo:lo + cnt]).all()
        assert np.array_equal(part.solution, dev.solution[lo:lo + cnt])
        wrap = capi.polyhedral_track_cells(sim, Ht, Hc, cells, cw, first=len(S) - 2, count=5)
        assert np.array_equal(wrap.solution[2:], dev.solution[:3])


@pytest.mark.parametrize("k", [1, 2, 4, 6, 10, 11, 12])
def test_system_sizes(oracle, sim, k):
    """n = k + 1 = 2 .. 13: the register-blocked LU / solve instantiations (n <= 12) and the generic
    fallback (n = 13) must reproduce the oracle; 24 paths of katsura(k) each."""
    out = []
    for api in (oracle, sim):
        td, H = straight_line(api, systems.katsura(k), 0.4 + 1.3j)
        out.append(H.track_batch(td.start_solutions()[:24]))
    assert_batches_match(*out)
    assert abs(int(out[0].accepted_steps.sum()) - int(out[1].accepted_steps.sum())) <= 0.02 * out[0].accepted_steps.sum()


def test_homogeneous_system_on_an_affine_chart(oracle, sim):
    """total_degree of a homogeneous system (two quadrics in P^2; reference src/total_degree.jl:46-108): tracked on a random
    affine chart v'x = 1 with G = s .* (x[1:n-1].^D .- x[n].^D); the chart row is part of F and G (start_systems.py).
    Four projective solutions, on the chart, zeros of F, and -- dehomogenised -- the solutions of the affine system."""
    from hcb200 import start_systems
    quadrics = lambda v: [v[0] ** 2 + 2 * v[1] ** 2 - 3 * v[2] ** 2 + v[0] * v[1], v[0] * v[1] - 2 * v[2] ** 2 + v[1] * v[2] + 0.5 * v[0] ** 2]
    F = make_system(lambda v, p: quadrics(v), 3)
    td = start_systems.total_degree(F, 0.4 + 1.3j)
    assert td.chart is not None and td.n_paths() == 4 and td.F.n_eqs == 3
    S = td.start_solutions()
    assert np.abs(S @ td.chart - 1).max() < 1e-14
    out = []
    for api in (oracle, sim):
        H = api.homotopy(capi.H_STRAIGHT_LINE, api.system(td.F), api.system(td.G), gamma=td.gamma, G_params=td.scaling, F_params=[])
        out.append(H.track_batch(S))
    assert_batches_match(*out)
    r = out[1]
    assert (r.return_code == 1).all() and np.abs(r.solution @ td.chart - 1).max() < 1e-12
    Fa = make_system(lambda v, p: quadrics([v[0], v[1], 1.0]), 2)
    ta, Ha = straight_line(sim, Fa, 0.4 + 1.3j)
    ra = Ha.track_batch(ta.start_solutions())
    A = r.solution[:, :2] / r.solution[:, 2:3]
    B = ra.solution[ra.return_code == 1]
    assert len(B) == 4
    d = np.abs(A[:, None, :] - B[None, :, :]).max(axis=2)
    assert (d.min(axis=1) < 1e-8).all() and len(set(d.argmin(axis=1))) == 4


def test_overdetermined_system_squared_up(oracle, sim):
    """total_degree of an overdetermined system (3 equations, 2 variables; reference src/total_degree.jl:66-92,
    src/systems/randomized_system.jl, src/overdetermined.jl): sorted by degree, squared up with a random [I A], the four
    paths of the 2 x 2 system end in the two common zeros and two excess solutions, which the check marks (code 14)."""
    from hcb200 import result, start_systems
    F = make_system(lambda v, p: [v[0] + v[1] - 3.0, v[0] ** 2 + v[1] ** 2 - 5.0, v[0] * v[1] - 2.0], 2)
    td = start_systems.total_degree(F, 0.4 + 1.3j)
    assert td.original is not None and td.F.n_eqs == 2 and list(td.degrees) == [2, 2] and td.A.shape == (2, 1)
    out = []
    for api in (oracle, sim):
        H = api.homotopy(capi.H_STRAIGHT_LINE, api.system(td.F), api.system(td.G), gamma=td.gamma, G_params=td.scaling, F_params=[])
        out.append(start_systems.excess_solution_check(td, H.track_batch(td.start_solutions())))
    assert_batches_match(*out)
    r = out[1]
    assert sorted(r.return_code.tolist()) == [1, 1, 14, 14]
    sols = sorted(np.round(r.solution[r.return_code == 1].real, 8).tolist())
    assert sols == [[1.0, 2.0], [2.0, 1.0]] and np.abs(r.solution[r.return_code == 1].imag).max() < 1e-10
    st = result.statistics(r)
    assert st.nonsingular == 2 and st.excess_solution == 2 and st.failed == 0


@pytest.mark.parametrize("window", [1, 8, 32, 64, 1000])
def test_segment_scheduler_invariants(sim, window):
    """The lowering the thread-per-path engine runs (hc_lower.h): every op's operands exist before its segment
    starts and are not produced inside it, segments are runs of one (class, sign) key, the fast format mirrors
    the packed one, nothing writes the input block -- for the tapes of all five configs."""
    import ctypes as C
    from hcb200 import workloads
    lib = sim.raw
    lib.hc_sim_check_lowering.restype = C.c_int32
    lib.hc_sim_check_lowering.argtypes = [C.POINTER(capi.ProgramDesc), C.c_int32, C.POINTER(C.c_int32)]
    progs = []
    for S in (systems.katsura(8), systems.cyclic(7), systems.tritangents(), systems.cyclooctane(), systems.biochem1()):
        progs += [S.eval_program, S.jac_program]
    for P in progs:
        keep = []
        d = sim._program_desc(P, keep)
        stats = (C.c_int32 * 3)()
        rc = lib.hc_sim_check_lowering(C.byref(d), window, stats)
        assert rc == 0, (rc, window)
        assert stats[0] >= 1 and 1 <= stats[1] <= stats[0]
        if window == 1:
            assert stats[1] == stats[0]          # tape order: one op per segment


def _random_system(rng, n, degs):
    """Sparse random polynomial system with complex coefficients: per equation one monomial of full degree, a few
    lower ones and a constant."""
    seed = int(rng.integers(1 << 30))

    def build(v, p):
        r = np.random.default_rng(seed)

        def monomial(deg):
            ex = r.multinomial(deg, np.ones(n) / n)
            m = complex(r.normal(), r.normal() if r.random() < 0.5 else 0.0)
            for k in range(n):
                if ex[k]:
                    m = m * v[k] ** int(ex[k])
            return m
        eqs = []
        for i in range(n):
            d = int(degs[i])
            e = monomial(d) + complex(r.normal(), r.normal())
            for _ in range(int(r.integers(2, 6))):
                e = e + monomial(int(r.integers(0, d + 1)))
            eqs.append(e)
        return eqs
    return make_system(build, n)


def test_random_systems_fuzz(oracle, sim):
    """80 random systems (n = 1..4, degrees 1..3, 500+ total-degree paths incl. diverging ones): the device code agrees
    with the oracle on every return code, singular flag and winding number and on the endpoints to 1e-8."""
    rng = np.random.default_rng(1)
    paths = 0
    for _ in range(80):
        n = int(rng.integers(1, 5))
        degs = rng.integers(1, 4, size=n)
        F = _random_system(rng, n, degs)
        gamma = np.exp(2j * np.pi * rng.random())
        tdo, Ho = straight_line(oracle, F, gamma)
        _, Hs = straight_line(sim, F, gamma)
        S = tdo.start_solutions()
        ro, rs = Ho.track_batch(S), Hs.track_batch(S)
        assert_batches_match(ro, rs)
        ok = ro.return_code == 1
        assert (np.abs(ro.accepted_steps - rs.accepted_steps)[ok] <= np.maximum(3, 0.1 * ro.accepted_steps[ok])).all()
        paths += len(S)
    assert paths > 500


def _track_polyhedral(api, ps):
    S, ci = ps.start_solutions()
    h = api.system(ps.F)
    Ht = api.homotopy(capi.H_TORIC, h, p=ps.start_coeffs)
    Hc = api.homotopy(capi.H_COEFFICIENT, h, p=ps.start_coeffs, q=ps.target_coeffs)
    return capi.polyhedral_track_batch(api, Ht, Hc, S, ci, ps.cell_weights())


def test_polyhedral_fuzz_and_agreement_with_total_degree(oracle, sim):
    """Random sparse systems (n = 2, 3): (1) the polyhedral driver of the device code agrees with the oracle path by
    path; (2) size-independent property: the polyhedral homotopy (mixed-volume many paths) and the total-degree homotopy
    (Bezout many paths) end in the same set of finite solutions."""
    from hcb200 import polyhedral as ph
    rng = np.random.default_rng(11)
    systems_seen = 0
    for _ in range(30):
        n = int(rng.integers(2, 4))
        F = _random_system(rng, n, rng.integers(1, 4, size=n))
        ps = ph.polyhedral(F)
        if ps.n_paths() == 0:
            continue
        rp = _track_polyhedral(sim, ps)
        assert_batches_match(_track_polyhedral(oracle, ps), rp)
        td, H = straight_line(sim, F, np.exp(2j * np.pi * rng.random()))
        rt = H.track_batch(td.start_solutions())
        A, B = rp.solution[rp.return_code == 1], rt.solution[rt.return_code == 1]
        assert len(A) == len(B), (len(A), len(B), ps.n_paths(), td.n_paths())
        if len(A):
            d = np.abs(A[:, None, :] - B[None, :, :]).max(axis=2)
            scale = np.maximum(1.0, np.abs(B).max(axis=1))[None, :]
            assert ((d / scale).min(axis=1) < 1e-8).all()          # every polyhedral endpoint is a total-degree endpoint
            assert len(set((d / scale).argmin(axis=1))) == len(A)  # and they are all different
        systems_seen += 1
    assert systems_seen >= 20


def test_parameter_homotopy_round_trip_fuzz(oracle, sim):
    """Random parameter homotopies F(x; p), p0 -> q -> p0 with the plain Tracker (mode 1): the device code matches the
    oracle and the round trip returns to the start solutions (the property reference steiner_higher_prec.jl:145-152 checks)."""
    rng = np.random.default_rng(5)
    done = 0
    for _ in range(25):
        n = int(rng.integers(1, 4))
        degs = rng.integers(1, 4, size=n)
        P = n + 1
        seed = int(rng.integers(1 << 30))

        def build(v, p, n=n, degs=degs, seed=seed):
            r = np.random.default_rng(seed)
            eqs = []
            for i in range(n):
                ex = r.multinomial(int(degs[i]), np.ones(n) / n)
                m = complex(r.normal(), r.normal())
                for k in range(n):
                    if ex[k]:
                        m = m * v[k] ** int(ex[k])
                lin = sum(complex(r.normal(), r.normal()) * v[k] for k in range(n))
                eqs.append(m + p[i] * lin + p[n] * p[i] + complex(r.normal(), r.normal()))   # parameters enter bilinearly
            return eqs
        F = make_system(build, n, P)
        p0 = rng.normal(size=P) + 1j * rng.normal(size=P)
        q = rng.normal(size=P) + 1j * rng.normal(size=P)
        td, H0 = straight_line(oracle, F, np.exp(2j * np.pi * rng.random()), p0)
        r0 = H0.track_batch(td.start_solutions())
        S = r0.solution[(r0.return_code == 1) & (r0.singular == 0)]
        if len(S) == 0:
            continue
        res = []
        for api in (oracle, sim):
            H = api.homotopy(capi.H_PARAMETER, api.system(F), p=p0, q=q)
            fwd = H.track_batch(S, mode=1)
            okf = fwd.return_code == 1
            back = H.track_batch(fwd.solution[okf], mode=1, t1=0.0, t0=1.0)
            res.append((fwd, back, okf))
        (fo, bo, oko), (fs, bs, oks) = res
        assert (fo.return_code == fs.return_code).all() and (bo.return_code == bs.return_code).all()
        if oko.any():
            assert np.abs(fo.solution[oko] - fs.solution[oko]).max() <= 1e-8 * max(1.0, np.abs(fo.solution[oko]).max())
            okb = bs.return_code == 1
            assert np.abs(bs.solution[okb] - S[oks][okb]).max() <= 1e-8 * max(1.0, np.abs(S).max())
        done += 1
    assert done >= 15


def test_change_parameters(oracle, sim):
    """reference test/tracker_test.jl:81-91 ("Change parameters"): start_parameters! / target_parameters! on an
    existing tracker (hc_homotopy_set_parameters) == a homotopy created with those parameters."""
    F = make_system(lambda v, p: [v[0] ** 2 - p[0], v[0] * v[1] - p[0] + p[1]], 2, 2)
    H = sim.homotopy(capi.H_PARAMETER, sim.system(F), p=[2.2, 3.2], q=[2.2, 3.2])
    H.set_parameters(p=[1, 0])
    H.set_parameters(q=[2, 4])
    r = H.track_batch([[1.0, 1.0]], mode=1)
    assert capi.TRACKER_CODES[r.return_code[0]] == "success"
    assert np.allclose(r.solution[0], [np.sqrt(2), -np.sqrt(2)])
    ref = sim.homotopy(capi.H_PARAMETER, sim.system(F), p=[1, 0], q=[2, 4]).track_batch([[1.0, 1.0]], mode=1)
    _same(ref, r)
    H.set_parameters(p=[2, 4], q=[1, 0])      # parameters!(T, p, q): the way back
    back = H.track_batch([r.solution[0]], mode=1)
    assert np.allclose(back.solution[0], [1, 1])
    td, Hs = straight_line(sim, systems.katsura(3), 0.4 + 1.3j)
    with pytest.raises(RuntimeError, match="straight-line"):
        Hs.set_parameters(p=[])
    Ho = oracle.homotopy(capi.H_PARAMETER, oracle.system(F), p=[1, 0], q=[2, 4])
    with pytest.raises(RuntimeError):
        Ho.set_parameters(p=[1, 0])           # the oracle has no such entry point
