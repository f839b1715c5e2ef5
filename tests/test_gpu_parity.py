"""Parity of the CUDA path (through the C ABI of libhc_b200.so) with the CPU oracle.

Bar (BASELINE.json north_star): identical return codes / solution classes, every endpoint within
1e-8 relative; operator-API values within 1e-12 relative (reference test/model_kit/e2e_test.jl:42-52)."""
import numpy as np
import pytest

import ref_systems
from helpers import assert_batches_match, assert_classes_match, rel_endpoint_error, straight_line, system_2x2
from hcb200 import capi, lib, systems
from hcb200.modelkit import make_system

pytestmark = pytest.mark.gpu


def track_td(api, F, gamma, tp=None, **kw):
    td, H = straight_line(api, F, gamma, tp)
    return H.track_batch(td.start_solutions(), **kw)


@pytest.mark.parametrize("name", ["katsura", "tritangents", "cyclooctane", "biochem1", "steiner"])
def test_operator_api(oracle, gpu, name):
    rng = np.random.default_rng(11)
    F, P = {"katsura": (systems.katsura(8), 0), "tritangents": (systems.tritangents(), 20), "cyclooctane": (systems.cyclooctane(), 36),
            "biochem1": (systems.biochem1(), 10), "steiner": (ref_systems.steiner_higher_prec()[0], 30)}[name]
    n = F.n_vars
    p = rng.normal(size=P) + 1j * rng.normal(size=P); q = rng.normal(size=P) + 1j * rng.normal(size=P)
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    tx = np.stack([x, rng.normal(size=n) + 1j * rng.normal(size=n), rng.normal(size=n) + 1j * rng.normal(size=n)])
    out = []
    for api in (oracle, gpu):
        H = api.homotopy(capi.H_PARAMETER, api.system(F), p=p, q=q)
        u, U = H.evaluate_and_jacobian(x, 0.41)
        out.append([u, U, H.evaluate(x, 0.41), H.evaluate_dd(x, x * 2.0 ** -54, 0.41)] + [H.taylor(K, tx[:K], 0.41) for K in (1, 2, 3)])
    for a, b in zip(*out):
        assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(a).max())


def test_straight_line_operator_api(oracle, gpu):
    rng = np.random.default_rng(2)
    x = rng.normal(size=9) + 1j * rng.normal(size=9)
    tx = np.stack([x, 0.3 * x + 0.1j, 0.01 * x * x])
    out = []
    for api in (oracle, gpu):
        td, H = straight_line(api, systems.katsura(8), 0.4 + 1.3j)
        u, U = H.evaluate_and_jacobian(x, 0.37)
        out.append([u, U, H.evaluate_dd(x, 0 * x, 0.37)] + [H.taylor(K, tx[:K], 0.37) for K in (1, 2, 3)])
    for a, b in zip(*out):
        assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(a).max())


def test_endgame_golden_cases(oracle, gpu):  # reference test/endgame_tracker_test.jl:4-45
    ro, rg = (track_td(api, system_2x2(), 0.4 + 1.3j) for api in (oracle, gpu))
    assert_batches_match(ro, rg)
    assert [capi.ENDGAME_CODES[c] for c in rg.return_code] == ["success", "success", "at_infinity", "at_infinity"]
    assert np.allclose(rg.valuation[2], [-1, -1], rtol=1e-3) and rg.accuracy[0] < 1e-12 and rg.residual[0] < 1e-12
    for d in (2, 3, 5):
        F = make_system(lambda v, p: [(v[0] - 10) ** d], 1)
        ro, rg = (track_td(api, F, np.exp(2j * np.pi * 0.123)) for api in (oracle, gpu))
        assert_batches_match(ro, rg)
        assert (rg.winding_number == d).all() and rg.singular.all() and (0 < rg.last_t).all() and (rg.last_t < 0.1).all()


def test_katsura8_config1(oracle, gpu):  # BASELINE.json configs[0]: 256 paths
    ro, rg = (track_td(api, systems.katsura(8), 0.4 + 1.3j, nthreads=8) for api in (oracle, gpu))
    assert_batches_match(ro, rg)
    assert (rg.return_code == 1).sum() == 256 and rg.singular.sum() == 0
    assert rg.residual.max() < 1e-12
    assert abs(int(ro.accepted_steps.sum()) - int(rg.accepted_steps.sum())) <= 0.02 * ro.accepted_steps.sum()


def test_cyclic7_total_degree(oracle, gpu):  # reference test/endgame_test.jl:2-5: 924 of 5040 paths finite
    ro, rg = (track_td(api, systems.cyclic(7), np.exp(2j * np.pi * 0.7133), nthreads=8) for api in (oracle, gpu))
    assert int((rg.return_code == 1).sum()) == 924
    assert_batches_match(ro, rg)


def test_steiner_and_four_bar(oracle, gpu):  # golden endpoint vectors, needs DoubleDouble
    F, g = ref_systems.steiner_higher_prec()
    H = gpu.homotopy(capi.H_PARAMETER, gpu.system(F), p=g["p"], q=g["q"])
    r = H.track_batch([g["s_p"]], mode=1)
    assert capi.TRACKER_CODES[r.return_code[0]] == "success" and r.extended_precision_used[0]
    assert np.allclose(r.solution[0], g["s_q"], rtol=1.5e-8, atol=0)
    back = H.track_batch([r.solution[0]], mode=1, t1=0.0, t0=1.0, omega_mu=[r.omega[0], r.mu[0]])
    assert np.allclose(back.solution[0], g["s_p"], rtol=1.5e-8, atol=0)
    F, g = ref_systems.four_bar()
    r = gpu.homotopy(capi.H_PARAMETER, gpu.system(F), p=g["p"], q=g["q"]).track_batch([g["s"]], mode=1)
    assert capi.TRACKER_CODES[r.return_code[0]] == "success"


def test_invalid_start_value(oracle, gpu):  # reference test/tracker_test.jl:93-105
    F = make_system(lambda v, p: [v[0] ** 2 + v[1] ** 2 - 3, 2 * v[0] ** 2 + 0.5 * v[0] * v[1] + 3 * v[1] ** 2 - 2], 2)
    td, H = straight_line(gpu, F, 1j)
    r = H.track_batch([[100, -100], [1, 1]], mode=1)
    assert capi.TRACKER_CODES[r.return_code[0]] == "terminated_invalid_startvalue"
    assert capi.TRACKER_CODES[r.return_code[1]] == "success"


def test_lane_refill_is_deterministic(gpu):
    """Size-independent property: replicas of the same start must give bit-identical results whatever
    lane / queue position they get (catches cross-lane aliasing and state leaking between paths)."""
    td, H = straight_line(gpu, systems.katsura(8), 0.4 + 1.3j)
    S = td.start_solutions()
    R = 40
    r = H.track_batch(np.tile(S, (R, 1)))
    assert (r.return_code == 1).all()
    sol = r.solution.reshape(R, 256, -1)
    assert (sol == sol[0][None]).all()
    assert (r.accepted_steps.reshape(R, 256) == r.accepted_steps[:256][None]).all()


def test_page_locked_reused_result_buffers(gpu):
    """hc_host_register'ed inputs and a reused result struct (what bench.py's e2e arm and a Julia host that keeps
    its Result vectors do) give bit-identical results to freshly allocated pageable buffers; stale contents of
    the reused arrays are overwritten."""
    td, H = straight_line(gpu, systems.katsura(6), 0.4 + 1.3j)
    S = np.ascontiguousarray(np.tile(td.start_solutions(), (8, 1)))
    ref = H.track_batch(S)
    out = capi.BatchResults.allocate(H.n, S.shape[0])
    for a in out.arrays():
        a.view(np.uint8)[...] = 0xA5
    pinned = lib.pin(S, *out.arrays())
    assert len(pinned) == 1 + len(out.arrays())
    lib.pin(S)  # registering twice is not an error
    try:
        for _ in range(2):
            r = H.track_batch(S, out=out)
            assert r is out
            for a, b in zip(ref.arrays(), out.arrays()):
                assert np.array_equal(a, b, equal_nan=a.dtype.kind in "fc")
    finally:
        lib.unpin(pinned)
    with pytest.raises(ValueError):
        H.track_batch(S[:5], out=out)


def _same(a, b):
    for x, y in zip(a.arrays(), b.arrays()):
        assert np.array_equal(x, y, equal_nan=x.dtype.kind in "fc")


def test_total_degree_starts_made_on_the_device(oracle, gpu):
    """hc_track_total_degree: start solutions from the path index on the device == the streamed-in
    values of TotalDegreeStartSolutionsIterator (src/total_degree.jl:235-262), whole range and a shard's sub-range."""
    td, H = straight_line(gpu, systems.katsura(8), 0.4 + 1.3j)
    S = td.start_solutions()
    ref = H.track_batch(S)
    _same(ref, capi.track_total_degree(H, td.degrees))
    _same(H.track_batch(S[100:200]), capi.track_total_degree(H, td.degrees, first=100, count=100))
    _, Ho = straight_line(oracle, systems.katsura(8), 0.4 + 1.3j)
    assert_batches_match(Ho.track_batch(S), capi.track_total_degree(H, td.degrees))
    # lane-group engine (n > 14): every lane of a group derives the digits itself
    F = systems.katsura(15)
    td, H = straight_line(gpu, F, 0.4 + 1.3j)
    S = td.start_solutions()[:96]
    _same(H.track_batch(S), capi.track_total_degree(H, td.degrees, first=0, count=96))
    with pytest.raises(RuntimeError, match="exceeds"):
        capi.track_total_degree(H, td.degrees, first=2 ** 15 - 1, count=2)


def test_sweep_entry_point(oracle, gpu):  # many_solve, src/solve.jl:815-881
    F = systems.biochem1()
    rng = np.random.default_rng(11)
    p1 = rng.normal(size=10) + 1j * rng.normal(size=10)
    td, H0 = straight_line(gpu, F, 0.4 + 1.3j, p1)
    r0 = H0.track_batch(td.start_solutions())
    starts = r0.solution[(r0.return_code == 1) & (r0.singular == 0)]
    k, M = len(starts), 300
    q = (systems.BIOCHEM1_PVALS[None, :] * np.exp(0.5 * rng.normal(size=(M, 10)))).astype(np.complex128)
    H = gpu.homotopy(capi.H_PARAMETER, gpu.system(F), p=p1, q=q[0])
    S = np.repeat(starts[None], M, axis=0).reshape(-1, 3)
    Q = np.repeat(q[:, None, :], k, axis=1).reshape(-1, 10)
    got = capi.track_sweep(H, starts, q)
    assert lib.timing().h2d_bytes < 16 * (3 * k + 10 * M) + 4096   # only S starts and one column per point cross the bus
    _same(H.track_batch(S, path_q=Q), got)
    Ho = oracle.homotopy(capi.H_PARAMETER, oracle.system(F), p=p1, q=q[0])
    assert_batches_match(Ho.track_batch(S, path_q=Q), got)


def test_parameter_sweep(oracle, gpu):  # BASELINE.json configs[4] at test size
    F = systems.biochem1()
    rng = np.random.default_rng(5)
    p1 = rng.normal(size=10) + 1j * rng.normal(size=10)
    td, H0 = straight_line(gpu, F, 0.4 + 1.3j, p1)
    r0 = H0.track_batch(td.start_solutions())
    starts = r0.solution[r0.return_code == 1]
    K = 500
    q = systems.BIOCHEM1_PVALS[None, :] * np.exp(0.5 * rng.normal(size=(K, 10)))
    S = np.repeat(starts[None], K, axis=0).reshape(-1, 3)
    Q = np.repeat(q[:, None, :], len(starts), axis=1).reshape(-1, 10).astype(np.complex128)
    ro, rg = (api.homotopy(capi.H_PARAMETER, api.system(F), p=p1, q=q[0]).track_batch(S, path_q=Q, nthreads=8) for api in (oracle, gpu))
    assert_batches_match(ro, rg)
    # residual check at the target parameters (property that holds at any size)
    ok = rg.return_code == 1
    assert rg.residual[ok].max() < 1e-9


def test_cyclic7_polyhedral_config2(oracle, gpu):  # BASELINE.json configs[1]: 924 mixed-volume paths
    from hcb200 import workloads
    w = workloads.cyclic_polyhedral(7)
    ro, rg = (w.track(api, w.build(api), nthreads=8) for api in (oracle, gpu))
    assert_batches_match(ro, rg)
    assert (rg.return_code == 1).sum() == 924 and rg.singular.sum() == 0
    assert len(np.unique(np.round(rg.solution, 6), axis=0)) == 924
    assert rg.residual.max() < 1e-10


@pytest.mark.parametrize("engine", ["auto", "group"])
def test_polyhedral_starts_made_on_the_device(oracle, gpu, monkeypatch, engine):
    """hc_polyhedral_track_cells (second half): only per-cell data (Hermite normal form, angles, moduli)
    crosses the bus, a lane makes its start solution from its path index (BinomialSystemSolver, reference
    src/binomial_system.jl:55-106, 238-261).  Same results as the oracle's run from the host-made start solutions, same
    results as the explicit-start call, an index range gives the slice, and the upload no longer grows with N."""
    from hcb200 import workloads
    if engine == "group":
        monkeypatch.setenv("HC_B200_ENGINE", "group")
    w = workloads.cyclic_polyhedral(7 if engine == "auto" else 5)
    assert w.cells is not None and int(w.cells["volume"].sum()) == w.N
    ro = w.track(oracle, w.build(oracle), nthreads=8)            # explicit starts (the oracle has no start generator)
    h = w.build(gpu)
    rg = w.track(gpu, h)                                         # hc_polyhedral_track_cells
    cells_bytes = lib.timing().h2d_bytes
    assert_batches_match(ro, rg)
    rx = capi.polyhedral_track_batch(gpu, h["H"], h["Hcoeff"], w.starts, w.cell_index, w.cell_weights)
    assert lib.timing().h2d_bytes - cells_bytes >= 16 * w.n * w.N - 8 * (w.n * w.n + 2 * w.n + 2) * len(w.cells["volume"])
    assert (rx.return_code == rg.return_code).all()
    assert np.abs(rx.solution - rg.solution).max() <= 1e-10 * np.abs(rx.solution).max()
    lo, hi = w.N // 3, w.N // 3 + w.N // 2
    part = w.slice(lo, hi).track(gpu, h)
    assert np.array_equal(part.solution, rg.solution[lo:hi]) and (part.return_code == rg.return_code[lo:hi]).all()
    # replicated batch: indices beyond the mixed volume wrap
    big = capi.polyhedral_track_cells(gpu, h["H"], h["Hcoeff"], w.cells, w.cell_weights, first=0, count=2 * w.N + 3)
    assert np.array_equal(big.solution[w.N:2 * w.N], rg.solution) and np.array_equal(big.solution[2 * w.N:], rg.solution[:3])


@pytest.mark.parametrize("k", [1, 2, 3, 5, 7, 9, 10, 11, 12, 15])
def test_system_sizes(oracle, gpu, k):
    """n = k + 1: every register-blocked LU instantiation class (n <= 12), the generic LU of the
    thread-per-path engine (n = 13) and the lane-group engine (n = 16); 64 paths of katsura(k) each."""
    out = []
    for api in (oracle, gpu):
        td, H = straight_line(api, systems.katsura(k), 0.4 + 1.3j)
        out.append(H.track_batch(td.start_solutions()[:64], nthreads=8))
    assert_batches_match(*out)
    assert abs(int(out[0].accepted_steps.sum()) - int(out[1].accepted_steps.sum())) <= 0.02 * out[0].accepted_steps.sum()


def test_group_engine_matches_thread_engine(oracle, gpu, monkeypatch):
    """Both engines run the same per-path algorithm: force the lane-group engine on a small system."""
    monkeypatch.setenv("HC_B200_ENGINE", "group")
    ro, rg = (track_td(api, systems.katsura(6), 0.4 + 1.3j, nthreads=8) for api in (oracle, gpu))
    assert_batches_match(ro, rg)
    assert lib.timing().lanes in (8, 32)


def test_tritangents_slice_config3(oracle, gpu):
    """BASELINE.json configs[2] at test size: the first 4096 total-degree paths of tritangents (n = 12): every path in
    the same class, identical nonsingular solutions, clustering-independent ResultStatistics identical
    (helpers.assert_classes_match; which terminated_* code a path that dies at t < 1e-9 reports may differ)."""
    from hcb200 import workloads
    w = workloads.tritangents_total_degree().subset(4096)
    ro, rg = (w.track(api, w.build(api), nthreads=8) for api in (oracle, gpu))
    rep = assert_classes_match(ro, rg)
    assert rep["statistics_got"]["nonsingular"] > 0


def test_cyclooctane_slice_config4(oracle, gpu):
    """BASELINE.json configs[3] at test size on the total-degree start system (n = 17, singular endpoints at infinity
    via the endgame + DoubleDouble): the first 2048 paths."""
    from hcb200 import workloads
    w = workloads.cyclooctane_total_degree().subset(2048)
    ro, rg = (w.track(api, w.build(api), nthreads=8) for api in (oracle, gpu))
    assert_classes_match(ro, rg)


@pytest.mark.parametrize("name", ["tritangents", "cyclooctane_td", "cyclooctane_polyhedral"])
def test_two_pass_batches_on_the_heavy_tailed_configs(oracle, gpu, monkeypatch, name):
    """Large batches of the heavy-tailed configs run in two passes (hc_api.cu: second_pass): the specialised
    thread-per-path kernel hands over the paths that need more than 120 endgame steps or extended precision, the
    lane-group engine tracks those again from their start solutions.  Same classes and nonsingular solutions as the
    oracle; which pass tracks a path depends on the path alone, so two runs are bit-identical; with the hand-over
    switched off the same classes come out of the single pass."""
    from hcb200 import workloads
    monkeypatch.setenv("HC_B200_JIT", "1")
    w = {"tritangents": lambda: workloads.tritangents_total_degree().subset(4096),
         "cyclooctane_td": lambda: workloads.cyclooctane_total_degree().subset(2048),
         "cyclooctane_polyhedral": lambda: workloads.cyclooctane_polyhedral().subset(2048)}[name]()
    ro = w.track(oracle, w.build(oracle), nthreads=8)
    h = w.build(gpu)
    rg = w.track(gpu, h)
    tm = lib.timing()
    assert tm.engine == 2 and 0 < tm.handoff_paths < 0.2 * w.N, (tm.engine, tm.handoff_paths)
    assert (rg.return_code > 0).all()                      # no path is left handed over
    assert_classes_match(ro, rg)
    again = w.track(gpu, h)
    _same(rg, again)
    monkeypatch.setenv("HC_B200_HANDOFF", "0")
    single = w.track(gpu, h)
    assert lib.timing().handoff_paths == 0
    assert_classes_match(single, rg)


def test_operator_api_batched(oracle, gpu):
    """hc_evaluate_batch on the device: 20 000 points in one call (several launches of 8192 blocks) == the oracle"""
    rng = np.random.default_rng(2)
    _, H = straight_line(gpu, systems.katsura(6), 0.4 + 1.3j)
    _, Ho = straight_line(oracle, systems.katsura(6), 0.4 + 1.3j)
    X = rng.normal(size=(20000, 7)) + 1j * rng.normal(size=(20000, 7))
    u, U = H.evaluate_batch(X, 0.37, jacobian=True)
    for k in (0, 8191, 8192, 19999):
        a, A = Ho.evaluate_and_jacobian(X[k], 0.37)
        assert np.abs(u[k] - a).max() <= 1e-12 * max(1.0, np.abs(a).max()) and np.abs(U[k] - A).max() <= 1e-12 * max(1.0, np.abs(A).max())
    assert np.abs(H.evaluate_batch(X[:100], 0.37) - u[:100]).max() == 0.0


def test_sweep_with_device_side_counts(gpu):
    """hc_track_sweep_counts (result post-processing on the device; reference many_solve with a counting
    `transform_result`, src/solve.jl:422-430): per parameter point [nonsingular, singular, real, at infinity, failed] ==
    the same counts taken on the host from the full PathResults of hc_track_sweep; only 20 bytes per point come back."""
    from hcb200 import workloads
    w = workloads.biochem_sweep(gpu, 4096)
    h = w.build(gpu)
    full = capi.track_sweep(h["H"], w.sweep_starts, w.sweep_q)
    counts = capi.track_sweep_counts(h["H"], w.sweep_starts, w.sweep_q)
    assert lib.timing().d2h_bytes == 20 * 4096
    S = len(w.sweep_starts)
    rc = full.return_code.reshape(4096, S)
    ok = rc == 1
    sing = full.singular.reshape(4096, S) != 0
    real = (np.abs(full.solution.imag).max(axis=1) < 1e-6).reshape(4096, S)
    want = np.stack([(ok & ~sing).sum(1), (ok & sing).sum(1), (ok & real).sum(1), np.isin(rc, (2, 3)).sum(1),
                     (~ok & ~np.isin(rc, (2, 3))).sum(1)], axis=1)
    assert np.array_equal(counts, want) and counts[:, 0].sum() > 0 and (counts.sum(axis=1) - counts[:, 2] == S).all()


def test_homogeneous_and_overdetermined_inputs(oracle, gpu):
    """The host-side wrappers of SURVEY.md 8f-3 end in ordinary square systems for the device: a homogeneous system on a
    random affine chart (reference src/total_degree.jl:94-108) and an overdetermined one squared up with [I A]
    (:66-92, src/overdetermined.jl) -- GPU vs oracle, the excess solutions marked."""
    from hcb200 import start_systems
    quad = make_system(lambda v, p: [v[0] ** 2 + 2 * v[1] ** 2 - 3 * v[2] ** 2 + v[0] * v[1], v[0] * v[1] - 2 * v[2] ** 2 + v[1] * v[2] + 0.5 * v[0] ** 2], 3)
    over = make_system(lambda v, p: [v[0] + v[1] - 3.0, v[0] ** 2 + v[1] ** 2 - 5.0, v[0] * v[1] - 2.0], 2)
    for F in (quad, over):
        td = start_systems.total_degree(F, 0.4 + 1.3j)
        out = []
        for api in (oracle, gpu):
            H = api.homotopy(capi.H_STRAIGHT_LINE, api.system(td.F), api.system(td.G), gamma=td.gamma, G_params=td.scaling, F_params=[])
            out.append(start_systems.excess_solution_check(td, H.track_batch(td.start_solutions())))
        assert_batches_match(*out)
        if td.chart is not None:
            assert (out[1].return_code == 1).all() and np.abs(out[1].solution @ td.chart - 1).max() < 1e-12
        else:
            assert sorted(out[1].return_code.tolist()) == [1, 1, 14, 14]


def test_set_parameters_between_batches(gpu):
    """start_parameters! / target_parameters! / parameters! (reference test/tracker_test.jl:81-91 "Change parameters"):
    hc_homotopy_set_parameters rewrites the device copies of p and q; the next batch equals a homotopy created with
    those parameters bit for bit."""
    F = make_system(lambda v, p: [v[0] ** 2 - p[0], v[0] * v[1] - p[0] + p[1]], 2, 2)
    H = gpu.homotopy(capi.H_PARAMETER, gpu.system(F), p=[2.2, 3.2], q=[2.2, 3.2])
    H.set_parameters(p=[1, 0])
    H.set_parameters(q=[2, 4])
    r = H.track_batch([[1.0, 1.0]], mode=1)
    assert capi.TRACKER_CODES[r.return_code[0]] == "success"
    assert np.allclose(r.solution[0], [np.sqrt(2), -np.sqrt(2)])
    ref = gpu.homotopy(capi.H_PARAMETER, gpu.system(F), p=[1, 0], q=[2, 4]).track_batch([[1.0, 1.0]], mode=1)
    _same(ref, r)
    H.set_parameters(p=[2, 4], q=[1, 0])      # parameters!(T, p, q): the way back
    back = H.track_batch([r.solution[0]], mode=1)
    assert np.allclose(back.solution[0], [1, 1])
    td, Hs = straight_line(gpu, systems.katsura(3), 0.4 + 1.3j)
    with pytest.raises(RuntimeError, match="straight-line"):
        Hs.set_parameters(p=[])


def test_bad_inputs_fail_with_a_message_not_a_crash(gpu):
    """C-ABI validation: missing target parameters, out-of-range cell indices"""
    from hcb200 import polyhedral as ph
    with pytest.raises(RuntimeError):
        gpu.homotopy(capi.H_PARAMETER, gpu.system(systems.biochem1()), p=np.ones(10))
    ps = ph.polyhedral(systems.cyclic(4))
    S, ci = ps.start_solutions()
    h = gpu.system(ps.F)
    Ht, Hc = gpu.homotopy(capi.H_TORIC, h, p=ps.start_coeffs), gpu.homotopy(capi.H_COEFFICIENT, h, p=ps.start_coeffs, q=ps.target_coeffs)
    bad = ci.copy(); bad[0] = len(ps.cell_weights()) + 3
    with pytest.raises(RuntimeError):
        capi.polyhedral_track_batch(gpu, Ht, Hc, S, bad, ps.cell_weights())
    assert "cell_index" in lib.last_error()
    assert (capi.polyhedral_track_batch(gpu, Ht, Hc, S, ci, ps.cell_weights()).return_code > 0).all()


def test_all_devices_through_one_call(gpu):
    """hc_init_devices: one host call drives every visible GPU -- the path index range is split into contiguous shards,
    each device copies its slice into the caller's arrays -- and the result equals the single-device run bit for bit
    (reference: one solve() call, all workers, results by path index; src/solve.jl:628-709).  On a one-GPU box this
    still exercises the sharded code path with a single shard."""
    import torch
    ndev = torch.cuda.device_count()
    from hcb200 import start_systems
    td = start_systems.total_degree(systems.katsura(6), 0.4 + 1.3j)
    S = np.tile(td.start_solutions(), (96, 1))

    def run(api):
        H = api.homotopy(capi.H_STRAIGHT_LINE, api.system(td.F), api.system(td.G), gamma=td.gamma, G_params=td.scaling, F_params=[])
        r = H.track_batch(S)
        return r, lib.timing().devices
    try:
        one, d1 = run(lib.load(devices=[0]))
        many, dn = run(lib.load(devices=list(range(ndev))))
        assert d1 == 1 and dn == ndev
        _same(one, many)
        assert (many.return_code == 1).all()
    finally:
        lib.load(devices=[0])


def test_cancel_stops_handing_out_paths(gpu):
    """hc_request_cancel (stop_early_cb / interrupt, reference src/solve.jl:685-707): a cancelled batch returns; the
    paths that were never started keep return_code 0 (tracking), the rest are complete results."""
    import threading
    import time
    td, H = straight_line(gpu, systems.katsura(8), 0.4 + 1.3j)
    S = np.tile(td.start_solutions(), (1200, 1))      # ~0.5 s of tracking
    t = threading.Timer(0.05, lambda: lib.load().raw.hc_request_cancel(1))
    t.start()
    t0 = time.perf_counter()
    r = H.track_batch(S)
    dt = time.perf_counter() - t0
    t.join()
    done = r.return_code != 0
    assert 0 < done.sum() < len(S), (int(done.sum()), dt)
    assert (r.return_code[done] == 1).all()
    full = H.track_batch(S[:256])                     # the flag is cleared on entry of the next call
    assert (full.return_code == 1).all()


def test_polyhedral_on_the_lane_group_engine(oracle, gpu, monkeypatch):
    """MODE_POLYHEDRAL (toric stage, re-weighted restart, coefficient stage; src/polyhedral.jl:414-530) on the lane-group
    engine: forced onto a small system, where the thread-per-path engine would run by default."""
    from hcb200 import polyhedral as ph
    monkeypatch.setenv("HC_B200_ENGINE", "group")
    ps = ph.polyhedral(systems.cyclic(5))
    S, ci = ps.start_solutions()
    res = []
    for api in (oracle, gpu):
        h = api.system(ps.F)
        res.append(capi.polyhedral_track_batch(api, api.homotopy(capi.H_TORIC, h, p=ps.start_coeffs),
                                               api.homotopy(capi.H_COEFFICIENT, h, p=ps.start_coeffs, q=ps.target_coeffs), S, ci, ps.cell_weights()))
    assert lib.timing().lanes in (8, 32)
    assert_batches_match(*res)
    assert (res[1].return_code == 1).sum() == 70


def test_cyclooctane_polyhedral_slice_config4(oracle, gpu):
    """BASELINE.json configs[3] as the benchmark runs it (benchmarks/cyclooctane.jl:27-29: solve(F0), polyhedral start
    system by default, src/solve.jl:121): n = 17, lane-group engine, the first 2048 of 32 768 mixed-volume paths."""
    from hcb200 import workloads
    w = workloads.cyclooctane_polyhedral().subset(2048)
    ro, rg = (w.track(api, w.build(api), nthreads=8) for api in (oracle, gpu))
    assert lib.timing().lanes in (8, 32)
    assert_classes_match(ro, rg)
    assert (rg.return_code == 1).sum() > 0
