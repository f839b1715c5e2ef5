"""The specialised (run-time compiled) kernels: per-system straight-line code generated from the reference tape
(csrc/hc_jitgen.h), compiled with NVRTC for sm_100a (csrc/hc_jit.h) and run by the lockstep thread-per-path kernel.

CPU part: the very same generated unit, compiled by g++ (tests/host_sim), against the oracle -- operator API
(reference test/model_kit/e2e_test.jl:42-52: evaluate / Jacobian / taylor agree to 1e-12) and whole tracked batches;
the NVRTC compile itself runs offline (HC_B200_NO_DEVICE).  GPU part: the kernels against the oracle, against the
interpreter engine, and -- built without FMA contraction -- against the host-compiled code bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import assert_batches_match, assert_classes_match, assert_last_point_on_path, compare_batches, straight_line, system_2x2
from hcb200 import capi, systems
from hcb200.modelkit import make_system

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture
def jit_on(monkeypatch):
    monkeypatch.setenv("HC_B200_JIT", "1")


def _rational_system():
    """ops beyond polynomials: division, inverse, integer powers (literal exponents > 3 and negative), x^0"""
    return make_system(lambda v, p: [v[0] ** 5 / (2.0 + v[1]) + p[0] * v[1] ** 4 - 1.0,
                                     (v[0] * v[1]) ** 2 + p[1] / v[0] + v[1] ** (-2) - 3.0 + v[0] ** 0], 2, n_params=2)


@pytest.mark.parametrize("name", ["katsura4_sl", "biochem_parameter", "cyclic4_toric", "rational_parameter"])
def test_generated_code_operator_api(oracle, sim, jit_on, name):
    rng = np.random.default_rng(3)
    out = []
    for api in (oracle, sim):
        t = 0.37
        if name == "katsura4_sl":
            _, H = straight_line(api, systems.katsura(4), 0.4 + 1.3j)
        elif name == "biochem_parameter":
            F = systems.biochem1()
            g = np.random.default_rng(5)
            H = api.homotopy(capi.H_PARAMETER, api.system(F), p=g.normal(size=10) + 1j * g.normal(size=10), q=g.normal(size=10) + 1j * g.normal(size=10))
        elif name == "rational_parameter":
            H = api.homotopy(capi.H_PARAMETER, api.system(_rational_system()), p=[0.3 + 0.1j, -1.2j], q=[1.0, 0.4 - 0.2j])
        else:
            from hcb200 import polyhedral as ph
            ps = ph.polyhedral(systems.cyclic(4))
            H = api.homotopy(capi.H_TORIC, api.system(ps.F), p=ps.start_coeffs)
            H.set_toric_weights(np.abs(np.random.default_rng(9).normal(size=H.P)) + 0.1)
        n = H.n
        x = rng.normal(size=n) + 1j * rng.normal(size=n)
        tx = np.stack([x, 0.3 * x + 0.1j, 0.01 * x * x])
        u, U = H.evaluate_and_jacobian(x, t)
        out.append([u, U, H.evaluate(x, t)] + [H.taylor(K, tx[:K], t) for K in (1, 2, 3)])
        rng = np.random.default_rng(3)
    for a, b in zip(*out):
        assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(a).max())


def test_generated_code_tracks_like_the_oracle(oracle, sim, jit_on):
    """whole batches through the specialised unit: total degree (straight line), endgame with winding numbers,
    polyhedral two-stage driver, parameter sweep with per-path parameter rows"""
    both = lambda build: [build(api) for api in (oracle, sim)]
    td_track = lambda F, gamma: (lambda api: (lambda tdH: tdH[1].track_batch(tdH[0].start_solutions()))(straight_line(api, F, gamma)))
    ro, rs = both(td_track(systems.katsura(5), 0.4 + 1.3j))
    assert_batches_match(ro, rs)
    assert (rs.return_code == 1).sum() == 32
    ro, rs = both(td_track(system_2x2(), 0.4 + 1.3j))
    assert_batches_match(ro, rs)
    ro, rs = both(td_track(make_system(lambda v, p: [(v[0] - 10) ** 3], 1), np.exp(0.77j)))
    assert_batches_match(ro, rs)
    assert (rs.winding_number == 3).all() and rs.singular.all()
    from hcb200 import polyhedral as ph
    ps = ph.polyhedral(systems.cyclic(5))
    S, ci = ps.start_solutions()

    def poly(api):
        h = api.system(ps.F)
        return capi.polyhedral_track_batch(api, api.homotopy(capi.H_TORIC, h, p=ps.start_coeffs),
                                           api.homotopy(capi.H_COEFFICIENT, h, p=ps.start_coeffs, q=ps.target_coeffs), S, ci, ps.cell_weights())
    ro, rs = both(poly)
    assert_batches_match(ro, rs)
    assert (rs.return_code == 1).sum() == 70
    F = systems.biochem1()
    rng = np.random.default_rng(5)
    p1 = rng.normal(size=10) + 1j * rng.normal(size=10)
    td, H0 = straight_line(oracle, F, 0.4 + 1.3j, p1)
    r0 = H0.track_batch(td.start_solutions())
    starts = r0.solution[r0.return_code == 1]
    q = systems.BIOCHEM1_PVALS[None, :] * np.exp(0.5 * rng.normal(size=(5, 10)))
    Sx = np.repeat(starts[None], 5, axis=0).reshape(-1, 3)
    Q = np.repeat(q[:, None, :], len(starts), axis=1).reshape(-1, 10).astype(np.complex128)
    ro, rs = both(lambda api: api.homotopy(capi.H_PARAMETER, api.system(F), p=p1, q=q[0]).track_batch(Sx, path_q=Q))
    assert_batches_match(ro, rs)
    Hs = sim.homotopy(capi.H_PARAMETER, sim.system(F), p=p1, q=q[0])
    assert_last_point_on_path(Hs, Hs.track_batch(starts))


def test_generated_code_n17_register_lu(oracle, sim, jit_on):
    """cyclooctane (n = 17) through the specialised unit: the register-blocked LU beyond n = 12 (two-word packed row
    permutation) and a polyhedral batch, 24 paths each"""
    from hcb200 import workloads
    for w in (workloads.cyclooctane_total_degree(24), workloads.cyclooctane_polyhedral().subset(24)):
        ro, rs = (w.track(api, w.build(api)) for api in (oracle, sim))
        assert_batches_match(ro, rs)


def test_generated_unit_equals_interpreter_classes(sim, monkeypatch):
    """the two device engines run the same tracker: same classes and endpoints on a batch with diverging paths"""
    F = systems.cyclic(5)
    out = []
    for jit in ("0", "1"):
        monkeypatch.setenv("HC_B200_JIT", jit)
        td, H = straight_line(sim, F, np.exp(2j * np.pi * 0.7133))
        out.append(H.track_batch(td.start_solutions()))
    assert_batches_match(*out)
    assert (out[1].return_code == 1).sum() == 70


def test_nvrtc_builds_the_specialised_kernel_offline():
    """NVRTC targets sm_100a without a device: generate + compile the katsura(4) kernel in a fresh process
    (HC_B200_NO_DEVICE lets the handles live in host memory) and check that a cubin came out."""
    code = r"""
import ctypes as C, sys
sys.path[:0] = [%r]
import hcb200
from hcb200 import capi, start_systems, systems
lib = C.CDLL(%r)
api = capi.CApi(lib, "hc_")
lib.hc_last_error.restype = C.c_char_p
lib.hc_jit_prepare.restype = C.c_int32
lib.hc_jit_prepare.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_double)]
td = start_systems.total_degree(systems.katsura(4), 0.4 + 1.3j)
H = api.homotopy(capi.H_STRAIGHT_LINE, api.system(td.F), api.system(td.G), gamma=td.gamma, G_params=td.scaling, F_params=[])
info = (C.c_double * 5)()
rc = lib.hc_jit_prepare(H.handle, 0, info)
assert rc == 0, lib.hc_last_error().decode()
print("cubin", int(info[0]), "hot", int(info[2]))
assert info[0] > 100000 and 0 < info[2] < 8192
""" % (ROOT, os.path.join(ROOT, "homotopycontinuation.jl_b200", "libhc_b200.so"))
    if not os.path.exists(os.path.join(ROOT, "homotopycontinuation.jl_b200", "libhc_b200.so")):
        import __graft_entry__ as ge
        ge.build()
    env = dict(os.environ, HC_B200_NO_DEVICE="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cubin" in r.stdout


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_jit_kernels_against_the_oracle(oracle, gpu, jit_on):
    from hcb200 import lib, workloads
    for w in (workloads.katsura8(1), workloads.cyclic_polyhedral(7, 1)):
        ro = w.track(oracle, w.build(oracle), nthreads=8)
        hg = w.build(gpu)
        rg = w.track(gpu, hg)
        assert lib.timing().lanes == 1
        assert_batches_match(ro, rg)
        assert (rg.return_code == 1).sum() == w.N
    w = workloads.katsura8(1)
    h = w.build(gpu)
    assert_last_point_on_path(h["H"], w.track(gpu, h))


@pytest.mark.gpu
def test_jit_sweep_with_per_path_parameters(oracle, gpu, jit_on):
    from hcb200 import workloads
    p1, starts = workloads.biochem_generic_start(oracle)
    w = workloads.biochem_sweep_from_starts(starts, p1, 64)
    ro = w.track(oracle, w.build(oracle), nthreads=8)
    rg = w.track(gpu, w.build(gpu))
    assert_batches_match(ro, rg)


@pytest.mark.gpu
def test_jit_large_batch_is_deterministic_and_equals_interpreter(gpu, monkeypatch):
    """auto policy: a batch above HC_B200_JIT_MIN_PATHS runs on the specialised kernel; replicas of the same start
    give bit-identical results whatever lane they land on, and the classes equal the interpreter engine's"""
    from hcb200 import lib
    td, H = straight_line(gpu, systems.katsura(6), 0.4 + 1.3j)
    S = td.start_solutions()
    R = 160
    monkeypatch.setenv("HC_B200_JIT", "0")
    ri = H.track_batch(np.tile(S, (2, 1)))
    monkeypatch.delenv("HC_B200_JIT")
    monkeypatch.setenv("HC_B200_JIT_MIN_PATHS", "4096")
    rj = H.track_batch(np.tile(S, (R, 1)))
    assert lib.timing().block >= 128 and (rj.return_code == 1).all()
    sol = rj.solution.reshape(R, len(S), -1)
    assert (sol == sol[0][None]).all()
    one = capi.BatchResults.allocate(H.n, 2 * len(S))
    for a, b in zip(one.arrays(), rj.arrays()):
        a[...] = b[:2 * len(S)]
    assert_batches_match(ri, one)


@pytest.mark.gpu
def test_gpu_is_bit_identical_to_the_host_compiled_code(sim, gpu, monkeypatch):
    """The kernels are built without FMA contraction (like Julia's arithmetic and the oracle's parity build), and then
    nothing but libm separates the GPU from the CPU: the specialised kernel reproduces the g++ build of the same unit
    (x86-64 baseline: no contraction) BIT FOR BIT -- return code, accepted / rejected steps, endpoint, t, accuracy -- on
    every path of a tritangents slice (diverging, dying at t < 1e-9, extended-precision and winding-number-4 singular
    paths included) except the singular endpoints of winding number 3, whose Cauchy endgame takes a cube root (cbrt of
    CUDA's libdevice and of glibc differ in the last bit); those still agree in code, winding number and, to the endgame's
    accuracy, endpoint.  (+, -, *, /, sqrt and fma are IEEE-exact on both sides.)  Measured first with
    tests/tools/gpu_parity_build.py: 739 of 768 paths bit-identical, the other 29 all of winding number 3.
    With contraction switched on (HC_B200_FMAD=1, + 4 ... 7 % paths/s) the classes still agree but the last bits do not."""
    from hcb200 import workloads
    monkeypatch.setenv("HC_B200_JIT", "1")
    monkeypatch.setenv("HC_B200_HANDOFF", "0")   # the kernel itself: no second pass on the (differently rounding) interpreter
    w = workloads.tritangents_total_degree().subset(768)
    rs = w.track(sim, w.build(sim))
    rg = w.track(gpu, w.build(gpu))
    assert (rs.return_code == rg.return_code).all(), np.flatnonzero(rs.return_code != rg.return_code)
    assert (rs.winding_number == rg.winding_number).all() and (rs.singular == rg.singular).all()
    exact = np.isin(rs.winding_number, (0, 1, 2, 4))      # nthroot(., m) is sqrt-only for these m
    assert exact.sum() > 0.9 * w.N and (rs.extended_precision_used[exact] != 0).any() and (rs.winding_number[exact] > 1).any()
    assert (rs.accepted_steps[exact] == rg.accepted_steps[exact]).all() and (rs.rejected_steps[exact] == rg.rejected_steps[exact]).all()
    assert np.array_equal(rs.solution[exact], rg.solution[exact], equal_nan=True)
    assert np.array_equal(rs.t[exact], rg.t[exact]) and np.array_equal(rs.accuracy[exact], rg.accuracy[exact], equal_nan=True)
    rest = ~exact
    if rest.any():
        tol = max(1e-6, 10 * float(np.nanmax(rs.accuracy[rest])))
        assert np.abs(rs.solution[rest] - rg.solution[rest]).max() < tol
    monkeypatch.setenv("HC_B200_FMAD", "1")
    rf = w.track(gpu, w.build(gpu))   # contraction on: same classes, different last bits
    assert_classes_match(rs, rf)
    assert not np.array_equal(rs.solution[exact], rf.solution[exact], equal_nan=True)


@pytest.mark.gpu
def test_interpreter_engine_is_bit_identical_to_its_host_build_too(sim, gpu, monkeypatch):
    """The ahead-of-time interpreter kernels are compiled -fmad=false as well: katsura(6) and the 72 total-degree paths
    of bio-chemical network 3 at the template parameters where, with contraction, path 49 jumped onto a neighbouring
    path on the GPU (12 steps to a duplicate solution instead of 85 steps to infinity; tests/tools/gpu_bio3_probe2.py)."""
    monkeypatch.setenv("HC_B200_JIT", "0")
    rng = np.random.default_rng(203)
    for _ in range(5):
        rng.random()
        pt = (rng.normal(size=8) + 1j * rng.normal(size=8)) / np.sqrt(2)
        g2 = np.exp(2j * np.pi * rng.random())
    for F, gamma, tp in ((systems.biochem2(), g2, pt), (systems.katsura(6), 0.4 + 1.3j, None)):
        out = []
        for api in (sim, gpu):
            td, H = straight_line(api, F, gamma, tp)
            out.append(H.track_batch(td.start_solutions()))
        rs, rg = out
        assert (rs.return_code == rg.return_code).all()
        assert (rs.accepted_steps == rg.accepted_steps).all() and (rs.rejected_steps == rg.rejected_steps).all()
        ok = rs.return_code == 1
        same = (rs.solution[ok] == rg.solution[ok]).all(axis=1)
        # (katsura(6): 63 of 64 endpoints are bit-identical, one differs by an ulp -- tests/tools/gpu_interp_bits.py)
        assert same.mean() >= 0.95 and np.abs(rs.solution[ok] - rg.solution[ok]).max() < 1e-15
        if F.n_vars == 3:
            assert same.all()
    assert rg.N == 64
