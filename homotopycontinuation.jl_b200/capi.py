"""ctypes binding of the C ABI declared in ``include/hc_b200.h``.

``CApi`` is parameterised by the loaded library and the symbol prefix so that the test harness
can drive the CPU oracle (``oracle/hc_oracle.h``, prefix ``orc_``) through the very same
plain-data structs.  Product code only ever instantiates it on ``libhc_b200.so`` (prefix
``hc_``) -- see ``lib.py``; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from dataclasses import dataclass

import numpy as np

from .modelkit import Program, System

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)
c_int64_p = C.POINTER(C.c_int64)


class ProgramDesc(C.Structure):
    _fields_ = [
        ("instructions", c_int32_p), ("n_instructions", C.c_int32),
        ("constants", c_double_p), ("n_constants", C.c_int32),
        ("param_offset", C.c_int32), ("n_params", C.c_int32),
        ("t_index", C.c_int32),
        ("var_offset", C.c_int32), ("n_vars", C.c_int32),
        ("u_assign", c_int32_p), ("n_u", C.c_int32),
        ("U_assign", c_int32_p), ("n_U", C.c_int32),
        ("out_dim", C.c_int32), ("tape_space", C.c_int32),
    ]


class HomotopyDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("F", C.c_void_p), ("G", C.c_void_p),
        ("gamma", C.c_double * 2),
        ("G_params", c_double_p), ("n_G_params", C.c_int32),
        ("F_params", c_double_p), ("n_F_params", C.c_int32),
        ("p", c_double_p), ("q", c_double_p), ("n_pq", C.c_int32),
    ]


class Options(C.Structure):
    """TrackerOptions + TrackerParameters + EndgameOptions + WeightedNormOptions
    (reference src/tracker.jl:45-62, 94-140; src/endgame_tracker.jl:47-72; src/norm.jl:36-40)."""
    _fields_ = [
        ("max_steps", C.c_int32), ("max_step_size", C.c_double), ("max_initial_step_size", C.c_double),
        ("extended_precision", C.c_int32), ("min_step_size", C.c_double), ("min_rel_step_size", C.c_double),
        ("a", C.c_double), ("beta_a", C.c_double), ("beta_omega_p", C.c_double), ("beta_tau", C.c_double),
        ("strict_beta_tau", C.c_double), ("min_newton_iters", C.c_int32),
        ("endgame_start", C.c_double), ("max_endgame_steps", C.c_int32), ("max_endgame_extended_steps", C.c_int32),
        ("min_cond", C.c_double), ("min_cond_growth", C.c_double), ("min_coord_growth", C.c_double),
        ("zero_is_at_infinity", C.c_int32), ("at_infinity_check", C.c_int32), ("only_nonsingular", C.c_int32),
        ("singular_min_accuracy", C.c_double), ("max_winding_number", C.c_int32),
        ("val_finite_tol", C.c_double), ("val_at_infinity_tol", C.c_double), ("sing_cond", C.c_double),
        ("sing_accuracy", C.c_double), ("scaling_threshold", C.c_double), ("refine_steps", C.c_int32),
        ("scale_min", C.c_double), ("scale_abs_min", C.c_double), ("scale_max", C.c_double),
    ]


class ResultsDesc(C.Structure):
    _fields_ = [
        ("return_code", c_int32_p), ("solution", c_double_p), ("t", c_double_p), ("accuracy", c_double_p),
        ("residual", c_double_p), ("singular", c_uint8_p), ("condition_jacobian", c_double_p),
        ("winding_number", c_int32_p), ("extended_precision", c_uint8_p), ("last_point", c_double_p),
        ("last_t", c_double_p), ("valuation", c_double_p), ("has_valuation", c_uint8_p), ("omega", c_double_p),
        ("mu", c_double_p), ("accepted_steps", c_int32_p), ("rejected_steps", c_int32_p), ("steps_eg", c_int32_p),
        ("extended_precision_used", c_uint8_p), ("counters", c_int64_p),
    ]


# EndgameTrackerCode order: reference src/endgame_tracker.jl:100-117
ENDGAME_CODES = [
    "tracking", "success", "at_infinity", "at_zero", "terminated_accuracy_limit",
    "terminated_invalid_startvalue", "terminated_invalid_startvalue_singular_jacobian",
    "terminated_ill_conditioned", "terminated_max_steps", "terminated_max_extended_steps",
    "terminated_max_winding_number", "terminated_step_size_too_small", "terminated_unknown",
    "post_check_failed", "excess_solution", "polyhedral_failed",
]
# TrackerCode order: reference src/tracker.jl:166-176
TRACKER_CODES = [
    "tracking", "success", "terminated_max_steps", "terminated_accuracy_limit",
    "terminated_ill_conditioned", "terminated_invalid_startvalue",
    "terminated_invalid_startvalue_singular_jacobian", "terminated_step_size_too_small",
    "terminated_unknown",
]

H_STRAIGHT_LINE, H_PARAMETER, H_COEFFICIENT, H_TORIC = 0, 1, 2, 3


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int32_p)


def _cflat(z, n=None):
    """complex array -> contiguous float64 (re, im interleaved)."""
    a = np.ascontiguousarray(np.asarray(z, dtype=np.complex128))
    if n is not None:
        assert a.size == n, (a.size, n)
    return a.view(np.float64).reshape(-1)


@dataclass
class BatchResults:
    """Struct-of-arrays PathResult batch (fields of reference src/path_result.jl:76-98)."""
    n: int
    N: int
    return_code: np.ndarray
    solution: np.ndarray          # (N, n) complex
    t: np.ndarray
    accuracy: np.ndarray
    residual: np.ndarray
    singular: np.ndarray
    condition_jacobian: np.ndarray
    winding_number: np.ndarray
    extended_precision: np.ndarray
    last_point: np.ndarray        # (N, n) complex
    last_t: np.ndarray
    valuation: np.ndarray         # (N, n)
    has_valuation: np.ndarray
    omega: np.ndarray
    mu: np.ndarray
    accepted_steps: np.ndarray
    rejected_steps: np.ndarray
    steps_eg: np.ndarray
    extended_precision_used: np.ndarray
    counters: np.ndarray          # (N, 8) int64

    @staticmethod
    def allocate(n: int, N: int) -> "BatchResults":
        f = lambda *s: np.zeros(s, dtype=np.float64)
        return BatchResults(
            n, N, np.zeros(N, np.int32), np.zeros((N, n), np.complex128), f(N), f(N), f(N),
            np.zeros(N, np.uint8), f(N), np.zeros(N, np.int32), np.zeros(N, np.uint8),
            np.zeros((N, n), np.complex128), f(N), f(N, n), np.zeros(N, np.uint8), f(N), f(N),
            np.zeros(N, np.int32), np.zeros(N, np.int32), np.zeros(N, np.int32), np.zeros(N, np.uint8),
            np.zeros((N, 8), np.int64))

    def arrays(self) -> list:
        return [getattr(self, f.name) for f in dataclasses.fields(self) if isinstance(getattr(self, f.name), np.ndarray)]

    def desc(self) -> ResultsDesc:
        u8 = lambda a: a.ctypes.data_as(c_uint8_p)
        return ResultsDesc(
            _ip(self.return_code), _dp(self.solution.view(np.float64)), _dp(self.t), _dp(self.accuracy),
            _dp(self.residual), u8(self.singular), _dp(self.condition_jacobian), _ip(self.winding_number),
            u8(self.extended_precision), _dp(self.last_point.view(np.float64)), _dp(self.last_t),
            _dp(self.valuation), u8(self.has_valuation), _dp(self.omega), _dp(self.mu),
            _ip(self.accepted_steps), _ip(self.rejected_steps), _ip(self.steps_eg),
            u8(self.extended_precision_used), self.counters.ctypes.data_as(c_int64_p))


class CApi:
    def __init__(self, lib: C.CDLL, prefix: str):
        self.lib, self.prefix = lib, prefix
        f = self._fn
        f("options_default", None, [C.POINTER(Options)])
        f("system_create", C.c_void_p, [C.POINTER(ProgramDesc), C.POINTER(ProgramDesc)])
        f("system_destroy", None, [C.c_void_p])
        f("homotopy_create", C.c_void_p, [C.POINTER(HomotopyDesc)])
        f("homotopy_destroy", None, [C.c_void_p])
        f("track_batch", C.c_int32, [C.c_void_p, C.POINTER(Options), C.c_int32, C.c_int64, c_double_p, c_double_p,
                                     c_double_p, c_double_p, c_double_p, c_double_p, C.POINTER(ResultsDesc), C.c_int32])
        f("polyhedral_track_batch", C.c_int32, [C.c_void_p, C.c_void_p, C.POINTER(Options), C.c_int64, c_double_p,
                                                c_int32_p, c_double_p, C.c_int32, C.POINTER(ResultsDesc), C.c_int32])
        f("evaluate", C.c_int32, [C.c_void_p, c_double_p, c_double_p, c_double_p])
        f("evaluate_dd", C.c_int32, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p])
        f("evaluate_and_jacobian", C.c_int32, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p])
        f("taylor", C.c_int32, [C.c_void_p, C.c_int32, c_double_p, c_double_p, c_double_p])
        f("toric_set_weights", C.c_int32, [C.c_void_p, c_double_p])
        f("homotopy_set_parameters", C.c_int32, [C.c_void_p, c_double_p, c_double_p], optional=True)
        # device-side start generation: entry points of libhc_b200 only (the oracle takes explicit starts)
        f("track_total_degree", C.c_int32, [C.c_void_p, C.POINTER(Options), c_int32_p, C.c_int64, C.c_int64,
                                            C.POINTER(ResultsDesc)], optional=True)
        f("polyhedral_track_cells", C.c_int32, [C.c_void_p, C.c_void_p, C.POINTER(Options), C.c_int64, C.c_int64, C.c_int32,
                                                c_int64_p, c_int64_p, c_double_p, c_double_p, c_double_p,
                                                C.POINTER(ResultsDesc)], optional=True)
        f("evaluate_batch", C.c_int32, [C.c_void_p, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p], optional=True)
        f("track_sweep_counts", C.c_int32, [C.c_void_p, C.POINTER(Options), C.c_int64, c_double_p, C.c_int64, c_double_p, C.c_double,
                                            c_int32_p], optional=True)
        f("unique_points_filter", C.c_int32, [C.c_int32, C.c_int64, c_double_p, C.c_int64, c_double_p, C.c_double, C.c_double,
                                              c_int64_p], optional=True)
        f("track_sweep", C.c_int32, [C.c_void_p, C.POINTER(Options), C.c_int64, c_double_p, C.c_int64, c_double_p,
                                     C.POINTER(ResultsDesc)], optional=True)

    def _fn(self, name, restype, argtypes, optional=False):
        fn = getattr(self.lib, self.prefix + name, None) if optional else getattr(self.lib, self.prefix + name)
        if fn is None:
            setattr(self, "_" + name, None)
            return
        fn.restype, fn.argtypes = restype, argtypes
        setattr(self, "_" + name, fn)

    # ---- options
    def default_options(self, **kw) -> Options:
        o = Options()
        self._options_default(C.byref(o))
        for k, v in kw.items():
            if not hasattr(o, k):
                raise TypeError(f"unknown option {k}")
            setattr(o, k, v)
        return o

    # ---- systems / homotopies
    def _program_desc(self, p: Program, keep: list) -> ProgramDesc:
        ins = np.ascontiguousarray(p.instructions, dtype=np.int32)
        cst = _cflat(p.constants)
        ua = np.ascontiguousarray(p.u_assign, dtype=np.int32)
        Ua = np.ascontiguousarray(p.U_assign, dtype=np.int32)
        keep += [ins, cst, ua, Ua]
        return ProgramDesc(_ip(ins), ins.shape[0], _dp(cst), len(p.constants), p.param_offset, p.n_params,
                           p.t_index, p.var_offset, p.n_vars, _ip(ua), ua.shape[0], _ip(Ua), Ua.shape[0],
                           p.out_dim, p.tape_space)

    def system(self, S: System) -> "SystemHandle":
        keep: list = []
        e = self._program_desc(S.eval_program, keep)
        j = self._program_desc(S.jac_program, keep)
        h = self._system_create(C.byref(e), C.byref(j))
        if not h:
            raise RuntimeError("system_create failed (unsupported op or device error)")
        return SystemHandle(self, h, S.n_eqs, S.n_vars, S.n_params)

    def homotopy(self, kind: int, F: "SystemHandle", G: "SystemHandle | None" = None, gamma=0j,
                 G_params=None, F_params=None, p=None, q=None) -> "HomotopyHandle":
        d = HomotopyDesc()
        keep = []
        d.kind, d.F, d.G = kind, F.handle, (G.handle if G is not None else None)
        d.gamma[0], d.gamma[1] = complex(gamma).real, complex(gamma).imag
        def setv(name, cnt, v, need=None):
            if v is None:
                setattr(d, name, None)
                if cnt: setattr(d, cnt, 0)
                return
            a = _cflat(v, need)
            keep.append(a)
            setattr(d, name, _dp(a))
            if cnt: setattr(d, cnt, a.size // 2)
        setv("G_params", "n_G_params", G_params)
        setv("F_params", "n_F_params", F_params)
        setv("p", "n_pq", p)
        setv("q", None, q)
        h = self._homotopy_create(C.byref(d))
        if not h:
            raise RuntimeError("homotopy_create failed")
        return HomotopyHandle(self, h, kind, F, G)


class SystemHandle:
    def __init__(self, api: CApi, handle, m, n, P):
        self.api, self.handle, self.m, self.n, self.P = api, handle, m, n, P

    def __del__(self):
        try:
            self.api._system_destroy(self.handle)
        except Exception:
            pass


class HomotopyHandle:
    def __init__(self, api: CApi, handle, kind, F: SystemHandle, G):
        self.api, self.handle, self.kind, self.F, self.G = api, handle, kind, F, G
        self.m, self.n, self.P = F.m, F.n, F.P

    def __del__(self):
        try:
            self.api._homotopy_destroy(self.handle)
        except Exception:
            pass

    def set_parameters(self, p=None, q=None):
        """start_parameters! / target_parameters! / parameters! (reference src/endgame_tracker.jl:831-840)."""
        if self.api._homotopy_set_parameters is None:
            raise RuntimeError("this library cannot change the parameters of a homotopy")
        pa = _cflat(p, self.P) if p is not None else None
        qa = _cflat(q, self.P) if q is not None else None
        rc = self.api._homotopy_set_parameters(self.handle, _dp(pa) if pa is not None else None, _dp(qa) if qa is not None else None)
        if rc:
            raise RuntimeError(f"homotopy_set_parameters failed ({rc}): {_last_error(self.api)}")

    # ---- operator API hooks (test hooks of the C ABI)
    def evaluate(self, x, t):
        xa, ta = _cflat(x, self.n), _cflat([t])
        u = np.zeros(self.m, np.complex128)
        rc = self.api._evaluate(self.handle, _dp(xa), _dp(ta), _dp(u.view(np.float64)))
        if rc: raise RuntimeError(f"evaluate failed ({rc})")
        return u

    def evaluate_dd(self, x_hi, x_lo, t):
        ha, la, ta = _cflat(x_hi, self.n), _cflat(x_lo, self.n), _cflat([t])
        u = np.zeros(self.m, np.complex128)
        rc = self.api._evaluate_dd(self.handle, _dp(ha), _dp(la), _dp(ta), _dp(u.view(np.float64)))
        if rc: raise RuntimeError(f"evaluate_dd failed ({rc})")
        return u

    def evaluate_and_jacobian(self, x, t):
        xa, ta = _cflat(x, self.n), _cflat([t])
        u = np.zeros(self.m, np.complex128)
        U = np.zeros(self.m * self.n, np.complex128)
        rc = self.api._evaluate_and_jacobian(self.handle, _dp(xa), _dp(ta), _dp(u.view(np.float64)), _dp(U.view(np.float64)))
        if rc: raise RuntimeError(f"evaluate_and_jacobian failed ({rc})")
        return u, U.reshape(self.n, self.m).T.copy()

    def evaluate_batch(self, X, t, jacobian: bool = False):
        """H(x_k, t) (and the Jacobians) at all rows of X at once (hc_evaluate_batch); falls back to single-point calls
        where the library has no batched entry (oracle)."""
        X = np.ascontiguousarray(np.asarray(X, dtype=np.complex128).reshape(-1, self.n))
        N = X.shape[0]
        if getattr(self.api, "_evaluate_batch", None) is None:
            if jacobian:
                r = [self.evaluate_and_jacobian(x, t) for x in X]
                return np.array([a for a, _ in r]).reshape(N, self.m), np.array([b for _, b in r]).reshape(N, self.m, self.n)
            return np.array([self.evaluate(x, t) for x in X]).reshape(N, self.m)
        ta = _cflat([t])
        u = np.zeros((N, self.m), np.complex128)
        U = np.zeros((N, self.n, self.m), np.complex128) if jacobian else None
        rc = self.api._evaluate_batch(self.handle, N, _dp(X.view(np.float64)), _dp(ta), _dp(u.view(np.float64)),
                                      _dp(U.view(np.float64)) if U is not None else None)
        if rc: raise RuntimeError(f"evaluate_batch failed ({rc}): {_last_error(self.api)}")
        return (u, U.transpose(0, 2, 1).copy()) if jacobian else u

    def taylor(self, K, tx, t):
        """tx: (K, n) rows x^0..x^{K-1}; returns the K-th Taylor coefficient of H(x(l), t+l)."""
        xa, ta = _cflat(np.asarray(tx).reshape(-1), K * self.n), _cflat([t])
        u = np.zeros(self.m, np.complex128)
        rc = self.api._taylor(self.handle, K, _dp(xa), _dp(ta), _dp(u.view(np.float64)))
        if rc: raise RuntimeError(f"taylor failed ({rc})")
        return u

    def set_toric_weights(self, w):
        wa = np.ascontiguousarray(w, dtype=np.float64)
        rc = self.api._toric_set_weights(self.handle, _dp(wa))
        if rc: raise RuntimeError(f"toric_set_weights failed ({rc})")

    # ---- batched tracking
    def track_batch(self, starts, options: Options | None = None, mode: int = 0, t1=1.0, t0=0.0,
                    path_p=None, path_q=None, omega_mu=None, nthreads: int = 1, out: BatchResults | None = None) -> BatchResults:
        """`out`: result arrays of an earlier call (or BatchResults.allocate) to be overwritten -- what a host
        that solves repeatedly does with its preallocated, page-locked result vectors."""
        starts = np.ascontiguousarray(np.asarray(starts, dtype=np.complex128).reshape(-1, self.n))
        N = starts.shape[0]
        opts = options if options is not None else self.api.default_options()
        res = _out_or_new(out, self.n, N)
        d = res.desc()
        t1a, t0a = _cflat([t1]), _cflat([t0])
        pp = _cflat(np.asarray(path_p).reshape(-1), N * self.P) if path_p is not None else None
        pq = _cflat(np.asarray(path_q).reshape(-1), N * self.P) if path_q is not None else None
        om = np.ascontiguousarray(omega_mu, dtype=np.float64).reshape(-1) if omega_mu is not None else None
        rc = self.api._track_batch(self.handle, C.byref(opts), mode, N, _dp(starts.view(np.float64)), _dp(t1a), _dp(t0a),
                                   _dp(pp) if pp is not None else None, _dp(pq) if pq is not None else None,
                                   _dp(om) if om is not None else None, C.byref(d), nthreads)
        if rc:
            raise RuntimeError(f"track_batch failed ({rc})")
        return res


def _last_error(api: CApi) -> str:
    try:
        fn = getattr(api.lib, api.prefix + "last_error")
        fn.restype = C.c_char_p
        return fn().decode()
    except AttributeError:
        return ""


def track_total_degree(H: "HomotopyHandle", degrees, first: int = 0, count: int | None = None,
                       options: Options | None = None, out: "BatchResults | None" = None) -> "BatchResults":
    """Paths first .. first + count - 1 of the total-degree start system, start solutions made on the device
    (hc_track_total_degree; order of TotalDegreeStartSolutionsIterator, reference src/total_degree.jl:235-262)."""
    api = H.api
    if api._track_total_degree is None:
        raise RuntimeError("this library has no device-side start generation")
    deg = np.ascontiguousarray(degrees, dtype=np.int32)
    if deg.shape != (H.n,):
        raise ValueError(f"need {H.n} degrees")
    total = int(np.prod(deg.astype(object)))
    N = total - first if count is None else int(count)
    opts = options if options is not None else api.default_options()
    res = _out_or_new(out, H.n, N)
    d = res.desc()
    rc = api._track_total_degree(H.handle, C.byref(opts), _ip(deg), int(first), N, C.byref(d))
    if rc:
        raise RuntimeError(f"track_total_degree failed ({rc}): {_last_error(api)}")
    return res


def track_sweep(H: "HomotopyHandle", starts, target_parameters, options: Options | None = None,
                out: "BatchResults | None" = None) -> "BatchResults":
    """The same S start solutions tracked to each of M target parameter vectors (reference many_solve,
    src/solve.jl:815-881).  Result row j * S + s = start s to parameter point j."""
    api = H.api
    if api._track_sweep is None:
        raise RuntimeError("this library has no sweep entry point")
    starts = np.ascontiguousarray(np.asarray(starts, dtype=np.complex128).reshape(-1, H.n))
    q = np.ascontiguousarray(np.asarray(target_parameters, dtype=np.complex128).reshape(-1, H.P))
    S, M = starts.shape[0], q.shape[0]
    opts = options if options is not None else api.default_options()
    res = _out_or_new(out, H.n, S * M)
    d = res.desc()
    rc = api._track_sweep(H.handle, C.byref(opts), S, _dp(starts.view(np.float64)), M, _dp(q.view(np.float64)), C.byref(d))
    if rc:
        raise RuntimeError(f"track_sweep failed ({rc}): {_last_error(api)}")
    return res


def _out_or_new(out: "BatchResults | None", n: int, N: int) -> BatchResults:
    if out is None:
        return BatchResults.allocate(n, N)
    if out.n != n or out.N != N:
        raise ValueError(f"out holds {out.N} paths of dimension {out.n}, the batch has {N} of dimension {n}")
    return out


def polyhedral_track_batch(api: CApi, Htoric: HomotopyHandle, Hcoeff: HomotopyHandle, starts, cell_index,
                           cell_weights, options: Options | None = None, nthreads: int = 1,
                           out: BatchResults | None = None) -> BatchResults:
    n, P = Htoric.n, Htoric.P
    starts = np.ascontiguousarray(np.asarray(starts, dtype=np.complex128).reshape(-1, n))
    N = starts.shape[0]
    ci = np.ascontiguousarray(cell_index, dtype=np.int32)
    cw = np.ascontiguousarray(cell_weights, dtype=np.float64).reshape(-1, P)
    opts = options if options is not None else api.default_options()
    res = _out_or_new(out, n, N)
    d = res.desc()
    rc = api._polyhedral_track_batch(Htoric.handle, Hcoeff.handle, C.byref(opts), N, _dp(starts.view(np.float64)),
                                     _ip(ci), _dp(cw), cw.shape[0], C.byref(d), nthreads)
    if rc:
        raise RuntimeError(f"polyhedral_track_batch failed ({rc})")
    return res


def polyhedral_track_cells(api: CApi, Htoric: HomotopyHandle, Hcoeff: HomotopyHandle, cells: dict, cell_weights,
                           first: int = 0, count: int | None = None, options: Options | None = None,
                           out: BatchResults | None = None) -> BatchResults:
    """Polyhedral batch whose start solutions are made on the device from per-cell data (hc_polyhedral_track_cells;
    `cells` = PolyhedralStart.binomial_data(): volume, H, mu, r).  Path k = start solution (first + k) mod mixed volume."""
    if api._polyhedral_track_cells is None:
        raise RuntimeError("device-side start generation is an entry point of libhc_b200")
    n, P = Htoric.n, Htoric.P
    vol = np.ascontiguousarray(cells["volume"], dtype=np.int64)
    Hm = np.ascontiguousarray(cells["H"], dtype=np.int64).reshape(len(vol), n, n)
    mu = np.ascontiguousarray(cells["mu"], dtype=np.float64).reshape(len(vol), n)
    r = np.ascontiguousarray(cells["r"], dtype=np.float64).reshape(len(vol), n)
    cw = np.ascontiguousarray(cell_weights, dtype=np.float64).reshape(len(vol), P)
    N = int(vol.sum()) - int(first) if count is None else int(count)
    opts = options if options is not None else api.default_options()
    res = _out_or_new(out, n, N)
    d = res.desc()
    rc = api._polyhedral_track_cells(Htoric.handle, Hcoeff.handle, C.byref(opts), int(first), N, len(vol),
                                     vol.ctypes.data_as(c_int64_p), Hm.ctypes.data_as(c_int64_p), _dp(mu), _dp(r), _dp(cw),
                                     C.byref(d))
    if rc:
        raise RuntimeError(f"polyhedral_track_cells failed ({rc}): {_last_error(api)}")
    return res


def track_sweep_counts(H: "HomotopyHandle", starts, target_params, options: Options | None = None, real_tol: float = 1e-6) -> np.ndarray:
    """many_solve with the results reduced on the device (hc_track_sweep_counts): (M, 5) int32 array, per parameter point
    [nonsingular, singular, real, at infinity, failed] over its start solutions."""
    api = H.api
    if getattr(api, "_track_sweep_counts", None) is None:
        raise RuntimeError("device-side result reduction is an entry point of libhc_b200")
    starts = np.ascontiguousarray(np.asarray(starts, dtype=np.complex128).reshape(-1, H.n))
    q = np.ascontiguousarray(np.asarray(target_params, dtype=np.complex128).reshape(-1, H.P))
    opts = options if options is not None else api.default_options()
    counts = np.zeros((q.shape[0], 5), dtype=np.int32)
    rc = api._track_sweep_counts(H.handle, C.byref(opts), starts.shape[0], _dp(starts.view(np.float64)), q.shape[0],
                                 _dp(q.view(np.float64)), float(real_tol), _ip(counts))
    if rc:
        raise RuntimeError(f"track_sweep_counts failed ({rc}): {_last_error(api)}")
    return counts
