"""The benchmark workloads of BASELINE.json, built for any C-ABI backend (GPU library or, in the
tests / CPU baseline, the oracle)."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

import os

from . import capi, flops, polyhedral, start_systems, systems

# mixed cells of the polyhedral configs (deterministic in the lifting seeds; enumerating them takes a minute for cyclic-7
# and much longer for cyclooctane with the pure-Python enumerator of polyhedral.py): shipped with the package
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


@dataclass
class Workload:
    name: str
    description: str
    n: int
    starts: np.ndarray                    # (N, n) complex
    mode: int = 0                         # 0 endgame tracker, 1 tracker, 2 polyhedral
    build: callable = None                # api -> dict(H=..., [Hcoeff=...])
    path_q: np.ndarray | None = None
    cell_index: np.ndarray | None = None
    cell_weights: np.ndarray | None = None
    costs: dict = field(default_factory=dict)
    expected: dict = field(default_factory=dict)
    # many-parameter sweeps (many_solve): the k start solutions and one target parameter row per POINT; `starts` and
    # `path_q` above are their per-path replication (path j * k + s = start s to point j)
    sweep_starts: np.ndarray | None = None
    sweep_q: np.ndarray | None = None
    # polyhedral batches: per-cell binomial data (PolyhedralStart.binomial_data()) -- libhc_b200 makes the start solutions
    # on the device from it (hc_polyhedral_track_cells); path k = start solution cells_first + k (mod mixed volume)
    cells: dict | None = None
    cells_first: int = 0

    @property
    def N(self):
        return int(self.starts.shape[0])

    def subset(self, count: int) -> "Workload":
        count = min(count, self.N)
        w = Workload(self.name, self.description, self.n, self.starts[:count], self.mode, self.build,
                     None if self.path_q is None else self.path_q[:count],
                     None if self.cell_index is None else self.cell_index[:count], self.cell_weights, self.costs, {})
        w.cells, w.cells_first = self.cells, self.cells_first
        return w

    def slice(self, lo: int, hi: int) -> "Workload":
        """Paths lo .. hi-1 (a shard of the batch; cell weights and homotopies are shared)."""
        w = Workload(self.name, self.description, self.n, self.starts[lo:hi], self.mode, self.build,
                     None if self.path_q is None else self.path_q[lo:hi],
                     None if self.cell_index is None else self.cell_index[lo:hi], self.cell_weights, self.costs, {})
        w.cells, w.cells_first = self.cells, self.cells_first + lo
        if self.sweep_starts is not None:
            k = len(self.sweep_starts)
            if lo % k == 0 and hi % k == 0:   # the shard holds whole parameter points
                w.sweep_starts, w.sweep_q = self.sweep_starts, self.sweep_q[lo // k:hi // k]
        return w

    def track(self, api, handles, options=None, nthreads=1, out=None):
        if self.sweep_starts is not None and getattr(api, "_track_sweep", None) is not None:
            return capi.track_sweep(handles["H"], self.sweep_starts, self.sweep_q, options, out=out)
        if self.mode == 2 and self.cells is not None and getattr(api, "_polyhedral_track_cells", None) is not None:
            return capi.polyhedral_track_cells(api, handles["H"], handles["Hcoeff"], self.cells, self.cell_weights,
                                               self.cells_first, self.N, options, out=out)
        if self.mode == 2:
            return capi.polyhedral_track_batch(api, handles["H"], handles["Hcoeff"], self.starts, self.cell_index,
                                               self.cell_weights, options, nthreads, out=out)
        return handles["H"].track_batch(self.starts, options=options, mode=self.mode, path_q=self.path_q, nthreads=nthreads,
                                        out=out)


def katsura8(replicas: int = 1) -> Workload:
    """BASELINE.json configs[0]: katsura(8) total-degree homotopy, 256 paths (x replicas for throughput)."""
    F = systems.katsura(8)
    td = start_systems.total_degree(F, 0.4 + 1.3j)   # gamma of reference test/endgame_tracker_test.jl:5

    def build(api):
        hF, hG = api.system(td.F), api.system(td.G)
        return {"H": api.homotopy(capi.H_STRAIGHT_LINE, hF, hG, gamma=td.gamma, G_params=td.scaling, F_params=[])}
    S = np.tile(td.start_solutions(), (replicas, 1))
    return Workload("katsura8", f"katsura(8) total-degree homotopy, 256 paths x {replicas} replicas", 9, S, 0, build,
                    costs=flops.homotopy_costs(td.F, td.G), expected={"success": 256 * replicas, "singular": 0})


def cyclic7_total_degree(replicas: int = 1) -> Workload:
    F = systems.cyclic(7)
    td = start_systems.total_degree(F, np.exp(2j * np.pi * 0.7133))

    def build(api):
        hF, hG = api.system(td.F), api.system(td.G)
        return {"H": api.homotopy(capi.H_STRAIGHT_LINE, hF, hG, gamma=td.gamma, G_params=td.scaling, F_params=[])}
    S = np.tile(td.start_solutions(), (replicas, 1))
    return Workload("cyclic7_td", f"cyclic-7 total-degree homotopy, 5040 paths x {replicas} replicas", 7, S, 0, build,
                    costs=flops.homotopy_costs(td.F, td.G), expected={"success": 924 * replicas})


def cyclic_polyhedral(n: int = 7, replicas: int = 1) -> Workload:
    """BASELINE.json configs[1]: cyclic-7 polyhedral start system, 924 mixed-volume paths (x replicas).
    Lifting and mixed cells are cached in the package's data/ directory (deterministic in the seeds; the enumeration
    takes about a minute for n = 7)."""
    ps = polyhedral.polyhedral(systems.cyclic(n), cache=os.path.join(_DATA, f"cyclic{n}_cells.json"))
    S, ci = ps.start_solutions()
    cw = ps.cell_weights()

    def build(api):
        h = api.system(ps.F)
        return {"H": api.homotopy(capi.H_TORIC, h, p=ps.start_coeffs),
                "Hcoeff": api.homotopy(capi.H_COEFFICIENT, h, p=ps.start_coeffs, q=ps.target_coeffs)}
    return Workload(f"cyclic{n}_polyhedral", f"cyclic-{n} polyhedral homotopy, {len(S)} mixed-volume paths x {replicas} replicas", n,
                    np.tile(S, (replicas, 1)), 2, build, cell_index=np.tile(ci, replicas), cell_weights=cw,
                    costs=flops.homotopy_costs(ps.F), expected={"success": len(S) * replicas}, cells=ps.binomial_data())


def tritangents_total_degree(limit: int | None = None) -> Workload:
    """BASELINE.json configs[2]: tritangents (reference test/test_systems.jl:31-54) with a fixed seeded
    real cubic (seed 3), total-degree homotopy, (2*2*3*4)^3 = 110 592 paths, 720 finite solutions."""
    F = systems.tritangents()
    c = np.random.default_rng(3).normal(size=20)
    td = start_systems.total_degree(F, 0.4 + 1.3j, c)

    def build(api):
        hF, hG = api.system(td.F), api.system(td.G)
        return {"H": api.homotopy(capi.H_STRAIGHT_LINE, hF, hG, gamma=td.gamma, G_params=td.scaling, F_params=c)}
    S = td.start_solutions()
    if limit is not None:
        S = S[:limit]
    return Workload("tritangents", f"tritangents total-degree homotopy, {len(S)} of 110592 paths", 12, S, 0, build,
                    costs=flops.homotopy_costs(td.F, td.G), expected={"success": 720 if limit is None else None})


def cyclooctane_parameters() -> np.ndarray:
    """A0 (2 x 17, column-major) and b0 ~ CN(0, 1), seed 4 (benchmarks/cyclooctane.jl:22-25)."""
    rng = np.random.default_rng(4)
    return (rng.normal(size=36) + 1j * rng.normal(size=36)) / np.sqrt(2)


def cyclooctane_total_degree(limit: int | None = None) -> Workload:
    """BASELINE.json configs[3] on the total-degree start system: 2^15 = 32 768 paths, 1408 finite
    solutions, the rest diverge (endgame at-infinity checks)."""
    F = systems.cyclooctane()
    p0 = cyclooctane_parameters()
    td = start_systems.total_degree(F, 0.4 + 1.3j, p0)

    def build(api):
        hF, hG = api.system(td.F), api.system(td.G)
        return {"H": api.homotopy(capi.H_STRAIGHT_LINE, hF, hG, gamma=td.gamma, G_params=td.scaling, F_params=p0)}
    S = td.start_solutions()
    if limit is not None:
        S = S[:limit]
    return Workload("cyclooctane_td", f"cyclooctane total-degree homotopy, {len(S)} of 32768 paths", 17, S, 0, build,
                    costs=flops.homotopy_costs(td.F, td.G), expected={"success": 1408 if limit is None else None})


def cyclooctane_polyhedral() -> Workload:
    """BASELINE.json configs[3]: cyclooctane (benchmarks/cyclooctane.jl:4-27) on the polyhedral start system."""
    ps = polyhedral.polyhedral(systems.cyclooctane(), target_parameters=cyclooctane_parameters(), seed_coeffs=14, seed_origin=15,
                               seed_lifting=16, cache=os.path.join(_DATA, "cyclooctane_cells.json"))
    S, ci = ps.start_solutions()
    cw = ps.cell_weights()

    def build(api):
        h = api.system(ps.F)
        return {"H": api.homotopy(capi.H_TORIC, h, p=ps.start_coeffs),
                "Hcoeff": api.homotopy(capi.H_COEFFICIENT, h, p=ps.start_coeffs, q=ps.target_coeffs)}
    return Workload("cyclooctane_polyhedral", f"cyclooctane polyhedral homotopy, {len(S)} mixed-volume paths", 17, S, 2, build,
                    cell_index=ci, cell_weights=cw, costs=flops.homotopy_costs(ps.F), expected={"success": 1408},
                    cells=ps.binomial_data())


def biochem_generic_start(api, seed: int = 5):
    """Generic parameters p1 ~ CN(0,1)^10 (seed 5) and the start solutions of bio-chemical network 1 at p1,
    obtained by one total-degree solve on `api` (reference: solve(F; target_parameters = p1))."""
    F = systems.biochem1()
    rng = np.random.default_rng(seed)
    p1 = rng.normal(size=10) + 1j * rng.normal(size=10)
    td = start_systems.total_degree(F, 0.4 + 1.3j, p1)
    hF, hG = api.system(td.F), api.system(td.G)
    H = api.homotopy(capi.H_STRAIGHT_LINE, hF, hG, gamma=td.gamma, G_params=td.scaling, F_params=p1)
    r = H.track_batch(td.start_solutions())
    ok = (r.return_code == 1) & (r.singular == 0)
    return p1, r.solution[ok]


def biochem_sweep(api, points: int, seed: int = 6, first: int = 0) -> Workload:
    """BASELINE.json configs[4]: parameter homotopy sweep of bio-chemical network 1
    (benchmarks/bio-chemical-rection-networks.jl:19-26): `points` target parameter vectors
    q_j = p_vals * exp(0.5 z_j), z_j ~ N(0,1)^10 (seed 6, j = first .. first + points - 1), each
    tracked from every generic start solution."""
    p1, starts = biochem_generic_start(api)
    return biochem_sweep_from_starts(starts, p1, points, seed, first)


def biochem_sweep_from_starts(starts: np.ndarray, p1: np.ndarray, points: int, seed: int = 6, first: int = 0) -> Workload:
    F = systems.biochem1()
    rng = np.random.default_rng(seed)
    z = rng.normal(size=(first + points, 10))[first:]       # the stream is indexed by parameter point: shards agree
    q = systems.BIOCHEM1_PVALS[None, :] * np.exp(0.5 * z)
    k = len(starts)
    S = np.repeat(starts[None], points, axis=0).reshape(-1, 3)
    Q = np.repeat(q[:, None, :], k, axis=1).reshape(-1, 10).astype(np.complex128)

    def build(api):
        return {"H": api.homotopy(capi.H_PARAMETER, api.system(F), p=p1, q=q[0])}
    return Workload("biochem_sweep", f"bio-chemical network 1 parameter sweep, {points} parameter points x {k} start solutions",
                    3, S, 0, build, path_q=Q, costs=flops.homotopy_costs(F),
                    sweep_starts=np.ascontiguousarray(starts, dtype=np.complex128), sweep_q=np.ascontiguousarray(q, dtype=np.complex128))


def specialised_kernel_builders():
    """(name, build(api) -> homotopy handle, hc_jit_prepare flags) of the BASELINE.json configs that run on the
    specialised (run-time compiled) kernels: lets `__graft_entry__.build()` fill the on-disk kernel cache ahead of the
    first batch -- NVRTC needs no device -- so that a bench or a solve on a fresh box loads cubins instead of compiling.
    Only the code matters for the cache key: parameter values, gamma and start solutions are run-time data."""
    def sl(F, tp=None):
        def build(api):
            td = start_systems.total_degree(F(), 0.4 + 1.3j, tp() if tp else None)
            return api.homotopy(capi.H_STRAIGHT_LINE, api.system(td.F), api.system(td.G), gamma=td.gamma, G_params=td.scaling,
                                F_params=td.target_parameters if td.target_parameters is not None else [])
        return build

    def cyclic7(api):
        ps = polyhedral.polyhedral(systems.cyclic(7), cache=os.path.join(_DATA, "cyclic7_cells.json"))
        h = api.system(ps.F)
        api.homotopy(capi.H_TORIC, h, p=ps.start_coeffs)
        return api.homotopy(capi.H_COEFFICIENT, h, p=ps.start_coeffs, q=ps.target_coeffs)

    def cyclooctane(api):
        ps = polyhedral.polyhedral(systems.cyclooctane(), target_parameters=cyclooctane_parameters(), seed_coeffs=14, seed_origin=15,
                                   seed_lifting=16, cache=os.path.join(_DATA, "cyclooctane_cells.json"))
        h = api.system(ps.F)
        api.homotopy(capi.H_TORIC, h, p=ps.start_coeffs)
        return api.homotopy(capi.H_COEFFICIENT, h, p=ps.start_coeffs, q=ps.target_coeffs)

    def biochem(api):
        z = np.zeros(10, dtype=np.complex128)
        return api.homotopy(capi.H_PARAMETER, api.system(systems.biochem1()), p=z, q=z)
    return [("cyclic7_polyhedral", cyclic7, 1), ("katsura8", sl(lambda: systems.katsura(8)), 0),
            ("tritangents", sl(systems.tritangents, lambda: np.random.default_rng(3).normal(size=20)), 0),
            ("biochem_sweep", biochem, 2),
            ("cyclooctane_td", sl(systems.cyclooctane, cyclooctane_parameters), 0), ("cyclooctane_polyhedral", cyclooctane, 1)]
