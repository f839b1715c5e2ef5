"""Where the end-to-end time of hc_track_batch goes: per-call wall time and the library's own phase timing
(HC_B200_VERBOSE=2 prints setup / launch / d2h / release per call) for pageable vs page-locked host buffers and
fresh vs reused result arrays.  GPU box only:  python scripts/e2e_breakdown.py [workload] [replicas]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
os.environ.setdefault("HC_B200_VERBOSE", "2")
import numpy as np  # noqa: E402
import hcb200  # noqa: E402,F401
from hcb200 import capi, lib  # noqa: E402
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cyclic7_polyhedral"
replicas = int(sys.argv[2]) if len(sys.argv) > 2 else 160
api = lib.load(0)
w = bench.make_workload(name, replicas, api)
w.starts = np.ascontiguousarray(w.starts, dtype=np.complex128)
handles = w.build(api)
opts = api.default_options()
print(f"{w.description}: {w.N} paths, pool={os.environ.get('HC_B200_POOL', '1')}", flush=True)


def run(tag, out, calls=3):
    for i in range(calls):
        t0 = time.perf_counter()
        w.track(api, handles, opts, out=out)
        dt = time.perf_counter() - t0
        tm = lib.timing()
        print(f"{tag} call {i}: wall {1e3 * dt:.1f} ms  (setup+h2d {tm.h2d_ms:.1f}, kernel {tm.kernel_ms:.1f}, d2h {tm.d2h_ms:.1f}, "
              f"other {1e3 * dt - tm.h2d_ms - tm.kernel_ms - tm.d2h_ms:.1f})", flush=True)


run("pageable, fresh result arrays", None)
out = capi.BatchResults.allocate(w.n, w.N)
run("pageable, reused result arrays", out)
ins = [w.sweep_starts, w.sweep_q] if w.sweep_starts is not None else [w.starts, w.path_q, w.cell_index]
t0 = time.perf_counter()
pinned = lib.pin(*ins, *out.arrays())
print(f"hc_host_register of {sum(a.nbytes for a in pinned) / 2**20:.0f} MiB: {1e3 * (time.perf_counter() - t0):.1f} ms")
run("page-locked, reused result arrays", out)
lib.unpin(pinned)
