"""Queue model of the persistent thread-per-path kernel on a heavy-tailed workload (CPU only, oracle data).

Every lane pulls the next path index when it is idle (the device queue); a path costs
    steps * T_STEP * (DDW if the path used extended precision else 1)
with per-path step counts from the oracle.  T_STEP is calibrated on the measured 16 384-path slice, where every path
has its own lane and the kernel time (2.45 s, profiles/r01s4_ncu_v5_tritangents16k.txt) is the slowest path (2 072
fp64 steps: an endgame that runs into max_endgame_steps); DDW is fitted so that the model reproduces the measured
full run (4.74 s, profiles/r01s5_bench_tritangents.json).  The model then says what path ordering or a faster
engine for the slow paths would buy.

usage: python tests/tools/tail_model.py [paths]"""
import heapq
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402
import hcb200  # noqa: E402,F401
import pyoracle  # noqa: E402
from hcb200 import workloads  # noqa: E402

LANES = 148 * 128
T_STEP = 0.4e-3


def queue_time(cost, lanes=LANES, order=None):
    """Makespan when `lanes` workers pull paths in the given order (default: index order)."""
    idx = np.arange(len(cost)) if order is None else order
    if len(cost) <= lanes:
        return float(cost.max())
    h = list(cost[idx[:lanes]])
    heapq.heapify(h)
    for k in idx[lanes:]:
        heapq.heappush(h, heapq.heappop(h) + cost[k])
    return float(max(h))


def main():
    limit = int(sys.argv[1]) if len(sys.argv) > 1 else None
    orc = pyoracle.load(fast=True)
    w = workloads.tritangents_total_degree(limit)
    t0 = time.perf_counter()
    r = w.track(orc, w.build(orc), nthreads=os.cpu_count())
    print(f"oracle: {w.N} paths in {time.perf_counter() - t0:.0f} s")
    steps = (r.accepted_steps + r.rejected_steps).astype(float)
    ext = r.extended_precision_used > 0
    n16 = min(16384, w.N)
    slow = int(np.argmax(steps[:n16]))
    t_step = 2.45 / steps[slow]
    print(f"slowest path of the first {n16}: {int(steps[slow])} steps (extended precision: {bool(ext[slow])}, code {int(r.return_code[slow])}, "
          f"{int(r.steps_eg[slow])} endgame steps) -> T_STEP = {1e3 * t_step:.2f} ms")
    measured = 4.74
    lo, hi = 1.0, 64.0
    for _ in range(40):   # fit the double-double weight to the measured full run
        mid = 0.5 * (lo + hi)
        if queue_time(steps * t_step * np.where(ext, mid, 1.0)) < measured: lo = mid
        else: hi = mid
    ddw = 0.5 * (lo + hi)
    cost = steps * t_step * np.where(ext, ddw, 1.0)
    total, longest = cost.sum(), cost.max()
    print(f"steps per path: mean {steps.mean():.0f}, max {int(steps.max())}; {float(ext.mean()):.3f} of the paths use extended precision; fitted DDW = {ddw:.1f}")
    print(f"work {total:.0f} lane-seconds = {total / LANES:.2f} s per lane, longest path {longest:.2f} s, share of the extended-precision paths {cost[ext].sum() / total:.2f}")
    print(f"index order (the kernel):  {queue_time(cost):.2f} s   (measured {measured} s)")
    print(f"longest path first:        {queue_time(cost, order=np.argsort(-cost)):.2f} s   (lower bound {max(longest, total / LANES):.2f} s)")
    slowp = cost > np.quantile(cost, 0.95)
    for f in (2, 4, 8):
        c2 = np.where(slowp, cost / f, cost)
        print(f"slowest 5 % of the paths {f} x faster (second engine): {queue_time(c2):.2f} s")
    for f in (2, 4):
        print(f"every step {f} x faster: {queue_time(cost / f):.2f} s")
    # two-phase run with the existing kernels (hcb200/two_phase.py): phase 1 = thread-per-path engine with max_steps
    # capped, phase 2 = paths that hit the cap, from scratch, one warp per path on the lane-group engine
    # (148 SMs x 8 warps; 0.42 ms per step measured at n = 17 in the tail, assumed 0.3 ms here at n = 12)
    t_group, slots = 0.3e-3, 148 * 8
    for cap in (150, 300, 600, 1000):
        w1 = np.where(ext, ddw, 1.0)
        p1 = queue_time(np.minimum(steps, cap) * t_step * w1)
        d = steps > cap
        p2 = queue_time(steps[d] * t_group * np.where(ext[d], ddw, 1.0), lanes=slots) if d.any() else 0.0
        print(f"two-phase, cap {cap:4d} steps: {int(d.sum()):5d} paths deferred, phase 1 {p1:.2f} s + phase 2 {p2:.2f} s = {p1 + p2:.2f} s")


if __name__ == "__main__":
    main()
