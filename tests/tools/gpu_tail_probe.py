"""GPU box tool: how long do the slow paths of a heavy-tailed workload take on each engine when they run alone?
Tracks the first <limit> paths on the specialised kernel, picks the paths with more than <cap> steps (or extended
precision), and tracks only those on every engine.  Usage: python tests/tools/gpu_tail_probe.py <workload> <limit> <cap>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT]
import numpy as np  # noqa: E402

import hcb200  # noqa: E402,F401
from hcb200 import lib, workloads  # noqa: E402


def main():
    name, limit, cap = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    api = lib.load(0)
    w = {"tritangents": workloads.tritangents_total_degree, "cyclooctane_td": workloads.cyclooctane_total_degree}[name](limit)
    h = w.build(api)
    os.environ["HC_B200_JIT"] = "1" if w.n <= 14 else "0"
    r = w.track(api, h)
    t_all = lib.timing().kernel_ms
    steps = r.accepted_steps + r.rejected_steps
    slow = np.flatnonzero((steps > cap) | (r.extended_precision_used > 0))
    print(f"{name}: {w.N} paths in {t_all:.0f} ms; {len(slow)} paths beyond {cap} steps or in extended precision "
          f"(max {steps.max()} steps; quantiles 50/90/99/99.9 % = {np.percentile(steps, [50, 90, 99, 99.9]).astype(int).tolist()})", flush=True)
    fast = np.flatnonzero(~((steps > cap) | (r.extended_precision_used > 0)))
    for label, env in (("specialised", {"HC_B200_JIT": "1", "HC_B200_JIT_MIN_PATHS": "1"}), ("interpreter tpp", {"HC_B200_JIT": "0", "HC_B200_ENGINE": "tpp"}),
                       ("group 8", {"HC_B200_JIT": "0", "HC_B200_ENGINE": "group", "HC_B200_GROUP": "8"}),
                       ("group 32", {"HC_B200_JIT": "0", "HC_B200_ENGINE": "group", "HC_B200_GROUP": "32"})):
        if w.n > 14 and not label.startswith("group"):
            continue
        for k in ("HC_B200_JIT", "HC_B200_ENGINE", "HC_B200_GROUP", "HC_B200_JIT_MIN_PATHS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        for what, idx in (("slow", slow), ("fast", fast)):
            if what == "fast" and label != "specialised":
                continue
            rr = h["H"].track_batch(w.starts[idx])
            tm = lib.timing()
            same = int((rr.return_code == r.return_code[idx]).sum())
            print(f"  {label:16s} {what}: {len(idx)} paths, kernel {tm.kernel_ms:8.1f} ms, grid {tm.grid} x {tm.block}, same codes {same}", flush=True)


if __name__ == "__main__":
    main()
