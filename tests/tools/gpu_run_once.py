"""GPU box tool (profiling driver): tracks one replicated workload `runs` times through the C ABI and prints the
kernel time.  Usage: python tests/tools/gpu_run_once.py <workload> <replicas> [runs]   (HC_B200_* select the engine)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT]
import numpy as np  # noqa: E402

import hcb200  # noqa: E402,F401
from hcb200 import lib, workloads  # noqa: E402


def main():
    name, reps = sys.argv[1], int(sys.argv[2])
    runs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    api = lib.load(0)
    mk = {"katsura8": workloads.katsura8, "cyclic7_polyhedral": lambda r: workloads.cyclic_polyhedral(7, r),
          "cyclic7_td": workloads.cyclic7_total_degree, "tritangents": lambda r: workloads.tritangents_total_degree(r if r > 1 else None),
          "cyclooctane_td": lambda r: workloads.cyclooctane_total_degree(r if r > 1 else None),
          "cyclooctane_polyhedral": lambda r: workloads.cyclooctane_polyhedral(),
          "biochem_sweep": lambda r: workloads.biochem_sweep(api, r * 1024)}[name]
    w = mk(reps)
    h = w.build(api)
    for _ in range(runs):
        r = w.track(api, h)
        tm = lib.timing()
        print(f"{name} x{reps} ({w.N} paths): kernel {tm.kernel_ms:.1f} ms = {w.N / tm.kernel_ms * 1e3:,.0f} paths/s, codes "
              f"{np.bincount(r.return_code).tolist()}, grid {tm.grid} x {tm.block}"
              + (f", second pass {tm.handoff_paths} paths {tm.handoff_ms:.1f} ms" if tm.handoff_paths else ""), flush=True)


if __name__ == "__main__":
    main()
