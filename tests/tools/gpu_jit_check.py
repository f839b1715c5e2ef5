"""GPU box tool: specialised (run-time compiled) kernels vs the interpreter engine vs the oracle on the small
configs, plus kernel timings of replicated batches.  Usage: python tests/tools/gpu_jit_check.py [workload ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402

import hcb200  # noqa: E402,F401
from hcb200 import lib, workloads  # noqa: E402
import pyoracle  # noqa: E402
from helpers import assert_batches_match  # noqa: E402


def run(w, api, jit):
    os.environ["HC_B200_JIT"] = jit
    h = w.build(api)
    t0 = time.perf_counter()
    r = w.track(api, h)
    dt = time.perf_counter() - t0
    tm = lib.timing()
    return r, dt, tm.kernel_ms


def main():
    names = sys.argv[1:] or ["katsura8", "cyclic7_polyhedral"]
    api = lib.load(0)
    orc = pyoracle.load()
    for name in names:
        mk = {"katsura8": workloads.katsura8, "cyclic7_polyhedral": lambda r: workloads.cyclic_polyhedral(7, r),
              "cyclic7_td": workloads.cyclic7_total_degree}[name]
        w = mk(1)
        ro = w.track(orc, w.build(orc))
        for jit in ("0", "1"):
            r, dt, kms = run(w, api, jit)
            assert_batches_match(ro, r)
            err = np.abs(r.solution - ro.solution)[ro.return_code == 1].max()
            print(f"{name} x1 jit={jit}: codes {np.bincount(r.return_code).tolist()} max|dx| vs oracle {err:.2e} "
                  f"steps {int(r.accepted_steps.sum())} (oracle {int(ro.accepted_steps.sum())}) call {dt:.2f} s kernel {kms:.1f} ms", flush=True)
        reps = int(os.environ.get("REPS", "160"))
        wr = mk(reps)
        for jit in ("0", "1"):
            best = None
            for _ in range(2):
                r, dt, kms = run(wr, api, jit)
                best = kms if best is None else min(best, kms)
            ok = int((r.return_code == 1).sum())
            print(f"{name} x{reps} ({wr.N} paths) jit={jit}: kernel {best:.1f} ms = {wr.N / best * 1e3:,.0f} paths/s, success {ok}", flush=True)


if __name__ == "__main__":
    main()
