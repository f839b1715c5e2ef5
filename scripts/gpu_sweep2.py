"""Kernel-only sweep over engines / occupancy / batch sizes (one process, resident batches).

usage: python scripts/gpu_sweep2.py WORKLOAD "ENV1=a,ENV2=b@replicas" ...
Each config: env settings (comma separated) @ replicas.  Prints paths/s of the tracker kernel
(hc_resident_run, CUDA events) and the class counts, one line per config.
"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np  # noqa: E402

sys.argv, argv = sys.argv[:1], sys.argv[1:]
import bench  # noqa: E402
import hcb200  # noqa: E402,F401
from hcb200 import capi, lib, sharding  # noqa: E402

KEYS = ("HC_B200_ENGINE", "HC_B200_GROUP", "HC_B200_BLOCK", "HC_B200_BLOCKS_PER_SM", "HC_B200_REFILL_MIN", "HC_B200_PATHS_PER_LANE",
        "HC_B200_TAPE_PAIRS", "HC_B200_POST_FRAC", "HC_B200_CARVEOUT", "HC_B200_SEG_WINDOW", "HC_B200_SEG_WINDOW_EVAL", "HC_B200_STAGE", "HC_B200_TPP_MAX_N", "HC_B200_SMEM_TAPE")


def main():
    wl = argv[0]
    api = lib.load(0)
    raw = api.raw
    cache = {}
    for spec in argv[1:]:
        envs, _, rep = spec.partition("@")
        rep = int(rep or 1)
        for k in KEYS:
            os.environ.pop(k, None)
        for kv in filter(None, envs.split(",")):
            k, v = kv.split("=")
            os.environ["HC_B200_" + k if not k.startswith("HC_") else k] = v
        if rep not in cache:
            cache.clear()
            cache[rep] = bench.make_workload(wl, rep, api)
        w = cache[rep]
        try:
            handles = w.build(api)
            opts = api.default_options()
            dp = lambda a: a.ctypes.data_as(capi.c_double_p)
            starts = np.ascontiguousarray(w.starts)
            t1 = np.array([1.0, 0.0]); t0 = np.array([0.0, 0.0])
            pq = np.ascontiguousarray(w.path_q).view(np.float64).reshape(-1) if w.path_q is not None else None
            ci = np.ascontiguousarray(w.cell_index, dtype=np.int32) if w.cell_index is not None else None
            cw = np.ascontiguousarray(w.cell_weights, dtype=np.float64) if w.cell_weights is not None else None
            h = raw.hc_resident_create(handles["H"].handle, handles["Hcoeff"].handle if "Hcoeff" in handles else None, C.byref(opts), w.mode,
                                       w.N, dp(starts.view(np.float64)), dp(t1), dp(t0), None, dp(pq) if pq is not None else None,
                                       ci.ctypes.data_as(capi.c_int32_p) if ci is not None else None, dp(cw) if cw is not None else None,
                                       cw.shape[0] if cw is not None else 0)
            if not h:
                print(spec, "create failed:", raw.hc_last_error().decode(), flush=True)
                continue
            h = C.c_void_p(h)
            ms = C.c_double()
            t_wall = time.perf_counter()
            assert raw.hc_resident_run(h, C.byref(ms)) == 0, raw.hc_last_error()
            times = []
            for _ in range(2):
                assert raw.hc_resident_run(h, C.byref(ms)) == 0, raw.hc_last_error()
                times.append(ms.value)
            res = capi.BatchResults.allocate(w.n, w.N)
            d = res.desc()
            assert raw.hc_resident_fetch(h, C.byref(d)) == 0
            raw.hc_resident_destroy(h)
            counts = sharding.class_counts(res)
            counts = {k: v for k, v in counts.items() if v}
            print(f"{wl} {spec}: N {w.N} paths/s {w.N / (np.mean(times) * 1e-3):.0f} ms {np.mean(times):.1f} counts {counts} "
                  f"steps/path {(res.accepted_steps.sum() + res.rejected_steps.sum()) / w.N:.1f} wall {time.perf_counter() - t_wall:.1f}s", flush=True)
        except Exception as e:  # noqa: BLE001
            print(spec, "failed:", repr(e), flush=True)


if __name__ == "__main__":
    main()
