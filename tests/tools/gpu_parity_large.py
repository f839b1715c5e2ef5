"""Class-count parity of the CUDA tracker against the CPU oracle on a large slice of a workload
(run on the GPU box): python tests/tools/gpu_parity_large.py tritangents 16384"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
sys.argv, argv = sys.argv[:1], sys.argv[1:]
import bench
import hcb200, pyoracle
from hcb200 import lib, sharding
name, count = argv[0], int(argv[1])
gpu, orc = lib.load(0), pyoracle.load(fast=True)
w = bench.make_workload(name, 1, gpu).subset(count)
out = {}
for tag, api in (("gpu", gpu), ("oracle", orc)):
    h = w.build(api)
    t0 = time.perf_counter()
    out[tag] = w.track(api, h, nthreads=os.cpu_count())
    print(tag, {k: v for k, v in sharding.class_counts(out[tag]).items() if v}, f"{time.perf_counter() - t0:.1f}s", flush=True)
g, o = out["gpu"], out["oracle"]
same = g.return_code == o.return_code
print("paths", w.N, "return codes differ on", int((~same).sum()))
ok = (g.return_code == 1) & (o.return_code == 1)
ns = ok & (o.singular == 0) & (g.singular == 0)
err = np.abs(g.solution[ns] - o.solution[ns]).max(axis=1) / np.maximum(1.0, np.abs(o.solution[ns]).max(axis=1))
print("nonsingular in both:", int(ns.sum()), "max rel endpoint deviation", float(err.max()) if ns.any() else None)
print("singular flag differs on", int((ok & (g.singular != o.singular)).sum()), "of", int(ok.sum()), "successful paths")
d = np.flatnonzero(~same)[:10]
for i in d:
    print("  path", i, "gpu", g.return_code[i], "oracle", o.return_code[i], "steps", g.accepted_steps[i], o.accepted_steps[i], "t", g.t[i], o.t[i])
