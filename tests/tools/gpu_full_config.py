"""GPU box tool: a whole config on the GPU vs the oracle (all paths): class-level comparison + ResultStatistics.
Usage: python tests/tools/gpu_full_config.py <tritangents|cyclooctane_td|cyclooctane_polyhedral> [limit]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import hcb200
from hcb200 import lib, result, workloads
import pyoracle
from helpers import CLASS_OF, compare_batches
name = sys.argv[1]
limit = int(sys.argv[2]) if len(sys.argv) > 2 else None
w = {"tritangents": workloads.tritangents_total_degree, "cyclooctane_td": workloads.cyclooctane_total_degree}.get(name, lambda l=None: workloads.cyclooctane_polyhedral())(limit) if name != "cyclooctane_polyhedral" else workloads.cyclooctane_polyhedral()
if limit and name == "cyclooctane_polyhedral":
    w = w.subset(limit)
api = lib.load(0)
t0 = time.perf_counter(); rg = w.track(api, w.build(api)); tg = time.perf_counter() - t0
orc = pyoracle.load(fast=True, native=True)
t0 = time.perf_counter(); ro = w.track(orc, w.build(orc), nthreads=os.cpu_count()); to = time.perf_counter() - t0
rep = compare_batches(ro, rg)
ca, cb = CLASS_OF[ro.return_code], CLASS_OF[rg.return_code]
ns_o = (ro.return_code == 1) & (ro.singular == 0); ns_g = (rg.return_code == 1) & (rg.singular == 0)
print(f"{name}: {w.N} paths; GPU call {tg:.2f} s (kernel {lib.timing().kernel_ms:.0f} ms, engine {lib.timing().engine}), oracle {to:.2f} s on {os.cpu_count()} cores (timing build)")
print("return codes oracle", np.bincount(ro.return_code).tolist())
print("return codes gpu   ", np.bincount(rg.return_code).tolist())
print("code mismatches", rep["code_mismatch"], "class mismatches", int((ca != cb).sum()), "success-class flips", int(((ca == 0) != (cb == 0)).sum()))
print("nonsingular sets identical:", bool((ns_o == ns_g).all()), "count", int(ns_g.sum()), "max rel endpoint deviation", rep["solution_nonsingular"])
print("statistics oracle", json.dumps(rep["statistics_ref"]))
print("statistics gpu   ", json.dumps(rep["statistics_got"]))
