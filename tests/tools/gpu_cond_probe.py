"""GPU box tool: where do condition_jacobian values of the CUDA path and the oracle differ most (cyclic-7 total degree)?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "host_sim")]
import numpy as np
import hcb200
from hcb200 import lib, workloads
import pyoracle, pysim
w = workloads.cyclic7_total_degree(1)
ro = w.track(pyoracle.load(), w.build(pyoracle.load()), nthreads=8)
rs = w.track(pysim.load(), w.build(pysim.load()))
api = lib.load(0)
rg = w.track(api, w.build(api))
ns = (ro.return_code == 1) & (ro.singular == 0)
for name, r in (("gpu", rg), ("sim", rs)):
    ratio = np.maximum(r.condition_jacobian[ns] / ro.condition_jacobian[ns], ro.condition_jacobian[ns] / r.condition_jacobian[ns])
    k = np.flatnonzero(ns)[np.argsort(-ratio)[:5]]
    print(name, "worst ratios", np.sort(ratio)[-5:])
    for i in k:
        print("  path", i, "cond oracle %.6e %s %.6e" % (ro.condition_jacobian[i], name, r.condition_jacobian[i]), "acc", ro.accuracy[i], r.accuracy[i], "steps", ro.accepted_steps[i], r.accepted_steps[i])
