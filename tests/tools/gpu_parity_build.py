"""GPU box tool: specialised kernel built with --fmad=false vs the g++ build of the same unit (tritangents slice):
which paths are not bit-identical, and what do they have in common?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "host_sim")]
import numpy as np
import hcb200
from hcb200 import lib, workloads
import pysim
os.environ["HC_B200_JIT"] = "1"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 768
w = workloads.tritangents_total_degree().subset(N)
sim = pysim.load()
rs = w.track(sim, w.build(sim))
os.environ["HC_B200_JIT_FLAGS"] = "--fmad=false"
api = lib.load(0)
rg = w.track(api, w.build(api))
same_code = rs.return_code == rg.return_code
same_steps = (rs.accepted_steps == rg.accepted_steps) & (rs.rejected_steps == rg.rejected_steps)
bit = same_code & same_steps & (rs.solution == rg.solution).all(axis=1) & (rs.t == rg.t)
print("paths", N, "codes identical", int(same_code.sum()), "steps identical", int(same_steps.sum()), "bit-identical", int(bit.sum()))
bad = np.flatnonzero(~bit)
print("not bit-identical:", len(bad))
for k in bad[:30]:
    print(f"  path {k}: code {rs.return_code[k]}/{rg.return_code[k]} steps {rs.accepted_steps[k]}+{rs.rejected_steps[k]} / {rg.accepted_steps[k]}+{rg.rejected_steps[k]} "
          f"winding {rs.winding_number[k]} ext {rs.extended_precision_used[k]} steps_eg {rs.steps_eg[k]} t {rs.t[k]:.3e}/{rg.t[k]:.3e} sing {rs.singular[k]}")
print("among bit-identical: winding>0:", int((rs.winding_number[bit] > 0).sum()), "ext used:", int(rs.extended_precision_used[bit].sum()), "codes", np.bincount(rs.return_code[bit]).tolist())
print("among different:     winding>0:", int((rs.winding_number[~bit] > 0).sum()), "ext used:", int(rs.extended_precision_used[~bit].sum()), "codes", np.bincount(rs.return_code[~bit]).tolist())
