"""GPU box tool: interpreter engine on the GPU vs its host build, katsura(6): which fields / paths differ in the last bits?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "host_sim")]
import numpy as np
import hcb200
from hcb200 import lib, systems
import pysim
from helpers import straight_line
os.environ["HC_B200_JIT"] = sys.argv[1] if len(sys.argv) > 1 else "0"
out = []
for api in (pysim.load(), lib.load(0)):
    td, H = straight_line(api, systems.katsura(6), 0.4 + 1.3j)
    out.append(H.track_batch(td.start_solutions()))
rs, rg = out
d = np.abs(rs.solution - rg.solution).max(axis=1)
bad = np.flatnonzero(d > 0)
print("paths with different endpoint bits:", len(bad), "of", rs.N, "max abs diff", d.max())
for k in bad[:8]:
    print(f"  path {k}: diff {d[k]:.2e} steps {rs.accepted_steps[k]}/{rg.accepted_steps[k]} acc {rs.accuracy[k]:.3e}/{rg.accuracy[k]:.3e} ext_used {rs.extended_precision_used[k]}/{rg.extended_precision_used[k]} ext {rs.extended_precision[k]}/{rg.extended_precision[k]} omega {rs.omega[k]:.6e}/{rg.omega[k]:.6e} last_t {rs.last_t[k]:.17g}/{rg.last_t[k]:.17g}")
print("last_point identical:", int((rs.last_point == rg.last_point).all(axis=1).sum()), "omega identical:", int((rs.omega == rg.omega).sum()), "mu identical:", int((rs.mu == rg.mu).sum()), "cond identical", int((rs.condition_jacobian == rg.condition_jacobian).sum()), "residual identical", int((rs.residual == rg.residual).sum()))
