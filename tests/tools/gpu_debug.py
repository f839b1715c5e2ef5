import sys, os, time; sys.path[:0] = ["/root/repo", "/root/repo/oracle", "/root/repo/tests"]
import numpy as np
import hcb200
from hcb200 import systems, capi, start_systems, lib
G = lib.load()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 148
td = start_systems.total_degree(systems.katsura(8), 0.4+1.3j)
hF, hG = G.system(td.F), G.system(td.G)
H = G.homotopy(capi.H_STRAIGHT_LINE, hF, hG, gamma=td.gamma, G_params=td.scaling, F_params=[])
S = td.start_solutions()
big = np.tile(S, (R, 1))
for rep in range(2):
    r = H.track_batch(big)
    tm = lib.timing()
    bad = np.nonzero(r.return_code != 1)[0]
    print("rep", rep, "grid", tm.grid, tm.block, "kernel ms", tm.kernel_ms, "bad paths:", bad[:20], "mod256:", bad[:20] % 256, "codes", r.return_code[bad][:20], flush=True)
    base = r.solution[:256]
    d = np.abs(r.solution.reshape(R, 256, -1) - base[None]).max(axis=(0, 2))
    print("  max replica deviation", d.max(), "steps differ:", (r.accepted_steps.reshape(R,256) != r.accepted_steps[:256][None]).sum(), "ext used", r.extended_precision_used.sum())
