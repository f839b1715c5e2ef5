import sys, os, time; sys.path[:0] = ["/root/repo", "/root/repo/oracle", "/root/repo/tests"]
import numpy as np
import hcb200
from hcb200 import systems, capi, start_systems, lib
G = lib.load()
td = start_systems.total_degree(systems.katsura(8), 0.4+1.3j)
hF, hG = G.system(td.F), G.system(td.G)
H = G.homotopy(capi.H_STRAIGHT_LINE, hF, hG, gamma=td.gamma, G_params=td.scaling, F_params=[])
S = td.start_solutions()
big = np.tile(S[5:6], (2048, 1))
def distinct(a):
    a = np.ascontiguousarray(a).reshape(len(a), -1)
    return len({row.tobytes() for row in a})
for ms in (1, 2, 3, 5, 10, 1000):
    o = G.default_options(max_steps=ms)
    r = H.track_batch(big, options=o, mode=1)
    print("max_steps", ms, "distinct: sol", distinct(r.solution), "omega", distinct(r.omega), "mu", distinct(r.mu), "t", distinct(r.t), "acc", distinct(r.accuracy), "tau", distinct(r.condition_jacobian), "steps", distinct(r.accepted_steps), flush=True)
    if distinct(r.solution) > 1:
        vals, cnt = np.unique(r.solution[:, 0], return_counts=True); print("  sol[0] variants", vals[:4], cnt[:4])
        vals, cnt = np.unique(r.omega, return_counts=True); print("  omega variants", vals[:4], cnt[:4])
