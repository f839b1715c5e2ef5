import sys, time; sys.path[:0] = ["/root/repo", "/root/repo/oracle", "/root/repo/tests"]
import numpy as np
import hcb200, pyoracle, ref_systems
from hcb200 import systems, capi, start_systems, lib
from hcb200.modelkit import make_system
O = pyoracle.load(); G = lib.load()
print("dfma GFLOP/s", G.raw.hc_dfma_peak(200000))
def sl(api, F, gamma, tp=None):
    td = start_systems.total_degree(F, gamma, tp)
    hF, hG = api.system(td.F), api.system(td.G)
    H = api.homotopy(capi.H_STRAIGHT_LINE, hF, hG, gamma=td.gamma, G_params=td.scaling, F_params=tp if tp is not None else [])
    return td, H
def compare(name, ro, rs):
    print(name, "codes equal:", (ro.return_code == rs.return_code).all(), np.bincount(ro.return_code), "maxdiff sol", np.abs(ro.solution-rs.solution)[ro.return_code==1].max() if (ro.return_code==1).any() else None,
          "steps", ro.accepted_steps.sum(), rs.accepted_steps.sum(), ro.rejected_steps.sum(), rs.rejected_steps.sum(), "sing", ro.singular.sum(), rs.singular.sum(), "wind", (ro.winding_number==rs.winding_number).all(), flush=True)
F2 = make_system(lambda v,p: [2.3*v[0]**2 + 1.2*v[1]**2 + 3*v[0] - 2*v[1] + 3, 2.3*v[0]**2 + 1.2*v[1]**2 + 5*v[0] + 2*v[1] - 5], 2)
res = []
for api in (O, G):
    td, H = sl(api, F2, 0.4+1.3j); res.append(H.track_batch(td.start_solutions()))
compare("2x2", *res)
F1 = make_system(lambda v,p: [(v[0]-10)**2], 1)
res = []
for api in (O, G):
    td, H = sl(api, F1, np.exp(2j*np.pi*0.37)); res.append(H.track_batch(td.start_solutions()))
compare("(x-10)^2", *res)
res = []
for api in (O, G):
    td, H = sl(api, systems.katsura(8), 0.4+1.3j)
    for rep in range(2):
        t=time.time(); r = H.track_batch(td.start_solutions(), nthreads=8); print("katsura8 256 paths wall", time.time()-t, flush=True)
    if api is G: tm = lib.timing(); print("timing", tm.h2d_ms, tm.kernel_ms, tm.d2h_ms, tm.grid, tm.block, tm.slab_bytes)
    res.append(r)
compare("katsura8", *res)
# throughput: replicate katsura8 starts R times
td, H = sl(G, systems.katsura(8), 0.4+1.3j)
S = td.start_solutions()
for R in (16, 148):
    big = np.tile(S, (R, 1))
    t=time.time(); r = H.track_batch(big); w=time.time()-t
    tm = lib.timing(); print("katsura8 x", R, "paths", len(big), "wall", w, "kernel ms", tm.kernel_ms, "paths/s (kernel)", len(big)/(tm.kernel_ms*1e-3), "grid", tm.grid, tm.block, np.bincount(r.return_code), flush=True)
t=time.time(); ro = sl(O, systems.katsura(8), 0.4+1.3j)[1].track_batch(np.tile(S,(4,1)), nthreads=8); print("oracle 8 threads paths/s", 1024/(time.time()-t))
F, g = ref_systems.steiner_higher_prec()
res = []
for api in (O, G):
    h = api.system(F); H = api.homotopy(capi.H_PARAMETER, h, p=g["p"], q=g["q"]); res.append(H.track_batch([g["s_p"]], mode=1))
compare("steiner", *res); print(np.abs(res[1].solution[0]-g["s_q"]).max(), res[1].extended_precision_used, capi.TRACKER_CODES[res[1].return_code[0]])
