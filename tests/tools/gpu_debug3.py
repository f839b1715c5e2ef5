"""Debug: 2x2 endgame golden case and katsura(3) on the selected engine, printed next to the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import hcb200, pyoracle
from hcb200 import capi, lib, systems
from helpers import straight_line, system_2x2
np.set_printoptions(linewidth=200, precision=6)
gpu, orc = lib.load(0), pyoracle.load()
for name, F in (("2x2", system_2x2()), ("katsura3", systems.katsura(3))):
    out = []
    for api in (orc, gpu):
        td, H = straight_line(api, F, 0.4 + 1.3j)
        out.append(H.track_batch(td.start_solutions()))
    ro, rg = out
    print(name, "engine", os.environ.get("HC_B200_ENGINE"), "lib", os.environ.get("HC_B200_LIB"))
    print(" codes  oracle", ro.return_code, "gpu", rg.return_code)
    print(" steps  oracle", ro.accepted_steps, "gpu", rg.accepted_steps)
    print(" counters gpu", rg.counters[:4].tolist())
    print(" sol oracle", ro.solution[:2].round(6).tolist())
    print(" sol gpu   ", rg.solution[:2].round(6).tolist())
