"""GPU box tool: bio-chemical network 3, trial 4, path 49 (jumps on the GPU?) under the different engines / builds."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "host_sim")]
import numpy as np
import hcb200
from hcb200 import capi, lib, systems
import pyoracle, pysim
from helpers import straight_line
F = systems.biochem2()
rng = np.random.default_rng(203)
for trial in range(5):
    gamma = np.exp(2j*np.pi*rng.random())
    pt = (rng.normal(size=8) + 1j*rng.normal(size=8))/np.sqrt(2)
    g2 = np.exp(2j*np.pi*rng.random())
def run(api, label):
    td, Ht = straight_line(api, F, g2, pt)
    S = td.start_solutions()
    r = Ht.track_batch(S)
    k = 49
    print(f"{label:28s} path 49: code {r.return_code[k]} steps {r.accepted_steps[k]}+{r.rejected_steps[k]} t {r.t[k]:.3e} |x| {np.abs(r.solution[k])}  all codes {np.bincount(r.return_code).tolist()}", flush=True)
    r1 = Ht.track_batch(S[49:50])
    print(f"{'':28s} alone  : code {r1.return_code[0]} steps {r1.accepted_steps[0]}+{r1.rejected_steps[0]}", flush=True)
run(pyoracle.load(), "oracle")
run(pysim.load(), "host build of device code")
gpu = lib.load(0)
os.environ["HC_B200_JIT"] = "0"; run(gpu, "gpu interpreter")
os.environ["HC_B200_JIT"] = "1"; run(gpu, "gpu specialised")
os.environ["HC_B200_JIT_FLAGS"] = "--fmad=false"; run(gpu, "gpu specialised --fmad=false")
