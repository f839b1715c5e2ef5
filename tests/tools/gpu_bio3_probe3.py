"""GPU box tool: operator API of the total-degree homotopy of bio-chemical network 2/3 -- interpreter on the GPU vs oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "host_sim")]
import numpy as np
import hcb200
from hcb200 import capi, lib, systems
import pyoracle, pysim
from helpers import straight_line
os.environ["HC_B200_JIT"] = "0"
F = systems.biochem2()
rng = np.random.default_rng(203)
for trial in range(5):
    gamma = np.exp(2j*np.pi*rng.random())
    pt = (rng.normal(size=8) + 1j*rng.normal(size=8))/np.sqrt(2)
    g2 = np.exp(2j*np.pi*rng.random())
r2 = np.random.default_rng(1)
x = r2.normal(size=3) + 1j*r2.normal(size=3)
tx = np.stack([x, 0.3*x + 0.1j, 0.01*x*x])
out = {}
for name, api in (("orc", pyoracle.load()), ("sim", pysim.load()), ("gpu", lib.load(0))):
    td, H = straight_line(api, F, g2, pt)
    u, U = H.evaluate_and_jacobian(x, 0.37)
    out[name] = [u, U, H.evaluate(x, 0.37)] + [H.taylor(K, tx[:K], 0.37) for K in (1, 2, 3)]
for name in ("sim", "gpu"):
    print(name, [float(np.abs(a - b).max() / max(1.0, np.abs(a).max())) for a, b in zip(out["orc"], out[name])])
