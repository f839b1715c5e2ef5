"""GPU box tool: bio-chemical network 3, the trial where the GPU's template solve delivered 7 nonsingular solutions."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import hcb200
from hcb200 import capi, lib, systems
import pyoracle
from helpers import straight_line
orc = pyoracle.load(); gpu = lib.load(0)
F = systems.biochem2(); pv = systems.BIOCHEM3_PVALS.astype(complex)
rng = np.random.default_rng(203)
np.set_printoptions(precision=6, linewidth=200)
for trial in range(12):
    gamma = np.exp(2j*np.pi*rng.random())
    pt = (rng.normal(size=8) + 1j*rng.normal(size=8))/np.sqrt(2)
    g2 = np.exp(2j*np.pi*rng.random())
    res = {}
    for name, api in (("orc", orc), ("gpu", gpu)):
        td, Ht = straight_line(api, F, g2, pt)
        res[name] = Ht.track_batch(td.start_solutions())
    a, b = res["orc"], res["gpu"]
    nsa = (a.return_code == 1) & (a.singular == 0); nsb = (b.return_code == 1) & (b.singular == 0)
    if nsa.sum() != nsb.sum() or (a.return_code != b.return_code).any():
        print("trial", trial, "nonsingular orc", int(nsa.sum()), "gpu", int(nsb.sum()), "codes orc", np.bincount(a.return_code).tolist(), "gpu", np.bincount(b.return_code).tolist())
        for k in np.flatnonzero((a.return_code != b.return_code) | (nsa != nsb)):
            for nm, o in (("orc", a), ("gpu", b)):
                print(f"   {nm} path {k}: code {o.return_code[k]} sing {o.singular[k]} cond {o.condition_jacobian[k]:.3e} acc {o.accuracy[k]:.2e} res {o.residual[k]:.2e} t {o.t[k]:.3e} wind {o.winding_number[k]} steps {o.accepted_steps[k]}+{o.rejected_steps[k]} ext {o.extended_precision_used[k]} |x| {np.abs(o.solution[k])}")
        print("   gpu nonsingular solutions:"); print(b.solution[nsb])
        print("   orc nonsingular solutions:"); print(a.solution[nsa])
