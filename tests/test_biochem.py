"""The parity contract of the reference's bio-chemical benchmark (benchmarks/bio-chemical-rection-networks.jl:128-140,
BASELINE.md section 2): for each of the four networks, solving the specific instance DIRECTLY and solving a generic
("template") instance first and then tracking its solutions to the specific parameters by a parameter homotopy must
find the same number of solutions -- "These numbers should all coincide".  Test size: a handful of random templates per
network instead of the benchmark's 1000; the histograms (number of solutions -> how often) must be equal and must be the
same on the oracle and on the device code."""
import numpy as np
import pytest

from helpers import straight_line
from hcb200 import capi, systems

NETWORKS = {
    1: (systems.biochem1, systems.BIOCHEM1_PVALS),
    2: (systems.biochem2, systems.BIOCHEM2_PVALS),
    3: (systems.biochem2, systems.BIOCHEM3_PVALS),
    4: (systems.biochem4, systems.BIOCHEM4_PVALS),
}


def n_solutions(r):
    """length(solutions(result)): nonsingular successes (src/result.jl:677-678, only_nonsingular = true)"""
    return int(((r.return_code == 1) & (r.singular == 0)).sum())


def histograms(api, which, templates, seed):
    F_of, pvals = NETWORKS[which]
    F = F_of()
    pv = pvals.astype(np.complex128)
    rng = np.random.default_rng(seed)
    direct, template = np.zeros(9, dtype=int), np.zeros(9, dtype=int)
    for _ in range(templates):
        gamma = np.exp(2j * np.pi * rng.random())
        td, H = straight_line(api, F, gamma, pv)                    # direct: total-degree solve at the specific parameters
        direct[n_solutions(H.track_batch(td.start_solutions()))] += 1
        pt = (rng.normal(size=len(pv)) + 1j * rng.normal(size=len(pv))) / np.sqrt(2)   # randn(ComplexF64, P)
        td, Ht = straight_line(api, F, np.exp(2j * np.pi * rng.random()), pt)
        rt = Ht.track_batch(td.start_solutions())
        starts = rt.solution[(rt.return_code == 1) & (rt.singular == 0)]
        if len(starts):
            again = api.homotopy(capi.H_PARAMETER, api.system(F), p=pt, q=pv).track_batch(starts)
            template[n_solutions(again)] += 1
        else:
            template[0] += 1
    return direct, template


@pytest.mark.parametrize("which", [1, 2, 3, 4])
def test_direct_and_template_histograms_coincide(oracle, sim, which):
    do, to = histograms(oracle, which, 4, 100 + which)
    ds, ts = histograms(sim, which, 4, 100 + which)
    assert do.sum() == 4
    if which != 4:
        assert (do == to).all(), (do, to)
    else:
        assert do[3] == 4 and to[2] == 4, (do, to)
    assert (ds == do).all() and (ts == to).all(), (ds, do, ts, to)


@pytest.mark.gpu
@pytest.mark.parametrize("which", [1, 2, 3, 4])
def test_direct_and_template_histograms_coincide_on_the_gpu(oracle, gpu, which):
    do, to = histograms(oracle, which, 12, 200 + which)
    dg, tg = histograms(gpu, which, 12, 200 + which)
    if which != 4:
        assert (do == to).all(), (do, to)
    assert (dg == do).all() and (tg == to).all(), (dg, do, tg, to)
