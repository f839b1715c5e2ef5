import numpy as np

import hcb200
from hcb200 import capi, start_systems, systems
from hcb200.modelkit import make_system


def straight_line(api, F, gamma, target_parameters=None):
    td = start_systems.total_degree(F, gamma, target_parameters)
    hF, hG = api.system(td.F), api.system(td.G)
    H = api.homotopy(capi.H_STRAIGHT_LINE, hF, hG, gamma=td.gamma, G_params=td.scaling,
                     F_params=target_parameters if target_parameters is not None else [])
    return td, H


def system_2x2():
    """reference test/endgame_tracker_test.jl:4"""
    return make_system(lambda v, p: [2.3 * v[0] ** 2 + 1.2 * v[1] ** 2 + 3 * v[0] - 2 * v[1] + 3,
                                     2.3 * v[0] ** 2 + 1.2 * v[1] ** 2 + 5 * v[0] + 2 * v[1] - 5], 2)


def rel_endpoint_error(a, b):
    """max over paths of ||a - b||_inf / max(1, ||b||_inf)"""
    a, b = np.asarray(a), np.asarray(b)
    return float((np.abs(a - b).max(axis=-1) / np.maximum(1.0, np.abs(b).max(axis=-1))).max())


def assert_batches_match(ref, got, rtol=1e-8, codes=True):
    """The parity bar of BASELINE.json: identical return codes / classes, endpoints within 1e-8 relative."""
    if codes:
        assert (ref.return_code == got.return_code).all(), (np.bincount(ref.return_code), np.bincount(got.return_code))
    ok = ref.return_code == 1
    assert (ref.singular[ok] == got.singular[ok]).all()
    assert (ref.winding_number == got.winding_number).all()
    ns = ok & (ref.singular == 0)
    if ns.any():
        assert rel_endpoint_error(got.solution[ns], ref.solution[ns]) < rtol
    sg = ok & (ref.singular == 1)
    if sg.any():  # singular endpoints are only accurate to the endgame's own estimate
        tol = max(1e-6, 10 * float(np.nanmax(ref.accuracy[sg])))
        assert rel_endpoint_error(got.solution[sg], ref.solution[sg]) < tol
