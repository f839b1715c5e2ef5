"""Known answers of the reference's own endgame / tracker tests (test/endgame_test.jl, test/tracker_test.jl:137-219),
checked on every backend behind the C ABI: the CPU oracle (pins the oracle), the device code compiled for the host
(kernel logic without a GPU) and, marked `gpu`, libhc_b200.so on a B200."""
import numpy as np
import pytest

import ref_systems
from helpers import straight_line
from hcb200 import capi

GAMMA = 0.4 + 1.3j   # reference test/endgame_tracker_test.jl:5


@pytest.fixture(params=["oracle", "sim", pytest.param("gpu", marks=pytest.mark.gpu)])
def api(request):
    return request.getfixturevalue(request.param)


def clusters(points, tol=1e-5, rtol=None):
    """Multiplicity clustering of endpoints (host side in the reference: src/result.jl:146-212)."""
    reps, counts = [], []
    for p in points:
        for i, q in enumerate(reps):
            if np.linalg.norm(p - q) < (tol if rtol is None else rtol * np.linalg.norm(q)):
                counts[i] += 1
                break
        else:
            reps.append(p); counts.append(1)
    return counts


def test_hyperbolic_6_6(api):  # test/endgame_test.jl:7-24
    td, H = straight_line(api, ref_systems.hyperbolic_6_6(), GAMMA)
    r = H.track_batch(td.start_solutions())
    assert td.n_paths() == 12
    assert (r.return_code == 1).all() and (r.winding_number == 3).all() and r.singular.all()
    assert sorted(clusters(r.solution)) == [6, 6]                      # 2 results of multiplicity 6
    assert np.allclose(np.abs(r.solution[:, 0]), 1, atol=1e-5) and np.allclose(r.solution[:, 1], 0, atol=1e-5)


def test_singular_1(api):  # test/endgame_test.jl:26-38
    td, H = straight_line(api, ref_systems.singular_1(), GAMMA)
    r = H.track_batch(td.start_solutions())
    assert (r.return_code == 1).all()
    assert int(r.singular.sum()) == 3 and sorted(clusters(r.solution[r.singular == 1])) == [3]   # 1 singular solution, multiplicity 3
    assert int((r.singular == 0).sum()) == 1                                                     # 1 nonsingular
    assert np.allclose(r.solution[r.singular == 1], [0, -1j], atol=1e-5)


def test_wilkinson_12(api):  # test/endgame_test.jl:40-47, endgame_options = (only_nonsingular = true,)
    td, H = straight_line(api, ref_systems.wilkinson(12), GAMMA)
    r = H.track_batch(td.start_solutions(), options=api.default_options(only_nonsingular=1))
    assert (r.return_code == 1).all()
    x = r.solution[:, 0]
    assert list(np.round(np.sort(x.real)).astype(int)) == list(range(1, 13))
    assert np.abs(x.imag).max() < 1e-4


@pytest.mark.parametrize("d", [2, 6])
def test_x_minus_10_to_the_d(api, d):  # test/endgame_test.jl:49-54
    from hcb200.modelkit import make_system
    td, H = straight_line(api, make_system(lambda v, p: [(v[0] - 10) ** d], 1), GAMMA)
    r = H.track_batch(td.start_solutions())
    assert int((r.winding_number == d).sum()) == d


@pytest.mark.parametrize("d", [2, 4, 6])
def test_winding_number_family(api, d):  # test/endgame_test.jl:63-70
    td, H = straight_line(api, ref_systems.winding_number_family(d), GAMMA)
    r = H.track_batch(td.start_solutions())
    assert td.n_paths() == (d + 1) ** 2
    assert int((r.return_code == 1).sum()) == d + 1


@pytest.mark.parametrize("gamma", [GAMMA, ref_systems.MOHAB_GAMMA])
def test_mohab_693(api, gamma):  # test/endgame_test.jl:77-110
    td, H = straight_line(api, ref_systems.mohab(), gamma)
    r = H.track_batch(td.start_solutions(), nthreads=8)
    assert td.n_paths() == 900 and list(td.degrees) == [9, 10, 10]
    ok = r.return_code == 1
    assert int((ok & (r.singular == 0)).sum()) == 693
    # all endpoints distinct, i.e. no path jumping (the closest pair of true solutions is 6e-9 apart, relatively)
    assert len(clusters(r.solution[ok], rtol=1e-10)) == 693


@pytest.mark.parametrize("db", [1.3e-3, -0.7e-3])
def test_invalid_startvalue_singular_jacobian(api, db):  # test/tracker_test.jl:137-219 (issue 454)
    F, start, b0 = ref_systems.pinned_framework()
    H = api.homotopy(capi.H_PARAMETER, api.system(F), p=[b0], q=[b0 + db])
    r = H.track_batch([start], mode=1)
    assert capi.TRACKER_CODES[r.return_code[0]] == "terminated_invalid_startvalue_singular_jacobian"
