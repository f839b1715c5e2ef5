import numpy as np

import hcb200
from hcb200 import capi, result, start_systems, systems
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


def _ratio(a, b, floor, q=None):
    """max (or quantile q) over entries of the factor between two positive quantities; values below `floor` (rounding-level
    numbers: an accuracy of 2e-17 vs 9e-17 says nothing) count as `floor`; NaN must match NaN, inf must match inf"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert (np.isnan(a) == np.isnan(b)).all(), "NaN pattern differs"
    ok = ~np.isnan(a)
    a, b = a[ok], b[ok]
    assert (np.isinf(a) == np.isinf(b)).all(), "inf pattern differs"
    fin = ~np.isinf(a)
    a, b = np.maximum(np.abs(a[fin]), floor), np.maximum(np.abs(b[fin]), floor)
    if not a.size:
        return 1.0
    r = np.maximum(a / b, b / a)
    return float(np.max(r)) if q is None else float(np.quantile(r, q))


def compare_batches(ref, got):
    """Deviation of every PathResult field (src/path_result.jl:76-98) of two batches, as a dict.  Numeric fields are
    compared on the paths where both runs ended with the same return code."""
    same = ref.return_code == got.return_code
    ok = same & (ref.return_code == 1)
    rep = {"paths": int(len(same)), "code_mismatch": int((~same).sum())}
    rep["singular_mismatch"] = int((ref.singular[same] != got.singular[same]).sum())
    rep["winding_mismatch"] = int((ref.winding_number[same] != got.winding_number[same]).sum())
    for f in ("extended_precision", "extended_precision_used", "has_valuation"):
        rep[f + "_mismatch"] = int((getattr(ref, f)[same] != getattr(got, f)[same]).sum())
    ns = ok & (ref.singular == 0)
    sg = ok & (ref.singular == 1) & (got.singular == 1)
    rep["solution_nonsingular"] = rel_endpoint_error(got.solution[ns], ref.solution[ns]) if ns.any() else 0.0
    rep["solution_singular"] = rel_endpoint_error(got.solution[sg], ref.solution[sg]) if sg.any() else 0.0
    rep["singular_accuracy_max"] = float(np.nanmax(ref.accuracy[sg])) if sg.any() else 0.0
    rep["t_success"] = float(np.abs(got.t[ok] - ref.t[ok]).max()) if ok.any() else 0.0
    # paths that did not succeed stop where a check fires: a step earlier or later moves t by a step size
    oth = same & ~ok & (ref.t > 0) & (got.t > 0)
    rep["t_other_log10"] = float(np.abs(np.log10(got.t[oth]) - np.log10(ref.t[oth])).max()) if oth.any() else 0.0
    # last_path_point is the point before the final step.  Where it sits depends on the step sizes, which depend on omega --
    # an estimate made from Newton updates of size 1e-16, i.e. rounding noise near a well-conditioned endpoint.  It is
    # comparable where both runs took the same steps AND arrived with the same omega (the same trajectory numerically).
    lp = ns & (ref.accepted_steps == got.accepted_steps) & (ref.rejected_steps == got.rejected_steps) & \
        (np.abs(ref.omega - got.omega) <= 1e-6 * np.abs(ref.omega))
    rep["last_point_same_steps"] = rel_endpoint_error(got.last_point[lp], ref.last_point[lp]) if lp.any() else 0.0
    rep["last_t_same_steps"] = float(np.abs(got.last_t[lp] - ref.last_t[lp]).max() / max(1e-300, np.abs(ref.last_t[lp]).max())) if lp.any() else 0.0
    rep["same_steps_paths"] = int(lp.sum())
    rep["accuracy_ratio"] = _ratio(got.accuracy[ns], ref.accuracy[ns], 1e-13)
    rep["residual_ratio"] = _ratio(got.residual[ns], ref.residual[ns], 1e-12)  # unscaled |H(x, 0)|: rounding level is eps x term size
    # The reference's Skeel row scaling is a step function (a row is scaled by 2^-e iff its exponent e exceeds
    # threshold + max row sum, linear_algebra.jl:432-459): two runs whose norm weights differ in the last digits can sit
    # on different sides of it, and the scaled condition number then differs by that power of two (seen: x 7.6 on 4 of
    # 924 cyclic-7 endpoints, GPU vs oracle, while the host build of the same code agrees to 1e-3).  Hence quantiles.
    rep["cond_ratio"] = _ratio(got.condition_jacobian[ns], ref.condition_jacobian[ns], 1.0, 0.98)
    rep["cond_ratio_median"] = _ratio(got.condition_jacobian[ns], ref.condition_jacobian[ns], 1.0, 0.5)
    rep["cond_ratio_max"] = _ratio(got.condition_jacobian[ns], ref.condition_jacobian[ns], 1.0)
    # the valuation is an estimate at the t where the path stopped: comparable where both runs stopped at the same t
    hv = same & (ref.has_valuation != 0) & (got.has_valuation != 0) & ~ns & (np.abs(ref.t - got.t) <= 1e-9 * np.abs(ref.t))
    rep["valuation_paths"] = int(hv.sum())
    if hv.any():
        a, b = got.valuation[hv], ref.valuation[hv]
        fin = np.isfinite(a) & np.isfinite(b)
        rep["valuation"] = float(np.abs(a[fin] - b[fin]).max()) if fin.any() else 0.0
    else:
        rep["valuation"] = 0.0
    tot = lambda r, f: int(getattr(r, f)[same].sum())
    for f in ("accepted_steps", "rejected_steps", "steps_eg"):
        rep[f + "_rel"] = abs(tot(ref, f) - tot(got, f)) / max(1, tot(ref, f))
        rep[f + "_abs"] = abs(tot(ref, f) - tot(got, f))
    rep["statistics_ref"] = result.statistics(ref).asdict()
    rep["statistics_got"] = result.statistics(got).asdict()
    return rep


CLASS_OF = np.array([3, 0, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 4, 2])  # EndgameTrackerCode -> success / at infinity / failed / tracking / excess


def assert_batches_match(ref, got, rtol=1e-8, codes=True, fields=True, classes=True):
    """The parity bar of BASELINE.json -- identical return codes and solution classes (ResultStatistics after
    multiplicity clustering, as the reference's Result counts them), endpoints within 1e-8 relative -- extended to
    every field of PathResult: t of finished paths exact (last_path_point and its t sit wherever the last step started,
    which rounding noise in omega moves by per cent even between two CPU builds -- `assert_last_point_on_path` checks
    those two fields through the homotopy instead), accuracy / residual / condition number of nonsingular endpoints within a factor 4 above rounding level
    (1e-13), valuations within 2e-3, precision flags exact, step counts within 2 %."""
    rep = compare_batches(ref, got)
    if codes:
        assert rep["code_mismatch"] == 0, (np.bincount(ref.return_code), np.bincount(got.return_code))
    assert rep["singular_mismatch"] == 0 and rep["winding_mismatch"] == 0, rep
    assert rep["solution_nonsingular"] < rtol, rep
    # singular endpoints are only accurate to the endgame's own estimate
    assert rep["solution_singular"] < max(1e-6, 10 * rep["singular_accuracy_max"]), rep
    if classes and rep["code_mismatch"] == 0:
        assert rep["statistics_ref"] == rep["statistics_got"], rep
    if fields:
        assert rep["has_valuation_mismatch"] == 0, rep
        assert rep["t_success"] == 0.0, rep
        lt = got.last_t[(got.return_code == 1) & (got.steps_eg > 0)]
        assert (lt > 0).all() and (lt <= 1).all() and np.isfinite(got.last_point[got.return_code == 1]).all(), rep
        assert rep["accuracy_ratio"] <= 4.0 and rep["residual_ratio"] <= 4.0, rep
        # (the median bar needs a population: on a handful of paths one endpoint on the edge of the Skeel step decides it)
        assert (rep["cond_ratio_median"] <= 1.1 or rep["paths"] < 16) and rep["cond_ratio"] <= 4.0 and rep["cond_ratio_max"] <= 64.0, rep
        assert rep["valuation"] <= 2e-3, rep
        # (tiny batches: a path may take a step or two more)
        assert rep["accepted_steps_rel"] <= 0.02 or rep["accepted_steps_abs"] <= 2 * rep["paths"], rep
        assert rep["steps_eg_rel"] <= 0.05 or rep["steps_eg_abs"] <= 2 * rep["paths"], rep
    return rep


def assert_classes_match(ref, got, rtol=1e-8, max_flips=0.02):
    """The bar for the heavy-tailed configs (tritangents, cyclooctane: > 90 % of the paths diverge, hundreds die inside
    the endgame at t < 1e-9): every path ends in the same CLASS (success / at infinity / failed; at most 0.2 % may
    swap between the last two) -- which of the terminated_* codes a dying path reports may differ at rounding level --, the nonsingular solutions are the same
    set within 1e-8, and so are all ResultStatistics counts that do not depend on clustering singular endpoints that
    are only accurate to ~1e-7 (those may differ by a cluster)."""
    rep = compare_batches(ref, got)
    ca, cb = CLASS_OF[ref.return_code], CLASS_OF[got.return_code]
    # a success is a success in both runs; between "at infinity" and "failed" a diverging path whose endgame runs
    # out of accuracy right where the at-infinity test would fire may land on either side (seen: 1 of 2048)
    # and so may, very rarely, a SINGULAR endpoint: its Cauchy endgame runs in extended precision against the
    # max_endgame_extended_steps budget (400), and finishing at step 399 or not is a last-bit matter (seen: 1 of 4096
    # tritangents paths -- success on the GPU, terminated_max_extended_steps on the oracle).  Nonsingular successes
    # never flip (asserted below through the nonsingular sets).
    flip = (ca == 0) != (cb == 0)
    assert flip.sum() <= max(1, 0.0005 * rep["paths"]), (np.bincount(ref.return_code), np.bincount(got.return_code))
    for k in np.flatnonzero(flip):
        assert (ref.singular[k] if ca[k] == 0 else got.singular[k]) == 1, k
    assert (ca != cb).sum() <= max(1, 0.002 * rep["paths"]), (np.bincount(ref.return_code), np.bincount(got.return_code))
    assert rep["code_mismatch"] <= max_flips * rep["paths"], rep
    ns_r = (ref.return_code == 1) & (ref.singular == 0)
    ns_g = (got.return_code == 1) & (got.singular == 0)
    assert (ns_r == ns_g).all(), (int(ns_r.sum()), int(ns_g.sum()))
    assert rep["solution_nonsingular"] < rtol, rep
    a, b = rep["statistics_ref"], rep["statistics_got"]
    for k in ("total", "nonsingular", "real_nonsingular", "excess_solution"):
        assert a[k] == b[k], (k, a, b)
    for k in ("at_infinity", "failed"):
        assert abs(a[k] - b[k]) <= max(1, 0.002 * rep["paths"]), (k, a, b)
    assert abs(a["singular_with_multiplicity"] - b["singular_with_multiplicity"]) <= max(2, 0.02 * a["singular_with_multiplicity"]), (a, b)
    return rep


def assert_last_point_on_path(H, res, tol=1e-9, limit=64):
    """last_path_point / last_t (src/path_result.jl:87-88) is a point ON the path: H(last_point, last_t) = 0 up to the
    corrector's accuracy.  Checks the two fields no cross-run comparison can (see assert_batches_match)."""
    idx = np.flatnonzero((res.return_code == 1) & (res.steps_eg > 0))[:limit]
    worst = 0.0
    for k in idx:
        u = H.evaluate(res.last_point[k], res.last_t[k])
        worst = max(worst, float(np.abs(u).max()) / max(1.0, float(np.abs(res.last_point[k]).max())))
    assert worst < tol, worst
    return worst
