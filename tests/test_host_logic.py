"""Host-side logic: tape compiler, start systems."""
import numpy as np

import hcb200
from hcb200 import modelkit, start_systems, systems
from hcb200.modelkit import make_system


def test_tape_layout_follows_the_reference_format():
    F = systems.katsura(3)
    for P in (F.eval_program, F.jac_program):
        ins = P.instructions
        assert ins.dtype == np.int32 and ins.shape[1] == 6
        assert ins[-1, 4] == modelkit.OP_STOP and (ins[:-1, 4] != modelkit.OP_STOP).all()
        C = len(P.constants)
        assert P.param_offset == C and P.var_offset == C + P.n_params and P.t_index == 0
        assert (ins[:-1, 5] > P.var_offset + P.n_vars).all()          # never writes into the input block
        assert ins[:, [0, 5]].min() >= 1 and ins[:, 5].max() <= P.tape_space
    assert F.jac_program.out_dim == 4 and F.jac_program.U_assign[:, 0].max() <= 16


def test_pow_int_keeps_literal_exponent():
    F = make_system(lambda x, p: [x[0] ** 5 + x[1]], 2)
    ins = F.eval_program.instructions
    row = ins[ins[:, 4] == modelkit.OP_POW_INT][0]
    assert row[1] == 5


def test_total_degree_start_solutions_order():
    td = start_systems.total_degree(make_system(lambda x, p: [x[0] ** 2 - 1, x[1] ** 3 - 2], 2), 1j)
    S = td.start_solutions()
    assert td.n_paths() == 6 and list(td.degrees) == [2, 3]
    assert np.allclose(S[0], [1, 1]) and np.allclose(S[1], [-1, 1])          # first index fastest
    assert np.allclose(S[2], [1, np.exp(2j * np.pi / 3)])
    assert np.allclose(td.scaling, [1.0, 2.0])


def test_support_coefficients_and_degrees():
    F = systems.tritangents()
    rng = np.random.default_rng(3)
    c = rng.normal(size=20)
    td = start_systems.total_degree(F, 0.4 + 1.3j, c)
    assert list(td.degrees) == [2, 2, 3, 4] * 3 and td.n_paths() == 110592   # benchmarks/tritangents.jl
    assert systems.cyclooctane().n_eqs == 17 and systems.cyclic(7).n_eqs == 7
