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


def test_hermite_normal_form_and_binomial_systems():
    """x^A = b has |det A| solutions, all valid (reference test/binomial_system_test.jl:107-120 checks residuals)."""
    from hcb200 import polyhedral as ph
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 5):
        for _ in range(5):
            A = rng.integers(-4, 5, size=(n, n))
            d = abs(ph._int_det(A))
            if d == 0:
                continue
            H, U = ph.hnf(A)
            assert (np.array(A, dtype=object).dot(U) == H).all()
            assert all(H[i][j] == 0 for i in range(n) for j in range(i + 1, n))
            assert all(0 <= H[i][j] < H[i][i] for i in range(n) for j in range(i))
            b = rng.normal(size=n) + 1j * rng.normal(size=n)
            X = ph.solve_binomial_system(A, b)
            assert X.shape == (n, d)
            for k in range(d):
                for j in range(n):
                    assert abs(np.prod(X[:, k] ** A[:, j].astype(float)) - b[j]) < 1e-9 * max(1, abs(b[j]))
            assert len(np.unique(np.round(X.T, 9), axis=0)) == d


def test_mixed_cells_give_the_mixed_volume():
    """Sum of the cell volumes == mixed volume: 6 (cyclic-3), 70 (cyclic-5, reference test/polyhedral_test.jl:38-46),
    and 924 for the cached cyclic-7 subdivision whose cells are re-verified here against the definition."""
    from hcb200 import polyhedral as ph
    assert ph.polyhedral(systems.cyclic(3)).n_paths() == 6
    ps = ph.polyhedral(systems.cyclic(5))
    assert ps.n_paths() == 70
    import os
    cache = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "homotopycontinuation.jl_b200", "data", "cyclic7_cells.json")
    ps = ph.polyhedral(systems.cyclic(7), cache=cache)
    assert ps.n_paths() == 924
    for cell in ps.cells:   # every cached cell satisfies the mixed-cell inequalities for the cached lifting
        E = np.array([ps.support[i][:, a] - ps.support[i][:, b] for i, (a, b) in enumerate(cell.indices)])
        assert abs(ph._int_det(E)) == cell.volume
        for i, A in enumerate(ps.support):
            v = A.T @ cell.normal + ps.lifting[i]
            a, b = cell.indices[i]
            assert abs(v[a] - cell.beta[i]) < 1e-7 and abs(v[b] - cell.beta[i]) < 1e-7
            rest = np.delete(v, [a, b])
            assert rest.size == 0 or rest.min() > cell.beta[i] + 1e-9
    w = ps.cell_weights()
    assert w.shape == (len(ps.cells), len(ps.start_coeffs)) and (w >= 0).all()


def test_mixed_volume_of_the_only_torus_example():
    """reference test/polyhedral_test.jl:23-36: 92 paths with the origins added (only_torus = false)."""
    from hcb200 import polyhedral as ph
    from hcb200.modelkit import make_system
    F = make_system(lambda v, p: [v[0] ** 3 * v[2] ** 15 + v[0] * v[1] * v[2] + v[1] ** 3 + v[2] ** 12,
                                  v[0] ** 2 * v[2] ** 9 + v[0] * v[1] ** 2 + v[1] * v[2] ** 3,
                                  v[0] ** 2 * v[1] * v[2] ** 5 + v[0] * v[2] ** 8 + v[1] ** 2], 3)
    assert ph.polyhedral(F).n_paths() == 92
