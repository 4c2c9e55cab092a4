"""Host-side problem classes: behaviour pinned by the reference's tests
(``tests/test_mesh.py``, ``test_constraint.py``, ``test_constrained_problem.py``,
``test_operator.py``) and, when the reference is present, direct equality with
its objects."""
import numpy as np
import pytest

import pararealml_b200 as ns
import refshim
from golden import cases
from pararealml_b200.constraint import from_nan_table, to_nan_table
from pararealml_b200.operators.fdm.lowering import boundary_tables, lower_problem


def test_mesh_shapes_and_coordinates():
    mesh = ns.Mesh([(-10.0, 10.0), (0.0, 50.0)], [0.1, 0.2])
    assert mesh.vertices_shape == (201, 251)
    assert mesh.cells_shape == (200, 250)
    assert np.allclose(mesh.vertex_axis_coordinates[0], np.linspace(-10, 10, 201))
    assert np.allclose(mesh.cell_center_axis_coordinates[1][:2], [0.1, 0.3])
    assert mesh.vertex_coordinate_grids[0].shape == (201, 251)
    assert mesh.all_index_coordinates(True, flatten=True).shape == (201 * 251, 2)
    assert np.isclose(mesh.volume, 1000.0)


def test_mesh_validation():
    with pytest.raises(ValueError):
        ns.Mesh([], [])
    with pytest.raises(ValueError):
        ns.Mesh([(0.0, 1.0)], [0.1, 0.1])
    with pytest.raises(ValueError):
        ns.Mesh([(1.0, 0.0)], [0.1])
    with pytest.raises(ValueError):
        ns.Mesh([(0.0, 1.0)], [-0.1])
    with pytest.raises(ValueError):
        ns.Mesh([(0.0, 1.0), (0.0, 1.0), (0.0, 1.0)], [0.1] * 3, ns.CoordinateSystem.POLAR)
    with pytest.raises(ValueError):
        ns.Mesh([(0.0, 1.0), (0.0, 7.0)], [0.1] * 2, ns.CoordinateSystem.POLAR)


def test_curvilinear_volumes():
    polar = ns.Mesh([(1.0, 2.0), (0.0, np.pi)], [0.1, 0.1], ns.CoordinateSystem.POLAR)
    assert np.isclose(polar.volume, (4.0 - 1.0) * np.pi / 2.0)
    sph = ns.Mesh(
        [(0.0, 2.0), (0.0, 2 * np.pi), (0.0, np.pi)], [0.5, 0.5, 0.5],
        ns.CoordinateSystem.SPHERICAL,
    )
    assert np.isclose(sph.volume, 4.0 / 3.0 * np.pi * 8.0)


def test_constraint_apply_and_multiply_and_add():
    mask = np.array([[True], [False], [True]])
    c = ns.Constraint(np.array([1.0, 2.0]), mask)
    a = np.zeros((3, 1))
    assert c.apply(a) is a
    assert np.array_equal(a[:, 0], [1.0, 0.0, 2.0])
    res = np.zeros((3, 1))
    c.multiply_and_add(np.full((3, 1), 10.0), -2.0, res)
    assert np.array_equal(res[:, 0], [8.0, 0.0, 6.0])
    with pytest.raises(ValueError):
        ns.Constraint(np.array([1.0]), mask)
    with pytest.raises(ValueError):
        c.apply(np.zeros((4, 1)))
    tab = to_nan_table(c, (3, 1))
    back = from_nan_table(tab)
    assert np.array_equal(back.mask, mask) and np.array_equal(back.values, c.values)


def test_dirichlet_corner_precedence_later_axis_wins():
    eq = ns.DiffusionEquation(2)
    mesh = ns.Mesh([(0.0, 1.0), (0.0, 1.0)], [0.5, 0.5])
    bcs = [
        (
            ns.DirichletBoundaryCondition(lambda x, t: np.full((len(x), 1), 1.0), is_static=True),
            ns.DirichletBoundaryCondition(lambda x, t: np.full((len(x), 1), 2.0), is_static=True),
        ),
        (
            ns.DirichletBoundaryCondition(lambda x, t: np.full((len(x), 1), 3.0), is_static=True),
            ns.NeumannBoundaryCondition(lambda x, t: np.zeros((len(x), 1)), is_static=True),
        ),
    ]
    cp = ns.ConstrainedProblem(eq, mesh, bcs)
    y = np.zeros((3, 3, 1))
    ns.apply_constraints_along_last_axis(cp.static_y_vertex_constraints, y)
    expected = np.array([[3.0, 1.0, 1.0], [3.0, 0.0, 0.0], [3.0, 2.0, 2.0]])
    assert np.array_equal(y[..., 0], expected)
    y2 = cp.apply_dirichlet_tables(np.zeros((3, 3, 1)), cp.dirichlet_face_tables())
    assert np.array_equal(y2, y)


def test_static_and_dynamic_boundary_conditions_mix():
    eq = ns.WaveEquation(1)
    mesh = ns.Mesh([(0.0, 1.0)], [0.25])
    bcs = [
        (
            ns.DirichletBoundaryCondition(lambda x, t: np.full((len(x), 2), t)),
            ns.NeumannBoundaryCondition(lambda x, t: np.full((len(x), 2), 5.0), is_static=True),
        )
    ]
    cp = ns.ConstrainedProblem(eq, mesh, bcs)
    assert not cp.are_all_boundary_conditions_static
    assert cp.are_there_boundary_conditions_on_y
    y_static, d_static = cp.static_boundary_vertex_constraints
    assert y_static[0, 0][0] is None and d_static[0, 0][1] is not None
    y_t, d_t = cp.create_boundary_constraints(True, 2.5)
    assert np.array_equal(y_t[0, 1][0].values, [2.5])
    assert d_t[0, 0][1] is d_static[0, 0][1] or np.array_equal(
        d_t[0, 0][1].values, d_static[0, 0][1].values
    )


def test_discretize_time_domain_rounding():
    t = ns.discretize_time_domain((0.0, 1.0), 0.3)
    assert np.allclose(t, [0.0, 0.3, 0.6, 0.9])
    t = ns.discretize_time_domain((1.0, 2.0), 0.25)
    assert np.array_equal(t, np.linspace(1.0, 2.0, 5))


def test_operator_rejects_non_positive_step():
    from pararealml_b200.operators.fdm import FDMOperator, RK4, ThreePointCentralDifferenceMethod

    with pytest.raises(ValueError):
        FDMOperator(RK4(), ThreePointCentralDifferenceMethod(), 0.0)
    with pytest.raises(ValueError):
        ThreePointCentralDifferenceMethod(-1.0)


def test_implicit_integrators_raise():
    from pararealml_b200.operators.fdm import BackwardEulerMethod, CrankNicolsonMethod

    with pytest.raises(NotImplementedError):
        BackwardEulerMethod()
    with pytest.raises(NotImplementedError):
        CrankNicolsonMethod()


def test_solution_container_and_diff():
    ivp = cases.lorenz(ns, 0.1)
    t = np.array([0.05, 0.1])
    y = np.arange(6.0).reshape(2, 3)
    sol = ns.Solution(ivp, t, y, d_t=0.05)
    lazy = ns.Solution(ivp, t, lambda: y + 1.0, d_t=0.05)
    assert np.array_equal(sol.discrete_y(), y)
    d = sol.diff([lazy])
    assert np.allclose(d.matching_time_points, t)
    assert np.allclose(d.differences[0], 1.0)
    with pytest.raises(ValueError):
        ns.Solution(ivp, t, np.zeros((3, 3)))


@pytest.mark.skipif(not refshim.available(), reason="reference not present")
@pytest.mark.parametrize(
    "case",
    [c for c in cases.FDM_CASES if "lorenz" not in c.name and "n_body" not in c.name
     and "population" not in c.name],
    ids=lambda c: c.name,
)
def test_boundary_tables_match_reference_constraints(case):
    """NaN-coded face tables of this package == tables derived from the
    reference's Constraint objects, static and (where present) dynamic."""
    ref = refshim.install()
    mine = case.build(ns).constrained_problem
    theirs = case.build(ref).constrained_problem
    for t in (None, 0.37):
        a = boundary_tables(mine, t)
        b = boundary_tables(theirs, t)
        for kind in range(2):
            for axis in range(len(a[kind])):
                for side in range(2):
                    ta, tb = a[kind][axis][side], b[kind][axis][side]
                    assert (ta is None) == (tb is None)
                    if ta is not None:
                        assert np.array_equal(ta, tb, equal_nan=True)
    la, lb = lower_problem(mine), lower_problem(theirs)
    assert la.neu_mask == lb.neu_mask and la.dir_mask == lb.dir_mask
    assert la.shape == lb.shape and la.kinds == lb.kinds
