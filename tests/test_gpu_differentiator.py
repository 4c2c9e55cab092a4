"""NumPy-in / NumPy-out differentiator entry points on the GPU against the
oracle (itself pinned by the reference's 67 differentiator tests), in all four
coordinate systems with Neumann constraints."""
import numpy as np
import pytest

import pararealml_b200 as ns
from oracle import differentiator as od
from pararealml_b200.operators.fdm import ThreePointCentralDifferenceMethod

pytestmark = pytest.mark.gpu

MESHES = {
    "cartesian_2d": lambda: ns.Mesh([(0.0, 1.0), (-1.0, 1.0)], [0.1, 0.25]),
    "cartesian_3d": lambda: ns.Mesh([(0.0, 1.0), (0.0, 2.0), (1.0, 2.0)], [0.2, 0.25, 0.125]),
    "polar": lambda: ns.Mesh([(1.0, 3.0), (0.0, np.pi)], [0.25, np.pi / 12], ns.CoordinateSystem.POLAR),
    "cylindrical": lambda: ns.Mesh([(1.0, 2.0), (0.0, np.pi), (0.0, 1.0)], [0.125, np.pi / 8, 0.2], ns.CoordinateSystem.CYLINDRICAL),
    "spherical": lambda: ns.Mesh([(1.0, 2.0), (0.0, np.pi), (0.3, 2.5)], [0.125, np.pi / 8, 0.2], ns.CoordinateSystem.SPHERICAL),
}


def random_dbc(rng, mesh, k):
    """Random, partially masked Neumann constraints on every face."""
    dbc = np.empty((mesh.dimensions, k), dtype=object)
    for axis in range(mesh.dimensions):
        face = tuple(1 if a == axis else n for a, n in enumerate(mesh.vertices_shape)) + (1,)
        for i in range(k):
            pair = []
            for side in range(2):
                choice = rng.integers(0, 3)
                if choice == 0:
                    pair.append(None)
                    continue
                mask = np.ones(face, dtype=bool) if choice == 1 else rng.random(face) > 0.4
                pair.append(ns.Constraint(rng.normal(size=int(mask.sum())), mask))
            dbc[axis, i] = tuple(pair)
    return dbc


@pytest.mark.parametrize("mesh_name", list(MESHES))
def test_leaf_operators_match_oracle(mesh_name):
    rng = np.random.default_rng(11)
    mesh = MESHES[mesh_name]()
    d = mesh.dimensions
    diff = ThreePointCentralDifferenceMethod()
    y = rng.normal(size=mesh.vertices_shape + (d,))
    dbc = random_dbc(rng, mesh, d)

    def close(a, b):
        scale = max(np.max(np.abs(b)), 1e-300)
        assert np.max(np.abs(a - b)) / scale <= 1e-13

    for axis in range(d):
        close(diff.gradient(y, mesh, axis, dbc), od.gradient(y, mesh, axis, dbc))
        for axis2 in range(d):
            close(diff.hessian(y, mesh, axis, axis2, dbc), od.hessian(y, mesh, axis, axis2, dbc))
    close(diff.laplacian(y, mesh, dbc), od.laplacian(y, mesh, dbc))
    close(diff.laplacian(y, mesh), od.laplacian(y, mesh))
    close(diff.divergence(y, mesh, dbc), od.divergence(y, mesh, dbc))
    for ind in range(1 if d == 2 else 3):
        close(diff.curl(y, mesh, ind, dbc), od.curl(y, mesh, ind, dbc))
    for ind in range(d):
        close(diff.vector_laplacian(y, mesh, ind, dbc), od.vector_laplacian(y, mesh, ind, dbc))


def test_argument_validation_matches_reference_errors():
    mesh = MESHES["cartesian_2d"]()
    diff = ThreePointCentralDifferenceMethod()
    y = np.zeros(mesh.vertices_shape + (2,))
    with pytest.raises(ValueError):
        diff.gradient(np.zeros((3, 3, 2)), mesh, 0)
    with pytest.raises(ValueError):
        diff.gradient(y, mesh, 2)
    with pytest.raises(ValueError):
        diff.hessian(y, mesh, 0, 2)
    with pytest.raises(ValueError):
        diff.divergence(np.zeros(mesh.vertices_shape + (3,)), mesh)
    with pytest.raises(ValueError):
        diff.curl(y, mesh, 1)
    with pytest.raises(ValueError):
        diff.laplacian(y, mesh, np.empty((1, 2), dtype=object))
    tiny = ns.Mesh([(0.0, 1.0), (0.0, 1.0)], [1.0, 0.25])
    with pytest.raises(ValueError):
        diff.gradient(np.zeros(tiny.vertices_shape + (1,)), tiny, 0)


@pytest.mark.parametrize("mesh_name", ["cartesian_2d", "polar", "cylindrical", "spherical"])
def test_anti_laplacian_matches_oracle_and_inverts_laplacian(mesh_name):
    rng = np.random.default_rng(5)
    mesh = MESHES[mesh_name]()
    eq = ns.DiffusionEquation(mesh.dimensions)
    bcs = [
        (
            ns.DirichletBoundaryCondition(lambda x, t: np.full((len(x), 1), 0.5), is_static=True),
            ns.DirichletBoundaryCondition(lambda x, t: 0.1 * x[:, :1], is_static=True),
        )
    ] * mesh.dimensions
    cp = ns.ConstrainedProblem(eq, mesh, bcs)
    y_c = cp.static_y_vertex_constraints
    rhs = rng.normal(size=mesh.vertices_shape + (1,))
    y_init = rng.random(rhs.shape)
    tol = 1e-9
    diff = ThreePointCentralDifferenceMethod(tol)
    got = diff.anti_laplacian(rhs, mesh, y_c, y_init=np.copy(y_init))
    want, sweeps = od.anti_laplacian(
        rhs, mesh, y_c, None, np.copy(y_init), tol=tol, return_sweeps=True
    )
    assert diff.last_sweeps == sweeps
    assert np.max(np.abs(got - want)) <= 1e-11 * max(1.0, np.max(np.abs(want)))
    interior = tuple([slice(1, -1)] * mesh.dimensions)
    lap = od.laplacian(got, mesh)
    assert np.max(np.abs((lap - rhs)[interior])) < 1e-5
