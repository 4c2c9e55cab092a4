"""Device-side initial conditions: the Gaussian
and marginal-Beta-product states evaluated by CUDA kernels against the host
classes (reference ``initial_condition.py:246-378``, SciPy densities)."""
import numpy as np
import pytest

import pararealml_b200 as ns
from pararealml_b200.operators.fdm import (
    RK4,
    FDMOperator,
    ThreePointCentralDifferenceMethod,
)
from pararealml_b200.operators.fdm import fdm_operator as fo

pytestmark = pytest.mark.gpu


def _planes_to_host(planes, shape, y_dim):
    return np.moveaxis(planes.cpu().numpy().reshape((y_dim,) + shape), 0, -1)


def _plan(cp):
    low = fo.lowered(cp)
    from pararealml_b200.operators.fdm import device as dv

    return low, dv.get_plan(low, **fo.plan_overrides(cp, low, None, True))


def _check_close(dev, host):
    # the exponent carries a few ulp of the Mahalanobis distance (up to ~1e3
    # in the tails; the host's matrix product may use other FMA contractions):
    # relative error of every single value <= 1e-12, of the state as a whole
    # <= 1e-15
    scale = np.max(np.abs(host))
    assert np.max(np.abs(dev - host)) <= 1e-15 * scale
    nz = np.abs(host) > 1e-300
    assert np.max(np.abs(dev[nz] - host[nz]) / np.abs(host[nz])) <= 1e-12


def test_gaussian_cartesian_3d_with_static_dirichlet_faces():
    eq = ns.BurgersEquation(3, 100.0)
    shape = (40, 44, 48)
    mesh = ns.Mesh([(0.0, 1.0)] * 3, [1.0 / (n - 1) for n in shape])
    neu = ns.NeumannBoundaryCondition(lambda x, t: np.zeros((len(x), 3)), is_static=True)
    dirichlet = ns.DirichletBoundaryCondition(
        lambda x, t: np.stack([x[:, 1], np.full(len(x), np.nan), x[:, 2] * 2.0], axis=-1),
        is_static=True,
    )
    cp = ns.ConstrainedProblem(eq, mesh, [(dirichlet, neu), (neu, neu), (neu, dirichlet)])
    cov = np.array([[0.05, 0.01, 0.0], [0.01, 0.07, 0.02], [0.0, 0.02, 0.04]])
    ic = ns.GaussianInitialCondition(
        cp,
        [(np.array([0.5, 0.4, 0.6]), cov), (np.full(3, 0.5), 0.05 * np.eye(3)),
         (np.array([0.2, 0.8, 0.5]), 0.1 * np.eye(3))],
        [0.3, -0.2, 0.1],
    )
    low, plan = _plan(cp)
    dev = _planes_to_host(ic.discrete_y_0_planes(plan), shape, 3)
    _check_close(dev, ic.discrete_y_0(True))


def test_gaussian_polar_mesh():
    eq = ns.ShallowWaterEquation(0.5)
    shape = (60, 90)
    mesh = ns.Mesh(
        [(4.0, 11.0), (0.5 * np.pi, 1.5 * np.pi)],
        [7.0 / (shape[0] - 1), np.pi / (shape[1] - 1)],
        ns.CoordinateSystem.POLAR,
    )
    bc = ns.NeumannBoundaryCondition(
        ns.vectorize_bc_function(lambda x, t: (0.0, None, None)), is_static=True
    )
    cp = ns.ConstrainedProblem(eq, mesh, [(bc, bc)] * 2)
    ic = ns.GaussianInitialCondition(
        cp,
        [(np.array([-6.0, 6.0]), 0.25 * np.eye(2)), (np.array([-5.0, 5.0]), np.eye(2)),
         (np.array([-7.0, 3.0]), np.array([[1.0, 0.3], [0.3, 2.0]]))],
        [1.0, 0.5, -2.0],
    )
    low, plan = _plan(cp)
    dev = _planes_to_host(ic.discrete_y_0_planes(plan), shape, 3)
    _check_close(dev, ic.discrete_y_0(True))


def test_marginal_beta_product_is_bit_identical():
    eq = ns.BurgersEquation(3, 100.0)
    shape = (24, 30, 36)
    mesh = ns.Mesh([(0.0, 1.0)] * 3, [1.0 / (n - 1) for n in shape])
    bc = ns.NeumannBoundaryCondition(lambda x, t: np.zeros((len(x), 3)), is_static=True)
    cp = ns.ConstrainedProblem(eq, mesh, [(bc, bc)] * 3)
    ic = ns.MarginalBetaProductInitialCondition(
        cp, [[(2.0, 5.0)] * 3, [(3.0, 3.0), (2.0, 4.0), (5.0, 2.0)], [(2.5, 2.5)] * 3],
        [1.0, -0.5, 0.25],
    )
    low, plan = _plan(cp)
    dev = _planes_to_host(ic.discrete_y_0_planes(plan), shape, 3)
    assert np.array_equal(dev, ic.discrete_y_0(True))


def test_solve_with_device_initial_condition_matches_host_path(monkeypatch):
    eq = ns.DiffusionEquation(2)
    shape = (64, 80)
    mesh = ns.Mesh([(0.0, 10.0)] * 2, [10.0 / (n - 1) for n in shape])
    d = ns.DirichletBoundaryCondition(lambda x, t: np.full((len(x), 1), 1.5), is_static=True)
    n_ = ns.NeumannBoundaryCondition(lambda x, t: np.zeros((len(x), 1)), is_static=True)
    cp = ns.ConstrainedProblem(eq, mesh, [(d, d), (n_, n_)])
    ic = ns.GaussianInitialCondition(cp, [(np.array([5.0, 5.0]), np.eye(2))], [1000.0])
    d_t = 0.2 * min(mesh.d_x) ** 2
    ivp = ns.InitialValueProblem(cp, (0.0, 5 * d_t), ic)
    op = FDMOperator(RK4(), ThreePointCentralDifferenceMethod(), d_t)
    monkeypatch.setattr(fo, "DEVICE_IC_MIN_CELLS", 0)
    assert op.prepare(ivp)[2] is None  # no host array of the initial state
    on_device = op.solve(ivp).discrete_y()
    monkeypatch.setattr(fo, "DEVICE_IC_MIN_CELLS", 1 << 40)
    assert op.prepare(ivp)[2] is not None
    on_host = op.solve(ivp).discrete_y()
    scale = np.max(np.abs(on_host))
    assert np.max(np.abs(on_device - on_host)) <= 1e-14 * scale
