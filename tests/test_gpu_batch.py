"""Batched solves: many initial conditions of one
problem in one device run, as the reference's ``SupervisedMLOperator``
generates its training data (``supervised_ml_operator.py:130-236``).  Every
member must equal the single solve bit-for-bit."""
import numpy as np
import pytest

import pararealml_b200 as ns
from pararealml_b200.operators.fdm import (
    RK4,
    ExplicitMidpointMethod,
    FDMOperator,
    ForwardEulerMethod,
    ThreePointCentralDifferenceMethod,
)
from pararealml_b200.operators.fdm import device as dv

pytestmark = pytest.mark.gpu


def _diffusion_cp(shape=(21, 21)):
    eq = ns.DiffusionEquation(2)
    mesh = ns.Mesh([(0.0, 10.0)] * 2, [10.0 / (n - 1) for n in shape])
    d = ns.DirichletBoundaryCondition(lambda x, t: np.full((len(x), 1), 1.5), is_static=True)
    n_ = ns.NeumannBoundaryCondition(lambda x, t: np.zeros((len(x), 1)), is_static=True)
    return ns.ConstrainedProblem(eq, mesh, [(d, d), (n_, n_)])


def _members(cp, count, t_end, y_dim=1):
    rng = np.random.default_rng(42)
    ivps = []
    for _ in range(count):
        y0 = rng.uniform(0.0, 2.0, cp.y_vertices_shape)
        ivps.append(ns.InitialValueProblem(
            cp, (0.0, t_end), ns.DiscreteInitialCondition(cp, y0, True)))
    return ivps


@pytest.mark.parametrize("integrator", [RK4, ExplicitMidpointMethod, ForwardEulerMethod])
def test_small_mesh_batch_is_one_launch_and_bit_identical(integrator):
    cp = _diffusion_cp()
    ivps = _members(cp, 7, 0.05)
    op = FDMOperator(integrator(), ThreePointCentralDifferenceMethod(), 1e-3)
    single = [op.solve(v).discrete_y() for v in ivps]
    before = dv.total_launches()
    batch = op.solve_batch(ivps)
    launches = dv.total_launches() - before
    # one time-loop launch for all members (+ layout conversions)
    assert launches <= 3
    for a, b in zip(single, batch):
        assert np.array_equal(a, b.discrete_y())


def test_multi_block_batch_is_bit_identical(monkeypatch):
    monkeypatch.setenv("PML_SMALL", "0")
    cp = _diffusion_cp((45, 50))
    ivps = _members(cp, 3, 0.02)
    op = FDMOperator(RK4(), ThreePointCentralDifferenceMethod(), 1e-3)
    single = [op.solve(v).discrete_y() for v in ivps]
    for a, b in zip(single, op.solve_batch(ivps)):
        assert np.array_equal(a, b.discrete_y())


def test_ode_batch():
    eq = ns.LorenzEquation()
    cp = ns.ConstrainedProblem(eq)
    rng = np.random.default_rng(1)
    ivps = [
        ns.InitialValueProblem(
            cp, (0.0, 0.5),
            ns.ContinuousInitialCondition(
                cp, (lambda y: (lambda _: y))(rng.uniform(0.5, 1.5, 3))))
        for _ in range(5)
    ]
    op = FDMOperator(RK4(), ThreePointCentralDifferenceMethod(), 1e-3)
    single = [op.solve(v).discrete_y() for v in ivps]
    for a, b in zip(single, op.solve_batch(ivps)):
        assert np.array_equal(a, b.discrete_y())


def test_unbatchable_members_are_solved_one_by_one():
    cp_a, cp_b = _diffusion_cp(), _diffusion_cp()
    ivps = _members(cp_a, 1, 0.01) + _members(cp_b, 1, 0.01)
    op = FDMOperator(RK4(), ThreePointCentralDifferenceMethod(), 1e-3)
    out = op.solve_batch(ivps)
    assert len(out) == 2
    assert np.array_equal(out[0].discrete_y(), op.solve(ivps[0]).discrete_y())
