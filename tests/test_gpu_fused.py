"""The fused stage-pair kernels (TMA-fed shared-memory plane rings, temporal
blocking of two stages) against the one-launch-per-stage kernels and the
oracle, on meshes that span several tiles and several chunks of the marching
axis, with partial tiles at the upper faces, all boundary-condition kinds and
both integrators that have stage pairs."""
import numpy as np
import pytest

import oracle
import pararealml_b200 as ns
from common import per_step_rel_err
from pararealml_b200.operators.fdm import (
    RK4,
    ExplicitMidpointMethod,
    FDMOperator,
    ForwardEulerMethod,
    ThreePointCentralDifferenceMethod,
)
from pararealml_b200.operators.fdm import device as dv

pytestmark = pytest.mark.gpu


def _zeros(k):
    return lambda x, t: np.zeros((len(x), k))


def burgers_3d(shape):
    eq = ns.BurgersEquation(3, 100)
    mesh = ns.Mesh(
        [(0.0, 1.0)] * 3, [1.0 / (n - 1) for n in shape]
    )
    bc = ns.NeumannBoundaryCondition(_zeros(3), is_static=True)
    cp = ns.ConstrainedProblem(eq, mesh, [(bc, bc)] * 3)
    rng = np.random.default_rng(3)
    y0 = rng.uniform(-1.0, 1.0, mesh.vertices_shape + (3,))
    return cp, y0, 2e-5


def convection_diffusion_3d_mixed(shape):
    """Dirichlet, Neumann and unconstrained faces, one of them dynamic."""
    eq = ns.ConvectionDiffusionEquation(3, [0.3, -0.2, 0.1], 0.05)
    mesh = ns.Mesh([(0.0, 1.0)] * 3, [1.0 / (n - 1) for n in shape])
    bcs = [
        (
            ns.DirichletBoundaryCondition(
                lambda x, t: np.full((len(x), 1), 0.5), is_static=True
            ),
            ns.NeumannBoundaryCondition(_zeros(1), is_static=True),
        ),
        (
            ns.NeumannBoundaryCondition(
                lambda x, t: np.full((len(x), 1), 0.25), is_static=True
            ),
            ns.DirichletBoundaryCondition(
                lambda x, t: x[:, :1] + x[:, 2:3], is_static=True
            ),
        ),
        (
            ns.DirichletBoundaryCondition(
                lambda x, t: np.full((len(x), 1), 1.0 + t), is_static=False
            ),
            ns.NeumannBoundaryCondition(_zeros(1), is_static=True),
        ),
    ]
    cp = ns.ConstrainedProblem(eq, mesh, bcs)
    rng = np.random.default_rng(5)
    y0 = rng.uniform(0.0, 1.0, mesh.vertices_shape + (1,))
    return cp, y0, 1e-5


def convection_diffusion_3d_static(shape):
    """Dirichlet, Neumann faces on every axis, all static."""
    eq = ns.ConvectionDiffusionEquation(3, [0.3, -0.2, 0.1], 0.05)
    mesh = ns.Mesh([(0.0, 1.0)] * 3, [1.0 / (n - 1) for n in shape])
    bcs = [
        (
            ns.DirichletBoundaryCondition(
                lambda x, t: np.full((len(x), 1), 0.5), is_static=True
            ),
            ns.NeumannBoundaryCondition(_zeros(1), is_static=True),
        ),
        (
            ns.NeumannBoundaryCondition(
                lambda x, t: x[:, :1] * 0.25, is_static=True
            ),
            ns.DirichletBoundaryCondition(
                lambda x, t: x[:, :1] + x[:, 2:3], is_static=True
            ),
        ),
        (
            ns.DirichletBoundaryCondition(
                lambda x, t: 1.0 + x[:, :1] * x[:, 1:2], is_static=True
            ),
            ns.NeumannBoundaryCondition(_zeros(1), is_static=True),
        ),
    ]
    cp = ns.ConstrainedProblem(eq, mesh, bcs)
    rng = np.random.default_rng(5)
    y0 = rng.uniform(0.0, 1.0, mesh.vertices_shape + (1,))
    return cp, y0, 1e-5


def cahn_hilliard_3d(shape):
    """An algebraic (LHS.Y) component next to the time-stepped one."""
    eq = ns.CahnHilliardEquation(3, gamma=0.5)
    mesh = ns.Mesh([(1.0, float(n)) for n in shape], [1.0] * 3)
    bc = ns.NeumannBoundaryCondition(_zeros(2), is_static=True)
    cp = ns.ConstrainedProblem(eq, mesh, [(bc, bc)] * 3)
    rng = np.random.default_rng(7)
    y0 = 0.05 * rng.uniform(-1.0, 1.0, mesh.vertices_shape + (2,))
    return cp, y0, 0.01


def diffusion_2d(shape):
    eq = ns.DiffusionEquation(2)
    mesh = ns.Mesh([(0.0, 10.0)] * 2, [10.0 / (n - 1) for n in shape])
    d = ns.DirichletBoundaryCondition(
        lambda x, t: np.full((len(x), 1), 1.5), is_static=True
    )
    n_ = ns.NeumannBoundaryCondition(_zeros(1), is_static=True)
    cp = ns.ConstrainedProblem(eq, mesh, [(d, d), (n_, n_)])
    rng = np.random.default_rng(11)
    y0 = rng.uniform(0.0, 2.0, mesh.vertices_shape + (1,))
    h = min(mesh.d_x)
    return cp, y0, 0.2 * h * h


def shallow_water_polar(shape):
    eq = ns.ShallowWaterEquation(0.5)
    mesh = ns.Mesh(
        [(4.0, 11.0), (0.5 * np.pi, 1.5 * np.pi)],
        [7.0 / (shape[0] - 1), np.pi / (shape[1] - 1)],
        ns.CoordinateSystem.POLAR,
    )
    bc = ns.NeumannBoundaryCondition(
        ns.vectorize_bc_function(lambda x, t: (0.0, None, None)), is_static=True
    )
    cp = ns.ConstrainedProblem(eq, mesh, [(bc, bc)] * 2)
    rng = np.random.default_rng(13)
    y0 = rng.uniform(-0.1, 0.1, mesh.vertices_shape + (3,))
    y0[..., 0] += 1.0
    return cp, y0, 1e-4


PROBLEMS = {
    # name: (builder, shape, tile "tx,ty", planes per chunk)
    "burgers_3d_multi_tile": (burgers_3d, (37, 35, 70), "32,16", "16"),
    "burgers_3d_small_tiles": (burgers_3d, (20, 21, 22), "8,4", "7"),
    "burgers_3d_default_tile": (burgers_3d, (40, 48, 64), None, None),
    "convection_diffusion_3d_mixed": (
        convection_diffusion_3d_mixed, (23, 19, 38), "16,8", "8"),
    "cahn_hilliard_3d": (cahn_hilliard_3d, (18, 21, 34), "16,8", "8"),
    "diffusion_2d_multi_tile": (diffusion_2d, (150, 600), "222,1", "32"),
    "diffusion_2d_small_tiles": (diffusion_2d, (45, 50), "16,1", "11"),
    "shallow_water_polar": (shallow_water_polar, (70, 300), "126,1", "24"),
}


def solve(problem, integrator, fuse, monkeypatch, n_steps=3):
    builder, shape, tile, zc = PROBLEMS[problem]
    monkeypatch.setenv("PML_SMALL", "0")
    monkeypatch.setenv("PML_FUSE", "1" if fuse else "0")
    for key, value in (("PML_FTILE", tile), ("PML_FZC", zc)):
        if fuse and value is not None:
            monkeypatch.setenv(key, value)
        else:
            monkeypatch.delenv(key, raising=False)
    cp, y0, d_t = builder(shape)
    ivp = ns.InitialValueProblem(
        cp, (0.0, n_steps * d_t), ns.DiscreteInitialCondition(cp, y0, True)
    )
    op = FDMOperator(integrator(), ThreePointCentralDifferenceMethod(), d_t)
    plan = op.prepare(ivp)[-1]
    assert (plan.fused is not None) == fuse
    before = dv.total_launches()
    y = op.solve(ivp).discrete_y()
    launches = dv.total_launches() - before
    return ivp, d_t, y, launches


@pytest.mark.parametrize("integrator", [RK4, ExplicitMidpointMethod],
                         ids=["rk4", "midpoint"])
@pytest.mark.parametrize("problem", list(PROBLEMS))
def test_fused_pairs_match_stage_kernels_and_oracle(problem, integrator, monkeypatch):
    ivp, d_t, fused, n_fused = solve(problem, integrator, True, monkeypatch)
    _, _, staged, n_staged = solve(problem, integrator, False, monkeypatch)
    assert np.isfinite(fused).all()
    # half as many launches: the pairs really ran
    assert n_fused < n_staged
    # same per-cell sequence of operations -> the same bits up to FMA
    # contraction choices of the compiler in the two instantiations
    assert per_step_rel_err(fused, staged) <= 1e-14
    name = "rk4" if integrator is RK4 else "explicit_midpoint"
    _, y_oracle = oracle.fdm_solve(ivp, name, d_t)
    assert per_step_rel_err(fused, y_oracle) <= 1e-12


@pytest.mark.parametrize(
    "problem", [p for p in PROBLEMS if p != "cahn_hilliard_3d"]
)
def test_euler_step_pairs_match_stage_kernels_and_oracle(problem, monkeypatch):
    """Forward Euler: two consecutive steps per launch of the stage-pair kernel
    (the odd last step runs the stage kernel)."""
    ivp, d_t, fused, n_fused = solve(problem, ForwardEulerMethod, True, monkeypatch, 5)
    _, _, staged, n_staged = solve(problem, ForwardEulerMethod, False, monkeypatch, 5)
    assert np.isfinite(fused).all()
    assert n_staged - n_fused == 2  # 2 pairs + 1 single against 5 launches
    assert per_step_rel_err(fused, staged) <= 1e-14
    _, y_oracle = oracle.fdm_solve(ivp, "forward_euler", d_t)
    assert per_step_rel_err(fused, y_oracle) <= 1e-12


@pytest.mark.parametrize(
    "rows,sync,variant", [("2", "0", "2"), ("1", "1", "2"), ("1", "0", "1")]
)
def test_pair_kernel_options_match_stage_kernels(rows, sync, variant, monkeypatch):
    """The non-default forms of the stage-pair kernels: register tiling over
    2 rows per thread, per-warp mbarrier arrivals instead of the block barrier,
    and the round-1 body (one thread per cell of the stage-A tile)."""
    monkeypatch.setenv("PML_FROWS", rows)
    monkeypatch.setenv("PML_FSYNC", sync)
    monkeypatch.setenv("PML_FVARIANT", variant)
    ivp, d_t, fused, n_fused = solve("burgers_3d_default_tile", RK4, True, monkeypatch)
    _, _, staged, n_staged = solve("burgers_3d_default_tile", RK4, False, monkeypatch)
    assert np.isfinite(fused).all() and n_fused < n_staged
    assert per_step_rel_err(fused, staged) <= 1e-14
    _, y_oracle = oracle.fdm_solve(ivp, "rk4", d_t)
    assert per_step_rel_err(fused, y_oracle) <= 1e-12
