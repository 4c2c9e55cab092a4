"""Named parity cases, buildable from EITHER host-class namespace.

``build(ns)`` receives a namespace exposing the reference's public names
(``Mesh``, ``ConstrainedProblem``, ``DiffusionEquation`` ...): the reference
package itself when golden vectors are generated
(``tests/golden/generate_golden.py``), ``pararealml_b200`` when the oracle or
the CUDA path is checked.  Scenarios follow the reference's examples
(``examples/*.py``) and its FDM operator tests
(``tests/operators/fdm/test_fdm_operator.py``) at sizes whose trajectories fit
small fixtures.
"""
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np


@dataclass
class FDMCase:
    name: str
    build: Callable  # ns -> ivp
    integrator: str  # forward_euler | explicit_midpoint | rk4
    d_t: float
    tol: float = 1e-3  # Jacobi tolerance of the differentiator
    seed: Optional[int] = None  # np.random.seed before solve (Jacobi start)
    stride: int = 1  # stored trajectory stride (the last step is always kept)
    rtol_traj: float = 1e-9
    tags: tuple = field(default_factory=tuple)


@dataclass
class PararealCase:
    name: str
    build: Callable  # ns -> ivp
    f: tuple  # (integrator, d_t)
    g: tuple
    tol: object  # float | sequence
    sizes: tuple = (1, 2, 4)
    stride: int = 1
    tags: tuple = field(default_factory=tuple)


def _zeros(k):
    return lambda x, t: np.zeros((len(x), k))


def _full(k, v):
    return lambda x, t: np.full((len(x), k), v)


# ---------------------------------------------------------------------------
# builders
# ---------------------------------------------------------------------------
def diffusion_1d_dynamic(ns, t_end=0.5):
    """examples/diffusion_1d_fdm.py with a shorter time interval."""
    eq = ns.DiffusionEquation(1, 1.5)
    mesh = ns.Mesh([(0.0, 10.0)], [0.1])
    bcs = [
        (
            ns.NeumannBoundaryCondition(_zeros(1)),
            ns.DirichletBoundaryCondition(
                lambda x, t: np.full((len(x), 1), t / 5.0)
            ),
        )
    ]
    cp = ns.ConstrainedProblem(eq, mesh, bcs)
    ic = ns.GaussianInitialCondition(
        cp, [(np.array([5.0]), np.array([[0.5]]))], [5.0]
    )
    return ns.InitialValueProblem(cp, (0.0, t_end), ic)


def diffusion_1d_static(ns, t_end=0.5):
    eq = ns.DiffusionEquation(1, 1.5)
    mesh = ns.Mesh([(0.0, 10.0)], [0.1])
    bcs = [
        (
            ns.NeumannBoundaryCondition(_zeros(1), is_static=True),
            ns.DirichletBoundaryCondition(_full(1, 0.3), is_static=True),
        )
    ]
    cp = ns.ConstrainedProblem(eq, mesh, bcs)
    ic = ns.GaussianInitialCondition(
        cp, [(np.array([5.0]), np.array([[0.5]]))], [5.0]
    )
    return ns.InitialValueProblem(cp, (0.0, t_end), ic)


def diffusion_1d_coarse_dynamic(ns):
    """tests/operators/fdm/test_fdm_operator.py:293-319."""
    eq = ns.DiffusionEquation(1, 1.5)
    mesh = ns.Mesh([(0.0, 10.0)], [1.0])
    bcs = [
        (
            ns.NeumannBoundaryCondition(_zeros(1)),
            ns.DirichletBoundaryCondition(
                lambda x, t: np.full((len(x), 1), t / 5.0)
            ),
        )
    ]
    cp = ns.ConstrainedProblem(eq, mesh, bcs)
    ic = ns.GaussianInitialCondition(
        cp, [(np.array([5.0]), np.array([[2.5]]))], [20.0]
    )
    return ns.InitialValueProblem(cp, (0.0, 10.0), ic)


def diffusion_2d(ns, t_end=0.1, d_x=0.5):
    """examples/diffusion_2d_parareal.py problem definition."""
    eq = ns.DiffusionEquation(2)
    mesh = ns.Mesh([(0.0, 10.0), (0.0, 10.0)], [d_x, d_x])
    bcs = [
        (
            ns.DirichletBoundaryCondition(_full(1, 1.5), is_static=True),
            ns.DirichletBoundaryCondition(_full(1, 1.5), is_static=True),
        ),
        (
            ns.NeumannBoundaryCondition(_zeros(1), is_static=True),
            ns.NeumannBoundaryCondition(_zeros(1), is_static=True),
        ),
    ]
    cp = ns.ConstrainedProblem(eq, mesh, bcs)
    ic = ns.GaussianInitialCondition(
        cp,
        [(np.array([5.0, 5.0]), np.array([[1.0, 0.0], [0.0, 1.0]]))],
        [1000.0],
    )
    return ns.InitialValueProblem(cp, (0.0, t_end), ic)


def convection_diffusion_2d_mixed(ns):
    """Mixed Dirichlet / Cauchy / partially constrained Neumann faces with
    corner precedence."""
    eq = ns.ConvectionDiffusionEquation(2, [0.7, -0.4], 0.8)
    mesh = ns.Mesh([(0.0, 4.0), (-1.0, 2.0)], [0.25, 0.2])

    def partial_flux(x, t):
        v = np.full((len(x), 1), 0.2)
        v[x[:, 0] > 2.0] = np.nan
        return v

    bcs = [
        (
            ns.DirichletBoundaryCondition(
                lambda x, t: 0.5 + 0.1 * x[:, 1:2], is_static=True
            ),
            ns.CauchyBoundaryCondition(
                _full(1, -0.25), _full(1, 0.05), is_static=True
            ),
        ),
        (
            ns.NeumannBoundaryCondition(partial_flux, is_static=True),
            ns.DirichletBoundaryCondition(
                lambda x, t: np.sin(x[:, :1]), is_static=True
            ),
        ),
    ]
    cp = ns.ConstrainedProblem(eq, mesh, bcs)
    ic = ns.GaussianInitialCondition(
        cp,
        [(np.array([2.0, 0.5]), np.array([[0.4, 0.1], [0.1, 0.3]]))],
        [3.0],
    )
    return ns.InitialValueProblem(cp, (0.0, 0.2), ic)


def wave_2d_dynamic(ns):
    """Two components, time dependent Dirichlet on one face, dynamic flux."""
    eq = ns.WaveEquation(2, 1.3)
    mesh = ns.Mesh([(-2.0, 2.0), (0.0, 3.0)], [0.2, 0.25])
    bcs = [
        (
            ns.DirichletBoundaryCondition(
                lambda x, t: np.stack(
                    [0.2 * np.sin(3.0 * t) * np.ones(len(x)),
                     np.full(len(x), np.nan)],
                    axis=-1,
                )
            ),
            ns.NeumannBoundaryCondition(
                lambda x, t: np.stack(
                    [0.1 * t * x[:, 1], np.zeros(len(x))], axis=-1
                )
            ),
        ),
        (
            ns.NeumannBoundaryCondition(_zeros(2), is_static=True),
            ns.DirichletBoundaryCondition(_zeros(2), is_static=True),
        ),
    ]
    cp = ns.ConstrainedProblem(eq, mesh, bcs)
    ic = ns.GaussianInitialCondition(
        cp,
        [(np.array([0.0, 1.5]), np.array([[0.2, 0.0], [0.0, 0.2]]))] * 2,
        [1.0, 0.0],
    )
    return ns.InitialValueProblem(cp, (0.0, 0.5), ic)


def cahn_hilliard_3d(ns, n=10, t_end=0.25):
    """examples/cahn_hilliard_3d_fdm.py on a smaller mesh; the initial state
    comes from a seeded generator and a plain second difference so that both
    namespaces see identical bits."""
    gamma = 0.5
    eq = ns.CahnHilliardEquation(3, gamma=gamma)
    mesh = ns.Mesh([(1.0, float(n))] * 3, [1.0] * 3)
    bcs = [
        (
            ns.NeumannBoundaryCondition(_zeros(2), is_static=True),
            ns.NeumannBoundaryCondition(_zeros(2), is_static=True),
        )
    ] * 3
    cp = ns.ConstrainedProblem(eq, mesh, bcs)
    rng = np.random.default_rng(7)
    y0 = 0.05 * rng.uniform(-1.0, 1.0, mesh.vertices_shape + (1,))
    lap = np.zeros_like(y0)
    for a in range(3):
        p = np.concatenate(
            [np.take(y0, [1], axis=a), y0, np.take(y0, [-2], axis=a)], axis=a
        )
        sl = [slice(None)] * 4
        lo, mid, hi = list(sl), list(sl), list(sl)
        lo[a], mid[a], hi[a] = slice(0, -2), slice(1, -1), slice(2, None)
        lap += p[tuple(hi)] - 2.0 * p[tuple(mid)] + p[tuple(lo)]
    y1 = y0**3 - y0 - gamma * lap
    ic = ns.DiscreteInitialCondition(
        cp, np.concatenate([y0, y1], axis=-1), True
    )
    return ns.InitialValueProblem(cp, (0.0, t_end), ic)


def shallow_water_polar(ns, nr=15, nth=20, t_end=0.025):
    """examples/shallow_water_polar_fdm.py on a smaller mesh."""
    eq = ns.ShallowWaterEquation(0.5)
    mesh = ns.Mesh(
        [(4.0, 11.0), (0.5 * np.pi, 1.5 * np.pi)],
        [7.0 / (nr - 1), np.pi / (nth - 1)],
        ns.CoordinateSystem.POLAR,
    )
    bc = ns.NeumannBoundaryCondition(
        ns.vectorize_bc_function(lambda x, t: (0.0, None, None)),
        is_static=True,
    )
    cp = ns.ConstrainedProblem(eq, mesh, [(bc, bc)] * 2)
    ic = ns.GaussianInitialCondition(
        cp,
        [(np.array([-6.0, 6.0]), np.array([[0.25, 0.0], [0.0, 0.25]]))] * 3,
        [1.0, 0.0, 0.0],
    )
    return ns.InitialValueProblem(cp, (0.0, t_end), ic)


def burgers_3d_spherical(ns):
    """examples/burgers_3d_fdm.py (spherical mesh), few steps."""
    eq = ns.BurgersEquation(3, 100)
    mesh = ns.Mesh(
        [(1.0, 5.0), (0.0, 2.0 * np.pi), (0.25 * np.pi, 0.75 * np.pi)],
        [0.5, np.pi / 10.0, np.pi / 10.0],
        ns.CoordinateSystem.SPHERICAL,
    )
    bc = ns.NeumannBoundaryCondition(_zeros(3), is_static=True)
    cp = ns.ConstrainedProblem(eq, mesh, [(bc, bc)] * 3)
    ic = ns.ContinuousInitialCondition(
        cp,
        lambda x: np.stack(
            [1.0 / x[:, 0] ** 2, np.zeros_like(x[:, 1]), np.zeros_like(x[:, 1])],
            axis=-1,
        ),
    )
    return ns.InitialValueProblem(cp, (0.0, 2.0), ic)


def burgers_3d_cartesian(ns, n=12, t_end=0.004):
    """K5's equation and boundary conditions on a small mesh."""
    eq = ns.BurgersEquation(3, 100)
    mesh = ns.Mesh([(0.0, 1.0)] * 3, [1.0 / (n - 1)] * 3)
    bc = ns.NeumannBoundaryCondition(_zeros(3), is_static=True)
    cp = ns.ConstrainedProblem(eq, mesh, [(bc, bc)] * 3)
    ic = ns.GaussianInitialCondition(
        cp,
        [(np.array([0.5, 0.5, 0.5]), 0.05 * np.eye(3))] * 3,
        [0.3, -0.2, 0.1],
    )
    return ns.InitialValueProblem(cp, (0.0, t_end), ic)


def diffusion_cylindrical(ns):
    """tests/operators/fdm/test_fdm_operator.py (cylindrical diffusion)."""
    eq = ns.DiffusionEquation(3, 0.7)
    mesh = ns.Mesh(
        [(1.0, 3.0), (0.0, np.pi), (-1.0, 1.0)],
        [0.25, np.pi / 8.0, 0.5],
        ns.CoordinateSystem.CYLINDRICAL,
    )
    bcs = [
        (
            ns.DirichletBoundaryCondition(_full(1, 0.1), is_static=True),
            ns.NeumannBoundaryCondition(_full(1, -0.05), is_static=True),
        ),
        (
            ns.NeumannBoundaryCondition(_zeros(1), is_static=True),
            ns.NeumannBoundaryCondition(_zeros(1), is_static=True),
        ),
        (
            ns.DirichletBoundaryCondition(_full(1, 0.2), is_static=True),
            ns.DirichletBoundaryCondition(_full(1, 0.0), is_static=True),
        ),
    ]
    cp = ns.ConstrainedProblem(eq, mesh, bcs)
    ic = ns.GaussianInitialCondition(
        cp, [(np.array([0.0, 2.0, 0.0]), 0.5 * np.eye(3))], [2.0]
    )
    return ns.InitialValueProblem(cp, (0.0, 0.05), ic)


def navier_stokes_2d(ns, d_x=0.25, t_end=0.1):
    """examples/navier_stokes_fdm.py on a coarser mesh (Jacobi start drawn
    from the seeded global NumPy stream)."""
    eq = ns.NavierStokesEquation(5000.0)
    mesh = ns.Mesh([(-2.5, 2.5), (0.0, 4.0)], [d_x, d_x])
    wall = ns.DirichletBoundaryCondition(
        ns.vectorize_bc_function(lambda x, t: [0.0, 0.0, None, None]),
        is_static=True,
    )
    inlet = ns.DirichletBoundaryCondition(
        ns.vectorize_bc_function(lambda x, t: [1.0, 0.1, None, None]),
        is_static=True,
    )
    cp = ns.ConstrainedProblem(eq, mesh, [(inlet, wall), (wall, wall)])
    ic = ns.ContinuousInitialCondition(cp, lambda x: np.zeros((len(x), 4)))
    return ns.InitialValueProblem(cp, (0.0, t_end), ic)


def all_leaves_2d(ns):
    """A user-defined system touching every leaf kind the FDM symbol mapper
    supports in 2D: t, x, y, gradient, hessian (same axis and mixed),
    laplacian and divergence (the reference cannot express a 2D curl symbol:
    ``Symbols`` builds ``y-curl`` with an empty shape for x_dimension == 2,
    differential_equation.py:44-50)."""

    class AllLeaves2D(ns.DifferentialEquation):
        def __init__(self):
            super().__init__(2, 2, [(0, 1)])

        @property
        def symbolic_equation_system(self):
            s = self.symbols
            return ns.SymbolicEquationSystem(
                [
                    0.3 * s.y_laplacian[0]
                    + 0.1 * s.y_hessian[1, 0, 1]
                    - 0.2 * s.y[1] * s.y_gradient[0, 1]
                    + s.t / 10.0 * (s.x[0] + s.x[1]) ** 2,
                    0.25 * s.y_hessian[1, 0, 0]
                    + 0.15 * s.y_hessian[1, 1, 1]
                    + 0.05 * s.y_hessian[0, 1, 0]
                    - 0.4 * s.y_divergence[0, 1]
                    - s.y[0] ** 2,
                ]
            )

    eq = AllLeaves2D()
    mesh = ns.Mesh([(-1.0, 1.0), (0.0, 1.5)], [0.2, 0.25])
    bcs = [
        (
            ns.NeumannBoundaryCondition(
                lambda x, t: np.stack(
                    [0.1 * np.ones(len(x)), np.full(len(x), np.nan)], axis=-1
                ),
                is_static=True,
            ),
            ns.DirichletBoundaryCondition(
                lambda x, t: np.stack(
                    [np.full(len(x), np.nan), 0.2 * x[:, 1]], axis=-1
                ),
                is_static=True,
            ),
        ),
        (
            ns.CauchyBoundaryCondition(
                _full(2, 0.05), _full(2, -0.1), is_static=True
            ),
            ns.NeumannBoundaryCondition(_zeros(2), is_static=True),
        ),
    ]
    cp = ns.ConstrainedProblem(eq, mesh, bcs)
    ic = ns.GaussianInitialCondition(
        cp,
        [(np.array([0.0, 0.7]), np.array([[0.2, 0.05], [0.05, 0.3]]))] * 2,
        [0.5, -0.3],
    )
    return ns.InitialValueProblem(cp, (0.0, 0.1), ic)


def all_leaves_3d(ns, cs_name="CARTESIAN"):
    """3D system with mixed hessians, divergence and all curl components; run
    in Cartesian, cylindrical and spherical coordinates."""

    class AllLeaves3D(ns.DifferentialEquation):
        def __init__(self):
            super().__init__(3, 3, [(0, 1, 2)])

        @property
        def symbolic_equation_system(self):
            s = self.symbols
            return ns.SymbolicEquationSystem(
                [
                    0.2 * s.y_laplacian[0]
                    + 0.05 * s.y_hessian[1, 0, 2]
                    + 0.1 * s.y_curl[0, 1, 2, 0]
                    - 0.1 * s.y[0] * s.y_gradient[0, 0],
                    0.2 * s.y_hessian[1, 1, 1]
                    + 0.1 * s.y_hessian[1, 2, 1]
                    + 0.07 * s.y_hessian[2, 0, 1]
                    + 0.1 * s.y_curl[0, 1, 2, 1]
                    - 0.3 * s.y_divergence[0, 1, 2],
                    0.15 * s.y_hessian[2, 2, 2]
                    + 0.15 * s.y_hessian[2, 0, 0]
                    + 0.1 * s.y_curl[0, 1, 2, 2]
                    + 0.02 * s.y_gradient[2, 1]
                    - 0.05 * s.y_gradient[1, 2]
                    + 0.01 * s.x[0] * s.t,
                ]
            )

    eq = AllLeaves3D()
    cs = ns.CoordinateSystem[cs_name]
    if cs_name == "CARTESIAN":
        mesh = ns.Mesh([(0.0, 1.0), (0.0, 1.2), (-0.5, 0.5)], [0.2, 0.2, 0.25])
    elif cs_name == "CYLINDRICAL":
        mesh = ns.Mesh(
            [(1.0, 2.0), (0.0, np.pi), (0.0, 1.0)],
            [0.2, np.pi / 6.0, 0.25],
            cs,
        )
    else:
        mesh = ns.Mesh(
            [(1.0, 2.0), (0.0, np.pi), (0.25 * np.pi, 0.75 * np.pi)],
            [0.2, np.pi / 6.0, np.pi / 8.0],
            cs,
        )
    flux = ns.NeumannBoundaryCondition(
        lambda x, t: np.stack(
            [0.1 * np.ones(len(x)), np.full(len(x), np.nan), np.zeros(len(x))],
            axis=-1,
        ),
        is_static=True,
    )
    value = ns.DirichletBoundaryCondition(
        lambda x, t: np.stack(
            [np.full(len(x), np.nan), 0.1 * np.ones(len(x)), 0.05 * x[:, 0]],
            axis=-1,
        ),
        is_static=True,
    )
    both = ns.CauchyBoundaryCondition(
        _full(3, 0.02), _full(3, -0.03), is_static=True
    )
    cp = ns.ConstrainedProblem(
        eq, mesh, [(flux, value), (both, flux), (value, flux)]
    )
    ic = ns.ContinuousInitialCondition(
        cp,
        lambda x: np.stack(
            [
                0.1 * np.sin(2.0 * x[:, 0]) * np.cos(x[:, 1]),
                0.2 * np.cos(x[:, 0] + x[:, 2]),
                0.1 * x[:, 1] * np.sin(3.0 * x[:, 2]),
            ],
            axis=-1,
        ),
    )
    return ns.InitialValueProblem(cp, (0.0, 0.05), ic)


def lorenz(ns, t_end=1.0):
    eq = ns.LorenzEquation()
    cp = ns.ConstrainedProblem(eq)
    ic = ns.ContinuousInitialCondition(cp, lambda _: np.ones(3))
    return ns.InitialValueProblem(cp, (0.0, t_end), ic)


def n_body(ns):
    eq = ns.NBodyGravitationalEquation(2, [4.0, 3.0, 5.0], g=1.0)
    cp = ns.ConstrainedProblem(eq)
    y0 = np.array(
        [0.0, 0.0, 3.0, 0.0, 0.0, 4.0, 0.0, 0.1, 0.0, -0.7, 0.6, 0.0]
    )
    ic = ns.ContinuousInitialCondition(cp, lambda _: y0)
    return ns.InitialValueProblem(cp, (0.0, 0.5), ic)


def population_growth(ns):
    eq = ns.PopulationGrowthEquation(0.02)
    cp = ns.ConstrainedProblem(eq)
    ic = ns.ContinuousInitialCondition(cp, lambda _: np.array([100.0]))
    return ns.InitialValueProblem(cp, (0.0, 1.0), ic)


FDM_CASES = [
    FDMCase("diffusion_1d_dynamic_rk4", diffusion_1d_dynamic, "rk4", 0.0025, stride=10),
    FDMCase("diffusion_1d_static_rk4", diffusion_1d_static, "rk4", 0.0025, stride=10),
    FDMCase("diffusion_1d_static_fe", diffusion_1d_static, "forward_euler", 0.0025, stride=10),
    FDMCase("diffusion_1d_coarse_dynamic_rk4", diffusion_1d_coarse_dynamic, "rk4", 0.5),
    FDMCase("diffusion_1d_coarse_dynamic_mid", diffusion_1d_coarse_dynamic, "explicit_midpoint", 0.5),
    FDMCase("diffusion_2d_rk4", diffusion_2d, "rk4", 1e-3, stride=10),
    FDMCase("diffusion_2d_midpoint", diffusion_2d, "explicit_midpoint", 1e-3, stride=10),
    FDMCase("diffusion_2d_fe", diffusion_2d, "forward_euler", 1e-3, stride=10),
    FDMCase("convection_diffusion_2d_mixed_rk4", convection_diffusion_2d_mixed, "rk4", 0.005, stride=5),
    FDMCase("wave_2d_dynamic_rk4", wave_2d_dynamic, "rk4", 0.01, stride=5),
    FDMCase("wave_2d_dynamic_mid", wave_2d_dynamic, "explicit_midpoint", 0.01, stride=5),
    FDMCase("cahn_hilliard_3d_rk4", cahn_hilliard_3d, "rk4", 0.05),
    FDMCase("shallow_water_polar_rk4", shallow_water_polar, "rk4", 0.0025),
    FDMCase("burgers_3d_spherical_rk4", burgers_3d_spherical, "rk4", 0.5),
    FDMCase("burgers_3d_cartesian_rk4", burgers_3d_cartesian, "rk4", 0.001),
    FDMCase("burgers_3d_cartesian_fe", burgers_3d_cartesian, "forward_euler", 0.001),
    FDMCase("diffusion_cylindrical_rk4", diffusion_cylindrical, "rk4", 0.005),
    FDMCase("navier_stokes_2d_rk4", navier_stokes_2d, "rk4", 0.05, seed=0, rtol_traj=1e-7, tags=("jacobi",)),
    FDMCase("all_leaves_2d_rk4", all_leaves_2d, "rk4", 0.01),
    FDMCase("all_leaves_3d_cartesian_rk4", lambda ns: all_leaves_3d(ns, "CARTESIAN"), "rk4", 0.01),
    FDMCase("all_leaves_3d_cylindrical_rk4", lambda ns: all_leaves_3d(ns, "CYLINDRICAL"), "rk4", 0.01),
    FDMCase("all_leaves_3d_spherical_mid", lambda ns: all_leaves_3d(ns, "SPHERICAL"), "explicit_midpoint", 0.01),
    FDMCase("lorenz_fe", lorenz, "forward_euler", 0.01, stride=10),
    FDMCase("lorenz_rk4", lorenz, "rk4", 0.01, stride=10),
    FDMCase("n_body_rk4", n_body, "rk4", 0.01, stride=5),
    FDMCase("population_growth_rk4", population_growth, "rk4", 1e-2, stride=10),
]

PARAREAL_CASES = [
    # examples/diffusion_2d_parareal.py, shorter interval (converges in 1 it.)
    PararealCase(
        "parareal_diffusion_2d_example",
        lambda ns: diffusion_2d(ns, t_end=0.4),
        ("rk4", 1e-3), ("rk4", 1e-2), 0.0025, stride=20,
    ),
    # coarse Forward Euler with a big step and a tight tolerance: needs
    # several corrective iterations
    PararealCase(
        "parareal_diffusion_2d_multi_iteration",
        lambda ns: diffusion_2d(ns, t_end=0.8),
        ("rk4", 2e-3), ("forward_euler", 5e-2), 1e-4, sizes=(1, 2, 4, 8),
        stride=20,
    ),
    PararealCase(
        "parareal_lorenz",
        lambda ns: lorenz(ns, t_end=0.8),
        ("rk4", 1e-3), ("forward_euler", 1e-2), [1e-3, 1e-3, 1e-3],
        stride=50,
    ),
    PararealCase(
        "parareal_burgers_3d",
        lambda ns: burgers_3d_cartesian(ns, n=8, t_end=0.016),
        ("rk4", 1e-3), ("forward_euler", 2e-3), 1e-6, sizes=(1, 2, 4, 8),
    ),
]

FDM_BY_NAME = {c.name: c for c in FDM_CASES}
PARAREAL_BY_NAME = {c.name: c for c in PARAREAL_CASES}
