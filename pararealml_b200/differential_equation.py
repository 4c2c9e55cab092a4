"""Symbolic (SymPy) definitions of time dependent differential equations.

Host-side input format of the hot path: the B200 code generator reads
``DifferentialEquation.symbolic_equation_system`` and emits CUDA C from it.
API mirror of the reference's ``pararealml/differential_equation.py``: symbol
names (``y_0``, ``y-gradient_0_1``, ``y-laplacian_2`` ...; reference :11-137),
``LHS`` kinds (:140-149) and the stock equations (:355-850) are identical so
either package's equation objects can be lowered.
"""
from abc import ABC, abstractmethod
from copy import copy, deepcopy
from enum import Enum
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
from sympy import Expr, Symbol, symarray


class Symbols:
    """All symbols an equation with the given dimensions may use."""

    def __init__(self, x_dimension: int, y_dimension: int):
        xd, yd = x_dimension, y_dimension
        self._t = Symbol("t")
        self._y = symarray("y", (yd,))
        self._x = None
        self._y_gradient = None
        self._y_hessian = None
        self._y_divergence = None
        self._y_curl = None
        self._y_laplacian = None
        self._y_vector_laplacian = None
        if xd:
            self._x = symarray("x", (xd,))
            self._y_gradient = symarray("y-gradient", (yd, xd))
            self._y_hessian = symarray("y-hessian", (yd, xd, xd))
            self._y_divergence = symarray("y-divergence", (yd,) * xd)
            if xd == 2:
                self._y_curl = symarray("y-curl", (yd, yd))
            elif xd == 3:
                self._y_curl = symarray("y-curl", (yd, yd, yd, 3))
            self._y_laplacian = symarray("y-laplacian", (yd,))
            self._y_vector_laplacian = symarray(
                "y-vector-laplacian", (yd,) * xd + (xd,)
            )

    @property
    def t(self) -> Symbol:
        return self._t

    @property
    def y(self) -> np.ndarray:
        return copy(self._y)

    @property
    def x(self) -> Optional[np.ndarray]:
        return copy(self._x)

    @property
    def y_gradient(self) -> Optional[np.ndarray]:
        return copy(self._y_gradient)

    @property
    def y_hessian(self) -> Optional[np.ndarray]:
        return copy(self._y_hessian)

    @property
    def y_divergence(self) -> Optional[np.ndarray]:
        return copy(self._y_divergence)

    @property
    def y_curl(self) -> Optional[np.ndarray]:
        return copy(self._y_curl)

    @property
    def y_laplacian(self) -> Optional[np.ndarray]:
        return copy(self._y_laplacian)

    @property
    def y_vector_laplacian(self) -> Optional[np.ndarray]:
        return copy(self._y_vector_laplacian)

    def all(self) -> set:
        out = {self._t, *self._y}
        for arr in (
            self._x,
            self._y_gradient,
            self._y_hessian,
            self._y_divergence,
            self._y_curl,
            self._y_laplacian,
            self._y_vector_laplacian,
        ):
            if arr is not None:
                out.update(arr.flatten())
        return out


class LHS(Enum):
    """What the left-hand side of an equation of the system is."""

    D_Y_OVER_D_T = 0
    Y = 1
    Y_LAPLACIAN = 2


class SymbolicEquationSystem:
    def __init__(
        self,
        rhs: Union[Sequence[Expr], np.ndarray],
        lhs_types: Optional[Sequence[LHS]] = None,
    ):
        if len(rhs) < 1:
            raise ValueError("an equation system needs at least one equation")
        if lhs_types is None:
            lhs_types = [LHS.D_Y_OVER_D_T] * len(rhs)
        if len(lhs_types) != len(rhs):
            raise ValueError(
                f"{len(rhs)} right-hand sides but {len(lhs_types)} "
                "left-hand side types"
            )
        self._rhs = copy(rhs)
        self._lhs_types = copy(lhs_types)
        self._by_type: Dict[LHS, List[int]] = {k: [] for k in LHS}
        for i, k in enumerate(lhs_types):
            self._by_type[k].append(i)

    @property
    def rhs(self):
        return copy(self._rhs)

    @property
    def lhs_types(self) -> Sequence[LHS]:
        return copy(self._lhs_types)

    def equation_indices_by_type(self, lhs_type: LHS) -> Sequence[int]:
        return copy(self._by_type[lhs_type])


class DifferentialEquation(ABC):
    def __init__(
        self,
        x_dimension: int,
        y_dimension: int,
        all_vector_field_indices: Optional[Sequence[Sequence[int]]] = None,
    ):
        if x_dimension < 0:
            raise ValueError("x dimension must be non-negative")
        if y_dimension < 1:
            raise ValueError("y dimension must be at least 1")
        for indices in all_vector_field_indices or ():
            if len(indices) != x_dimension:
                raise ValueError(
                    f"vector field {indices} must have {x_dimension} "
                    "components"
                )
            if any(not 0 <= i < y_dimension for i in indices):
                raise ValueError(
                    f"vector field indices {indices} out of range"
                )
        self._x_dimension = x_dimension
        self._y_dimension = y_dimension
        self._all_vector_field_indices = deepcopy(all_vector_field_indices)
        self._symbols = Symbols(x_dimension, y_dimension)
        self._validate()

    @property
    def x_dimension(self) -> int:
        return self._x_dimension

    @property
    def y_dimension(self) -> int:
        return self._y_dimension

    @property
    def symbols(self) -> Symbols:
        return self._symbols

    @property
    def all_vector_field_indices(self):
        return deepcopy(self._all_vector_field_indices)

    @property
    @abstractmethod
    def symbolic_equation_system(self) -> SymbolicEquationSystem:
        """rhs[i] defines dy_i/dt, y_i or laplacian(y_i) per lhs_types[i]."""

    def _validate(self):
        system = self.symbolic_equation_system
        if len(system.rhs) != self._y_dimension:
            raise ValueError(
                f"{len(system.rhs)} equations for {self._y_dimension} unknowns"
            )
        legal = self._symbols.all()
        for i, expr in enumerate(system.rhs):
            if not expr.free_symbols <= legal:
                raise ValueError(
                    f"equation {i} uses unknown symbols "
                    f"{expr.free_symbols - legal}"
                )
        n_dt = len(system.equation_indices_by_type(LHS.D_Y_OVER_D_T))
        if self._x_dimension:
            if n_dt == 0:
                raise ValueError(
                    "a PDE system needs at least one D_Y_OVER_D_T equation"
                )
        elif n_dt != self._y_dimension:
            raise ValueError("ODE systems may only have D_Y_OVER_D_T equations")


def _require_pde(x_dimension: int):
    if x_dimension <= 0:
        raise ValueError("x dimension must be at least 1")


class PopulationGrowthEquation(DifferentialEquation):
    def __init__(self, r: float = 0.01):
        self._r = r
        super().__init__(0, 1)

    @property
    def symbolic_equation_system(self):
        return SymbolicEquationSystem([self._r * self._symbols.y[0]])


class LotkaVolterraEquation(DifferentialEquation):
    def __init__(self, alpha=2.0, beta=0.04, gamma=1.06, delta=0.02):
        if min(alpha, beta, gamma, delta) < 0.0:
            raise ValueError("coefficients must be non-negative")
        self._alpha, self._beta = alpha, beta
        self._gamma, self._delta = gamma, delta
        super().__init__(0, 2)

    @property
    def symbolic_equation_system(self):
        prey, pred = self._symbols.y
        return SymbolicEquationSystem(
            [
                self._alpha * prey - self._beta * prey * pred,
                self._delta * prey * pred - self._gamma * pred,
            ]
        )


class LorenzEquation(DifferentialEquation):
    def __init__(self, sigma=10.0, rho=28.0, beta=8.0 / 3.0):
        if min(sigma, rho, beta) < 0.0:
            raise ValueError("coefficients must be non-negative")
        self._sigma, self._rho, self._beta = sigma, rho, beta
        super().__init__(0, 3)

    @property
    def symbolic_equation_system(self):
        c, h, v = self._symbols.y
        return SymbolicEquationSystem(
            [
                self._sigma * (h - c),
                c * (self._rho - v) - h,
                c * h - self._beta * v,
            ]
        )


class SIREquation(DifferentialEquation):
    def __init__(self, beta=0.2, gamma=0.1):
        if beta < 0.0 or gamma < 0.0:
            raise ValueError("beta and gamma must be non-negative")
        self._beta, self._gamma = beta, gamma
        super().__init__(0, 3)

    @property
    def symbolic_equation_system(self):
        s, i, r = self._symbols.y
        n = s + i + r
        return SymbolicEquationSystem(
            [
                -self._beta * s * i / n,
                self._beta * s * i / n - self._gamma * i,
                self._gamma * i,
            ]
        )


class VanDerPolEquation(DifferentialEquation):
    def __init__(self, mu=1.0):
        if mu < 0.0:
            raise ValueError("mu must be non-negative")
        self._mu = mu
        super().__init__(0, 2)

    @property
    def symbolic_equation_system(self):
        u, v = self._symbols.y
        return SymbolicEquationSystem([v, self._mu * (1.0 - u**2) * v - u])


class NBodyGravitationalEquation(DifferentialEquation):
    """State = all positions (object-major) followed by all velocities."""

    def __init__(self, n_dims: int, masses: Sequence[float], g=6.6743e-11):
        if n_dims not in (2, 3):
            raise ValueError("n_dims must be 2 or 3")
        if len(masses) < 2:
            raise ValueError("at least 2 masses are needed")
        if any(m <= 0.0 for m in masses):
            raise ValueError("masses must be positive")
        self._dims = n_dims
        self._masses = tuple(masses)
        self._n_objects = len(masses)
        self._g = g
        super().__init__(0, 2 * len(masses) * n_dims)

    @property
    def spatial_dimension(self) -> int:
        return self._dims

    @property
    def masses(self) -> Tuple[float, ...]:
        return copy(self._masses)

    @property
    def n_objects(self) -> int:
        return self._n_objects

    @property
    def symbolic_equation_system(self):
        y = np.array(self._symbols.y, dtype=object)
        d, n = self._dims, self._n_objects
        half = n * d
        rhs = np.empty(self._y_dimension, dtype=object)
        rhs[:half] = y[half:]
        pull = np.zeros((n, n, d), dtype=object)
        for i in range(n):
            pos_i = y[i * d : (i + 1) * d]
            for j in range(i + 1, n):
                pos_j = y[j * d : (j + 1) * d]
                sep = pos_j - pos_i
                dist = np.power(np.power(sep, 2).sum(axis=-1), 0.5)
                f_ij = (self._g * self._masses[i] * self._masses[j]) * (
                    sep / np.power(dist, 3)
                )
                pull[i, j, :] = f_ij
                pull[j, i, :] = -f_ij
            rhs[half + i * d : half + (i + 1) * d] = (
                pull[i, :, :].sum(axis=0) / self._masses[i]
            )
        return SymbolicEquationSystem(rhs)


class DiffusionEquation(DifferentialEquation):
    def __init__(self, x_dimension: int, d: float = 1.0):
        _require_pde(x_dimension)
        self._d = d
        super().__init__(x_dimension, 1)

    @property
    def symbolic_equation_system(self):
        return SymbolicEquationSystem([self._d * self._symbols.y_laplacian[0]])


class ConvectionDiffusionEquation(DifferentialEquation):
    def __init__(self, x_dimension: int, velocity: Sequence[float], d=1.0):
        _require_pde(x_dimension)
        if len(velocity) != x_dimension:
            raise ValueError("velocity must have one entry per x dimension")
        self._velocity = copy(velocity)
        self._d = d
        super().__init__(x_dimension, 1)

    @property
    def symbolic_equation_system(self):
        sym = self._symbols
        return SymbolicEquationSystem(
            [
                self._d * sym.y_laplacian[0]
                - np.dot(self._velocity, sym.y_gradient[0, :])
            ]
        )


class WaveEquation(DifferentialEquation):
    def __init__(self, x_dimension: int, c: float = 1.0):
        _require_pde(x_dimension)
        self._c = c
        super().__init__(x_dimension, 2)

    @property
    def symbolic_equation_system(self):
        sym = self._symbols
        return SymbolicEquationSystem(
            [sym.y[1], (self._c**2) * sym.y_laplacian[0]]
        )


class CahnHilliardEquation(DifferentialEquation):
    def __init__(self, x_dimension: int, d: float = 0.1, gamma: float = 0.01):
        _require_pde(x_dimension)
        self._d = d
        self._gamma = gamma
        super().__init__(x_dimension, 2)

    @property
    def symbolic_equation_system(self):
        sym = self._symbols
        return SymbolicEquationSystem(
            [
                self._d * sym.y_laplacian[1],
                sym.y[0] ** 3 - sym.y[0] - self._gamma * sym.y_laplacian[0],
            ],
            [LHS.D_Y_OVER_D_T, LHS.Y],
        )


class BurgersEquation(DifferentialEquation):
    def __init__(self, x_dimension: int, re: float = 4000.0):
        _require_pde(x_dimension)
        self._re = re
        super().__init__(x_dimension, x_dimension, [tuple(range(x_dimension))])

    @property
    def symbolic_equation_system(self):
        sym = self._symbols
        return SymbolicEquationSystem(
            [
                (1.0 / self._re) * sym.y_laplacian[i]
                - np.dot(sym.y, sym.y_gradient[i, :])
                for i in range(self._x_dimension)
            ]
        )


class ShallowWaterEquation(DifferentialEquation):
    """y = (height perturbation, velocity_0, velocity_1), 2 spatial dims."""

    def __init__(self, h: float, b=0.01, v=0.1, f=0.0, g=9.80665):
        self._h, self._b, self._v, self._f, self._g = h, b, v, f, g
        super().__init__(2, 3, [(1, 2)])

    @property
    def symbolic_equation_system(self):
        s = self._symbols
        y, grad, lap = s.y, s.y_gradient, s.y_laplacian
        return SymbolicEquationSystem(
            [
                -self._h * s.y_divergence[1, 2]
                - y[0] * grad[1, 0]
                - y[1] * grad[0, 0]
                - y[0] * grad[2, 1]
                - y[2] * grad[0, 1],
                self._v * lap[1]
                - y[1] * grad[1, 0]
                - y[2] * grad[1, 1]
                - self._g * grad[0, 0]
                - self._b * y[1]
                + self._f * y[2],
                self._v * lap[2]
                - y[1] * grad[2, 0]
                - y[2] * grad[2, 1]
                - self._g * grad[0, 1]
                - self._b * y[2]
                - self._f * y[1],
            ]
        )


class NavierStokesEquation(DifferentialEquation):
    """2D vorticity / stream function / velocity formulation."""

    def __init__(self, re: float = 4000.0):
        self._re = re
        super().__init__(2, 4, [(2, 3)])

    @property
    def symbolic_equation_system(self):
        s = self._symbols
        return SymbolicEquationSystem(
            [
                (1.0 / self._re) * s.y_laplacian[0]
                - np.dot(s.y[2:], s.y_gradient[0, :]),
                -s.y[0],
                s.y_gradient[1, 1],
                -s.y_gradient[1, 0],
            ],
            [LHS.D_Y_OVER_D_T, LHS.Y_LAPLACIAN, LHS.Y, LHS.Y],
        )
