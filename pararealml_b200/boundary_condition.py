"""Boundary conditions (host side; user callables evaluated on the CPU).

API mirror of the reference's ``pararealml/boundary_condition.py``.  A
condition function maps ``(x[n, x_dim], t) -> [n, y_dim]`` with NaN meaning
"this component is not constrained here".
"""
from abc import ABC, abstractmethod
from typing import Callable, Optional, Sequence

import numpy as np

VectorizedBoundaryConditionFunction = Callable[
    [np.ndarray, Optional[float]], np.ndarray
]


class BoundaryCondition(ABC):
    def __init__(
        self, has_y_condition: bool, has_d_y_condition: bool, is_static: bool
    ):
        self._has_y = has_y_condition
        self._has_d_y = has_d_y_condition
        self._static = is_static

    @property
    def has_y_condition(self) -> bool:
        return self._has_y

    @property
    def has_d_y_condition(self) -> bool:
        return self._has_d_y

    @property
    def is_static(self) -> bool:
        return self._static

    @abstractmethod
    def y_condition(self, x: np.ndarray, t: Optional[float]) -> np.ndarray:
        """Values of y on the boundary points ``x``."""

    @abstractmethod
    def d_y_condition(self, x: np.ndarray, t: Optional[float]) -> np.ndarray:
        """Values of the derivative of y along the boundary's axis."""


class DirichletBoundaryCondition(BoundaryCondition):
    def __init__(
        self,
        y_condition: VectorizedBoundaryConditionFunction,
        is_static: bool = False,
    ):
        super().__init__(True, False, is_static)
        self._y_fn = y_condition

    def y_condition(self, x, t):
        return self._y_fn(x, t)

    def d_y_condition(self, x, t):
        raise RuntimeError("a Dirichlet condition has no derivative part")


class NeumannBoundaryCondition(BoundaryCondition):
    def __init__(
        self,
        d_y_condition: VectorizedBoundaryConditionFunction,
        is_static: bool = False,
    ):
        super().__init__(False, True, is_static)
        self._d_y_fn = d_y_condition

    def y_condition(self, x, t):
        raise RuntimeError("a Neumann condition has no value part")

    def d_y_condition(self, x, t):
        return self._d_y_fn(x, t)


class CauchyBoundaryCondition(BoundaryCondition):
    def __init__(
        self,
        y_condition: VectorizedBoundaryConditionFunction,
        d_y_condition: VectorizedBoundaryConditionFunction,
        is_static: bool = False,
    ):
        super().__init__(True, True, is_static)
        self._y_fn = y_condition
        self._d_y_fn = d_y_condition

    def y_condition(self, x, t):
        return self._y_fn(x, t)

    def d_y_condition(self, x, t):
        return self._d_y_fn(x, t)


class ConstantBoundaryCondition(BoundaryCondition):
    """Space and time independent per-component constants (None = free)."""

    def __init__(
        self,
        constant_y_conditions: Optional[Sequence[Optional[float]]],
        constant_d_y_conditions: Optional[Sequence[Optional[float]]],
    ):
        if constant_y_conditions is None and constant_d_y_conditions is None:
            raise ValueError("both constant condition sequences are None")
        super().__init__(
            constant_y_conditions is not None,
            constant_d_y_conditions is not None,
            True,
        )
        self._y_consts = constant_y_conditions
        self._d_y_consts = constant_d_y_conditions

    @staticmethod
    def _tile(consts, n):
        row = np.array(
            [np.nan if c is None else c for c in consts], dtype=float
        )
        return np.tile(row, (n, 1))

    def y_condition(self, x, t):
        if not self._y_consts:
            raise RuntimeError("no constant condition on y")
        return self._tile(self._y_consts, len(x))

    def d_y_condition(self, x, t):
        if not self._d_y_consts:
            raise RuntimeError("no constant condition on the derivative of y")
        return self._tile(self._d_y_consts, len(x))


class ConstantValueBoundaryCondition(ConstantBoundaryCondition):
    def __init__(self, constant_y_conditions: Sequence[Optional[float]]):
        super().__init__(constant_y_conditions, None)


class ConstantFluxBoundaryCondition(ConstantBoundaryCondition):
    def __init__(self, constant_d_y_conditions: Sequence[Optional[float]]):
        super().__init__(None, constant_d_y_conditions)


def vectorize_bc_function(
    bc_function: Callable[
        [Sequence[float], Optional[float]], Sequence[Optional[float]]
    ]
) -> VectorizedBoundaryConditionFunction:
    """Row-by-row wrapper; ``None`` entries become NaN."""

    def vectorized(x: np.ndarray, t: Optional[float]) -> np.ndarray:
        return np.array(
            [bc_function(x[i], t) for i in range(len(x))], dtype=float
        )

    return vectorized
