"""Initial conditions (host side; produce the ``y_0`` array once).

API mirror of the reference's ``pararealml/initial_condition.py``.  Static
Dirichlet values are written into vertex-oriented discretisations straight
from the NaN-coded face tables (O(surface)) instead of through full-grid
masks; the result is identical (reference :86-89, :226-229).
"""
from abc import ABC, abstractmethod
from copy import deepcopy
from typing import Callable, Optional, Sequence, Tuple

import numpy as np
from scipy.interpolate import interpn
from scipy.stats import beta, multivariate_normal

from pararealml_b200.constraint import apply_constraints_along_last_axis
from pararealml_b200.mesh import to_cartesian_coordinates

VectorizedInitialConditionFunction = Callable[
    [Optional[np.ndarray]], np.ndarray
]


def _apply_static_dirichlet(cp, y: np.ndarray) -> np.ndarray:
    if not cp.differential_equation.x_dimension:
        return y
    if hasattr(cp, "apply_dirichlet_tables"):
        return cp.apply_dirichlet_tables(y, cp.dirichlet_face_tables(None))
    return apply_constraints_along_last_axis(
        cp.static_y_vertex_constraints, y
    )


class InitialCondition(ABC):
    @abstractmethod
    def y_0(self, x: Optional[np.ndarray]) -> np.ndarray:
        """Initial values at the points ``x`` ((n, x_dim); None for ODEs)."""

    @abstractmethod
    def discrete_y_0(
        self, vertex_oriented: Optional[bool] = None
    ) -> np.ndarray:
        """Initial values on the mesh vertices or cell centres (a copy)."""


class DiscreteInitialCondition(InitialCondition):
    def __init__(
        self,
        cp,
        y_0: np.ndarray,
        vertex_oriented: Optional[bool] = None,
        interpolation_method: str = "linear",
    ):
        if cp.differential_equation.x_dimension and vertex_oriented is None:
            raise ValueError("vertex orientation is required for PDEs")
        if y_0.shape != cp.y_shape(vertex_oriented):
            raise ValueError(
                f"y_0 shape {y_0.shape} != problem shape "
                f"{cp.y_shape(vertex_oriented)}"
            )
        self._cp = cp
        self._y_0 = np.array(y_0, dtype=float, copy=True)
        self._vertex_oriented = vertex_oriented
        self._interpolation_method = interpolation_method
        if vertex_oriented:
            _apply_static_dirichlet(cp, self._y_0)

    def y_0(self, x):
        if not self._cp.differential_equation.x_dimension:
            return np.copy(self._y_0)
        return interpn(
            self._cp.mesh.axis_coordinates(self._vertex_oriented),
            self._y_0,
            x,
            method=self._interpolation_method,
            bounds_error=False,
            fill_value=None,
        )

    def discrete_y_0_view(self, vertex_oriented=None):
        """The stored array itself (no copy; callers must not write to it) or
        None if it would have to be interpolated.  Lets the B200 operators
        upload a multi-GB initial state without an extra host copy."""
        if vertex_oriented is None:
            vertex_oriented = self._vertex_oriented
        if (
            not self._cp.differential_equation.x_dimension
            or vertex_oriented == self._vertex_oriented
        ):
            return self._y_0
        return None

    def discrete_y_0(self, vertex_oriented=None):
        if vertex_oriented is None:
            vertex_oriented = self._vertex_oriented
        if (
            not self._cp.differential_equation.x_dimension
            or vertex_oriented == self._vertex_oriented
        ):
            return np.copy(self._y_0)
        y = self.y_0(self._cp.mesh.all_index_coordinates(vertex_oriented))
        if vertex_oriented:
            _apply_static_dirichlet(self._cp, y)
        return y


class ConstantInitialCondition(DiscreteInitialCondition):
    def __init__(self, cp, constant_y_0s: Sequence[float]):
        y_dim = cp.differential_equation.y_dimension
        if len(constant_y_0s) != y_dim:
            raise ValueError(
                f"{len(constant_y_0s)} constants for {y_dim} components"
            )
        y = np.empty(cp.y_shape(True))
        y[...] = np.asarray(constant_y_0s, dtype=float)
        super().__init__(cp, y, True)


class ContinuousInitialCondition(InitialCondition):
    def __init__(
        self,
        cp,
        y_0_func: VectorizedInitialConditionFunction,
        multipliers: Optional[Sequence[float]] = None,
    ):
        y_dim = cp.differential_equation.y_dimension
        if multipliers is None:
            self._multipliers = np.ones(y_dim)
        else:
            if len(multipliers) != y_dim:
                raise ValueError(
                    f"{len(multipliers)} multipliers for {y_dim} components"
                )
            self._multipliers = np.array(multipliers)
        self._cp = cp
        self._y_0_func = y_0_func
        self._discrete = {
            True: self._discretise(True),
            False: self._discretise(False),
        }

    def y_0(self, x):
        return np.multiply(self._y_0_func(x), self._multipliers)

    def discrete_y_0_view(self, vertex_oriented=None):
        return self._discrete[bool(vertex_oriented)]

    def discrete_y_0(self, vertex_oriented=None):
        return np.copy(self._discrete[bool(vertex_oriented)])

    def _discretise(self, vertex_oriented: bool) -> np.ndarray:
        cp = self._cp
        eq = cp.differential_equation
        if not eq.x_dimension:
            y = np.array(self.y_0(None))
            if y.shape != cp.y_shape():
                raise ValueError(
                    f"initial condition returned shape {y.shape}, expected "
                    f"{cp.y_shape()}"
                )
            return y
        x = cp.mesh.all_index_coordinates(vertex_oriented, flatten=True)
        y = self.y_0(x)
        if y.shape != (len(x), eq.y_dimension):
            raise ValueError(
                f"initial condition returned shape {y.shape}, expected "
                f"{(len(x), eq.y_dimension)}"
            )
        y = np.ascontiguousarray(y.reshape(cp.y_shape(vertex_oriented)))
        if vertex_oriented:
            _apply_static_dirichlet(cp, y)
        return y

    def _to_cartesian(self, x: np.ndarray) -> np.ndarray:
        parts = to_cartesian_coordinates(
            [x[:, i] for i in range(x.shape[1])],
            self._cp.mesh.coordinate_system_type,
        )
        return np.stack(parts, axis=-1)


class GaussianInitialCondition(ContinuousInitialCondition):
    """Each component is a multivariate normal PDF over Cartesian space."""

    def __init__(
        self,
        cp,
        means_and_covs: Sequence[Tuple[np.ndarray, np.ndarray]],
        multipliers: Optional[Sequence[float]] = None,
    ):
        eq = cp.differential_equation
        if not eq.x_dimension:
            raise ValueError("Gaussian initial conditions need a PDE")
        if len(means_and_covs) != eq.y_dimension:
            raise ValueError(
                f"{len(means_and_covs)} (mean, cov) pairs for "
                f"{eq.y_dimension} components"
            )
        for mean, cov in means_and_covs:
            if mean.shape != (eq.x_dimension,):
                raise ValueError(f"bad mean shape {mean.shape}")
            if cov.shape != (eq.x_dimension, eq.x_dimension):
                raise ValueError(f"bad covariance shape {cov.shape}")
        self._means_and_covs = deepcopy(means_and_covs)
        super().__init__(cp, self._pdf, multipliers)

    def _pdf(self, x):
        xc = self._to_cartesian(x)
        out = np.empty((len(x), len(self._means_and_covs)))
        for i, (mean, cov) in enumerate(self._means_and_covs):
            out[:, i] = multivariate_normal.pdf(xc, mean=mean, cov=cov)
        return out


class MarginalBetaProductInitialCondition(ContinuousInitialCondition):
    """Each component is a product of per-axis Beta PDFs."""

    def __init__(
        self,
        cp,
        all_alphas_and_betas: Sequence[Sequence[Tuple[float, float]]],
        multipliers: Optional[Sequence[float]] = None,
    ):
        eq = cp.differential_equation
        if len(all_alphas_and_betas) != eq.y_dimension:
            raise ValueError(
                f"{len(all_alphas_and_betas)} parameter sequences for "
                f"{eq.y_dimension} components"
            )
        if any(len(ab) != eq.x_dimension for ab in all_alphas_and_betas):
            raise ValueError(
                f"every parameter sequence needs {eq.x_dimension} entries"
            )
        self._all_alphas_and_betas = deepcopy(all_alphas_and_betas)
        super().__init__(cp, self._pdf, multipliers)

    def _pdf(self, x):
        xc = self._to_cartesian(x)
        cols = []
        for params in self._all_alphas_and_betas:
            col = np.prod(
                [
                    beta.pdf(xc[:, k : k + 1], a, b)
                    for k, (a, b) in enumerate(params)
                ],
                axis=0,
            )
            cols.append(col)
        return np.concatenate(cols, axis=-1)


def vectorize_ic_function(
    ic_function: Callable[[Optional[Sequence[float]]], Sequence[float]]
) -> VectorizedInitialConditionFunction:
    def vectorized(x: Optional[np.ndarray]) -> np.ndarray:
        if x is None:
            return np.array(ic_function(None))
        return np.array([ic_function(x[i]) for i in range(len(x))])

    return vectorized
