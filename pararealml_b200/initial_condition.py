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
from scipy.linalg import eigh as scipy_eigh
from scipy.stats import beta, multivariate_normal

from pararealml_b200.constraint import apply_constraints_along_last_axis
from pararealml_b200.mesh import to_cartesian_coordinates

def torch_empty(n: int, device):
    import torch

    return torch.empty(n, dtype=torch.float64, device=device)


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
        # discretised on first use (the reference does it in the constructor,
        # initial_condition.py:212-229; the values are the same): a large mesh
        # whose state is evaluated on the device never pays for the host arrays
        self._discrete = {}

    def y_0(self, x):
        return np.multiply(self._y_0_func(x), self._multipliers)

    def _discrete_y_0(self, vertex_oriented) -> np.ndarray:
        key = bool(vertex_oriented)
        if key not in self._discrete:
            self._discrete[key] = self._discretise(key)
        return self._discrete[key]

    def discrete_y_0_view(self, vertex_oriented=None):
        return self._discrete_y_0(vertex_oriented)

    def discrete_y_0(self, vertex_oriented=None):
        return np.copy(self._discrete_y_0(vertex_oriented))

    def discrete_y_0_planes(self, plan):
        """Vertex-oriented initial state as component planes on the device of
        ``plan`` (a ``DevicePlan`` of this problem, whose static Dirichlet
        tables are applied), evaluated by a CUDA kernel -- or None when this
        initial condition has no device form (arbitrary Python callables)."""
        return None

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

    def discrete_y_0_planes(self, plan):
        """The densities evaluated by ``pml_ic_gaussian`` with SciPy's
        formula: exp(-0.5 (rank log(2 pi) + log pdet(cov) + |(x - mean) U|^2))
        where U = eigenvectors * sqrt(1 / eigenvalues) of the covariance."""
        from pararealml_b200 import _native
        from pararealml_b200.operators.fdm import device as dv

        y_dim = len(self._means_and_covs)
        if y_dim > _native.IC_MAX_COMPONENTS:
            return None
        d = self._cp.differential_equation.x_dimension
        params = (_native.IcGaussianParams * y_dim)()
        for c, (mean, cov) in enumerate(self._means_and_covs):
            s, u = scipy_eigh(np.asarray(cov, dtype=float), lower=True)
            eps = 1e6 * np.finfo(s.dtype).eps * np.max(np.abs(s))
            if np.min(s) < -eps or np.any(np.abs(s) <= eps):
                return None  # singular covariance: left to SciPy on the host
            whiten = np.multiply(u, np.sqrt(1.0 / s))
            for j in range(d):
                params[c].mean[j] = float(mean[j])
                for k in range(d):
                    params[c].whiten[j * d + k] = float(whiten[j, k])
            params[c].log_norm = float(d * np.log(2.0 * np.pi) + np.sum(np.log(s)))
            params[c].multiplier = float(self._multipliers[c])
        mesh, keep = dv.ic_mesh(self._cp.mesh, plan.device)
        planes = torch_empty(y_dim * plan.n_cells, plan.device)
        _native.check(
            _native.lib().pml_ic_gaussian(
                mesh, y_dim, params, planes.data_ptr(), dv.stream_ptr()
            )
        )
        dv.count_launch()
        plan.apply_static_dirichlet(planes)
        del keep
        return planes


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

    def discrete_y_0_planes(self, plan):
        """Cartesian meshes: the 1-D Beta densities of every axis are
        evaluated on the host (n values each, the same SciPy calls as
        ``_pdf``) and multiplied on the device (``pml_ic_separable``) in the
        order ``np.prod`` multiplies them -- bit-identical to the host
        discretisation."""
        import ctypes

        from pararealml_b200 import _native
        from pararealml_b200.mesh import CoordinateSystem
        from pararealml_b200.operators.fdm import device as dv

        mesh_obj = self._cp.mesh
        y_dim = len(self._all_alphas_and_betas)
        if (mesh_obj.coordinate_system_type != CoordinateSystem.CARTESIAN
                or y_dim > _native.IC_MAX_COMPONENTS):
            return None
        d = mesh_obj.dimensions
        axes = mesh_obj.vertex_axis_coordinates
        keep, ptrs = [], (ctypes.c_void_p * (y_dim * d))()
        for c, params in enumerate(self._all_alphas_and_betas):
            for k, (a, b) in enumerate(params):
                vec = dv.to_device(beta.pdf(axes[k], a, b), plan.device)
                keep.append(vec)
                ptrs[c * d + k] = vec.data_ptr()
        mult = (ctypes.c_double * y_dim)(*[float(m) for m in self._multipliers])
        mesh, keep_mesh = dv.ic_mesh(mesh_obj, plan.device)
        planes = torch_empty(y_dim * plan.n_cells, plan.device)
        _native.check(
            _native.lib().pml_ic_separable(
                mesh, y_dim, ptrs, mult, planes.data_ptr(), dv.stream_ptr()
            )
        )
        dv.count_launch()
        plan.apply_static_dirichlet(planes)
        del keep, keep_mesh
        return planes


def vectorize_ic_function(
    ic_function: Callable[[Optional[Sequence[float]]], Sequence[float]]
) -> VectorizedInitialConditionFunction:
    def vectorized(x: Optional[np.ndarray]) -> np.ndarray:
        if x is None:
            return np.array(ic_function(None))
        return np.array([ic_function(x[i]) for i in range(len(x))])

    return vectorized
