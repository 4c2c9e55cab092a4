"""Three-point central difference differentiator on B200.

Mirrors the public surface of the reference's
``pararealml/operators/fdm/numerical_differentiator.py``:
``ThreePointCentralDifferenceMethod(tol)`` with the NumPy-in / NumPy-out
methods ``gradient``, ``hessian``, ``divergence``, ``curl``, ``laplacian``,
``vector_laplacian`` (:114-870) and ``anti_laplacian`` (:872-927), including
their argument validation (:929-996).  Each call lowers the request to a small
synthetic equation system whose right-hand sides are the requested leaves,
generates the CUDA kernels for it (same template and primitives the fused
stage kernels use) and evaluates it on the device.  Inside ``FDMOperator`` the
differentiator is not called per symbol at all: its stencils are fused into
the stage kernels.
"""
from abc import ABC
from typing import Optional, Sequence, Union

import numpy as np
import sympy as sp
import torch

from pararealml_b200.constraint import Constraint, to_nan_table
from pararealml_b200.mesh import Mesh
from pararealml_b200.operators.fdm import device as dv
from pararealml_b200.operators.fdm.lowering import LoweredProblem


class NumericalDifferentiator(ABC):
    def __init__(self, tol: float = 1e-3):
        if tol < 0.0:
            raise ValueError("tolerance must be non-negative")
        self._tol = tol


def _check_shape(array: np.ndarray, mesh: Mesh, name: str = "y"):
    if array.shape[:-1] != mesh.vertices_shape:
        raise ValueError(
            f"{name} shape up to second to last axis {array.shape[:-1]} must "
            f"match mesh vertices shape {mesh.vertices_shape}"
        )


def _check_vector_field(array: np.ndarray, mesh: Mesh):
    _check_shape(array, mesh)
    if array.shape[-1] != mesh.dimensions:
        raise ValueError(
            f"y value vector length ({array.shape[-1]}) must match number of "
            f"x dimensions ({mesh.dimensions})"
        )


def _check_dbc(dbc: Optional[np.ndarray], x_dim: int, k: int) -> np.ndarray:
    if dbc is None:
        return np.empty((x_dim, k), dtype=object)
    if dbc.shape != (x_dim, k):
        raise ValueError(
            f"expected derivative boundary constraints shape to be "
            f"{(x_dim, k)} but got {dbc.shape}"
        )
    return dbc


def _lower_request(mesh: Mesh, k: int, rhs, kinds, dbc, y_constraints=None):
    """Synthetic lowered problem for a direct differentiator call."""
    shape = tuple(int(n) for n in mesh.vertices_shape)
    if len(shape) > 3:
        raise NotImplementedError(
            "the B200 FDM kernels support at most 3 spatial dimensions"
        )
    coord = mesh.coordinate_system_type.name
    low = LoweredProblem(
        shape=shape,
        d_x=tuple(float(h) for h in mesh.d_x),
        coord=coord,
        y_dim=k,
        rhs=list(rhs),
        kinds=list(kinds),
    )
    low.coords = [
        np.ascontiguousarray(a, dtype=np.float64)
        for a in mesh.vertex_axis_coordinates
    ]
    aux = [None] * 4
    if coord != "CARTESIAN":
        aux[0] = 1.0 / low.coords[0]
        if coord == "SPHERICAL":
            aux[1] = np.sin(low.coords[2])
            aux[2] = np.cos(low.coords[2])
            aux[3] = 1.0 / np.sin(low.coords[2])
    low.aux = aux
    for axis in range(len(shape)):
        face_shape = tuple(1 if a == axis else n for a, n in enumerate(shape))
        for side in range(2):
            f = axis * 2 + side
            low.face_static[f] = True
            low.face_cells[f] = int(np.prod(face_shape))
            cs = [
                None if dbc[axis, i] is None else dbc[axis, i][side]
                for i in range(k)
            ]
            if any(c is not None for c in cs):
                tab = np.stack(
                    [to_nan_table(c, face_shape + (1,))[..., 0] for c in cs],
                    axis=-1,
                )
                low.neu_mask |= 1 << f
                low.static_neu[f] = np.ascontiguousarray(tab).reshape(-1)
            if y_constraints is not None:
                tabs = []
                for c in y_constraints:
                    full = to_nan_table(c, shape + (1,))[..., 0]
                    idx = [slice(None)] * len(shape)
                    idx[axis] = slice(-1, None) if side else slice(0, 1)
                    tabs.append(full[tuple(idx)])
                tab = np.stack(tabs, axis=-1)
                if not np.all(np.isnan(tab)):
                    low.dir_mask |= 1 << f
                    low.static_dir[f] = np.ascontiguousarray(tab).reshape(-1)
    return low


class ThreePointCentralDifferenceMethod(NumericalDifferentiator):
    """Second order central differences with the reference's boundary rules:
    zero ghost cells for first derivatives, Neumann-extrapolated ghosts for
    second differences (numerical_differentiator.py:1012-1095, 1188-1242)."""

    def __init__(self, tol: float = 1e-3):
        super().__init__(tol)

    # -- evaluation of leaf symbols on the device --------------------------
    def _evaluate(self, y: np.ndarray, mesh: Mesh, dbc, leaf_names):
        if any(n <= 2 for n in y.shape[:-1]):
            bad = [a for a, n in enumerate(y.shape[:-1]) if n <= 2][0]
            raise ValueError(
                f"y must contain at least 3 points along x-axis ({bad})"
            )
        k = y.shape[-1]
        m = len(leaf_names)
        # the synthetic system has one equation per component: pad with zeros
        rhs = [sp.Symbol(name) for name in leaf_names]
        rhs += [sp.Integer(0)] * (k - m)
        kinds = ["D_Y_OVER_D_T"] * k
        low = _lower_request(mesh, k, rhs, kinds, dbc)
        plan = dv.get_plan(low)
        plan.bind_tables(low)
        u = dv.upload_state(y, low.n_cells, k)
        out = torch.empty(
            k * low.n_cells, dtype=torch.float64, device=plan.device
        )
        plan.eval_rhs(u, out)
        res = dv.soa_to_aos(out[: m * low.n_cells], low.n_cells, m)
        return res.cpu().numpy().reshape(y.shape[:-1] + (m,))

    def gradient(self, y, mesh, x_axis, derivative_boundary_constraints=None):
        _check_shape(y, mesh)
        if not 0 <= x_axis < mesh.dimensions:
            raise ValueError(
                f"x-axis ({x_axis}) must be non-negative and less than number "
                f"of x dimensions ({mesh.dimensions})"
            )
        k = y.shape[-1]
        dbc = _check_dbc(derivative_boundary_constraints, mesh.dimensions, k)
        return self._evaluate(
            y, mesh, dbc, [f"y-gradient_{i}_{x_axis}" for i in range(k)]
        )

    def hessian(
        self, y, mesh, x_axis1, x_axis2, derivative_boundary_constraints=None
    ):
        _check_shape(y, mesh)
        if not (
            0 <= x_axis1 < mesh.dimensions and 0 <= x_axis2 < mesh.dimensions
        ):
            raise ValueError(
                f"both first x-axis ({x_axis1}) and second x-axis ({x_axis2}) "
                "must be non-negative and less than number of x dimensions "
                f"({mesh.dimensions})"
            )
        k = y.shape[-1]
        dbc = _check_dbc(derivative_boundary_constraints, mesh.dimensions, k)
        return self._evaluate(
            y, mesh, dbc,
            [f"y-hessian_{i}_{x_axis1}_{x_axis2}" for i in range(k)],
        )

    def divergence(self, y, mesh, derivative_boundary_constraints=None):
        _check_vector_field(y, mesh)
        k = y.shape[-1]
        dbc = _check_dbc(derivative_boundary_constraints, mesh.dimensions, k)
        name = "y-divergence_" + "_".join(str(i) for i in range(k))
        return self._evaluate(y, mesh, dbc, [name])

    def curl(self, y, mesh, curl_ind=0, derivative_boundary_constraints=None):
        _check_vector_field(y, mesh)
        if not 2 <= mesh.dimensions <= 3:
            raise ValueError(
                f"number of x dimensions ({mesh.dimensions}) must be 2 or 3"
            )
        if mesh.dimensions == 2 and curl_ind != 0:
            raise ValueError(f"curl index ({curl_ind}) must be 0 for 2D curl")
        if not 0 <= curl_ind < mesh.dimensions:
            raise ValueError(
                f"curl index ({curl_ind}) must be non-negative and less than "
                f"number of x dimensions ({mesh.dimensions})"
            )
        k = y.shape[-1]
        dbc = _check_dbc(derivative_boundary_constraints, mesh.dimensions, k)
        name = "y-curl_" + "_".join(str(i) for i in range(k))
        if mesh.dimensions == 3:
            name += f"_{curl_ind}"
        return self._evaluate(y, mesh, dbc, [name])

    def laplacian(self, y, mesh, derivative_boundary_constraints=None):
        _check_shape(y, mesh)
        k = y.shape[-1]
        dbc = _check_dbc(derivative_boundary_constraints, mesh.dimensions, k)
        return self._evaluate(
            y, mesh, dbc, [f"y-laplacian_{i}" for i in range(k)]
        )

    def vector_laplacian(
        self, y, mesh, vector_laplacian_ind,
        derivative_boundary_constraints=None,
    ):
        _check_vector_field(y, mesh)
        if not 0 <= vector_laplacian_ind < mesh.dimensions:
            raise ValueError(
                f"vector Laplacian index ({vector_laplacian_ind}) must be "
                "non-negative and less than number of x dimensions "
                f"({mesh.dimensions})"
            )
        k = y.shape[-1]
        dbc = _check_dbc(derivative_boundary_constraints, mesh.dimensions, k)
        name = (
            "y-vector-laplacian_"
            + "_".join(str(i) for i in range(k))
            + f"_{vector_laplacian_ind}"
        )
        return self._evaluate(y, mesh, dbc, [name])

    # -- Jacobi anti-Laplacian --------------------------------------------------
    def anti_laplacian(
        self,
        laplacian: np.ndarray,
        mesh: Mesh,
        y_constraints: Union[Sequence[Optional[Constraint]], np.ndarray],
        derivative_boundary_constraints: Optional[np.ndarray] = None,
        y_init: Optional[np.ndarray] = None,
        max_sweeps: int = 0,
    ) -> np.ndarray:
        """Jacobi solve of ``laplacian(y) = rhs`` on the device; sweeps until
        the 2-norm of the update drops to ``tol`` (reference :872-927).
        ``y_constraints`` must only constrain boundary vertices."""
        _check_shape(laplacian, mesh, "Laplacian")
        k = laplacian.shape[-1]
        dbc = _check_dbc(derivative_boundary_constraints, mesh.dimensions, k)
        if y_init is None:
            y0 = np.random.random(laplacian.shape)
        else:
            if y_init.shape != laplacian.shape:
                raise ValueError
            y0 = y_init
        if not np.all(np.array(laplacian.shape[:-1]) > 2):
            raise ValueError(
                "y must contain at least 3 points along all x axes"
            )
        constraints = (
            [None] * k if y_constraints is None else list(y_constraints)
        )
        interior = tuple([slice(1, -1)] * mesh.dimensions)
        for c in constraints:
            if c is not None and np.any(
                c.mask.reshape(laplacian.shape[:-1])[interior]
            ):
                raise NotImplementedError(
                    "the B200 Jacobi kernel supports value constraints on "
                    "boundary vertices only"
                )
        low = _lower_request(
            mesh, k, [sp.Integer(0)] * k, ["Y_LAPLACIAN"] * k, dbc, constraints
        )
        plan = dv.get_plan(low)
        plan.bind_tables(low)
        rhs = dv.upload_state(laplacian, low.n_cells, k)
        init = torch.from_numpy(np.ascontiguousarray(y0)).to(plan.device)
        out = torch.empty(k * low.n_cells, dtype=torch.float64, device=plan.device)
        self.last_sweeps = plan.jacobi(rhs, init, out, self._tol, max_sweeps)
        res = dv.soa_to_aos(out, low.n_cells, k)
        return res.cpu().numpy().reshape(laplacian.shape)
