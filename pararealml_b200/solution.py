"""Result container of ``Operator.solve`` (host side).

API mirror of the reference's ``pararealml/solution.py`` (:25-292) without the
matplotlib plots (out of scope).  ``discrete_y`` may be
backed lazily by a callable so that a device-resident trajectory is only
copied to the host when somebody reads it.
"""
from typing import Callable, List, NamedTuple, Optional, Sequence, Union

import numpy as np
from scipy.interpolate import interpn

from pararealml_b200.constraint import apply_constraints_along_last_axis


class Diffs(NamedTuple):
    matching_time_points: np.ndarray
    differences: Sequence[np.ndarray]


class Solution:
    def __init__(
        self,
        ivp,
        t_coordinates: np.ndarray,
        discrete_y: Union[np.ndarray, Callable[[], np.ndarray]],
        vertex_oriented: Optional[bool] = None,
        d_t: Optional[float] = None,
        copy: bool = True,
    ):
        if t_coordinates.ndim != 1:
            raise ValueError("t coordinates must be one dimensional")
        if len(t_coordinates) == 0:
            raise ValueError("t coordinates must not be empty")
        cp = ivp.constrained_problem
        if cp.differential_equation.x_dimension and vertex_oriented is None:
            raise ValueError("vertex orientation is required for PDEs")
        self._expected_shape = (len(t_coordinates),) + cp.y_shape(
            vertex_oriented
        )
        self._ivp = ivp
        self._t = np.copy(t_coordinates)
        self._t.setflags(write=False)
        self._vertex_oriented = vertex_oriented
        if callable(discrete_y):
            self._y = None
            self._y_source = discrete_y
        else:
            self._check_shape(discrete_y)
            self._y = np.copy(discrete_y) if copy else discrete_y
            self._y_source = None
        if d_t is None:
            d_t = 0.0 if len(self._t) == 1 else self._t[1] - self._t[0]
        self._d_t = d_t

    def _check_shape(self, y):
        if y.shape != self._expected_shape:
            raise ValueError(
                f"solution shape {y.shape} != expected {self._expected_shape}"
            )

    def _materialise(self) -> np.ndarray:
        if self._y is None:
            y = self._y_source()
            self._check_shape(y)
            self._y = y
            self._y_source = None
        return self._y

    @property
    def initial_value_problem(self):
        return self._ivp

    @property
    def vertex_oriented(self) -> Optional[bool]:
        return self._vertex_oriented

    @property
    def d_t(self) -> float:
        return self._d_t

    @property
    def t_coordinates(self) -> np.ndarray:
        return self._t

    def y(self, x: Optional[np.ndarray] = None, interpolation_method="linear"):
        cp = self._ivp.constrained_problem
        eq = cp.differential_equation
        data = self._materialise()
        if not eq.x_dimension:
            return np.copy(data)
        vals = interpn(
            cp.mesh.axis_coordinates(self._vertex_oriented),
            np.moveaxis(data, 0, -2),
            x,
            method=interpolation_method,
            bounds_error=False,
            fill_value=None,
        )
        vals = np.moveaxis(vals, -2, 0).reshape(
            (len(self._t),) + x.shape[:-1] + (eq.y_dimension,)
        )
        return np.ascontiguousarray(vals)

    def discrete_y(
        self, vertex_oriented: Optional[bool] = None, interpolation_method="linear"
    ) -> np.ndarray:
        if vertex_oriented is None:
            vertex_oriented = self._vertex_oriented
        cp = self._ivp.constrained_problem
        if (
            not cp.differential_equation.x_dimension
            or self._vertex_oriented == vertex_oriented
        ):
            return np.copy(self._materialise())
        y = self.y(
            cp.mesh.all_index_coordinates(vertex_oriented),
            interpolation_method,
        )
        if vertex_oriented:
            apply_constraints_along_last_axis(
                cp.static_y_vertex_constraints, y
            )
        return y

    def diff(self, solutions: Sequence["Solution"], atol: float = 1e-8):
        """Differences to other solutions at the time points all share."""
        if len(solutions) == 0:
            raise ValueError("at least one solution to compare to is needed")
        on_device = self._device_diff(solutions, atol)
        if on_device is not None:
            return on_device
        mine = self._materialise()
        others = [s.discrete_y(self._vertex_oriented) for s in solutions]
        grids = [self._t] + [s.t_coordinates for s in solutions]
        steps = [self._d_t] + [s.d_t for s in solutions]
        shortest = int(np.argmin([len(g) for g in grids]))
        matched: List[float] = []
        diffs: List[List[np.ndarray]] = [[] for _ in solutions]
        for i, t in enumerate(grids[shortest]):
            where = []
            for j, g in enumerate(grids):
                if j == shortest:
                    where.append(i)
                    continue
                k = int(round((t - g[0]) / steps[j]))
                if 0 <= k < len(g) and np.isclose(t, g[k], atol=atol, rtol=0.0):
                    where.append(k)
                else:
                    break
            if len(where) != len(grids):
                continue
            matched.append(t)
            for j, other in enumerate(others):
                diffs[j].append(other[where[j + 1]] - mine[where[0]])
        return Diffs(np.array(matched), [np.array(d) for d in diffs])

    def _matching_indices(self, solutions, atol):
        """(matched times, per solution (self first) the index of every
        matched time point) -- the time-point matching rule of ``diff``."""
        grids = [self._t] + [s.t_coordinates for s in solutions]
        steps = [self._d_t] + [s.d_t for s in solutions]
        shortest = int(np.argmin([len(g) for g in grids]))
        matched: List[float] = []
        indices: List[List[int]] = [[] for _ in grids]
        for i, t in enumerate(grids[shortest]):
            where = []
            for j, g in enumerate(grids):
                if j == shortest:
                    where.append(i)
                    continue
                k = int(round((t - g[0]) / steps[j]))
                if 0 <= k < len(g) and np.isclose(t, g[k], atol=atol, rtol=0.0):
                    where.append(k)
                else:
                    break
            if len(where) != len(grids):
                continue
            matched.append(t)
            for j, k in enumerate(where):
                indices[j].append(k)
        return np.array(matched), indices

    def _device_diff(self, solutions, atol):
        """``diff`` for trajectories that are still resident in HBM (lazy
        solutions of ``FDMOperator.device_resident_solution``): the matching
        steps are subtracted on the device and only the differences cross
        PCIe.  None if any solution has already been copied to the host."""
        everyone = [self] + list(solutions)
        if any(
            getattr(s, "device_trajectory", None) is None
            or getattr(s, "_y", None) is not None
            or s.vertex_oriented != self._vertex_oriented
            or s.device_trajectory.shape[1] != self.device_trajectory.shape[1]
            for s in everyone
        ):
            return None
        import torch

        from pararealml_b200 import _native
        from pararealml_b200.operators.fdm import device as dv

        matched, indices = self._matching_indices(solutions, atol)
        mine = self.device_trajectory
        state = mine.shape[1]
        y_dim = self._expected_shape[-1]
        n_cells = state // y_dim
        lib = _native.lib()
        out = []
        for j, other in enumerate(solutions):
            planes = torch.empty(
                (len(matched), state), dtype=torch.float64, device=mine.device
            )
            for row, (k_other, k_mine) in enumerate(zip(indices[j + 1], indices[0])):
                # other - mine (the subtraction kernel of the Parareal
                # correction: a - b over one state)
                _native.check(
                    lib.pml_parareal_correction(
                        other.device_trajectory[k_other].data_ptr(),
                        mine[k_mine].data_ptr(), planes[row].data_ptr(),
                        state, dv.stream_ptr(),
                    )
                )
            host = torch.empty(planes.shape, dtype=torch.float64, pin_memory=True)
            if len(matched):
                aos = dv.soa_to_aos(planes, n_cells, y_dim, len(matched))
                host.copy_(aos, non_blocking=True)
                torch.cuda.current_stream().synchronize()
            out.append(
                host.numpy().reshape((len(matched),) + self._expected_shape[1:])
            )
        return Diffs(matched, out)

    def generate_plots(self, **kwargs):
        raise NotImplementedError(
            "plotting is outside the scope of the B200 hot path"
        )
