"""Uniform hyper-rectangular meshes in Cartesian and curvilinear coordinates.

Host-side, set-up only.  Mirrors the public surface of the reference's
``pararealml/mesh.py`` (``Mesh`` :20-347, coordinate helpers :498-628) so that
either implementation's objects can be handed to the B200 operators.  Unlike
the reference, the full coordinate grids are created lazily as read-only
broadcast views, so a 512^3 mesh costs O(n) host memory until somebody asks
for a materialised array.
"""
from enum import Enum
from typing import Sequence, Tuple

import numpy as np

SpatialDomainInterval = Tuple[float, float]


class CoordinateSystem(Enum):
    CARTESIAN = 0
    POLAR = 1
    CYLINDRICAL = 2
    SPHERICAL = 3


class Mesh:
    """A grid with per-axis uniform spacing.  Axis order for curvilinear
    systems is (r, theta[, z | phi]) as in the reference (mesh.py:63-94)."""

    def __init__(
        self,
        x_intervals: Sequence[SpatialDomainInterval],
        d_x: Sequence[float],
        coordinate_system_type: CoordinateSystem = CoordinateSystem.CARTESIAN,
    ):
        n = len(x_intervals)
        if n == 0:
            raise ValueError("a mesh needs at least one spatial interval")
        if n != len(d_x):
            raise ValueError(
                f"got {n} spatial intervals but {len(d_x)} step sizes"
            )
        for lo, hi in x_intervals:
            if hi <= lo:
                raise ValueError(
                    f"empty or inverted spatial interval ({lo}, {hi})"
                )
        if any(h <= 0.0 for h in d_x):
            raise ValueError("spatial step sizes must be positive")

        cs = coordinate_system_type
        if cs != CoordinateSystem.CARTESIAN:
            if x_intervals[0][0] < 0:
                raise ValueError("r must be non-negative")
            if n < 2 or x_intervals[1][0] < 0.0 or (
                x_intervals[1][1] > 2.0 * np.pi
            ):
                raise ValueError("theta must lie within [0, 2 pi]")
            if cs == CoordinateSystem.POLAR and n != 2:
                raise ValueError("polar meshes are 2 dimensional")
            if cs != CoordinateSystem.POLAR and n != 3:
                raise ValueError(
                    "cylindrical and spherical meshes are 3 dimensional"
                )
            if cs == CoordinateSystem.SPHERICAL and (
                x_intervals[2][0] < 0.0 or x_intervals[2][1] > np.pi
            ):
                raise ValueError("phi must lie within [0, pi]")

        self._x_intervals = tuple((lo, hi) for lo, hi in x_intervals)
        self._d_x = tuple(d_x)
        self._cs = cs
        self._dims = n

        self._shapes = {}
        self._axes = {}
        for vo in (True, False):
            shape = tuple(
                round((hi - lo) / h + vo)
                for (lo, hi), h in zip(self._x_intervals, self._d_x)
            )
            self._shapes[vo] = shape
            axes = []
            for (lo, hi), h, m in zip(self._x_intervals, self._d_x, shape):
                if not vo:
                    lo, hi = lo + h / 2.0, hi - h / 2.0
                a = np.linspace(lo, hi, m)
                a.setflags(write=False)
                axes.append(a)
            self._axes[vo] = tuple(axes)
        self._grids = {}

    # -- plain attributes -------------------------------------------------
    @property
    def x_intervals(self):
        return self._x_intervals

    @property
    def d_x(self):
        return self._d_x

    @property
    def coordinate_system_type(self) -> CoordinateSystem:
        return self._cs

    @property
    def dimensions(self) -> int:
        return self._dims

    @property
    def vertices_shape(self) -> Tuple[int, ...]:
        return self._shapes[True]

    @property
    def cells_shape(self) -> Tuple[int, ...]:
        return self._shapes[False]

    @property
    def vertex_axis_coordinates(self):
        return self._axes[True]

    @property
    def cell_center_axis_coordinates(self):
        return self._axes[False]

    @property
    def vertex_coordinate_grids(self):
        return self.coordinate_grids(True)

    @property
    def cell_center_coordinate_grids(self):
        return self.coordinate_grids(False)

    def shape(self, vertex_oriented: bool) -> Tuple[int, ...]:
        return self._shapes[bool(vertex_oriented)]

    def axis_coordinates(self, vertex_oriented: bool):
        return self._axes[bool(vertex_oriented)]

    def coordinate_grids(self, vertex_oriented: bool):
        """Per-axis coordinate grids ('ij' indexing) as read-only broadcast
        views (no per-point storage)."""
        vo = bool(vertex_oriented)
        if vo not in self._grids:
            shape = self._shapes[vo]
            grids = []
            for i, a in enumerate(self._axes[vo]):
                view_shape = [1] * self._dims
                view_shape[i] = len(a)
                g = np.broadcast_to(a.reshape(view_shape), shape)
                grids.append(g)
            self._grids[vo] = tuple(grids)
        return self._grids[vo]

    def cartesian_coordinate_grids(self, vertex_oriented: bool):
        return tuple(
            to_cartesian_coordinates(
                self.coordinate_grids(vertex_oriented), self._cs
            )
        )

    def all_index_coordinates(
        self, vertex_oriented: bool, flatten: bool = False
    ) -> np.ndarray:
        """Array (*shape, dims) (or (N, dims)) of every mesh point."""
        coords = np.stack(self.coordinate_grids(vertex_oriented), axis=-1)
        return coords.reshape((-1, self._dims)) if flatten else coords

    def boundary_index_coordinates(
        self, vertex_oriented: bool, axis: int, upper: bool
    ) -> np.ndarray:
        """Coordinates of the points of one boundary face, shape
        (*shape with 1 along ``axis``, dims).  The coordinate along ``axis``
        is the domain boundary itself (also for cell-oriented faces), which
        is what the reference evaluates boundary conditions on
        (constrained_problem.py:394-401)."""
        vo = bool(vertex_oriented)
        axes = list(self._axes[vo])
        axes[axis] = np.array([self._axes[True][axis][-1 if upper else 0]])
        grids = np.meshgrid(*axes, indexing="ij")
        return np.stack(grids, axis=-1)

    def unit_vector_grids(self, vertex_oriented: bool):
        return tuple(
            np.stack(v, axis=-1)
            for v in unit_vectors_at(
                self.coordinate_grids(vertex_oriented), self._cs
            )
        )

    # -- geometry ---------------------------------------------------------
    @property
    def volume(self) -> float:
        iv = self._x_intervals
        if self._cs == CoordinateSystem.CARTESIAN:
            return float(np.prod([hi - lo for lo, hi in iv]))
        (r0, r1), (th0, th1) = iv[0], iv[1]
        if self._cs == CoordinateSystem.SPHERICAL:
            p0, p1 = iv[2]
            return (
                (r1**3 - r0**3) / 3.0 * (th1 - th0) * (np.cos(p0) - np.cos(p1))
            )
        area = (r1**2 - r0**2) * (th1 - th0) / 2.0
        if self._dims == 2:
            return area
        return area * (iv[2][1] - iv[2][0])

    @property
    def boundary_sizes(self):
        iv = self._x_intervals
        if self._cs == CoordinateSystem.CARTESIAN:
            lengths = [hi - lo for lo, hi in iv]
            vol = np.prod(lengths)
            return tuple((vol / ell,) * 2 for ell in lengths)
        (r0, r1), (th0, th1) = iv[0], iv[1]
        dth = th1 - th0
        if self._cs == CoordinateSystem.SPHERICAL:
            p0, p1 = iv[2]
            dcos = np.cos(p0) - np.cos(p1)
            ring = (r1**2 - r0**2) / 2.0
            return (
                (r0**2 * dth * dcos, r1**2 * dth * dcos),
                (ring * (p1 - p0),) * 2,
                (ring * dth * np.sin(p0), ring * dth * np.sin(p1)),
            )
        r_faces = (r0 * dth, r1 * dth)
        th_faces = (r1 - r0,) * 2
        if self._dims == 2:
            return (r_faces, th_faces)
        dz = iv[2][1] - iv[2][0]
        return (
            (r_faces[0] * dz, r_faces[1] * dz),
            (th_faces[0] * dz, th_faces[1] * dz),
            ((r1**2 - r0**2) * dth / 2.0,) * 2,
        )


def unit_vectors_at(x, coordinate_system_type: CoordinateSystem):
    """Cartesian components of the local orthonormal basis at ``x``."""
    cs = coordinate_system_type
    if cs == CoordinateSystem.CARTESIAN:
        out = []
        for i in range(len(x)):
            vec = [np.zeros_like(x[i])] * len(x)
            vec[i] = np.ones_like(x[i])
            out.append(vec)
        return out
    th = x[1]
    s, c = np.sin(th), np.cos(th)
    if cs == CoordinateSystem.POLAR:
        return [[c, s], [-s, c]]
    zero = np.zeros_like(th)
    if cs == CoordinateSystem.CYLINDRICAL:
        return [[c, s, zero], [-s, c, zero], [zero, zero, np.ones_like(th)]]
    if cs == CoordinateSystem.SPHERICAL:
        sp, cp = np.sin(x[2]), np.cos(x[2])
        return [[sp * c, sp * s, cp], [-s, c, zero], [cp * c, cp * s, -sp]]
    raise ValueError(f"unsupported coordinate system {cs}")


def to_cartesian_coordinates(x, from_coordinate_system_type: CoordinateSystem):
    cs = from_coordinate_system_type
    if cs == CoordinateSystem.CARTESIAN:
        return x
    if cs == CoordinateSystem.POLAR:
        return [x[0] * np.cos(x[1]), x[0] * np.sin(x[1])]
    if cs == CoordinateSystem.CYLINDRICAL:
        return [x[0] * np.cos(x[1]), x[0] * np.sin(x[1]), x[2]]
    if cs == CoordinateSystem.SPHERICAL:
        return [
            x[0] * np.sin(x[2]) * np.cos(x[1]),
            x[0] * np.sin(x[2]) * np.sin(x[1]),
            x[0] * np.cos(x[2]),
        ]
    raise ValueError(f"unsupported coordinate system {cs}")


def from_cartesian_coordinates(x, to_coordinate_system_type: CoordinateSystem):
    cs = to_coordinate_system_type
    if cs == CoordinateSystem.CARTESIAN:
        return x
    rho = np.sqrt(x[0] ** 2 + x[1] ** 2)
    if cs == CoordinateSystem.POLAR:
        return [rho, np.arctan2(x[1], x[0])]
    if cs == CoordinateSystem.CYLINDRICAL:
        return [rho, np.arctan2(x[1], x[0]), x[2]]
    if cs == CoordinateSystem.SPHERICAL:
        return [
            np.sqrt(x[0] ** 2 + x[1] ** 2 + x[2] ** 2),
            np.arctan2(x[1], x[0]),
            np.arctan2(rho, x[2]),
        ]
    raise ValueError(f"unsupported coordinate system {cs}")
