"""Differential equation + mesh + boundary conditions (host side, set-up).

API mirror of the reference's ``pararealml/constrained_problem.py``
(:16-476).  The primary representation here is the *NaN-coded face table*
(one dense array per boundary face and condition kind, NaN = unconstrained),
which is what the B200 kernels consume; the reference's ``Constraint`` object
arrays are derived from the tables on demand.  Full-grid artefacts
(``static_y_vertex_constraints``) are built lazily so that a 512^3 problem
does not allocate per-vertex masks unless a caller asks for them.
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np

from pararealml_b200.boundary_condition import BoundaryCondition
from pararealml_b200.constraint import Constraint, from_nan_table
from pararealml_b200.mesh import Mesh

BoundaryConditionPair = Tuple[BoundaryCondition, BoundaryCondition]

# tables[axis][side] -> ndarray (*face_shape, y_dim) or None
FaceTables = List[List[Optional[np.ndarray]]]


class ConstrainedProblem:
    def __init__(
        self,
        diff_eq,
        mesh: Optional[Mesh] = None,
        boundary_conditions: Optional[Sequence[BoundaryConditionPair]] = None,
    ):
        self._diff_eq = diff_eq
        x_dim = diff_eq.x_dimension
        y_dim = diff_eq.y_dimension
        self._static_tables = {}
        self._static_constraints = {}
        self._y_vertex_constraints = None

        if not x_dim:
            self._mesh = None
            self._bcs = None
            self._y_vertices_shape = self._y_cells_shape = (y_dim,)
            self._all_static = np.bool_(False)
            self._any_on_y = np.bool_(False)
            return

        if mesh is None:
            raise ValueError("a PDE needs a mesh")
        if mesh.dimensions != x_dim:
            raise ValueError(
                f"mesh has {mesh.dimensions} dimensions, the equation {x_dim}"
            )
        if boundary_conditions is None:
            raise ValueError("a PDE needs boundary conditions")
        if len(boundary_conditions) != x_dim:
            raise ValueError(
                f"got {len(boundary_conditions)} boundary condition pairs "
                f"for {x_dim} spatial dimensions"
            )
        self._mesh = mesh
        self._bcs = tuple(boundary_conditions)
        self._y_vertices_shape = mesh.vertices_shape + (y_dim,)
        self._y_cells_shape = mesh.cells_shape + (y_dim,)
        self._all_static = np.all(
            [lo.is_static and hi.is_static for lo, hi in self._bcs]
        )
        self._any_on_y = np.any(
            [lo.has_y_condition or hi.has_y_condition for lo, hi in self._bcs]
        )
        # static tables are evaluated eagerly (errors in user callables
        # surface at construction time, like in the reference)
        for vo in (True, False):
            self._static_tables[vo] = self._evaluate_tables(vo, None, None)

    # -- simple accessors --------------------------------------------------
    @property
    def differential_equation(self):
        return self._diff_eq

    @property
    def mesh(self) -> Optional[Mesh]:
        return self._mesh

    @property
    def boundary_conditions(self):
        return self._bcs

    @property
    def y_vertices_shape(self) -> Tuple[int, ...]:
        return self._y_vertices_shape

    @property
    def y_cells_shape(self) -> Tuple[int, ...]:
        return self._y_cells_shape

    @property
    def are_all_boundary_conditions_static(self) -> np.bool_:
        return self._all_static

    @property
    def are_there_boundary_conditions_on_y(self) -> np.bool_:
        return self._any_on_y

    def y_shape(self, vertex_oriented: Optional[bool] = None):
        return (
            self._y_vertices_shape if vertex_oriented else self._y_cells_shape
        )

    # -- NaN-coded face tables (what the kernels consume) -------------------
    def boundary_tables(
        self, vertex_oriented: bool, t: Optional[float] = None
    ) -> Tuple[Optional[FaceTables], Optional[FaceTables]]:
        """(y tables, d_y tables); entry ``[axis][side]`` is an array of shape
        (*mesh shape with 1 along axis, y_dim) with NaN where unconstrained,
        or None if the face has no condition of that kind (or it is dynamic
        and ``t`` is None)."""
        if not self._diff_eq.x_dimension:
            return None, None
        vo = bool(vertex_oriented)
        if t is None or self._all_static:
            return self._static_tables[vo]
        return self._evaluate_tables(vo, t, self._static_tables[vo])

    def _evaluate_tables(self, vo, t, static):
        x_dim = self._diff_eq.x_dimension
        y_dim = self._diff_eq.y_dimension
        y_tabs: FaceTables = [[None, None] for _ in range(x_dim)]
        d_tabs: FaceTables = [[None, None] for _ in range(x_dim)]
        for axis, pair in enumerate(self._bcs):
            for side, bc in enumerate(pair):
                if not bc.is_static and t is None:
                    continue
                if bc.is_static and static is not None:
                    y_tabs[axis][side] = static[0][axis][side]
                    d_tabs[axis][side] = static[1][axis][side]
                    continue
                coords = self._mesh.boundary_index_coordinates(
                    vo, axis, bool(side)
                )
                x = coords.reshape((-1, x_dim))
                for has, fn, out in (
                    (bc.has_y_condition, bc.y_condition, y_tabs),
                    (bc.has_d_y_condition, bc.d_y_condition, d_tabs),
                ):
                    if not has:
                        continue
                    vals = np.asarray(fn(x, t), dtype=float)
                    if vals.shape != (len(x), y_dim):
                        raise ValueError(
                            "boundary condition function returned shape "
                            f"{vals.shape}, expected {(len(x), y_dim)}"
                        )
                    tab = vals.reshape(coords.shape[:-1] + (y_dim,))
                    tab.setflags(write=False)
                    out[axis][side] = tab
        return y_tabs, d_tabs

    def dirichlet_face_tables(self, t: Optional[float] = None):
        """Vertex-oriented y tables (shortcut for ``boundary_tables(True,
        t)[0]``)."""
        return self.boundary_tables(True, t)[0]

    # -- reference-compatible Constraint views -------------------------------
    @staticmethod
    def _tables_to_constraints(tables: FaceTables, y_dim: int) -> np.ndarray:
        out = np.empty((len(tables), y_dim), dtype=object)
        for axis, (lo, hi) in enumerate(tables):
            for i in range(y_dim):
                out[axis, i] = (
                    None if lo is None else from_nan_table(lo[..., i : i + 1]),
                    None if hi is None else from_nan_table(hi[..., i : i + 1]),
                )
        return out

    def create_boundary_constraints(
        self, vertex_oriented: bool, t: Optional[float] = None
    ) -> Tuple[Optional[np.ndarray], Optional[np.ndarray]]:
        """Two (x_dim, y_dim) object arrays of (lower, upper) ``Constraint``
        pairs for y and for its derivative (reference :303-348)."""
        if not self._diff_eq.x_dimension:
            return None, None
        vo = bool(vertex_oriented)
        if t is None or self._all_static:
            return self.static_boundary_constraints(vo)
        y_dim = self._diff_eq.y_dimension
        y_tabs, d_tabs = self.boundary_tables(vo, t)
        return (
            self._tables_to_constraints(y_tabs, y_dim),
            self._tables_to_constraints(d_tabs, y_dim),
        )

    def static_boundary_constraints(self, vertex_oriented: bool):
        if not self._diff_eq.x_dimension:
            return None
        vo = bool(vertex_oriented)
        if vo not in self._static_constraints:
            y_dim = self._diff_eq.y_dimension
            y_tabs, d_tabs = self._static_tables[vo]
            pair = (
                self._tables_to_constraints(y_tabs, y_dim),
                self._tables_to_constraints(d_tabs, y_dim),
            )
            pair[0].setflags(write=False)
            pair[1].setflags(write=False)
            self._static_constraints[vo] = pair
        return self._static_constraints[vo]

    @property
    def static_boundary_vertex_constraints(self):
        return self.static_boundary_constraints(True)

    @property
    def static_boundary_cell_constraints(self):
        return self.static_boundary_constraints(False)

    @property
    def static_y_vertex_constraints(self) -> Optional[np.ndarray]:
        if not self._diff_eq.x_dimension:
            return None
        if self._y_vertex_constraints is None:
            c = self.create_y_vertex_constraints(
                self.static_boundary_constraints(True)[0]
            )
            c.setflags(write=False)
            self._y_vertex_constraints = c
        return self._y_vertex_constraints

    def create_y_vertex_constraints(
        self, y_boundary_vertex_constraints: Optional[np.ndarray]
    ) -> Optional[np.ndarray]:
        """One full-grid ``Constraint`` per component; faces are written in
        the order axis 0 lower, axis 0 upper, axis 1 lower, ... so later faces
        win on shared edges (reference :262-301)."""
        x_dim = self._diff_eq.x_dimension
        if not x_dim or y_boundary_vertex_constraints is None:
            return None
        y_dim = self._diff_eq.y_dimension
        out = np.empty(y_dim, dtype=object)
        grid = np.empty(self._y_vertices_shape[:-1] + (1,))
        for i in range(y_dim):
            grid.fill(np.nan)
            for axis in range(x_dim):
                idx = [slice(None)] * (x_dim + 1)
                for side, c in enumerate(
                    y_boundary_vertex_constraints[axis, i]
                ):
                    if c is None:
                        continue
                    idx[axis] = slice(-1, None) if side else slice(0, 1)
                    c.apply(grid[tuple(idx)])
            out[i] = from_nan_table(grid)
        return out

    def apply_dirichlet_tables(self, y: np.ndarray, y_tables) -> np.ndarray:
        """In-place Dirichlet overwrite of a vertex-oriented ``y`` straight
        from face tables, O(surface); same precedence as
        ``create_y_vertex_constraints``."""
        if y_tables is None:
            return y
        for axis, pair in enumerate(y_tables):
            for side, tab in enumerate(pair):
                if tab is None:
                    continue
                idx = [slice(None)] * y.ndim
                idx[axis] = slice(-1, None) if side else slice(0, 1)
                face = y[tuple(idx)]
                np.copyto(face, tab, where=~np.isnan(tab))
        return y
