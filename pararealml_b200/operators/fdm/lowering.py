"""Host-side lowering of a constrained problem to plain kernel inputs.

Reads what ``FDMOperator.solve`` reads from the problem in the reference
(``fdm_operator.py:48-231``: the SymPy system, the mesh, the static and
dynamic boundary constraints) and turns it into: a ``ProblemSpec`` for the
code generator, NaN-coded face tables (per axis and side, channels-last,
NaN = unconstrained) and 1-D coordinate vectors.  Works on this package's
``ConstrainedProblem`` (table-first fast path) and, duck-typed, on the
reference's (through its ``Constraint`` objects).
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from pararealml_b200.constraint import to_nan_table
from pararealml_b200.operators.fdm.codegen import ProblemSpec


@dataclass
class LoweredProblem:
    shape: Tuple[int, ...]
    d_x: Tuple[float, ...]
    coord: str
    y_dim: int
    rhs: list
    kinds: List[str]
    neu_mask: int = 0
    dir_mask: int = 0
    #: faces whose (static) Neumann table is 0.0 for every cell and component
    #: (zero-flux walls): the kernels then need no table look-ups there
    neu_zero_mask: int = 0
    # face (axis * 2 + side) -> is the boundary condition static
    face_static: Dict[int, bool] = field(default_factory=dict)
    face_cells: Dict[int, int] = field(default_factory=dict)
    static_neu: Dict[int, np.ndarray] = field(default_factory=dict)
    static_dir: Dict[int, np.ndarray] = field(default_factory=dict)
    coords: List[np.ndarray] = field(default_factory=list)
    aux: List[Optional[np.ndarray]] = field(default_factory=list)
    all_static: bool = True

    @property
    def n_cells(self) -> int:
        return int(np.prod(self.shape)) if self.shape else 1

    @property
    def n_dims(self) -> int:
        return len(self.shape)

    def kind_indices(self, kind: str) -> List[int]:
        return [i for i, k in enumerate(self.kinds) if k == kind]

    @property
    def has_dynamic(self) -> bool:
        return not self.all_static

    def spec(self, **overrides) -> ProblemSpec:
        return ProblemSpec(
            shape=self.shape,
            d_x=self.d_x,
            coord=self.coord,
            y_dim=self.y_dim,
            rhs=self.rhs,
            kinds=self.kinds,
            neu_mask=self.neu_mask,
            dir_mask=self.dir_mask,
            neu_zero_mask=self.neu_zero_mask & self.neu_mask,
            **overrides,
        )


def _tables_from_constraints(pairs_by_axis, shape, y_dim):
    """(x_dim, y_dim) object array of (lower, upper) Constraint pairs ->
    tables[axis][side] (or None when every component is None)."""
    out = []
    for axis in range(len(shape)):
        face_shape = tuple(1 if a == axis else n for a, n in enumerate(shape))
        sides = []
        for side in range(2):
            cs = [
                None if pairs_by_axis[axis, i] is None
                else pairs_by_axis[axis, i][side]
                for i in range(y_dim)
            ]
            if all(c is None for c in cs):
                sides.append(None)
                continue
            tab = np.stack(
                [to_nan_table(c, face_shape + (1,))[..., 0] for c in cs],
                axis=-1,
            )
            sides.append(tab)
        out.append(sides)
    return out


def boundary_tables(cp, t: Optional[float]):
    """(y tables, derivative tables) at time ``t`` for either package's
    constrained problem."""
    if hasattr(cp, "boundary_tables"):
        return cp.boundary_tables(True, t)
    y_pairs, d_pairs = cp.create_boundary_constraints(True, t)
    shape = cp.mesh.vertices_shape
    y_dim = cp.differential_equation.y_dimension
    return (
        _tables_from_constraints(y_pairs, shape, y_dim),
        _tables_from_constraints(d_pairs, shape, y_dim),
    )


def _flat(tab: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(tab, dtype=np.float64).reshape(-1)


def lower_problem(cp) -> LoweredProblem:
    eq = cp.differential_equation
    system = eq.symbolic_equation_system
    kinds = [k.name for k in system.lhs_types]
    rhs = list(system.rhs)
    if not eq.x_dimension:
        return LoweredProblem(
            shape=(), d_x=(), coord="CARTESIAN", y_dim=eq.y_dimension,
            rhs=rhs, kinds=kinds, coords=[], aux=[None] * 4,
        )
    mesh = cp.mesh
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
        y_dim=eq.y_dimension,
        rhs=rhs,
        kinds=kinds,
        all_static=bool(cp.are_all_boundary_conditions_static),
    )
    low.coords = [
        np.ascontiguousarray(a, dtype=np.float64)
        for a in mesh.vertex_axis_coordinates
    ]
    aux: List[Optional[np.ndarray]] = [None] * 4
    if coord != "CARTESIAN":
        aux[0] = 1.0 / low.coords[0]
        if coord == "SPHERICAL":
            phi = low.coords[2]
            aux[1] = np.sin(phi)
            aux[2] = np.cos(phi)
            aux[3] = 1.0 / np.sin(phi)
    low.aux = aux

    static_y, static_d = boundary_tables(cp, None)
    for axis, pair in enumerate(cp.boundary_conditions):
        for side, bc in enumerate(pair):
            f = axis * 2 + side
            low.face_static[f] = bool(bc.is_static)
            low.face_cells[f] = int(np.prod(shape)) // shape[axis]
            if bc.has_d_y_condition:
                low.neu_mask |= 1 << f
                if bc.is_static:
                    low.static_neu[f] = _flat(static_d[axis][side])
                    # (NaN = unconstrained cell compares unequal)
                    if np.all(low.static_neu[f] == 0.0):
                        low.neu_zero_mask |= 1 << f
            if bc.has_y_condition:
                low.dir_mask |= 1 << f
                if bc.is_static:
                    low.static_dir[f] = _flat(static_y[axis][side])
    return low


def dynamic_tables(cp, low: LoweredProblem, times: Sequence[float]):
    """Evaluates the dynamic boundary conditions at every time of ``times``
    (host; the conditions are user Python callables, reference
    fdm_operator.py:199-231).  Returns (neu, dir): face -> array
    (len(times), face_cells * y_dim)."""
    neu: Dict[int, np.ndarray] = {}
    dirichlet: Dict[int, np.ndarray] = {}
    dyn_faces = [f for f, s in low.face_static.items() if not s]
    for f in dyn_faces:
        n = low.face_cells[f] * low.y_dim
        if low.neu_mask >> f & 1:
            neu[f] = np.empty((len(times), n))
        if low.dir_mask >> f & 1:
            dirichlet[f] = np.empty((len(times), n))
    for k, t in enumerate(times):
        y_tabs, d_tabs = boundary_tables(cp, float(t))
        for f in dyn_faces:
            axis, side = divmod(f, 2)
            if f in neu:
                neu[f][k] = _flat(d_tabs[axis][side])
            if f in dirichlet:
                dirichlet[f][k] = _flat(y_tabs[axis][side])
    return neu, dirichlet


def apply_dirichlet_host(cp, y: np.ndarray, t: Optional[float]) -> np.ndarray:
    """In-place Dirichlet overwrite of a host array at time ``t`` (set-up
    only: initial state of a solve with dynamic conditions, reference
    fdm_operator.py:56-63)."""
    y_tabs, _ = boundary_tables(cp, t)
    for axis, pair in enumerate(y_tabs):
        for side, tab in enumerate(pair):
            if tab is None:
                continue
            idx = [slice(None)] * y.ndim
            idx[axis] = slice(-1, None) if side else slice(0, 1)
            face = y[tuple(idx)]
            np.copyto(face, tab, where=~np.isnan(tab))
    return y
