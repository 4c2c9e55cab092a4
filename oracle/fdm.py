"""NumPy oracle of ``FDMOperator.solve``.

Restates ``pararealml/operators/fdm/fdm_operator.py`` (:48-231),
``fdm_symbol_mapper.py`` (:45-158) and ``operators/symbol_mapper.py``
(:160-253) of the reference: symbol names are parsed into leaf evaluators
that call the oracle differentiator, the right-hand sides are lambdified to
NumPy, and the explicit integrators step the state.  Works on either
package's (duck-typed) ``InitialValueProblem``.  Test infrastructure only.
"""
import numpy as np
import sympy as sp

from oracle import differentiator as diff
from oracle.integrator import STEPPERS, _constrained


def time_grid(t_interval, d_t):
    """reference operator.py:60-74"""
    t_0 = t_interval[0]
    steps = int(round((t_interval[1] - t_0) / d_t))
    return np.linspace(t_0, t_0 + steps * d_t, steps + 1)


class OracleSolution:
    def __init__(self, t, y, d_t):
        self.t_coordinates = t
        self._y = y
        self.d_t = d_t
        self.vertex_oriented = True

    def discrete_y(self, vertex_oriented=None):
        return np.copy(self._y)


def _leaf_evaluators(cp):
    """symbol -> f(t, y, dbc_of_t) (reference symbol_mapper.py:160-220 and
    fdm_symbol_mapper.py:45-148; the vector Laplacian branch is never stored
    there, so such symbols raise KeyError exactly like the reference)."""
    eq = cp.differential_equation
    mesh = cp.mesh
    x_dim = eq.x_dimension
    symbols = set.union(*[e.free_symbols for e in eq.symbolic_equation_system.rhs])
    leaves = {}
    for sym in symbols:
        tokens = sym.name.split("_")
        kind = tokens[0]
        idx = [int(s) for s in tokens[1:]]

        def comp_slice(ids):
            contiguous = all(ids[i] + 1 == ids[i + 1] for i in range(len(ids) - 1))
            return slice(ids[0], ids[-1] + 1) if contiguous else list(ids)

        if kind == "t":
            leaves[sym] = lambda t, y, dbc: np.array([t])
        elif kind == "y":
            leaves[sym] = lambda t, y, dbc, i=idx[0]: y[..., i : i + 1]
        elif kind == "x":
            leaves[sym] = lambda t, y, dbc, a=idx[0]: (
                mesh.vertex_coordinate_grids[a][..., np.newaxis]
            )
        elif kind == "y-gradient":
            leaves[sym] = lambda t, y, dbc, i=idx[0], a=idx[1]: diff.gradient(
                y[..., i : i + 1], mesh, a, dbc[:, i : i + 1]
            )
        elif kind == "y-hessian":
            leaves[sym] = (
                lambda t, y, dbc, i=idx[0], a=idx[1], b=idx[2]: diff.hessian(
                    y[..., i : i + 1], mesh, a, b, dbc[:, i : i + 1]
                )
            )
        elif kind == "y-laplacian":
            leaves[sym] = lambda t, y, dbc, i=idx[0]: diff.laplacian(
                y[..., i : i + 1], mesh, dbc[:, i : i + 1]
            )
        elif kind == "y-divergence":
            leaves[sym] = lambda t, y, dbc, s=comp_slice(idx): diff.divergence(
                y[..., s], mesh, dbc[:, s]
            )
        elif kind == "y-curl":
            ids, ind = (idx, 0) if x_dim == 2 else (idx[:-1], idx[-1])
            leaves[sym] = lambda t, y, dbc, s=comp_slice(ids), k=ind: diff.curl(
                y[..., s], mesh, k, dbc[:, s]
            )
    return leaves


def _rhs_function(cp, leaves, indices):
    rhs = cp.differential_equation.symbolic_equation_system.rhs
    exprs = [rhs[i] for i in indices]
    syms = set()
    for e in exprs:
        syms.update(e.free_symbols)
    syms = list(syms)
    evaluators = [leaves[s] for s in syms]  # KeyError for vector Laplacians
    fn = sp.lambdify([syms], exprs, "numpy")

    def evaluate(t, y, dbc):
        parts = fn([ev(t, y, dbc) for ev in evaluators])
        return np.concatenate(
            [np.broadcast_to(p, y.shape[:-1] + (1,)) for p in parts], axis=-1
        )

    return evaluate


def _constraint_functions(cp):
    """(y constraints of t, derivative constraints of t), reference
    fdm_operator.py:167-231 (the per-step caches are an implementation detail
    and are dropped)."""
    if not cp.differential_equation.x_dimension:
        return (lambda t: None), (lambda t: None)
    if cp.are_all_boundary_conditions_static:
        y_c = cp.static_y_vertex_constraints
        d_c = cp.static_boundary_vertex_constraints[1]
        return (lambda t: y_c), (lambda t: d_c)
    cache = {}

    def bcs(t):
        if t not in cache:
            cache.clear()
            cache[t] = cp.create_boundary_constraints(True, t)
        return cache[t]

    def d_c(t):
        return bcs(t)[1]

    if not cp.are_there_boundary_conditions_on_y:
        y_static = cp.static_y_vertex_constraints
        return (lambda t: y_static), d_c
    return (lambda t: cp.create_y_vertex_constraints(bcs(t)[0])), d_c


def fdm_solve(ivp, integrator, d_t, tol=1e-3, stats=None):
    """Returns ``(t[1:], y)`` with ``y.shape == (n_steps, *mesh, y_dim)``
    (reference fdm_operator.py:48-165)."""
    step = STEPPERS[integrator]
    cp = ivp.constrained_problem
    eq = cp.differential_equation
    system = eq.symbolic_equation_system
    lhs_enum = type(system.lhs_types[0])
    dt_idx = list(system.equation_indices_by_type(lhs_enum["D_Y_OVER_D_T"]))
    alg_idx = list(system.equation_indices_by_type(lhs_enum["Y"]))
    lap_idx = list(system.equation_indices_by_type(lhs_enum["Y_LAPLACIAN"]))

    t = time_grid(ivp.t_interval, d_t)
    y = np.empty((len(t) - 1,) + cp.y_vertices_shape)
    y_i = ivp.initial_condition.discrete_y_0(True)

    if eq.x_dimension and not cp.are_all_boundary_conditions_static:
        init_bcs = cp.create_boundary_constraints(True, t[0])
        _constrained(cp.create_y_vertex_constraints(init_bcs[0]), y_i)

    leaves = _leaf_evaluators(cp)
    rhs_dt = _rhs_function(cp, leaves, dt_idx)
    rhs_alg = _rhs_function(cp, leaves, alg_idx) if alg_idx else None
    rhs_lap = _rhs_function(cp, leaves, lap_idx) if lap_idx else None
    y_c, d_c = _constraint_functions(cp)

    def f(t_, y_):
        out = np.zeros(y_.shape)
        out[..., dt_idx] = rhs_dt(t_, y_, d_c(t_))
        return out

    for i, t_i in enumerate(t[:-1]):
        y_next = step(y_i, t_i, d_t, f, y_c)
        if alg_idx:
            c = y_c(t_i + d_t)
            c = None if c is None else c[alg_idx]
            y_next[..., alg_idx] = _constrained(c, rhs_alg(t_i, y_i, d_c(t_i)))
        if lap_idx:
            c = y_c(t_i + d_t)
            c = None if c is None else c[lap_idx]
            dc = d_c(t_i + d_t)
            dc = None if dc is None else dc[:, lap_idx]
            sol, sweeps = diff.anti_laplacian(
                rhs_lap(t_i, y_i, d_c(t_i)), cp.mesh, c, dc, tol=tol,
                return_sweeps=True,
            )
            if stats is not None:
                stats.setdefault("jacobi_sweeps", []).append(sweeps)
            y_next[..., lap_idx] = sol
        y[i] = y_i = y_next
    return t[1:], y


class OracleFDMOperator:
    """Operator-shaped wrapper (``d_t``, ``vertex_oriented``, ``solve``) so the
    oracle can serve as ``f``/``g`` of a Parareal driver in tests."""

    def __init__(self, integrator: str, d_t: float, tol: float = 1e-3):
        if d_t <= 0.0:
            raise ValueError("time step size must be greater than 0")
        self._integrator = integrator
        self._d_t = d_t
        self._tol = tol

    @property
    def d_t(self):
        return self._d_t

    @property
    def vertex_oriented(self):
        return True

    def solve(self, ivp, parallel_enabled=True):
        t, y = fdm_solve(ivp, self._integrator, self._d_t, self._tol)
        return OracleSolution(t, y, self._d_t)
