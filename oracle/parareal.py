"""NumPy oracle of ``PararealOperator.solve`` for P ranks, emulated serially.

Restates ``pararealml/operators/parareal/parareal_operator.py`` of the
reference (``solve`` :102-197, ``_should_terminate`` :53-100).  In the
reference every rank holds identical copies of the border points, coarse end
points and corrections and differs only in the fine slice it integrates, so
one process can emulate all ranks by looping over the slices.  Test
infrastructure only.
"""
from typing import Sequence

import numpy as np

from oracle.fdm import time_grid


def should_terminate(condition, old_end_points, new_end_points) -> bool:
    """reference :53-100"""
    if callable(condition):
        return condition(old_end_points, new_end_points)
    y_dim = old_end_points.shape[-1]
    if isinstance(condition, Sequence):
        if len(condition) != y_dim:
            raise ValueError(
                f"{len(condition)} tolerances for {y_dim} y dimensions"
            )
        tolerances = np.array(condition)
    else:
        tolerances = np.array([condition] * y_dim)
    worst = np.empty(y_dim)
    for c in range(y_dim):
        m = 0.0
        for new, old in zip(new_end_points[..., c], old_end_points[..., c]):
            m = np.maximum(m, np.sqrt(np.square(new - old).mean()))
        worst[c] = m
    return all(worst < tolerances)


def parareal_solve(
    ivp, f, g, termination_condition, size, make_sub_ivp,
    max_iterations=None,
):
    """Emulates ``size`` ranks.  ``make_sub_ivp(cp, (t0, t1), y0)`` builds a
    sub-IVP whose initial condition is the discrete state ``y0`` (with static
    Dirichlet values re-applied, reference :153-162).  Returns
    ``(t, y_fine, n_iterations)``."""
    t_interval = ivp.t_interval
    delta_t = (t_interval[1] - t_interval[0]) / size
    for op, name in ((f, "fine"), (g, "coarse")):
        if not np.isclose(delta_t, op.d_t * round(delta_t / op.d_t)):
            raise ValueError(
                f"{name} operator time step size ({op.d_t}) must be a "
                f"divisor of sub-IVP time slice length ({delta_t})"
            )
    vo = f.vertex_oriented
    cp = ivp.constrained_problem
    y_shape = cp.y_shape(vo)
    borders = np.linspace(t_interval[0], t_interval[1], size + 1)

    coarse_ends = g.solve(ivp).discrete_y(vo)[
        np.rint((borders[1:] - t_interval[0]) / g.d_t).astype(int) - 1, ...
    ]
    border_points = np.concatenate(
        [ivp.initial_condition.discrete_y_0(vo)[np.newaxis], coarse_ends]
    )
    fine = [None] * size
    corrections = np.empty((size, *y_shape))
    iterations = 0
    limit = size if max_iterations is None else min(size, max_iterations)
    for i in range(limit):
        iterations += 1
        for rank in range(size):
            sub = make_sub_ivp(
                cp, (borders[rank], borders[rank + 1]), border_points[rank]
            )
            fine[rank] = f.solve(sub, False).discrete_y(vo)
            corrections[rank] = fine[rank][-1] - coarse_ends[rank]
        old_ends = np.copy(border_points[1:])
        for j in range(i, size):
            if j > i:
                sub = make_sub_ivp(
                    cp, (borders[j], borders[j + 1]), border_points[j]
                )
                coarse_ends[j] = g.solve(sub).discrete_y(vo)[-1]
            border_points[j + 1] = coarse_ends[j] + corrections[j]
        if should_terminate(
            termination_condition, old_ends, border_points[1:]
        ):
            break
    for rank in range(size):
        fine[rank] = fine[rank] + (border_points[rank + 1] - fine[rank][-1])
    t = time_grid(t_interval, f.d_t)[1:]
    return t, np.concatenate(fine, axis=0), iterations
