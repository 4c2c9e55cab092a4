"""Per-rank form of the Parareal oracle -- TEST / BENCHMARK INFRASTRUCTURE ONLY.

``oracle.parareal.parareal_solve`` emulates all ranks in one process (cheap,
used for parity).  This module restates the same reference routine
(``pararealml/operators/parareal/parareal_operator.py:102-197``) the way the
reference executes it under ``mpirun``: SPMD, one OS process per time slice,
every rank running the serial coarse sweep redundantly and exchanging data
with two ``Allgather`` calls (:165, :193).  The communicator is a
``torch.distributed`` gloo group standing in for ``MPI.COMM_WORLD`` (there is
no MPI runtime in the image).  ``bench.py --impl reference --gpus N`` times it
on the host cores; ``tests/test_parareal_gloo.py`` checks it against the
golden trajectories.
"""
import numpy as np

from oracle.fdm import time_grid
from oracle.parareal import should_terminate


class GlooComm:
    """The three members of ``MPI.COMM_WORLD`` the reference uses."""

    def __init__(self):
        import torch.distributed as dist

        self._dist = dist
        self.size = dist.get_world_size()
        self.rank = dist.get_rank()

    def allgather(self, send: np.ndarray, recv: np.ndarray):
        import torch

        out = torch.from_numpy(recv).view(-1)
        src = torch.from_numpy(np.ascontiguousarray(send)).view(-1)
        self._dist.all_gather_into_tensor(out, src)

    def barrier(self):
        self._dist.barrier()


def parareal_rank_solve(comm, ivp, f, g, termination_condition, make_sub_ivp,
                        max_iterations=None):
    """What one rank of the reference computes.  Returns
    ``(t, y_fine (every rank holds the whole trajectory), n_iterations)``."""
    size, rank = comm.size, comm.rank
    t_interval = ivp.t_interval
    delta_t = (t_interval[1] - t_interval[0]) / size
    for op, name in ((f, "fine"), (g, "coarse")):
        # reference :114-123
        if not np.isclose(delta_t, op.d_t * round(delta_t / op.d_t)):
            raise ValueError(
                f"{name} operator time step size ({op.d_t}) must be a "
                f"divisor of sub-IVP time slice length ({delta_t})"
            )
    vo = f.vertex_oriented
    cp = ivp.constrained_problem
    y_shape = cp.y_shape(vo)
    borders = np.linspace(t_interval[0], t_interval[1], size + 1)  # :129-131

    # :133-146 -- full-interval coarse solve on every rank
    coarse_ends = g.solve(ivp).discrete_y(vo)[
        np.rint((borders[1:] - t_interval[0]) / g.d_t).astype(int) - 1, ...
    ]
    border_points = np.concatenate(
        [ivp.initial_condition.discrete_y_0(vo)[np.newaxis], coarse_ends]
    )
    sub_fine = None
    corrections = np.empty((size, *y_shape))
    iterations = 0
    limit = size if max_iterations is None else min(size, max_iterations)
    for i in range(limit):  # :151
        iterations += 1
        sub = make_sub_ivp(cp, (borders[rank], borders[rank + 1]), border_points[rank])
        sub_fine = f.solve(sub, False).discrete_y(vo)  # :163
        correction = sub_fine[-1] - coarse_ends[rank]
        comm.allgather(correction, corrections)  # :165
        old_ends = np.copy(border_points[1:])
        for j in range(i, size):  # :168-186, redundantly on every rank
            if j > i:
                sub = make_sub_ivp(cp, (borders[j], borders[j + 1]), border_points[j])
                coarse_ends[j] = g.solve(sub).discrete_y(vo)[-1]
            border_points[j + 1] = coarse_ends[j] + corrections[j]
        if should_terminate(termination_condition, old_ends, border_points[1:]):
            break
    t = time_grid(t_interval, f.d_t)[1:]
    y_fine = np.empty((len(t), *y_shape))
    sub_fine = sub_fine + (border_points[rank + 1] - sub_fine[-1])  # :192
    comm.allgather(sub_fine, y_fine)  # :193
    return t, y_fine, iterations
