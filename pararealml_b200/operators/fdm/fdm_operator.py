"""``FDMOperator``: the drop-in finite difference solver on B200.

Same constructor and ``solve(ivp, parallel_enabled=True) -> Solution`` as the
reference (``pararealml/operators/fdm/fdm_operator.py:27-77``).  Underneath,
the problem is lowered once (``lowering.py``), the SymPy right-hand side is
emitted into the fixed CUDA template and compiled with NVRTC (``codegen.py``,
``csrc/fdm_template.cuh``), and the whole time loop runs on the device through
the C ABI (``pml_fdm_run``): the state, the stage buffers and the trajectory
stay in HBM for all steps; the host only evaluates user Python callables
(dynamic boundary conditions) and draws the Jacobi start values from NumPy's
global generator like the reference does (numerical_differentiator.py:908-909).
"""
import os
from typing import Optional, Tuple

import numpy as np
import torch

from pararealml_b200.operator import Operator, discretize_time_domain
from pararealml_b200.operators.fdm import codegen
from pararealml_b200.operators.fdm import device as dv
from pararealml_b200.operators.fdm.lowering import (
    LoweredProblem,
    apply_dirichlet_host,
    dynamic_tables,
    lower_problem,
)
from pararealml_b200.operators.fdm.numerical_differentiator import (
    NumericalDifferentiator,
)
from pararealml_b200.operators.fdm.numerical_integrator import (
    NumericalIntegrator,
)
from pararealml_b200.solution import Solution

# device bytes one trajectory chunk may take before the solve is pipelined
TRAJECTORY_CHUNK_BYTES = int(
    os.environ.get("PML_TRAJ_CHUNK_BYTES", str(8 << 30))
)
# time steps per upload of dynamic boundary tables
DYNAMIC_BC_CHUNK_STEPS = 256
# meshes of at least this many vertices evaluate initial conditions that have
# a device form (Gaussian, marginal Beta product) on the GPU
DEVICE_IC_MIN_CELLS = int(os.environ.get("PML_DEVICE_IC_MIN_CELLS", str(1 << 18)))


def lowered(cp) -> LoweredProblem:
    """Lowering is cached on the constrained problem (Parareal solves the
    same problem many times)."""
    low = getattr(cp, "_pml_b200_lowered", None)
    if low is None:
        low = lower_problem(cp)
        try:
            cp._pml_b200_lowered = low
        except AttributeError:
            pass
    return low


def has_device_initial_condition(ic, low: LoweredProblem) -> bool:
    """True if the initial state can be evaluated on the device
    (``discrete_y_0_planes``) and the mesh is large enough for that to pay."""
    from pararealml_b200.initial_condition import ContinuousInitialCondition

    method = getattr(type(ic), "discrete_y_0_planes", None)
    return bool(
        method is not None
        and method is not ContinuousInitialCondition.discrete_y_0_planes
        and low.n_dims and low.all_static
        and low.n_cells >= DEVICE_IC_MIN_CELLS
    )


def plan_overrides(cp, low: LoweredProblem, y0: Optional[np.ndarray],
                   dirichlet_satisfied: bool = False) -> dict:
    """Code generation switches of the stage kernels for this problem.
    ``dirichlet_satisfied``: the initial state carries the static Dirichlet
    values by construction (device-side initial conditions)."""
    for sym in set().union(*[e.free_symbols for e in low.rhs]):
        if sym.name.startswith("y-vector-laplacian"):
            # the reference's symbol mapper never stores this evaluator
            # (symbol_mapper.py:215-218) and fails the same way
            raise KeyError(
                f"{sym}: like the reference, FDMOperator has no evaluator for "
                "vector-Laplacian symbols (ThreePointCentralDifferenceMethod."
                "vector_laplacian itself works; write the equation with "
                "y-laplacian / y-gradient / y-hessian symbols instead)"
            )
    passthrough = False
    n_other = len(low.kinds) - len(low.kind_indices("D_Y_OVER_D_T"))
    if n_other and low.all_static and low.n_dims:
        # the stage inputs of the non-dt components equal y itself when y
        # already satisfies the (static) Dirichlet values
        if low.dir_mask == 0 or dirichlet_satisfied:
            passthrough = True
        elif y0 is not None:
            probe = apply_dirichlet_host(cp, np.array(y0, copy=True), None)
            passthrough = bool(np.array_equal(probe, y0))
    return {
        "passthrough": passthrough,
        "fused": codegen.default_fused(
            low.shape, low.y_dim, len(low.kind_indices("D_Y_OVER_D_T")),
            passthrough,
        ),
        "small_threads": codegen.default_small(low.shape),
        "zrep": codegen.default_zrep(low.shape),
    }


class FDMOperator(Operator):
    def __init__(
        self,
        integrator: NumericalIntegrator,
        differentiator: NumericalDifferentiator,
        d_t: float,
    ):
        super().__init__(d_t, True)
        family = getattr(integrator, "kernel_family", "")
        if family not in ("forward_euler", "explicit_midpoint", "rk4"):
            raise NotImplementedError(
                f"{type(integrator).__name__} has no fused stage kernels; the "
                "B200 FDM path provides ForwardEulerMethod, "
                "ExplicitMidpointMethod and RK4 and does not fall back to the "
                "CPU"
            )
        self._integrator = integrator
        self._differentiator = differentiator
        self._family = family
        #: Jacobi sweeps per time step of the most recent solve (systems with
        #: LHS.Y_LAPLACIAN equations)
        self.last_jacobi_sweeps = None
        #: optional cap on Jacobi sweeps per step (0 = run to tolerance as the
        #: reference does)
        self.max_jacobi_sweeps = 0
        #: extension: keep the whole trajectory
        #: in HBM and copy it to the host only when the returned ``Solution``
        #: is read; its component planes are ``solution.device_trajectory``
        self.device_resident_solution = False
        #: extension: cut the mesh into slabs of
        #: axis 0, one per rank of ``spatial_group`` (default: all ranks of the
        #: initialised ``torch.distributed`` group), halo planes exchanged
        #: after every kernel launch; every rank returns the whole ``Solution``
        #: (gathered lazily, on first read, when ``gather_slabs`` is False)
        self.spatial_decomposition = False
        self.spatial_group = None
        self.gather_slabs = True
        #: slab solver and local trajectory (component planes of this rank's
        #: slab incl. halo planes) of the most recent decomposed solve
        self.last_slab_solver = None
        self.last_slab_trajectory = None

    # ------------------------------------------------------------------
    # plan selection
    # ------------------------------------------------------------------
    def _plan_for(self, cp, low: LoweredProblem, y0: Optional[np.ndarray]):
        return dv.get_plan(low, **plan_overrides(cp, low, y0))

    # ------------------------------------------------------------------
    # device-resident integration (also used by the Parareal fast path)
    # ------------------------------------------------------------------
    def integrate_on_device(
        self,
        cp,
        plan: "dv.DevicePlan",
        y0_planes: torch.Tensor,
        t: np.ndarray,
        traj: torch.Tensor,
        jacobi_starts: Optional[torch.Tensor] = None,
    ):
        """Advances ``y0_planes`` (component planes) through the steps
        starting at ``t[:-1]`` and writes step j to ``traj[j]``.
        ``jacobi_starts`` (systems with Y_LAPLACIAN equations, static boundary
        conditions): start values already drawn with ``_draw_jacobi_starts``
        and resident on the device, one per step."""
        low = plan.low
        n_steps = len(t) - 1
        d_t = self._d_t
        tol = getattr(self._differentiator, "_tol", 1e-3)
        sweeps_log = []
        if low.all_static or low.n_dims == 0:
            plan.bind_tables(low)
            jac = jacobi_starts
            if jac is None:
                jac = self._draw_jacobi_starts(plan, n_steps)
            s = plan.run(
                self._family, y0_planes, traj, t[:-1], d_t, 0, jac, tol,
                self.max_jacobi_sweeps,
            )
            if s is not None:
                sweeps_log.append(s)
        else:
            y_prev = y0_planes
            for first in range(0, n_steps, DYNAMIC_BC_CHUNK_STEPS):
                last = min(first + DYNAMIC_BC_CHUNK_STEPS, n_steps)
                starts = t[first:last]
                times = np.empty(3 * len(starts))
                times[0::3] = starts
                times[1::3] = starts + d_t / 2.0
                times[2::3] = starts + d_t
                dyn_neu, dyn_dir = dynamic_tables(cp, low, times)
                plan.bind_tables(low, dyn_neu, dyn_dir)
                jac = self._draw_jacobi_starts(plan, last - first)
                s = plan.run(
                    self._family, y_prev, traj[first:last], starts, d_t, 0,
                    jac, tol, self.max_jacobi_sweeps,
                )
                if s is not None:
                    sweeps_log.append(s)
                y_prev = traj[last - 1]
        self.last_jacobi_sweeps = (
            np.concatenate(sweeps_log) if sweeps_log else None
        )

    @staticmethod
    def _draw_jacobi_starts(plan, n_steps) -> Optional[torch.Tensor]:
        if not plan.n_lap:
            return None
        low = plan.low
        draws = np.empty((n_steps,) + low.shape + (plan.n_lap,))
        for j in range(n_steps):
            # one draw per time step from the global stream, like
            # numerical_differentiator.py:908-909
            draws[j] = np.random.random(low.shape + (plan.n_lap,))
        return torch.from_numpy(draws).to(plan.device)

    def prepare(self, ivp) -> Tuple:
        """Host set-up shared by ``solve`` and the benchmarks: time grid,
        initial state (with the dynamic constraints of t0 applied,
        fdm_operator.py:56-63), lowering and plan."""
        cp = ivp.constrained_problem
        t = discretize_time_domain(ivp.t_interval, self._d_t)
        low = lowered(cp)
        ic = ivp.initial_condition
        if has_device_initial_condition(ic, low):
            # the state at t0 is evaluated on the device (``initial_planes``);
            # no host array of it is ever formed
            plan = dv.get_plan(low, **plan_overrides(cp, low, None, True))
            return cp, t, None, low, plan
        dynamic = bool(low.n_dims and not low.all_static)
        view = getattr(ic, "discrete_y_0_view", None)
        y0 = None if (dynamic or view is None) else view(True)
        if y0 is None:
            y0 = ic.discrete_y_0(True)
        if dynamic:
            apply_dirichlet_host(cp, y0, float(t[0]))
        plan = self._plan_for(cp, low, y0)
        return cp, t, y0, low, plan

    @staticmethod
    def initial_planes(ivp, low: LoweredProblem, plan, y0: Optional[np.ndarray]):
        """The state at t0 as component planes on the device: uploaded from
        the host array ``y0`` of ``prepare``, or -- ``y0`` None -- evaluated
        by the initial condition's CUDA kernel."""
        if y0 is None:
            ic = ivp.initial_condition
            planes = ic.discrete_y_0_planes(plan)
            if planes is not None:
                return planes
            y0 = ic.discrete_y_0(True)  # no device form after all
        return dv.upload_state(y0, low.n_cells, low.y_dim)

    def solve_on_device(self, ivp, y0_planes: Optional[torch.Tensor] = None):
        """Solves with the whole trajectory kept in HBM.  Returns
        ``(t[1:], trajectory planes (n_steps, y_dim * n_cells))``."""
        cp, t, y0, low, plan = self.prepare(ivp)
        if y0_planes is None:
            y0_planes = self.initial_planes(ivp, low, plan, y0)
        traj = torch.empty(
            (len(t) - 1, low.y_dim * low.n_cells),
            dtype=torch.float64, device=plan.device,
        )
        self.integrate_on_device(cp, plan, y0_planes, t, traj)
        return t[1:], traj

    # ------------------------------------------------------------------
    # the drop-in entry point
    # ------------------------------------------------------------------
    def _solve_lazily(self, ivp) -> Solution:
        """The trajectory stays on the device as component planes
        ``(n_steps, y_dim * n_cells)``; reading the ``Solution`` converts it
        to the reference's channels-last layout and copies it to the host in
        chunks through pinned memory."""
        cp = ivp.constrained_problem
        t, traj = self.solve_on_device(ivp)
        low = lowered(cp)
        n_steps, state = traj.shape
        y_shape = tuple(cp.y_vertices_shape)

        def materialise() -> np.ndarray:
            host = torch.empty((n_steps, state), dtype=torch.float64, pin_memory=True)
            chunk = max(1, min(n_steps, (1 << 30) // (state * 8)))
            need_stage = low.y_dim > 1 and low.n_cells > 1
            for first in range(0, n_steps, chunk):
                part = traj[first : first + chunk]
                if need_stage:
                    part = dv.soa_to_aos(part, low.n_cells, low.y_dim, part.shape[0])
                host[first : first + part.shape[0]].copy_(part, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return host.numpy().reshape((n_steps,) + y_shape)

        sol = Solution(
            ivp, t, materialise, vertex_oriented=True, d_t=self._d_t, copy=False
        )
        sol.device_trajectory = traj
        return sol

    def _solve_decomposed(self, ivp) -> Solution:
        from pararealml_b200.operators.fdm.slab import SlabSolver

        cp = ivp.constrained_problem
        t = discretize_time_domain(ivp.t_interval, self._d_t)
        solver = SlabSolver(lowered(cp), self._family, self.spatial_group)
        y0 = ivp.initial_condition.discrete_y_0(True)
        y_dev = solver.local_planes(y0)
        traj = torch.empty(
            (len(t) - 1, solver.state), dtype=torch.float64, device=y_dev.device
        )
        solver.integrate(y_dev, t, self._d_t, traj)
        self.last_slab_solver = solver
        self.last_slab_trajectory = traj

        def gather() -> np.ndarray:
            return solver.gather(traj)

        y = gather() if self.gather_slabs else gather
        return Solution(
            ivp, t[1:], y, vertex_oriented=True, d_t=self._d_t, copy=False
        )

    def solve_batch(self, ivps, parallel_enabled: bool = True):
        """Solves several IVPs of ONE constrained problem over ONE time
        interval that differ only in their initial conditions -- what the
        data generation of the reference's ``SupervisedMLOperator`` does with
        its oracle operator (supervised_ml_operator.py:130-236, one
        ``solve`` per perturbed sub-IVP, fanned out over host processes) --
        as one batched device run: small meshes
        integrate all members in a single launch, one thread block per member.
        Returns the list of ``Solution`` objects, each bit-identical to
        ``solve(ivp)``.  IVPs that cannot be batched are solved one by one."""
        ivps = list(ivps)
        if not ivps:
            return []
        first = ivps[0]
        cp = first.constrained_problem
        low = lowered(cp)
        batchable = (
            len(ivps) > 1
            and not self.spatial_decomposition
            and all(v.constrained_problem is cp for v in ivps)
            and all(tuple(v.t_interval) == tuple(first.t_interval) for v in ivps)
            and not low.kind_indices("Y_LAPLACIAN")
            and (low.all_static or low.n_dims == 0)
        )
        if not batchable:
            return [self.solve(v, parallel_enabled) for v in ivps]
        t = discretize_time_domain(first.t_interval, self._d_t)
        n_steps, batch = len(t) - 1, len(ivps)
        state = low.y_dim * low.n_cells
        y0s = np.stack([v.initial_condition.discrete_y_0(True) for v in ivps])
        # one plan for the batch: a component that is not time-stepped may
        # only be passed through if every member satisfies the Dirichlet values
        plans = {id(self._plan_for(cp, low, y0)): self._plan_for(cp, low, y0)
                 for y0 in y0s}
        if len(plans) > 1:
            return [self.solve(v, parallel_enabled) for v in ivps]
        plan = next(iter(plans.values()))
        dev = dv.require_cuda()
        flat = dv._from_numpy(y0s.reshape(-1))
        y_dev = dv.aos_to_soa(
            dv._to_device_staged(flat, dev), low.n_cells, low.y_dim, batch
        ).view(batch, state)
        traj = torch.empty((batch, n_steps, state), dtype=torch.float64, device=dev)
        plan.bind_tables(low)
        plan.run_batch(self._family, y_dev, traj, t[:-1], self._d_t)
        host = torch.empty((batch, n_steps, state), dtype=torch.float64, pin_memory=True)
        aos = dv.soa_to_aos(traj.view(-1), low.n_cells, low.y_dim, batch * n_steps)
        host.view(-1).copy_(aos, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        y_shape = tuple(cp.y_vertices_shape)
        ys = host.numpy().reshape((batch, n_steps) + y_shape)
        return [
            Solution(v, t[1:], ys[b], vertex_oriented=True, d_t=self._d_t, copy=False)
            for b, v in enumerate(ivps)
        ]

    def solve(self, ivp, parallel_enabled: bool = True) -> Solution:
        if self.spatial_decomposition:
            return self._solve_decomposed(ivp)
        if self.device_resident_solution:
            return self._solve_lazily(ivp)
        cp, t, y0, low, plan = self.prepare(ivp)
        n_steps = len(t) - 1
        state = low.y_dim * low.n_cells
        y_shape = cp.y_vertices_shape
        host = torch.empty(
            (n_steps, state), dtype=torch.float64, pin_memory=True
        )
        y_prev = self.initial_planes(ivp, low, plan, y0)

        chunk = max(1, min(n_steps, TRAJECTORY_CHUNK_BYTES // (state * 8)))
        if state * 8 >= (64 << 20) and n_steps >= 2:
            # large states: several chunks so that the device-to-host copy of
            # one chunk overlaps the computation of the next
            chunk = min(chunk, max(1, -(-n_steps // 4)))
        n_buf = 1 if chunk >= n_steps else 2
        f64 = dict(dtype=torch.float64, device=plan.device)
        bufs = [torch.empty((chunk, state), **f64) for _ in range(n_buf)]
        need_stage = low.y_dim > 1 and low.n_cells > 1
        stages = (
            [torch.empty((chunk, state), **f64) for _ in range(n_buf)]
            if need_stage else None
        )
        compute = torch.cuda.current_stream()
        copier = torch.cuda.Stream() if n_buf > 1 else compute
        drained = [None] * n_buf
        sweeps = []
        for k, first in enumerate(range(0, n_steps, chunk)):
            last = min(first + chunk, n_steps)
            b = k % n_buf
            if drained[b] is not None:
                compute.wait_event(drained[b])
            traj = bufs[b][: last - first]
            self.integrate_on_device(cp, plan, y_prev, t[first : last + 1], traj)
            if self.last_jacobi_sweeps is not None:
                sweeps.append(self.last_jacobi_sweeps)
            y_prev = traj[last - first - 1]
            filled = torch.cuda.Event()
            filled.record(compute)
            with torch.cuda.stream(copier):
                copier.wait_event(filled)
                src = traj
                if need_stage:
                    src = dv.soa_to_aos(
                        traj, low.n_cells, low.y_dim, last - first,
                        out=stages[b][: last - first],
                    )
                host[first:last].copy_(src, non_blocking=True)
                drained[b] = torch.cuda.Event()
                drained[b].record(copier)
        copier.synchronize()
        compute.synchronize()
        self.last_jacobi_sweeps = np.concatenate(sweeps) if sweeps else None
        y = host.numpy().reshape((n_steps,) + tuple(y_shape))
        return Solution(
            ivp, t[1:], y, vertex_oriented=True, d_t=self._d_t, copy=False
        )
