"""Benchmark of the B200 FDM + Parareal hot path (driver contract).

N = 1   : fp64 RK4 ``FDMOperator`` time steps of the 3-D Burgers equation on a
          512^3 mesh (BASELINE.json configs[4] grid, the largest configuration
          that fits one GPU); a "step" is one RK4 time step (4 fused stage
          kernels).  ``value`` is timed with the state resident in HBM,
          ``e2e`` goes through ``FDMOperator.solve`` with host buffers.
N > 1   : ``PararealOperator`` (fine RK4, coarse ForwardEuler) on the same
          problem with one time slice per GPU (weak scaling: fixed fine steps
          per slice); a "step" is one Parareal solve and ``value`` counts the
          cell-steps of the fine-resolution trajectory it produces.

``--impl reference`` times the oracle port of the reference's NumPy path
(``oracle/``; the reference itself is pure Python and cannot travel to the GPU
box) on a bounded sample of the same workload on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fp64 FDM cell-steps/s (3-D Burgers 512^3, RK4)"
UNIT = "Gcell-steps/s"


def metric_name(workload, n):
    if workload == "burgers_3d" and n == 512:
        return METRIC
    return f"fp64 FDM cell-steps/s ({workload} {n}, RK4)"


def parareal_metric(workload, n):
    """N > 1: a Parareal solve, one time slice per GPU -- its own metric name
    (the value still counts cell-steps of the fine trajectory per second)."""
    return (f"fp64 Parareal cell-steps/s ({workload} {n}, fine RK4 / coarse "
            "ForwardEuler, one time slice per GPU)")


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--grid", type=int, default=int(os.environ.get("PML_BENCH_GRID", "0")),
                   help="vertices per axis (0 = the workload's named size)")
    p.add_argument("--workload", default="burgers_3d", choices=list(WORKLOADS))
    p.add_argument("--e2e-steps", type=int, default=8)
    p.add_argument("--slice-steps", type=int, default=16,
                   help="fine steps per Parareal time slice")
    p.add_argument("--coarse-ratio", type=int, default=4,
                   help="coarse step = ratio * fine step")
    p.add_argument("--parareal-tol", type=float, default=1e-6,
                   help="RMS end point update tolerance (states are O(1))")
    p.add_argument("--cpu-grid", type=int, default=96)
    p.add_argument("--reference-timeout", type=float, default=240.0,
                   help="wall-clock bound (s) of the reference arm's host work")
    p.add_argument("--cpu-parareal-grid", type=int, default=40,
                   help="vertices per axis of the host-process Parareal sample")
    p.add_argument("--jacobi-sweeps", type=int, default=100,
                   help="cap on Jacobi sweeps per step (navier_stokes_2d; the "
                        "reference iterates to tolerance, 1e3..1e7 sweeps)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-workloads", action="store_true",
                   help="N = 1: skip the other named configurations")
    p.add_argument("--no-parity", action="store_true",
                   help="skip the parity checks that precede the timed region")
    p.add_argument("--parity-grid-3d", type=int, default=256)
    p.add_argument("--parity-grid-2d", type=int, default=2048)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-stiff", action="store_true",
                   help="N > 1: skip the Parareal solve at the harder setting")
    p.add_argument("--stiff-coarse-ratio", type=int, default=8)
    p.add_argument("--stiff-tol", type=float, default=1e-9)
    p.add_argument("--no-spatial", action="store_true",
                   help="N > 1: skip the slab-decomposed solve")
    p.add_argument("--spatial-steps", type=int, default=20)
    return p.parse_args()


# ---------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------
def burgers_problem(ns, n, n_steps, d_t=None):
    """K5: BurgersEquation(3, 100) on [0,1]^3 with n^3 vertices, zero-flux
    boundaries, GaussianInitialCondition (mean 0.5, covariance 0.05 I per
    component).  The time step keeps the explicit
    scheme stable (d_t <= d_x^2 Re / 6).  On the B200 operators a mesh of this
    size evaluates the Gaussian on the device (``pml_ic_gaussian``); the
    reference evaluates it with SciPy on the host."""
    eq = ns.BurgersEquation(3, 100.0)
    h = 1.0 / (n - 1)
    mesh = ns.Mesh([(0.0, 1.0)] * 3, [h] * 3)
    bc = ns.NeumannBoundaryCondition(
        lambda x, t: np.zeros((len(x), 3)), is_static=True
    )
    cp = ns.ConstrainedProblem(eq, mesh, [(bc, bc)] * 3)
    if d_t is None:
        d_t = 0.1 * h * h * 100.0 / 6.0
    ic = ns.GaussianInitialCondition(
        cp, [(np.full(3, 0.5), 0.05 * np.eye(3))] * 3, [0.3, -0.2, 0.1]
    )
    ivp = ns.InitialValueProblem(cp, (0.0, n_steps * d_t), ic)
    return ivp, d_t


def with_host_state(ns, ivp, planes, low, dv):
    """The same IVP with its initial state as a host array behind a
    ``DiscreteInitialCondition`` (the end-to-end legs copy their input from
    host memory; the state itself may have been evaluated on the device)."""
    aos = dv.soa_to_aos(planes, low.n_cells, low.y_dim)
    y0 = aos.cpu().numpy().reshape(tuple(low.shape) + (low.y_dim,))
    cp = ivp.constrained_problem
    return ns.InitialValueProblem(
        cp, ivp.t_interval, ns.DiscreteInitialCondition(cp, y0, True)
    )


def cahn_hilliard_problem(ns, n, n_steps, d_t=None):
    """K3: CahnHilliardEquation(3, gamma=0.5) on an n^3 mesh with unit
    spacing, zero-flux boundaries, uniform random concentration
    (examples/cahn_hilliard_3d_fdm.py)."""
    gamma = 0.5
    eq = ns.CahnHilliardEquation(3, gamma=gamma)
    mesh = ns.Mesh([(1.0, float(n))] * 3, [1.0] * 3)
    bc = ns.NeumannBoundaryCondition(
        lambda x, t: np.zeros((len(x), 2)), is_static=True
    )
    cp = ns.ConstrainedProblem(eq, mesh, [(bc, bc)] * 3)
    rng = np.random.default_rng(0)
    y0 = np.empty((n, n, n, 2))
    c = 0.05 * rng.uniform(-1.0, 1.0, (n, n, n))
    lap = np.zeros_like(c)
    for a in range(3):
        p = np.concatenate(
            [np.take(c, [1], axis=a), c, np.take(c, [-2], axis=a)], axis=a
        )
        idx = [slice(None)] * 3
        lo, mid, hi = list(idx), list(idx), list(idx)
        lo[a], mid[a], hi[a] = slice(0, -2), slice(1, -1), slice(2, None)
        lap += p[tuple(hi)] - 2.0 * p[tuple(mid)] + p[tuple(lo)]
    y0[..., 0] = c
    y0[..., 1] = c**3 - c - gamma * lap
    if d_t is None:
        d_t = 0.05
    ic = ns.DiscreteInitialCondition(cp, y0, True)
    return ns.InitialValueProblem(cp, (0.0, n_steps * d_t), ic), d_t


def shallow_water_problem(ns, n, n_steps, d_t=None):
    """K4a: ShallowWaterEquation(0.5) on a polar n x n mesh
    (examples/shallow_water_polar_fdm.py)."""
    eq = ns.ShallowWaterEquation(0.5)
    mesh = ns.Mesh(
        [(4.0, 11.0), (0.5 * np.pi, 1.5 * np.pi)],
        [7.0 / (n - 1), np.pi / (n - 1)],
        ns.CoordinateSystem.POLAR,
    )
    bc = ns.NeumannBoundaryCondition(
        lambda x, t: np.stack(
            [np.zeros(len(x)), np.full(len(x), np.nan), np.full(len(x), np.nan)],
            axis=-1,
        ),
        is_static=True,
    )
    cp = ns.ConstrainedProblem(eq, mesh, [(bc, bc)] * 2)
    r, th = np.meshgrid(*mesh.vertex_axis_coordinates, indexing="ij")
    x, y = r * np.cos(th), r * np.sin(th)
    y0 = np.zeros((n, n, 3))
    y0[..., 0] = np.exp(-0.5 * ((x + 6.0) ** 2 + (y - 6.0) ** 2) / 0.25) / (
        2.0 * np.pi * 0.25
    )
    if d_t is None:
        h = min(7.0 / (n - 1), 4.0 * np.pi / (n - 1))
        d_t = 0.05 * h * h / 0.1
    ic = ns.DiscreteInitialCondition(cp, y0, True)
    return ns.InitialValueProblem(cp, (0.0, n_steps * d_t), ic), d_t


def diffusion_2d_problem(ns, n, n_steps, d_t=None):
    """K2 scaled: examples/diffusion_2d_parareal.py on an n x n mesh."""
    eq = ns.DiffusionEquation(2)
    h = 10.0 / (n - 1)
    mesh = ns.Mesh([(0.0, 10.0)] * 2, [h, h])
    dirichlet = ns.DirichletBoundaryCondition(
        lambda x, t: np.full((len(x), 1), 1.5), is_static=True
    )
    neumann = ns.NeumannBoundaryCondition(
        lambda x, t: np.zeros((len(x), 1)), is_static=True
    )
    cp = ns.ConstrainedProblem(
        eq, mesh, [(dirichlet, dirichlet), (neumann, neumann)]
    )
    x = np.linspace(0.0, 10.0, n)
    g = np.exp(-0.5 * (x - 5.0) ** 2) / np.sqrt(2.0 * np.pi)
    y0 = (1000.0 * g[:, None] * g[None, :])[..., None]
    if d_t is None:
        d_t = 0.2 * h * h
    ic = ns.DiscreteInitialCondition(cp, y0, True)
    return ns.InitialValueProblem(cp, (0.0, n_steps * d_t), ic), d_t


def navier_stokes_problem(ns, n, n_steps, d_t=None):
    """K4b: NavierStokesEquation(5000) (vorticity / stream function / velocity,
    LHS = [D_Y_OVER_D_T, Y_LAPLACIAN, Y, Y]) with the boundary conditions of
    examples/navier_stokes_fdm.py on an n x n mesh, fluid at rest."""
    eq = ns.NavierStokesEquation(5000.0)
    mesh = ns.Mesh([(-2.5, 2.5), (0.0, 4.0)], [5.0 / (n - 1), 4.0 / (n - 1)])
    wall = ns.DirichletBoundaryCondition(
        ns.vectorize_bc_function(lambda x, t: [0.0, 0.0, None, None]),
        is_static=True,
    )
    inlet = ns.DirichletBoundaryCondition(
        ns.vectorize_bc_function(lambda x, t: [1.0, 0.1, None, None]),
        is_static=True,
    )
    cp = ns.ConstrainedProblem(eq, mesh, [(inlet, wall), (wall, wall)])
    if d_t is None:
        d_t = 0.1 * (4.0 / (n - 1)) ** 2 * 5000.0 / 4.0
        d_t = min(d_t, 0.02 * 4.0 / (n - 1))
    ic = ns.DiscreteInitialCondition(cp, np.zeros((n, n, 4)), True)
    return ns.InitialValueProblem(cp, (0.0, n_steps * d_t), ic), d_t


# name -> (builder, default vertices per axis, y_dim, spatial dims, label)
WORKLOADS = {
    "burgers_3d": (burgers_problem, 512, 3, 3,
                   "BurgersEquation(3, Re=100), zero-flux boundaries, Gaussian initial condition"),
    "cahn_hilliard_3d": (cahn_hilliard_problem, 256, 2, 3,
                         "CahnHilliardEquation(3, gamma=0.5), zero-flux boundaries, random concentration"),
    "shallow_water_polar": (shallow_water_problem, 4096, 3, 2,
                            "ShallowWaterEquation(0.5) on a polar mesh, zero-flux height"),
    "diffusion_2d": (diffusion_2d_problem, 2048, 1, 2,
                     "DiffusionEquation(2), Dirichlet 1.5 / zero-flux boundaries"),
    "navier_stokes_2d": (navier_stokes_problem, 4096, 4, 2,
                         "NavierStokesEquation(Re=5000), inlet / no-slip walls, "
                         "Jacobi stream-function solve capped per step"),
}


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    QUERY = (
        "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
        "clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,"
        "clocks_event_reasons.sw_power_cap"
    )
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            parts = [s.strip() for s in line.split(",")]
            if len(parts) >= 6:
                self.rows.append(parts)

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(self.NAMES, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {
            "sm_mhz": float(np.median(sm)),
            "sm_max_mhz": float(max(mx)),
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


# ---------------------------------------------------------------------------
# CPU baseline (oracle port of the reference's NumPy path)
# ---------------------------------------------------------------------------
def cpu_baseline(workload, n, n_steps):
    import oracle
    import pararealml_b200 as ns

    builder, _, _, dims, _ = WORKLOADS[workload]
    ivp, d_t = builder(ns, n, n_steps)
    t0 = time.perf_counter()
    oracle.fdm_solve(ivp, "rk4", d_t)
    dt = time.perf_counter() - t0
    return n**dims * n_steps / dt / 1e9, dt


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """dram bytes per step from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    return {}


# ---------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------
#: launcher variables a spawned host worker must not inherit: under
#: ``torch.distributed.run`` TORCHELASTIC_USE_AGENT_STORE makes every
#: ``init_process_group`` a TCPStore *client* of the agent's store
_LAUNCHER_ENV = (
    "RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK",
    "ROLE_RANK", "ROLE_WORLD_SIZE", "GROUP_WORLD_SIZE", "MASTER_ADDR",
    "MASTER_PORT", "ROLE_NAME", "OMP_NUM_THREADS",
)


def reference_namespace():
    """``(ns, fdm_ops, parareal_cls, kind)``: the UNMODIFIED reference from
    ``baseline/_ref`` (offline pip install made by ``build()``; or
    ``PML_REFERENCE_ROOT``) behind the import shim of ``tests/refshim.py``
    when present -- ``kind`` "reference" -- else ``None`` and the caller falls
    back to the oracle port.  ``/root/reference`` is never read here."""
    if os.environ.get("PML_BENCH_PORT_ONLY") == "1":
        return None
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import refshim
    except ImportError:
        return None
    root = refshim.REFERENCE_ROOT
    if os.path.abspath(root).startswith("/root/reference") or not refshim.available():
        return None
    try:
        ref = refshim.install()
        from pararealml.operators import fdm as ref_fdm
        from pararealml.operators.parareal import PararealOperator as RefParareal
    except Exception:  # an unusable install must not take the arm down
        return None
    return ref, ref_fdm, RefParareal, refshim


def _reference_parareal_worker(rank, world, port, n, slice_steps, ratio, tol,
                               steps, warmup, out):
    """One host process = one time slice of the reference's mpirun layout."""
    for k in list(os.environ):
        if k in _LAUNCHER_ENV or k.startswith("TORCHELASTIC_"):
            del os.environ[k]
    import torch
    import torch.distributed as dist

    torch.set_num_threads(1)
    store = dist.TCPStore("127.0.0.1", port, world, is_master=(rank == 0))
    dist.init_process_group("gloo", store=store, rank=rank, world_size=world)

    found = reference_namespace()
    if found is not None:
        ref, ref_fdm, RefParareal, refshim = found
        refshim.set_comm(refshim.GlooComm())
        ivp, d_t = burgers_problem(ref, n, world * slice_steps)
        tcd = ref_fdm.ThreePointCentralDifferenceMethod
        f = ref_fdm.FDMOperator(ref_fdm.RK4(), tcd(), d_t)
        g = ref_fdm.FDMOperator(ref_fdm.ForwardEulerMethod(), tcd(), d_t * ratio)
        p = RefParareal(f, g, tol)
        count = [0]
        inner = p._should_terminate

        def counting(old, new):
            count[0] += 1
            return inner(old, new)

        p._should_terminate = counting

        def solve():
            count[0] = 0
            p.solve(ivp)
            return count[0]

        kind = "reference"
    else:
        import oracle
        import pararealml_b200 as ns
        from oracle.parareal_ranks import GlooComm, parareal_rank_solve

        ivp, d_t = burgers_problem(ns, n, world * slice_steps)
        f = oracle.OracleFDMOperator("rk4", d_t)
        g = oracle.OracleFDMOperator("forward_euler", d_t * ratio)
        comm = GlooComm()

        def sub_ivp(cp, interval, y0):
            return ns.InitialValueProblem(
                cp, interval, ns.DiscreteInitialCondition(cp, y0, True)
            )

        def solve():
            return parareal_rank_solve(comm, ivp, f, g, tol, sub_ivp)[2]

        kind = "port"

    iterations = 0
    for _ in range(warmup):
        solve()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        iterations = solve()
    dist.barrier()
    if rank == 0:
        out.put((time.perf_counter() - t0, iterations, kind))
    dist.destroy_process_group()


def run_reference_parareal(args):
    """N > 1: the reference's Parareal on N host processes, one per time
    slice, the way ``mpirun -n N`` runs it (Makefile:36-37): the unmodified
    reference from ``baseline/_ref`` with a gloo group behind the ``mpi4py``
    shim, else the oracle port (no MPI runtime exists in the image)."""
    import socket

    import torch.multiprocessing as mp

    world = args.gpus
    n = args.cpu_parareal_grid
    steps = max(1, min(args.steps, 3))
    warmup = min(args.warmup, 1)
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    spawned = mp.spawn(
        _reference_parareal_worker,
        args=(world, port, n, args.slice_steps, args.coarse_ratio,
              args.parareal_tol, steps, warmup, out),
        nprocs=world, join=False,
    )
    # watchdog: the arm must print a line even if a worker hangs
    deadline = time.monotonic() + args.reference_timeout
    error = None
    try:
        while not spawned.join(timeout=1.0):
            if time.monotonic() > deadline:
                error = f"host workers exceeded {args.reference_timeout} s"
                for proc in spawned.processes:
                    if proc.is_alive():
                        proc.kill()
                break
    except Exception as exc:
        error = f"{type(exc).__name__}: {str(exc)[:300]}"
    if error is None and not out.empty():
        dt, iterations, kind = out.get()
    else:
        error = error or "no result from the host workers"
        dt, iterations, kind = float("nan"), 0, "port"
    total_steps = world * args.slice_steps
    value = None if error else n**3 * total_steps * steps / dt / 1e9
    wall_ms = None if error else dt / steps * 1e3
    impl = (
        "the unmodified reference PararealOperator (baseline/_ref, gloo "
        "Allgather behind the mpi4py shim)"
        if kind == "reference"
        else "oracle port of the reference PararealOperator (gloo Allgather)"
    )
    sample = (
        f"{impl}: f = NumPy FDM RK4, g = ForwardEuler at {args.coarse_ratio} "
        f"d_t, {world} host processes standing in for mpirun -n {world} "
        f"({os.cpu_count()} host cores), 3-D Burgers on {n}^3 (bounded sample "
        f"of the 512^3 workload), {world} slices x {args.slice_steps} fine "
        f"steps, {iterations} Parareal iterations per solve, {steps} timed "
        f"solves"
    )
    line = {
        "impl": "reference",
        "metric": parareal_metric(args.workload, 512),
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": steps,
        "warmup": warmup,
        "ms_per_step": wall_ms,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": f"PararealOperator(f=FDM RK4 d_t, g=FDM ForwardEuler "
                        f"{args.coarse_ratio} d_t) on 3-D Burgers, {n}^3 sample "
                        f"of the 512^3 workload, {world} time slices x "
                        f"{args.slice_steps} fine steps, tol {args.parareal_tol}",
            "parareal_iterations": iterations,
            "wall_ms_per_solve": wall_ms,
        },
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": world, "kind": kind,
            "sample": sample,
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if error:
        line["error"] = error
    print(json.dumps(line), flush=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.gpus > 1:
        run_reference_parareal(args)
        return
    n = args.cpu_grid
    found = reference_namespace()
    if found is not None:
        ref, ref_fdm, _, _ = found
        ivp, d_t = burgers_problem(ref, n, 1)
        op = ref_fdm.FDMOperator(
            ref_fdm.RK4(), ref_fdm.ThreePointCentralDifferenceMethod(), d_t
        )
        ns = ref

        def solve(sub):
            return op.solve(sub).discrete_y()[-1]

        kind = "reference"
        what = "the unmodified reference FDMOperator (baseline/_ref)"
    else:
        import pararealml_b200 as ns
        from oracle.fdm import fdm_solve

        ivp, d_t = burgers_problem(ns, n, 1)

        def solve(sub):
            return fdm_solve(sub, "rk4", d_t)[1][-1]

        kind = "port"
        what = "oracle port of the reference NumPy FDMOperator"

    cp = ivp.constrained_problem
    y = ivp.initial_condition.discrete_y_0(True)

    def one_step(y_in):
        sub = ns.InitialValueProblem(
            cp, (0.0, d_t), ns.DiscreteInitialCondition(cp, y_in, True)
        )
        return solve(sub)

    # bounded: the whole run must end within a few minutes on any host
    budget = time.monotonic() + args.reference_timeout
    steps_done = 0
    for _ in range(args.warmup):
        y = one_step(y)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        y = one_step(y)
        steps_done += 1
        if time.monotonic() > budget:
            break
    dt = time.perf_counter() - t0
    value = n**3 * steps_done / dt / 1e9
    sample = (
        f"{what} (RK4, ThreePointCentralDifferenceMethod), 3-D Burgers on "
        f"{n}^3 (bounded sample of the 512^3 workload), {steps_done} steps, "
        f"single process ({os.cpu_count()} host cores present; the "
        "reference's NumPy stencils are single-threaded)"
    )
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": steps_done,
        "warmup": args.warmup,
        "ms_per_step": dt / steps_done * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"3-D Burgers RK4 FDM, {n}^3 sample of the 512^3 workload"},
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": sample,
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# parity checks at benchmark scale (outside every timed region; the oracle is
# used here only as the checker)
# ---------------------------------------------------------------------------
def _rel(a, b):
    scale = float(np.max(np.abs(b)))
    return float(np.max(np.abs(a - b)) / (scale if scale else 1.0))


def parity_single_gpu(args, ns, FDMOperator, RK4, TCD, dv, torch):
    """N = 1: (1) one RK4 step of the 256^3 Burgers problem and of the 2048^2
    polar shallow-water problem through ``FDMOperator.solve`` against the
    oracle (<= 1e-12, SURVEY.md section 8d "single step at full size where RAM
    allows"); (2) one step of the benchmark mesh itself, fused stage-pair
    kernels against the one-launch-per-stage kernels on the device
    (<= 1e-14) -- this exercises the chunking, the grid and the > 2^31-byte
    offsets of the timed run."""
    import oracle
    from pararealml_b200.operators.fdm.fdm_operator import plan_overrides

    out = []
    for workload, n in (("burgers_3d", args.parity_grid_3d),
                        ("shallow_water_polar", args.parity_grid_2d)):
        if n <= 0:
            continue
        builder = WORKLOADS[workload][0]
        ivp, d_t = builder(ns, n, 1)
        op = FDMOperator(RK4(), TCD(), d_t)
        y = op.solve(ivp).discrete_y()
        plan = op.prepare(ivp)[-1]
        t0 = time.perf_counter()
        _, y_ref = oracle.fdm_solve(ivp, "rk4", d_t)
        secs = time.perf_counter() - t0
        err = _rel(y[0], y_ref[0])
        out.append({
            "case": f"{workload} {'x'.join([str(n)] * WORKLOADS[workload][3])}, "
                    "one RK4 step, FDMOperator.solve vs oracle.fdm_solve",
            "kernels": "fused stage pairs" if plan.fused is not None else "stage kernels",
            "max_rel_err": err, "tolerance": 1e-12, "ok": bool(err <= 1e-12),
            "oracle_seconds": round(secs, 1),
        })
        del y, y_ref
    # the benchmark mesh: fused pairs vs stage kernels, on the device
    builder, n_default = WORKLOADS[args.workload][:2]
    n = args.grid or n_default
    ivp, d_t = builder(ns, n, 1)
    op = FDMOperator(RK4(), TCD(), d_t)
    if args.workload == "navier_stokes_2d":
        op.max_jacobi_sweeps = args.jacobi_sweeps
    cp, t, y0, low, plan = op.prepare(ivp)
    if plan.fused is not None:
        ov = plan_overrides(cp, low, y0, y0 is None)
        ov["fused"] = None
        plan_u = dv.get_plan(low, **ov)
        y_dev = op.initial_planes(ivp, low, plan, y0)
        res = []
        for pl in (plan, plan_u):
            traj = torch.empty((1, low.y_dim * low.n_cells), dtype=torch.float64,
                               device="cuda")
            op.integrate_on_device(cp, pl, y_dev, t, traj)
            res.append(traj)
        torch.cuda.synchronize()
        scale = float(res[1].abs().max().item())
        err = float((res[0] - res[1]).abs().max().item()) / (scale if scale else 1.0)
        out.append({
            "case": f"{args.workload} {'x'.join(str(v) for v in low.shape)}, one "
                    "RK4 step on the device, fused stage pairs vs stage kernels",
            "max_rel_err": err, "tolerance": 1e-14, "ok": bool(err <= 1e-14),
        })
        del res, traj, y_dev
        torch.cuda.empty_cache()
    return out


def parity_multi_gpu(world, rank, ns, FDMOperator, RK4, FE, TCD, Parareal, torch):
    """N > 1, over the live NCCL group: (1) the multi-rank golden trajectories
    of the unmodified reference (tests/golden, P = world) -- trajectory error
    and iteration count; (2) a Parareal solve on a mesh that runs the fused
    stage-pair kernels against the oracle's P-rank emulation; (3) the
    slab-decomposed solve against the undecomposed one."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from common import load_golden, per_step_rel_err
    from golden import cases

    kinds = {"rk4": RK4, "forward_euler": FE}
    out = []
    for name in ("parareal_burgers_3d", "parareal_diffusion_2d_multi_iteration"):
        case = cases.PARAREAL_BY_NAME[name]
        if world not in case.sizes:
            continue
        ivp = case.build(ns)
        p = Parareal(FDMOperator(kinds[case.f[0]](), TCD(), case.f[1]),
                     FDMOperator(kinds[case.g[0]](), TCD(), case.g[1]), case.tol)
        y = p.solve(ivp).discrete_y()
        g = load_golden(name)
        err = per_step_rel_err(y[g[f"steps_{world}"]], g[f"y_{world}"])
        expected = int(g[f"iterations_{world}"])
        out.append({
            "case": f"{name}, P = {world}, NCCL, vs the reference's golden trajectory",
            "max_rel_err": err, "tolerance": 1e-12,
            "iterations": p.last_iterations, "iterations_expected": expected,
            "ok": bool(err <= 1e-12 and p.last_iterations == expected),
        })
    # fused kernels + NCCL hand-off against the oracle's rank emulation
    shape = (24, 28, 32)
    eq = ns.BurgersEquation(3, 100.0)
    mesh = ns.Mesh([(0.0, 1.0)] * 3, [1.0 / (m - 1) for m in shape])
    bc = ns.NeumannBoundaryCondition(lambda x, t: np.zeros((len(x), 3)), is_static=True)
    cp = ns.ConstrainedProblem(eq, mesh, [(bc, bc)] * 3)
    rng = np.random.default_rng(17)
    y0 = rng.uniform(-1.0, 1.0, shape + (3,))
    d_t = 2e-5
    ivp = ns.InitialValueProblem(cp, (0.0, world * 4 * d_t),
                                 ns.DiscreteInitialCondition(cp, y0, True))
    f = FDMOperator(RK4(), TCD(), d_t)
    g = FDMOperator(FE(), TCD(), 2 * d_t)
    p = Parareal(f, g, 1e-9)
    y = p.solve(ivp).discrete_y()
    plan = f.prepare(ivp)[-1]

    def sub_ivp(cp_, interval, y_start):
        return ns.InitialValueProblem(
            cp_, interval, ns.DiscreteInitialCondition(cp_, y_start, True))

    _, y_ref, its = oracle.parareal_solve(
        ivp, oracle.OracleFDMOperator("rk4", d_t),
        oracle.OracleFDMOperator("forward_euler", 2 * d_t), 1e-9, world, sub_ivp)
    err = per_step_rel_err(y, y_ref)
    out.append({
        "case": f"Parareal 3-D Burgers {'x'.join(map(str, shape))}, P = {world}, "
                "NCCL, vs oracle.parareal_solve",
        "kernels": "fused stage pairs" if plan.fused is not None else "stage kernels",
        "max_rel_err": err, "tolerance": 1e-12,
        "iterations": p.last_iterations, "iterations_expected": int(its),
        "ok": bool(err <= 1e-12 and p.last_iterations == its),
    })
    # slabs of axis 0 against the undecomposed solve (same kernels, one GPU)
    shape = (16 * world, 24, 32)
    mesh = ns.Mesh([(0.0, 1.0)] * 3, [1.0 / (m - 1) for m in shape])
    cp = ns.ConstrainedProblem(eq, mesh, [(bc, bc)] * 3)
    y0 = np.random.default_rng(19).uniform(-1.0, 1.0, shape + (3,))
    ivp = ns.InitialValueProblem(cp, (0.0, 3 * d_t),
                                 ns.DiscreteInitialCondition(cp, y0, True))
    whole = FDMOperator(RK4(), TCD(), d_t).solve(ivp).discrete_y()
    cut = FDMOperator(RK4(), TCD(), d_t)
    cut.spatial_decomposition = True
    y = cut.solve(ivp).discrete_y()
    err = per_step_rel_err(y, whole)
    out.append({
        "case": f"slab-decomposed 3-D Burgers {'x'.join(map(str, shape))}, "
                f"{world} slabs, NCCL halo exchange, vs the undecomposed solve",
        "max_rel_err": err, "tolerance": 1e-14, "ok": bool(err <= 1e-14),
    })
    return out


# ---------------------------------------------------------------------------
# modelled DRAM traffic and the other named configurations
# ---------------------------------------------------------------------------
def _stencil_components(exprs):
    """y components the given right-hand sides read (any leaf kind)."""
    comps = set()
    for e in exprs:
        for sym in e.free_symbols:
            tokens = sym.name.split("_")
            kind, idx = tokens[0], [int(v) for v in tokens[1:]]
            if kind in ("t", "x"):
                continue
            if kind in ("y", "y-gradient", "y-hessian", "y-laplacian"):
                comps.add(idx[0])
            elif kind == "y-divergence":
                comps.update(idx)
            elif kind == "y-curl":
                comps.update(idx if len(idx) == 2 else idx[:-1])
            elif kind == "y-vector-laplacian":
                comps.update(idx[:-1])
    return comps


def modelled_bytes_per_cell_step(low, plan):
    """Compulsory DRAM bytes per cell and RK4 step of the schedule the plan
    runs (MODELLED from the plan, not measured: halo re-reads, boundary
    tables and coordinate vectors are not counted).  Fused stage pairs move
    7 doubles per time-stepped component (1+2: r y, w u3, w acc; 3+4: r u3,
    r y, r acc, w y+).  Stage kernels move, per stage, every component a
    right-hand side reads once, the step-start value and the accumulator of
    the time-stepped components, and the stage's outputs; components that are
    not time-stepped are read in place when the plan passes them through."""
    dt = low.kind_indices("D_Y_OVER_D_T")
    other = [i for i in range(low.y_dim) if i not in dt]
    n_dt = len(dt)
    if plan.fused is not None:
        return 8 * 7 * n_dt
    rhs_dt = [low.rhs[i] for i in dt]
    rhs_aux = [low.rhs[i] for i in other]
    read_dt = _stencil_components(rhs_dt)
    read_first = read_dt | _stencil_components(rhs_aux) | set(dt)
    passthrough = bool(plan.spec.passthrough)
    carried = 0 if passthrough else len(other)  # copied through every stage
    doubles = 0
    # stage 1: reads y (all needed comps once), writes u, acc, aux outputs
    doubles += len(read_first) + 2 * n_dt + len(other) + carried
    # stages 2, 3: read stage input comps, y and acc of the dt comps; write u, acc
    per_mid = len(read_dt | (set() if passthrough else set(other))) + 2 * n_dt
    per_mid += 2 * n_dt + carried
    # (dt comps that no right-hand side reads are not loaded as stencil input)
    doubles += 2 * per_mid
    # stage 4: as above without the u / acc writes
    doubles += len(read_dt) + 2 * n_dt + n_dt
    return 8 * doubles


def run_workload_line(name, n, steps, warmup, args, ns, FDMOperator, RK4, TCD, dv,
                      torch, peak):
    """One bounded device-resident measurement of a named configuration."""
    builder, _, y_dim, dims, label = WORKLOADS[name]
    cells = n**dims
    total = warmup + steps
    ivp, d_t = builder(ns, n, total)
    op = FDMOperator(RK4(), TCD(), d_t)
    jacobi = name == "navier_stokes_2d"
    if jacobi:
        op.max_jacobi_sweeps = args.jacobi_sweeps
        np.random.seed(0)
    cp, t, y0, low, plan = op.prepare(ivp)
    y_dev = op.initial_planes(ivp, low, plan, y0)
    traj = torch.empty((total, y_dim * cells), dtype=torch.float64, device="cuda")
    op.integrate_on_device(cp, plan, y_dev, t[: warmup + 1], traj[:warmup])
    # the reference draws the Jacobi start of every step from NumPy's global
    # stream on the host; drawn and uploaded before the timed region
    starts = op._draw_jacobi_starts(plan, steps) if jacobi else None
    torch.cuda.synchronize()
    l0 = dv.total_launches()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    op.integrate_on_device(cp, plan, traj[warmup - 1], t[warmup:], traj[warmup:],
                           jacobi_starts=starts)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    launches = (dv.total_launches() - l0) // steps
    finite = bool(torch.isfinite(traj[-1]).all().item())
    real = modelled_bytes_per_cell_step(low, plan)
    alg = 128 * y_dim
    line = {
        "name": name, "workload": f"{label}, {'x'.join([str(n)] * dims)} mesh, RK4",
        "value": cells / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms,
        "steps": steps, "launches_per_step": launches,
        "kernels": "fused stage pairs" if plan.fused is not None else "stage kernels",
        "algorithmic_bytes_per_cell_step": alg,
        "frac": alg * cells / (ms * 1e-3) / 1e9 / peak,
        "modelled_bytes_per_cell_step": real,
        "frac_real": real * cells / (ms * 1e-3) / 1e9 / peak,
        "finite": finite,
    }
    if jacobi:
        sweeps = float(np.mean(op.last_jacobi_sweeps))
        line["jacobi_sweeps_per_step_cap"] = args.jacobi_sweeps
        line["jacobi_sweeps_per_step"] = sweeps
        line["note"] = ("ms_per_step = explicit stages + the capped Jacobi "
                        "stream-function solve (the reference iterates to "
                        "tolerance: not reference-matching at this size); frac "
                        "counts 24 B per cell-sweep on top of the stages")
        jb = 24 * sweeps
        line["frac"] = (alg + jb) * cells / (ms * 1e-3) / 1e9 / peak
        line["frac_real"] = (real + jb) * cells / (ms * 1e-3) / 1e9 / peak
        # the Jacobi kernel alone: a fixed number of sweeps, tolerance 0
        n_sw = 400
        rhs = torch.zeros(cells, dtype=torch.float64, device="cuda")
        init = torch.rand(cells, dtype=torch.float64, device="cuda")
        out = torch.empty(y_dim * cells, dtype=torch.float64, device="cuda")
        plan.jacobi(rhs, init, out, 0.0, 16)
        torch.cuda.synchronize()
        a.record()
        done = plan.jacobi(rhs, init, out, 0.0, n_sw)
        b.record()
        torch.cuda.synchronize()
        ms_sw = a.elapsed_time(b) / max(done, 1)
        line["jacobi"] = {
            "sweeps": done, "ms_per_sweep": ms_sw,
            "value": cells / (ms_sw * 1e-3) / 1e9, "unit": "Gcell-sweeps/s",
            "frac": 24 * cells / (ms_sw * 1e-3) / 1e9 / peak,
            "bytes_per_cell_sweep": 24,
        }
    del traj, y_dev
    torch.cuda.empty_cache()
    return line


def small_mesh_latency(ns, FDMOperator, RK4, TCD, torch):
    """K1 (examples/diffusion_1d_fdm.py, 101 vertices, dynamic boundary
    conditions) and the K2 example's fine solver (21 x 21): microseconds per
    time step through ``FDMOperator.solve`` (single-block time loop)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden import cases

    out = []
    for label, ivp, d_t in (
        ("K1 examples/diffusion_1d_fdm.py (101 vertices, dynamic BCs, T = 10)",
         cases.diffusion_1d_dynamic(ns, 10.0), 0.0025),
        ("K1 static-boundary twin", cases.diffusion_1d_static(ns, 10.0), 0.0025),
        ("K2 example fine solver (21 x 21, T = 40, d_t = 1e-3)",
         cases.diffusion_2d(ns, 40.0), 1e-3),
    ):
        op = FDMOperator(RK4(), TCD(), d_t)
        op.solve(ivp)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sol = op.solve(ivp)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n_steps = len(sol.t_coordinates)
        out.append({"case": label, "steps": n_steps,
                    "us_per_step": dt / n_steps * 1e6, "wall_s": dt})
    return out


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import pararealml_b200 as ns
    from pararealml_b200.operators.fdm import (
        RK4,
        FDMOperator,
        ForwardEulerMethod,
        ThreePointCentralDifferenceMethod,
    )
    from pararealml_b200.operators.fdm import device as dv
    from pararealml_b200.operators.parareal import PararealOperator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    builder, n_default, y_dim, dims, label = WORKLOADS[args.workload]
    n = args.grid or n_default
    cells = n**dims
    grid = [n] * dims
    peak, peak_src = peak_hbm()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    TCD = ThreePointCentralDifferenceMethod
    if world == 1:
        parity = None
        if not args.no_parity:
            try:
                parity = parity_single_gpu(args, ns, FDMOperator, RK4, TCD, dv, torch)
            except Exception as exc:  # the measured line must survive
                parity = [{"error": f"{type(exc).__name__}: {str(exc)[:300]}", "ok": False}]
            torch.cuda.empty_cache()
        # ---- device-resident RK4 steps --------------------------------
        total = args.warmup + args.steps
        ivp, d_t = builder(ns, n, total)
        op = FDMOperator(RK4(), ThreePointCentralDifferenceMethod(), d_t)
        if args.workload == "navier_stokes_2d":
            op.max_jacobi_sweeps = args.jacobi_sweeps
            np.random.seed(0)
        cp, t, y0, low, plan = op.prepare(ivp)
        y_dev = op.initial_planes(ivp, low, plan, y0)
        ivp_host = with_host_state(ns, ivp, y_dev, low, dv)
        traj = torch.empty((total, y_dim * cells), dtype=torch.float64, device="cuda")
        op.integrate_on_device(cp, plan, y_dev, t[: args.warmup + 1], traj[: args.warmup])
        barrier()
        launches0 = dv.total_launches()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clocks:
            start.record()
            op.integrate_on_device(
                cp, plan, traj[args.warmup - 1], t[args.warmup:], traj[args.warmup:]
            )
            stop.record()
            barrier()
        ms = start.elapsed_time(stop)
        launches = dv.total_launches() - launches0
        finite = bool(torch.isfinite(traj[-1]).all().item())
        value = cells * args.steps / (ms * 1e-3) / 1e9
        alg_bytes_step = 128 * y_dim * cells
        sweeps = None
        if op.last_jacobi_sweeps is not None:
            # + 24 B per cell and Jacobi sweep
            sweeps = float(np.mean(op.last_jacobi_sweeps))
            alg_bytes_step += int(24 * sweeps * cells)
        achieved = alg_bytes_step * args.steps / (ms * 1e-3) / 1e9
        del traj, y_dev
        torch.cuda.empty_cache()

        # ---- end to end through FDMOperator.solve (host buffers) -----------
        e2e = None
        if not args.no_e2e:
            d_t_e = d_t
            ivp_e = ns.InitialValueProblem(
                ivp_host.constrained_problem, (0.0, args.e2e_steps * d_t_e),
                ivp_host.initial_condition,
            )
            op_e = FDMOperator(RK4(), ThreePointCentralDifferenceMethod(), d_t_e)
            if args.workload == "navier_stokes_2d":
                op_e.max_jacobi_sweeps = args.jacobi_sweeps
            op_e.solve(ivp_e)  # warm-up (pinned buffers, plan)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sol = op_e.solve(ivp_e)
            torch.cuda.synchronize()
            dt_e = time.perf_counter() - t0
            state_bytes = cells * y_dim * 8
            e2e = {
                "value": cells * args.e2e_steps / dt_e / 1e9,
                "unit": UNIT,
                "h2d_bytes_per_step": state_bytes // args.e2e_steps,
                "d2h_bytes_per_step": state_bytes,
                "steps": args.e2e_steps,
                "note": "FDMOperator.solve(ivp): H2D of y0, device time loop, "
                        "D2H of every step of the trajectory into the returned "
                        "Solution",
            }
            del sol
            # the same solve with the trajectory left in HBM (lazy Solution):
            # only the final state is read back
            try:
                op_e.device_resident_solution = True
                last = torch.empty(y_dim * cells, dtype=torch.float64, pin_memory=True)
                dt_l = None
                for _ in range(2):  # the first pass primes the allocators
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    sol = op_e.solve(ivp_e)
                    last.copy_(sol.device_trajectory[-1], non_blocking=True)
                    torch.cuda.synchronize()
                    dt_l = time.perf_counter() - t0
                    del sol
                e2e["device_resident"] = {
                    "value": cells * args.e2e_steps / dt_l / 1e9,
                    "unit": UNIT,
                    "h2d_bytes_per_step": state_bytes // args.e2e_steps,
                    "d2h_bytes_per_step": state_bytes // args.e2e_steps,
                    "note": "FDMOperator.device_resident_solution = True: the "
                            "trajectory stays in HBM behind a lazy Solution, only "
                            "the final state is copied to the host",
                }
            except Exception as exc:  # the eager line must survive
                e2e["device_resident"] = {"error": f"{type(exc).__name__}: {exc}"}
                last = None

            del last
        cpu = None
        if not args.no_cpu_baseline and args.workload != "navier_stokes_2d":
            # (the oracle's Jacobi solve iterates to tolerance: no bounded
            # Navier-Stokes sample exists)
            cpu_n = args.cpu_grid if dims == 3 else 8 * args.cpu_grid
            v, secs = cpu_baseline(args.workload, cpu_n, 2)
            cpu = {
                "value": v, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"oracle port of the reference NumPy RK4 FDM path, "
                          f"{args.workload} on {'x'.join([str(cpu_n)] * dims)}, "
                          f"2 steps, {secs:.1f} s, single process "
                          f"({os.cpu_count()} host cores present)",
            }
        profile = load_traffic()
        traffic = profile.get(f"{args.workload}_{n}_rk4_step_dram_bytes")
        kernels = profile.get(f"{args.workload}_{n}_kernels")
        modelled = modelled_bytes_per_cell_step(low, plan)
        workloads, latency = None, None
        if not args.no_workloads and args.workload == "burgers_3d":
            workloads = []
            for w_name, w_n, w_steps in (
                ("cahn_hilliard_3d", 256, 20), ("shallow_water_polar", 4096, 10),
                ("navier_stokes_2d", 4096, 3), ("diffusion_2d", 2048, 50),
            ):
                try:
                    workloads.append(run_workload_line(
                        w_name, w_n, w_steps, 3, args, ns, FDMOperator, RK4, TCD,
                        dv, torch, peak))
                except Exception as exc:
                    workloads.append({"name": w_name,
                                      "error": f"{type(exc).__name__}: {str(exc)[:300]}"})
            try:
                latency = small_mesh_latency(ns, FDMOperator, RK4, TCD, torch)
            except Exception as exc:
                latency = [{"error": f"{type(exc).__name__}: {str(exc)[:300]}"}]
        line = {
            "metric": metric_name(args.workload, n),
            "value": value,
            "unit": UNIT,
            "n_gpus": 1,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": f"FDMOperator(RK4, ThreePointCentralDifferenceMethod) "
                            f"on {label}, {'x'.join(str(v) for v in grid)} mesh",
                "name": args.workload,
                "grid": grid, "y_dim": y_dim, "d_t": d_t,
                "cache": f"state ({cells * y_dim * 8 / 1e9:.2f} GB) is larger "
                         "than the 126 MB L2",
                "finite": finite,
            },
            "clocks": clocks.summary(),
            "roofline": {
                "bound": "hbm",
                "achieved": achieved,
                "peak": peak,
                "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": traffic,
                "peak_source": peak_src,
                "algorithmic_bytes_per_cell_step": alg_bytes_step // cells,
                "modelled_bytes_per_cell_step": modelled,
                "frac_real_modelled": modelled * cells * args.steps / (ms * 1e-3) / 1e9 / peak,
                "traffic_source": "ncu capture of an earlier run (profiles/traffic.json); "
                                  "the modelled figure is computed from this run's plan",
                "jacobi_sweeps_per_step": sweeps,
                "scope": "one time step = all stage(-pair) launches of the "
                         "timed region; achieved = algorithmic bytes of a step "
                         "/ step time, traffic = measured DRAM bytes of a step "
                         "(ncu, profiles/traffic.json)",
                "kernels_ncu": kernels,
                "launches_per_step": launches // max(args.steps, 1),
            },
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": launches,
            "parity": parity,
            "workloads": workloads,
            "small_mesh_latency": latency,
        }
        print(json.dumps(line), flush=True)
        return

    # ---- N > 1: Parareal, one time slice per GPU ----------------------------
    parity = None
    if not args.no_parity:
        try:
            parity = parity_multi_gpu(world, rank, ns, FDMOperator, RK4,
                                      ForwardEulerMethod, TCD, PararealOperator, torch)
        except Exception as exc:  # the measured line must survive
            parity = [{"error": f"{type(exc).__name__}: {str(exc)[:300]}", "ok": False}]
        torch.cuda.empty_cache()
    s_steps = args.slice_steps
    total_steps = world * s_steps
    ivp, d_t = builder(ns, n, total_steps)
    f = FDMOperator(RK4(), ThreePointCentralDifferenceMethod(), d_t)
    g = FDMOperator(
        ForwardEulerMethod(), ThreePointCentralDifferenceMethod(),
        d_t * args.coarse_ratio,
    )
    p = PararealOperator(f, g, args.parareal_tol, gather_trajectory=False)
    # value: the initial state is resident in HBM before the timed region
    cp_f, _, y0_f, low_f, plan_f = f.prepare(ivp)
    y0_planes = f.initial_planes(ivp, low_f, plan_f, y0_f)
    # the serial alternative: one GPU stepping the fine operator through one
    # slice (x world slices); timed here on every rank, max over ranks
    t_slice = np.arange(s_steps + 1) * d_t
    scratch = torch.empty((s_steps, y_dim * cells), dtype=torch.float64, device="cuda")
    f.integrate_on_device(cp_f, plan_f, y0_planes, t_slice[:3], scratch[:2])
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    f.integrate_on_device(cp_f, plan_f, y0_planes, t_slice, scratch)
    f1.record()
    barrier()
    fine_ms = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(fine_ms, op=dist.ReduceOp.MAX)
    fine_slice_ms = float(fine_ms.item())
    del scratch
    torch.cuda.empty_cache()
    for _ in range(args.warmup):
        p.solve_on_device(ivp, y0_planes)
    barrier()
    launches0 = dv.total_launches()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        start.record()
        for _ in range(args.steps):
            p.solve_on_device(ivp, y0_planes)
        stop.record()
        barrier()
    ms = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = dv.total_launches() - launches0
    iterations = p.last_iterations

    # ---- the same solve at a harder setting (coarser coarse propagator, tighter
    # tolerance): how wall time and iteration count move when convergence is
    # not nearly free -------------------------------------------------------
    stiff = None
    if not args.no_stiff:
        try:
            p.last_slice_trajectory = None
            torch.cuda.empty_cache()
            g_s = FDMOperator(
                ForwardEulerMethod(), ThreePointCentralDifferenceMethod(),
                d_t * args.stiff_coarse_ratio,
            )
            p_s = PararealOperator(f, g_s, args.stiff_tol, gather_trajectory=False)
            p_s.solve_on_device(ivp, y0_planes)  # warm-up (coarse plan)
            barrier()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record()
            p_s.solve_on_device(ivp, y0_planes)
            q1.record()
            barrier()
            ms_q = torch.tensor([q0.elapsed_time(q1)], dtype=torch.float64, device="cuda")
            dist.all_reduce(ms_q, op=dist.ReduceOp.MAX)
            stiff = {
                "coarse_ratio": args.stiff_coarse_ratio, "tol": args.stiff_tol,
                "wall_ms_per_solve": float(ms_q.item()),
                "iterations": p_s.last_iterations,
                "bound_p_over_k": world / max(p_s.last_iterations, 1),
                "speedup_vs_serial_fine": fine_slice_ms * world / float(ms_q.item()),
                "update_norms": [[float(v) for v in row] for row in p_s.last_update_norms],
                "wasted_speculative_fine_steps_rank0": p_s.last_wasted_fine_steps,
            }
            p_s.last_slice_trajectory = None
            del p_s
        except Exception as exc:  # the main line must survive
            stiff = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}

    e2e = None
    # the timed solves' slice trajectory (slice_steps x 3.2 GB) must not stay
    # resident next to the one the end-to-end solve allocates
    p.last_slice_trajectory = None
    # the remaining legs start from host memory
    ivp = with_host_state(ns, ivp, y0_planes, low_f, dv)
    del y0_planes
    torch.cuda.empty_cache()
    if not args.no_e2e:
        # the reference's final Allgather would put the whole trajectory
        # (world x slice_steps x 3.2 GB) on every rank; the trajectory stays
        # sharded instead and every rank reads back its own time slice
        pe = PararealOperator(f, g, args.parareal_tol, gather_trajectory=False)
        state_bytes = cells * y_dim * 8
        # read back through a bounded pinned staging buffer (8 ranks x 51 GB of
        # pinned host memory would be unreasonable): every byte of the slice
        # crosses PCIe, the host copy is not retained
        chunk = min(s_steps, 4)
        host = torch.empty((chunk, y_dim * cells), dtype=torch.float64,
                           pin_memory=True)
        barrier()
        t0 = time.perf_counter()
        sol = pe.solve(ivp)
        traj_local = pe.last_slice_trajectory
        for first in range(0, s_steps, chunk):
            part = traj_local[first:first + chunk]
            host[: part.shape[0]].copy_(part, non_blocking=True)
        barrier()
        dt_e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt_e, op=dist.ReduceOp.MAX)
        e2e = {
            "value": cells * total_steps / float(dt_e.item()) / 1e9,
            "unit": UNIT,
            "h2d_bytes_per_step": state_bytes * world,
            "d2h_bytes_per_step": state_bytes * total_steps,
            "note": "PararealOperator(gather_trajectory=False).solve(ivp): "
                    "H2D of y0 on every rank, Parareal iterations, D2H of "
                    "every rank's own slice of the trajectory (component "
                    "planes) through a pinned staging buffer; bytes are "
                    "whole-job totals per solve",
        }
        del sol, host
    # ---- the same mesh cut into slabs of axis 0, one per GPU (strong scaling
    # of the single-GPU workload; extension beyond the reference) -------------
    spatial = None
    if not args.no_spatial:
        try:
            from pararealml_b200.operators.fdm.fdm_operator import lowered
            from pararealml_b200.operators.fdm.slab import SlabSolver

            torch.cuda.empty_cache()
            k_steps = max(args.spatial_steps, 1)
            solver = SlabSolver(lowered(ivp.constrained_problem), "rk4")
            y_loc = solver.local_planes(ivp.initial_condition.discrete_y_0_view(True))  # (host state)
            traj_s = torch.empty((args.warmup + k_steps, solver.state),
                                 dtype=torch.float64, device="cuda")
            t_s = np.arange(args.warmup + k_steps + 1) * d_t
            solver.integrate(y_loc, t_s[: args.warmup + 1], d_t, traj_s[: args.warmup])
            barrier()
            launches_s0 = dv.total_launches()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            solver.integrate(traj_s[args.warmup - 1], t_s[args.warmup:], d_t,
                             traj_s[args.warmup:])
            s1.record()
            barrier()
            ms_s = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device="cuda")
            dist.all_reduce(ms_s, op=dist.ReduceOp.MAX)
            ms_s = float(ms_s.item())
            finite = bool(torch.isfinite(solver.owned(traj_s[-1:])).all().item())
            spatial = {
                "metric": "fp64 FDM cell-steps/s, one 512^3 RK4 solve cut into "
                          "slabs of axis 0 (one per GPU, 2 halo planes exchanged "
                          "over NCCL after every stage-pair launch)",
                "value": cells * k_steps / (ms_s * 1e-3) / 1e9,
                "unit": UNIT,
                "ms_per_step": ms_s / k_steps,
                "steps": k_steps,
                "scaling": "strong",
                "planes_per_rank": solver.z1 - solver.z0,
                "halo_bytes_per_launch_and_neighbour": 8 * y_dim * 2 * solver.plane,
                "kernel_launches_per_rank": dv.total_launches() - launches_s0,
                "finite": finite,
            }
            del traj_s, y_loc, solver
        except Exception as exc:  # the Parareal line must survive
            spatial = {"error": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        value = cells * total_steps * args.steps / (ms * 1e-3) / 1e9
        wall_ms = ms / args.steps
        serial_ms = fine_slice_ms * world
        line = {
            "metric": parareal_metric(args.workload, n),
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": f"PararealOperator(f=FDM RK4 d_t, g=FDM ForwardEuler "
                            f"{args.coarse_ratio} d_t) on {label}, "
                            f"{'x'.join(str(v) for v in grid)} mesh, {world} time "
                            f"slices x {s_steps} fine steps, tol {args.parareal_tol}",
                "name": args.workload,
                "grid": grid, "y_dim": y_dim, "d_t": d_t,
                "parareal_iterations": iterations,
                "parareal_update_norms": [
                    [float(v) for v in row] for row in p.last_update_norms
                ],
                "peak_hbm_allocated_gb_rank0": round(
                    torch.cuda.max_memory_allocated() / 1e9, 1),
                "cache": f"state ({cells * y_dim * 8 / 1e9:.2f} GB) is larger "
                         "than the 126 MB L2",
            },
            "clocks": clocks.summary(),
            "roofline": {
                "bound": "hbm",
                "achieved": 128 * y_dim * value,
                "peak": peak * world,
                "unit": "GB/s",
                "frac": 128 * y_dim * value / (peak * world),
                "traffic": None,
                "peak_source": peak_src,
                "note": "useful fine-trajectory bytes only; Parareal repeats "
                        "fine solves once per iteration",
            },
            "cpu_baseline": None,
            "e2e": e2e,
            "gpu_launches": launches,
            "parareal": {
                "wall_ms_per_solve": wall_ms,
                "fine_slice_ms": fine_slice_ms,
                "serial_fine_ms": serial_ms,
                "speedup_vs_serial_fine": serial_ms / wall_ms,
                "slices": world,
                "fine_steps_per_slice": s_steps,
                "iterations": iterations,
                "bound_p_over_k": world / max(iterations, 1),
                "harder_setting": stiff,
                "note": "serial_fine_ms = the fine operator stepping all "
                        f"{total_steps} steps on one GPU (the slice solve timed "
                        "on every rank, max over ranks, x slices); the Parareal "
                        "speed-up over it is bounded by P / K",
            },
            "parity": parity,
            "spatial_decomposition": spatial,
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def main():
    args = parse_args()
    # the timed region continues from the last warm-up step: at least one
    args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
