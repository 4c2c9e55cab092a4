"""Parity of the CUDA FDM path (through FDMOperator.solve -> C ABI) with the
reference's trajectories (golden fixtures) and with the oracle on the same
inputs.  fp64 tolerances: <= 1e-12 relative per step, <= 1e-9 on the final
state (BASELINE.json north_star); chaotic ODE cases are compared per step over
their short horizon with 1e-10."""
import numpy as np
import pytest

import oracle
import pararealml_b200 as ns
from common import load_golden, per_step_rel_err, rel_err
from golden import cases
from pararealml_b200.operators.fdm import (
    RK4,
    ExplicitMidpointMethod,
    FDMOperator,
    ForwardEulerMethod,
    ThreePointCentralDifferenceMethod,
)

pytestmark = pytest.mark.gpu

INTEGRATORS = {
    "rk4": RK4,
    "explicit_midpoint": ExplicitMidpointMethod,
    "forward_euler": ForwardEulerMethod,
}


def make_operator(case):
    return FDMOperator(
        INTEGRATORS[case.integrator](),
        ThreePointCentralDifferenceMethod(case.tol),
        case.d_t,
    )


def step_tolerance(case):
    if "jacobi" in case.tags:
        return case.rtol_traj
    return 1e-10 if case.name.startswith("lorenz") else 1e-12


@pytest.mark.parametrize("case", cases.FDM_CASES, ids=lambda c: c.name)
def test_cuda_fdm_matches_reference_golden(case):
    g = load_golden(case.name)
    ivp = case.build(ns)
    if case.seed is not None:
        np.random.seed(case.seed)
    sol = make_operator(case).solve(ivp)
    y = sol.discrete_y()
    assert y.shape == (int(g["n_steps"]),) + g["y"].shape[1:]
    assert np.array_equal(sol.t_coordinates[g["steps"]], g["t"])
    assert per_step_rel_err(y[g["steps"]], g["y"]) <= step_tolerance(case)
    assert rel_err(y[-1], g["y"][-1]) <= 1e-9


@pytest.mark.parametrize(
    "case",
    [c for c in cases.FDM_CASES if "jacobi" not in c.tags],
    ids=lambda c: c.name,
)
def test_cuda_fdm_matches_oracle_every_step(case):
    ivp = case.build(ns)
    y = make_operator(case).solve(ivp).discrete_y()
    _, y_oracle = oracle.fdm_solve(ivp, case.integrator, case.d_t, case.tol)
    assert per_step_rel_err(y, y_oracle) <= step_tolerance(case)


def test_navier_stokes_jacobi_sweep_counts_match_oracle():
    case = cases.FDM_BY_NAME["navier_stokes_2d_rk4"]
    ivp = case.build(ns)
    np.random.seed(3)
    op = make_operator(case)
    y = op.solve(ivp).discrete_y()
    stats = {}
    np.random.seed(3)
    _, y_oracle = oracle.fdm_solve(ivp, "rk4", case.d_t, case.tol, stats=stats)
    assert list(op.last_jacobi_sweeps) == stats["jacobi_sweeps"]
    assert per_step_rel_err(y, y_oracle) <= 1e-9


def test_chunked_pipelined_solve_equals_single_chunk(monkeypatch):
    """The trajectory pipeline (device chunk ring + copy stream) must not
    change results."""
    from pararealml_b200.operators.fdm import fdm_operator

    case = cases.FDM_BY_NAME["wave_2d_dynamic_rk4"]
    ivp = case.build(ns)
    whole = make_operator(case).solve(ivp).discrete_y()
    state_bytes = whole[0].size * 8
    monkeypatch.setattr(fdm_operator, "TRAJECTORY_CHUNK_BYTES", 7 * state_bytes)
    chunked = make_operator(case).solve(ivp).discrete_y()
    assert np.array_equal(whole, chunked)


def test_reference_problem_objects_are_accepted():
    """Drop-in: an IVP built from the REFERENCE's own classes (when present)
    is solved by the B200 operator."""
    import refshim

    if not refshim.available():
        pytest.skip("reference not present on this machine")
    ref = refshim.install()
    case = cases.FDM_BY_NAME["convection_diffusion_2d_mixed_rk4"]
    g = load_golden(case.name)
    y = make_operator(case).solve(case.build(ref)).discrete_y()
    assert per_step_rel_err(y[g["steps"]], g["y"]) <= 1e-12


def test_solve_on_device_keeps_the_trajectory_in_hbm():
    import torch

    case = cases.FDM_BY_NAME["burgers_3d_cartesian_rk4"]
    g = load_golden(case.name)
    ivp = case.build(ns)
    t, traj = make_operator(case).solve_on_device(ivp)
    assert traj.is_cuda and traj.dtype == torch.float64
    n = 12**3
    planes = traj.cpu().numpy().reshape(len(t), 3, n)
    y = np.moveaxis(planes, 1, 2).reshape((len(t), 12, 12, 12, 3))
    assert per_step_rel_err(y[g["steps"]], g["y"]) <= 1e-12


def test_staged_upload_of_large_states(monkeypatch):
    """Large initial states go to the device through two pinned staging
    buffers (chunked, overlapped with the DMA); forced here with tiny chunks."""
    from pararealml_b200.operators.fdm import device as dv

    monkeypatch.setattr(dv, "_UPLOAD_CHUNK", 1000)
    monkeypatch.setattr(dv, "_UPLOAD_STAGE", [])
    rng = np.random.default_rng(1)
    y = rng.normal(size=(17, 19, 23, 3))
    planes = dv.upload_state(y, 17 * 19 * 23, 3).cpu().numpy()
    expected = np.moveaxis(y.reshape(-1, 3), 1, 0).reshape(-1)
    assert np.array_equal(planes, expected)


def test_device_resident_solution_is_read_lazily():
    """the trajectory stays in HBM until the
    Solution is read; values equal the eager path's."""
    import torch

    case = cases.FDM_BY_NAME["shallow_water_polar_rk4"]
    ivp = case.build(ns)
    eager = make_operator(case).solve(ivp).discrete_y()
    op = make_operator(case)
    op.device_resident_solution = True
    sol = op.solve(ivp)
    assert sol.device_trajectory.is_cuda
    assert sol.device_trajectory.shape == (len(eager), eager[0].size)
    assert sol._y is None  # nothing copied yet
    assert np.array_equal(sol.discrete_y(), eager)
    g = load_golden(case.name)
    assert per_step_rel_err(sol.discrete_y()[g["steps"]], g["y"]) <= 1e-12


def test_diff_of_device_resident_solutions_stays_on_the_device():
    """Solution.diff (reference solution.py:182-262) between lazy solutions
    subtracts the matching steps in HBM; only the differences are copied."""
    case = cases.FDM_BY_NAME["shallow_water_polar_rk4"]
    ivp = case.build(ns)
    ops = []
    for ratio in (1, 2, 5):
        op = FDMOperator(
            INTEGRATORS[case.integrator](),
            ThreePointCentralDifferenceMethod(case.tol), case.d_t * ratio,
        )
        ops.append(op)
    eager = [op.solve(ivp) for op in ops]
    expected = eager[0].diff(eager[1:])
    for op in ops:
        op.device_resident_solution = True
    lazy = [op.solve(ivp) for op in ops]
    got = lazy[0].diff(lazy[1:])
    assert all(s._y is None for s in lazy)  # nothing was materialised
    assert np.array_equal(got.matching_time_points, expected.matching_time_points)
    assert len(got.matching_time_points) == 1  # t = 0.025 only
    assert np.isfinite(expected.differences[1]).all()
    for a, b in zip(got.differences, expected.differences):
        assert a.shape == b.shape and np.array_equal(a, b)
    # a solution that has been read falls back to the host arrays
    lazy[1].discrete_y()
    again = lazy[0].diff(lazy[1:])
    for a, b in zip(again.differences, expected.differences):
        assert np.array_equal(a, b)


@pytest.mark.parametrize(
    "case_name",
    ["diffusion_2d_rk4", "wave_2d_dynamic_mid", "cahn_hilliard_3d_rk4",
     "all_leaves_3d_spherical_mid", "navier_stokes_2d_rk4", "lorenz_rk4"],
)
def test_kernel_variants_agree(case_name, monkeypatch):
    """The three execution strategies of the stage arithmetic -- one launch
    per stage, fused stage pairs (shared-memory plane rings) and the
    single-block time loop for small meshes -- give the same trajectory."""
    case = cases.FDM_BY_NAME[case_name]
    results = {}
    for name, env in (
        ("per_stage", {"PML_SMALL": "0", "PML_FUSE": "0"}),
        ("fused", {"PML_SMALL": "0", "PML_FUSE": "1"}),
        ("single_block", {"PML_SMALL": "1", "PML_FUSE": "0"}),
    ):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ivp = case.build(ns)
        if case.seed is not None:
            np.random.seed(case.seed)
        results[name] = make_operator(case).solve(ivp).discrete_y()
    tol = 1e-9 if "jacobi" in case.tags else 1e-13
    assert per_step_rel_err(results["fused"], results["per_stage"]) <= tol
    assert per_step_rel_err(results["single_block"], results["per_stage"]) <= tol
