"""Device-resident Parareal (both operators FDMOperator, states as component
planes in HBM, updates through the pml_parareal_* kernels): one rank, and 2 / 4
ranks sharing cuda:0 through a gloo group (host-staged hand-off; the NCCL
hand-off itself is exercised by ``bench.py --gpus N``)."""
import numpy as np
import pytest

from dist_util import run_distributed

pytestmark = pytest.mark.gpu


def _solve(case_name, world_size, gather=True, speculative=True):
    import pararealml_b200 as ns
    from golden import cases
    from pararealml_b200.operators.fdm import (
        RK4,
        FDMOperator,
        ForwardEulerMethod,
        ThreePointCentralDifferenceMethod,
    )
    from pararealml_b200.operators.parareal import PararealOperator

    kinds = {"rk4": RK4, "forward_euler": ForwardEulerMethod}
    case = cases.PARAREAL_BY_NAME[case_name]
    ivp = case.build(ns)
    f = FDMOperator(kinds[case.f[0]](), ThreePointCentralDifferenceMethod(), case.f[1])
    g = FDMOperator(kinds[case.g[0]](), ThreePointCentralDifferenceMethod(), case.g[1])
    p = PararealOperator(f, g, case.tol, gather_trajectory=gather)
    p.speculative_fine_solves = speculative
    return case, p, p.solve(ivp)


def _check(case, p, sol, world_size):
    from common import load_golden, per_step_rel_err

    g = load_golden(case.name)
    y = sol.discrete_y()
    assert p.last_iterations == int(g[f"iterations_{world_size}"])
    assert p.last_slice_trajectory is not None and p.last_slice_trajectory.is_cuda
    tol = 1e-10 if "lorenz" in case.name else 1e-12
    err = per_step_rel_err(y[g[f"steps_{world_size}"]], g[f"y_{world_size}"])
    assert err <= tol, err


def _worker(rank, world_size, case_name):
    import torch

    torch.cuda.set_device(0)
    case, p, sol = _solve(case_name, world_size)
    _check(case, p, sol, world_size)


CASES = [
    "parareal_diffusion_2d_example",
    "parareal_diffusion_2d_multi_iteration",
    "parareal_lorenz",
    "parareal_burgers_3d",
]


@pytest.mark.parametrize("case_name", CASES)
def test_device_parareal_single_rank(case_name):
    case, p, sol = _solve(case_name, 1)
    _check(case, p, sol, 1)


@pytest.mark.parametrize("world_size", [2, 4])
@pytest.mark.parametrize("case_name", CASES[1:])
def test_device_parareal_multi_rank_on_one_gpu(case_name, world_size):
    run_distributed(_worker, world_size, (case_name,))


def _speculative_worker(rank, world_size, case_name):
    """Multi-block kernels (no single-block time loop): the ranks launch the
    next iteration's fine solve step by step while the convergence test is in
    flight, and drop it when the test says stop."""
    import os

    import torch

    os.environ["PML_SMALL"] = "0"
    torch.cuda.set_device(0)
    case, p, sol = _solve(case_name, world_size)
    _check(case, p, sol, world_size)
    # the same solve without speculation: bit-identical trajectory
    _, q, sol_q = _solve(case_name, world_size, speculative=False)
    assert np.array_equal(sol.discrete_y(), sol_q.discrete_y())
    assert q.last_iterations == p.last_iterations


@pytest.mark.parametrize("world_size", [2, 4])
@pytest.mark.parametrize(
    "case_name", ["parareal_diffusion_2d_multi_iteration", "parareal_burgers_3d"]
)
def test_device_parareal_speculative_fine_solves(case_name, world_size):
    run_distributed(_speculative_worker, world_size, (case_name,))


def _nccl_worker(rank, world_size, case_name):
    import torch

    assert torch.cuda.current_device() == rank
    case, p, sol = _solve(case_name, world_size)
    _check(case, p, sol, world_size)
    assert p.last_slice_trajectory.device.index == rank


@pytest.mark.parametrize("case_name", CASES[1:])
def test_device_parareal_nccl_one_rank_per_gpu(case_name):
    """The production layout: one rank per GPU, NCCL send/recv of the slice
    states and an NCCL all-reduce(MAX) convergence check (needs >= 2 GPUs)."""
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world_size = 4 if n >= 4 else 2
    run_distributed(_nccl_worker, world_size, (case_name,), backend="nccl")


def test_lazy_sharded_trajectory():
    case, p, sol = _solve("parareal_burgers_3d", 1, gather=False)
    _check(case, p, sol, 1)
