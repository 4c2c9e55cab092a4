"""Spatial slab decomposition (extension): several
ranks, each owning a slab of axis 0 plus two halo planes per neighbour, must
reproduce the undecomposed solve.  Ranks share cuda:0 through a gloo group
here (halo planes staged through the host); one rank per GPU over NCCL where
two GPUs exist."""
import numpy as np
import pytest

from dist_util import run_distributed

pytestmark = pytest.mark.gpu


def _problem(name):
    import pararealml_b200 as ns
    import test_gpu_fused as tf

    builder, shape = {
        "burgers_3d": (tf.burgers_3d, (26, 21, 22)),
        # enough planes per slab for the edge planes to be launched (and
        # exchanged) ahead of the interior
        "burgers_3d_tall": (tf.burgers_3d, (72, 20, 30)),
        "convection_diffusion_3d_mixed": (tf.convection_diffusion_3d_static, (23, 19, 38)),
        "diffusion_2d": (tf.diffusion_2d, (45, 50)),
        "shallow_water_polar": (tf.shallow_water_polar, (70, 300)),
    }[name]
    cp, y0, d_t = builder(shape)
    ivp = ns.InitialValueProblem(
        cp, (0.0, 3 * d_t), ns.DiscreteInitialCondition(cp, y0, True)
    )
    return ns, ivp, d_t


def _operator(integrator, d_t):
    from pararealml_b200.operators.fdm import (
        RK4, ExplicitMidpointMethod, FDMOperator, ForwardEulerMethod,
        ThreePointCentralDifferenceMethod,
    )

    kinds = {"rk4": RK4, "midpoint": ExplicitMidpointMethod, "euler": ForwardEulerMethod}
    return FDMOperator(kinds[integrator](), ThreePointCentralDifferenceMethod(), d_t)


def _worker(rank, world_size, name, integrator, fuse):
    import os

    import torch

    os.environ["PML_SMALL"] = "0"
    os.environ["PML_FUSE"] = fuse
    if torch.cuda.device_count() >= world_size and os.environ.get("PML_TEST_NCCL") == "1":
        torch.cuda.set_device(rank)
    else:
        torch.cuda.set_device(0)
    from common import per_step_rel_err

    ns, ivp, d_t = _problem(name)
    whole = _operator(integrator, d_t).solve(ivp).discrete_y()
    op = _operator(integrator, d_t)
    op.spatial_decomposition = True
    sol = op.solve(ivp)
    y = sol.discrete_y()
    assert y.shape == whole.shape
    assert np.isfinite(y).all()
    solver = op.last_slab_solver
    assert solver.size == world_size and solver.z1 - solver.z0 >= 2
    # same kernels, same operation order per cell (the warp-level choice of the
    # boundary-free instantiation may differ: FMA contraction, <= 1 ulp)
    assert per_step_rel_err(y, whole) <= 1e-14
    lazy = _operator(integrator, d_t)
    lazy.spatial_decomposition = True
    lazy.gather_slabs = False
    assert np.array_equal(lazy.solve(ivp).discrete_y(), y)


@pytest.mark.parametrize("world_size", [2, 3])
@pytest.mark.parametrize(
    "name,integrator,fuse",
    [
        ("burgers_3d", "rk4", "1"),
        ("burgers_3d", "rk4", "0"),
        ("burgers_3d_tall", "rk4", "1"),
        ("burgers_3d_tall", "midpoint", "1"),
        ("burgers_3d", "midpoint", "1"),
        ("convection_diffusion_3d_mixed", "rk4", "1"),
        ("diffusion_2d", "rk4", "1"),
        ("diffusion_2d", "euler", "1"),
        ("shallow_water_polar", "rk4", "1"),
    ],
)
def test_slabs_reproduce_the_undecomposed_solve(name, integrator, fuse, world_size):
    run_distributed(_worker, world_size, (name, integrator, fuse))


def _split_worker(rank, world_size):
    import torch

    torch.cuda.set_device(0)
    ns, ivp, d_t = _problem("burgers_3d_tall")
    op = _operator("rk4", d_t)
    op.spatial_decomposition = True
    op.solve(ivp)
    edges, rest = op.last_slab_solver.edge_ranges()
    assert len(edges) == (1 if rank in (0, world_size - 1) else 2)
    assert rest[1] - rest[0] >= 8


def test_tall_slabs_launch_their_edge_planes_first():
    run_distributed(_split_worker, 3, ())


def test_plane_range_launches_equal_one_launch(monkeypatch):
    """pml_fdm_phase_planes: launches over disjoint plane ranges, in any order,
    write what one launch over the whole mesh writes (a plane may pass through
    another of the three unrolled copies of the loop body, whose FMA
    contraction can differ: <= 1 ulp per operation, like between slabs)."""
    import ctypes

    import torch

    from pararealml_b200 import _native
    from pararealml_b200.operators.fdm import codegen
    from pararealml_b200.operators.fdm import device as dv
    from pararealml_b200.operators.fdm.fdm_operator import lowered

    monkeypatch.setenv("PML_SMALL", "0")
    monkeypatch.setenv("PML_FZC", "16")
    ns, ivp, d_t = _problem("burgers_3d_tall")
    low = lowered(ivp.constrained_problem)
    plan = dv.get_plan(
        low, passthrough=False, small_threads=0, zrep=1,
        fused=codegen.default_fused(low.shape, low.y_dim, low.y_dim, False),
    )
    assert plan.spec.fused is not None
    plan.bind_tables(low)
    ws = plan.workspace()
    lib = _native.lib()
    y0 = dv.upload_state(ivp.initial_condition.discrete_y_0(True), low.n_cells, low.y_dim)
    n0 = low.shape[0]
    results = []
    for ranges in ([(0, n0)], [(n0 - 9, n0), (0, 5), (5, 40), (40, n0 - 9)]):
        y, fresh = y0, ctypes.c_void_p()
        bufs = {t.data_ptr(): t for t in plan._ws_bufs.values()}
        y_next = torch.zeros_like(y0)
        for phase in range(2):
            for z_begin, z_end in ranges:
                _native.check(lib.pml_fdm_phase_planes(
                    plan.handle, _native.INTEGRATOR_CODES["rk4"], ctypes.byref(ws),
                    y0.data_ptr(), y_next.data_ptr(), 0.0, float(d_t), 0, phase,
                    z_begin, z_end, ctypes.byref(fresh), dv.stream_ptr()))
        torch.cuda.synchronize()
        results.append(y_next.cpu().numpy())
    assert np.isfinite(results[0]).all() and np.abs(results[0]).max() > 0
    err = np.abs(results[0] - results[1]).max() / np.abs(results[0]).max()
    assert err <= 1e-14
    with pytest.raises(RuntimeError, match="plane range"):
        _native.check(lib.pml_fdm_phase_planes(
            plan.handle, _native.INTEGRATOR_CODES["rk4"], ctypes.byref(ws),
            y0.data_ptr(), y_next.data_ptr(), 0.0, float(d_t), 0, 0, 5, n0 + 1,
            ctypes.byref(fresh), dv.stream_ptr()))


def test_slabs_over_nccl_one_rank_per_gpu(monkeypatch):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    monkeypatch.setenv("PML_TEST_NCCL", "1")
    run_distributed(_worker, 2, ("burgers_3d_tall", "rk4", "1"), backend="nccl")
