"""Multi-rank Parareal on CPU: world_size 2 and 4 ``gloo`` process groups run
the pipelined driver (generic host path) with the oracle's FDM operator as
fine and coarse solver; trajectories and iteration counts must equal the
reference's (golden fixtures made with a fake-MPI run of the unmodified
reference)."""
import numpy as np
import pytest

from dist_util import run_distributed


def _worker(rank, world_size, case_name):
    import oracle
    import pararealml_b200 as ns
    from common import load_golden, per_step_rel_err
    from golden import cases
    from pararealml_b200.operators.parareal import PararealOperator

    case = cases.PARAREAL_BY_NAME[case_name]
    g = load_golden(case.name)
    ivp = case.build(ns)
    p = PararealOperator(
        oracle.OracleFDMOperator(*case.f), oracle.OracleFDMOperator(*case.g),
        case.tol,
    )
    y = p.solve(ivp).discrete_y()
    assert p.last_iterations == int(g[f"iterations_{world_size}"]), (
        p.last_iterations, int(g[f"iterations_{world_size}"]))
    assert len(y) == int(g[f"n_steps_{world_size}"])
    err = per_step_rel_err(y[g[f"steps_{world_size}"]], g[f"y_{world_size}"])
    assert err <= 1e-12, err


@pytest.mark.parametrize("world_size", [2, 4])
@pytest.mark.parametrize(
    "case_name",
    ["parareal_diffusion_2d_multi_iteration", "parareal_lorenz", "parareal_burgers_3d"],
)
def test_parareal_matches_reference_on_gloo_ranks(case_name, world_size):
    run_distributed(_worker, world_size, (case_name,))


def _spmd_oracle_worker(rank, world_size, case_name):
    import oracle
    import pararealml_b200 as ns
    from common import load_golden, per_step_rel_err
    from golden import cases
    from oracle.parareal_ranks import GlooComm, parareal_rank_solve

    case = cases.PARAREAL_BY_NAME[case_name]
    g = load_golden(case.name)
    ivp = case.build(ns)

    def sub_ivp(cp, interval, y0):
        return ns.InitialValueProblem(
            cp, interval, ns.DiscreteInitialCondition(cp, y0, True)
        )

    _, y, iterations = parareal_rank_solve(
        GlooComm(), ivp, oracle.OracleFDMOperator(*case.f),
        oracle.OracleFDMOperator(*case.g), case.tol, sub_ivp,
    )
    assert iterations == int(g[f"iterations_{world_size}"])
    err = per_step_rel_err(y[g[f"steps_{world_size}"]], g[f"y_{world_size}"])
    assert err <= 1e-12, err


def test_spmd_oracle_used_by_the_reference_arm_matches_golden():
    """``oracle/parareal_ranks.py`` (timed by ``bench.py --impl reference
    --gpus N``) reproduces the reference's multi-rank trajectories."""
    run_distributed(
        _spmd_oracle_worker, 2, ("parareal_diffusion_2d_multi_iteration",)
    )


def _callable_worker(rank, world_size):
    import oracle
    import pararealml_b200 as ns
    from golden import cases
    from pararealml_b200.operators.parareal import PararealOperator

    case = cases.PARAREAL_BY_NAME["parareal_diffusion_2d_multi_iteration"]
    ivp = case.build(ns)
    seen = []

    def condition(old, new):
        seen.append((old.shape, new.shape))
        return len(seen) >= 2

    p = PararealOperator(
        oracle.OracleFDMOperator(*case.f), oracle.OracleFDMOperator(*case.g),
        condition,
    )
    lazy = PararealOperator(
        oracle.OracleFDMOperator(*case.f), oracle.OracleFDMOperator(*case.g),
        condition, gather_trajectory=False,
    )
    y = p.solve(ivp).discrete_y()
    assert p.last_iterations == 2
    assert seen[0][0] == (world_size, 21, 21, 1)
    seen.clear()
    y_lazy = lazy.solve(ivp).discrete_y()
    assert np.array_equal(y, y_lazy)


def test_callable_termination_and_lazy_gather_on_two_ranks():
    run_distributed(_callable_worker, 2)


def test_single_rank_and_validation():
    import oracle
    import pararealml_b200 as ns
    from common import load_golden, per_step_rel_err
    from golden import cases
    from pararealml_b200.operators.parareal import PararealOperator

    case = cases.PARAREAL_BY_NAME["parareal_lorenz"]
    g = load_golden(case.name)
    ivp = case.build(ns)
    f = oracle.OracleFDMOperator(*case.f)
    gg = oracle.OracleFDMOperator(*case.g)
    p = PararealOperator(f, gg, case.tol)
    y = p.solve(ivp).discrete_y()
    assert per_step_rel_err(y[g["steps_1"]], g["y_1"]) <= 1e-12
    # serial mode is the fine operator itself (reference :105-106)
    assert np.array_equal(
        p.solve(ivp, parallel_enabled=False).discrete_y(),
        f.solve(ivp).discrete_y(),
    )
    with pytest.raises(ValueError):
        PararealOperator(oracle.OracleFDMOperator("rk4", 0.3), gg, 1e-3).solve(ivp)
    with pytest.raises(ValueError):
        PararealOperator(f, oracle.OracleFDMOperator("rk4", 0.3), 1e-3).solve(ivp)
    with pytest.raises(ValueError):
        PararealOperator(f, gg, [1e-3, 1e-3]).solve(ivp)
