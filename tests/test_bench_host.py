"""The parts of ``bench.py`` that run without a GPU: the reference arm (the
oracle port timed on host cores) and the JSON contract of its line, and the
workload builders."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONTRACT_KEYS = {
    "impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
    "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
    "cpu_baseline", "e2e",
}


def _run(*args, env=None):
    res = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args],
        capture_output=True, text=True, cwd=ROOT, timeout=600,
        env={**os.environ, **(env or {})},
    )
    assert res.returncode == 0, res.stderr[-2000:]
    return res.stdout.strip().splitlines()


def test_reference_arm_single_process_line():
    lines = _run("--gpus", "1", "--steps", "2", "--warmup", "1", "--cpu-grid", "16")
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert CONTRACT_KEYS <= set(d)
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2
    assert d["value"] > 0 and d["unit"] == "Gcell-steps/s" and d["dtype"] == "f64"
    # the unmodified reference when baseline/_ref holds it, else the port
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_parareal_on_host_processes():
    """N > 1: rank 0 runs the reference's Parareal on N host processes."""
    lines = _run("--gpus", "2", "--steps", "1", "--warmup", "0",
                 "--cpu-parareal-grid", "12", "--slice-steps", "4", "--coarse-ratio", "2")
    d = json.loads(lines[-1])
    assert CONTRACT_KEYS <= set(d)
    assert d["n_gpus"] == 2 and d["cpu_baseline"]["cores"] == 2
    assert 1 <= d["config"]["parareal_iterations"] <= 2
    assert "2 time slices x 4 fine steps" in d["config"]["workload"]


def test_reference_arm_falls_back_to_the_oracle_port():
    lines = _run("--gpus", "1", "--steps", "1", "--warmup", "1", "--cpu-grid", "12",
                 env={"PML_BENCH_PORT_ONLY": "1"})
    assert json.loads(lines[0])["cpu_baseline"]["kind"] == "port"


def test_reference_arm_under_the_drivers_launcher():
    """The driver starts the arm with ``torch.distributed.run`` like the GPU
    arm; the host workers must not inherit the launcher's rendezvous (round 1:
    they hung as clients of the agent store until the driver's kill)."""
    from dist_util import free_port

    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
         "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
         "--master-port", str(free_port()), os.path.join(ROOT, "bench.py"),
         "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
         "--cpu-parareal-grid", "12", "--slice-steps", "4", "--coarse-ratio", "2"],
        capture_output=True, text=True, cwd=ROOT, timeout=300,
    )
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert CONTRACT_KEYS <= set(d) and "error" not in d
    assert d["n_gpus"] == 2 and d["value"] > 0
    assert d["config"]["wall_ms_per_solve"] > 0


def test_reference_arm_other_ranks_exit_without_work():
    lines = _run("--gpus", "2", "--steps", "1", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert lines == []


@pytest.mark.parametrize(
    "workload,n,y_dim",
    [("burgers_3d", 10, 3), ("cahn_hilliard_3d", 8, 2), ("shallow_water_polar", 12, 3),
     ("diffusion_2d", 12, 1), ("navier_stokes_2d", 12, 4)],
)
def test_workload_builders(workload, n, y_dim):
    sys.path.insert(0, ROOT)
    import bench
    import pararealml_b200 as ns

    builder, _, c, dims, _ = bench.WORKLOADS[workload]
    assert c == y_dim
    ivp, d_t = builder(ns, n, 3)
    y0 = ivp.initial_condition.discrete_y_0(True)
    assert y0.shape == (n,) * dims + (y_dim,) and np.isfinite(y0).all()
    assert d_t > 0 and np.isclose(ivp.t_interval[1], 3 * d_t)
