"""Pins the oracle to the reference's own known-answer tests: the reference's
unit test files for the differentiator (67 tests with hand-computed expected
arrays in all four coordinate systems), the explicit integrators and the FDM
operator are executed UNMODIFIED with the oracle swapped in
(``ref_alias_plugin.py``).  Needs ``/root/reference`` (build container only)."""
import os
import subprocess
import sys

import pytest

import refshim

HERE = os.path.dirname(os.path.abspath(__file__))
REF_TESTS = os.path.join(refshim.REFERENCE_ROOT, "tests", "operators", "fdm")

pytestmark = pytest.mark.skipif(
    not refshim.available(), reason="reference checkout not present"
)


def _run(test_file, extra, env_extra=None):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(
        [HERE, os.path.dirname(HERE), env.get("PYTHONPATH", "")]
    )
    env.update(env_extra or {})
    cmd = [
        sys.executable, "-m", "pytest", "-q", "-p", "ref_alias_plugin",
        "-p", "no:cacheprovider", "--rootdir", "/tmp",
        os.path.join(REF_TESTS, test_file),
    ] + extra
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, cwd="/tmp")
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    return res.stdout


def test_reference_differentiator_tests_pass_on_oracle():
    out = _run("test_numerical_differentiator.py", [])
    assert "67 passed" in out, out[-500:]


def test_reference_explicit_integrator_tests_pass_on_oracle():
    out = _run(
        "test_numerical_integrator.py",
        ["-k", "forward_euler_method or explicit_midpoint or test_rk4"],
    )
    assert "6 passed" in out, out[-500:]


def test_reference_fdm_operator_tests_pass_on_oracle():
    out = _run(
        "test_fdm_operator.py",
        ["-k", "not conserves_density"],  # Crank-Nicolson: out of scope
        {"PML_ALIAS_FDM_OPERATOR": "1"},
    )
    assert "10 passed" in out, out[-500:]
