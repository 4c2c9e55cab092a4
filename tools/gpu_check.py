"""Diagnostic run of every golden FDM case on the GPU (prints a table, never
stops at the first failure).  Usage: python tools/gpu_check.py [case ...]"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import pararealml_b200 as ns  # noqa: E402
from common import load_golden, per_step_rel_err  # noqa: E402
from golden import cases  # noqa: E402
from pararealml_b200.operators.fdm import (  # noqa: E402
    RK4,
    ExplicitMidpointMethod,
    FDMOperator,
    ForwardEulerMethod,
    ThreePointCentralDifferenceMethod,
)

INTEGRATORS = {
    "rk4": RK4,
    "explicit_midpoint": ExplicitMidpointMethod,
    "forward_euler": ForwardEulerMethod,
}


def main():
    wanted = set(sys.argv[1:])
    bad = 0
    for c in cases.FDM_CASES:
        if wanted and c.name not in wanted:
            continue
        try:
            g = load_golden(c.name)
            ivp = c.build(ns)
            if c.seed is not None:
                np.random.seed(c.seed)
            op = FDMOperator(
                INTEGRATORS[c.integrator](),
                ThreePointCentralDifferenceMethod(c.tol),
                c.d_t,
            )
            t0 = time.time()
            y = op.solve(ivp).discrete_y()
            dt = time.time() - t0
            err = per_step_rel_err(y[g["steps"]], g["y"])
            ok = err <= c.rtol_traj
            bad += not ok
            extra = ""
            if op.last_jacobi_sweeps is not None:
                extra = f" sweeps={list(op.last_jacobi_sweeps)}"
            print(f"{'OK  ' if ok else 'FAIL'} {c.name:40s} err={err:.3e} {dt:.2f}s{extra}", flush=True)
        except Exception:
            bad += 1
            print(f"EXC  {c.name}", flush=True)
            traceback.print_exc()
    print("failures:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
