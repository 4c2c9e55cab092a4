"""Latency-bound configurations: K1 (examples/diffusion_1d_fdm.py as shipped)
and K2's fine/coarse solves on the 21x21 mesh; prints wall time per solve and
us per step with and without the single-block time-loop kernel."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import pararealml_b200 as ns
from golden import cases
from pararealml_b200.operators.fdm import RK4, FDMOperator, ThreePointCentralDifferenceMethod


def run(name, ivp, d_t):
    op = FDMOperator(RK4(), ThreePointCentralDifferenceMethod(), d_t)
    op.solve(ivp)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sol = op.solve(ivp)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = len(sol.t_coordinates)
    print(f"{name}: {n} steps, {dt:.3f} s wall, {dt / n * 1e6:.2f} us/step "
          f"(PML_SMALL={os.environ.get('PML_SMALL', '1')})", flush=True)


if __name__ == "__main__":
    run("K1 diffusion_1d (dynamic BCs, 101 cells)", cases.diffusion_1d_dynamic(ns, 10.0), 0.0025)
    run("K1 static twin", cases.diffusion_1d_static(ns, 10.0), 0.0025)
    run("K2 fine 21x21 (T=40, dt=1e-3)", cases.diffusion_2d(ns, 40.0), 1e-3)
    run("K2 coarse 21x21 (T=40, dt=1e-2)", cases.diffusion_2d(ns, 40.0), 1e-2)
    run("Lorenz ODE (T=10, dt=1e-3)", cases.lorenz(ns, 10.0), 1e-3)
