"""Forward Euler steps per second on the 512^3 Burgers workload (the coarse
propagator of the Parareal benchmark): python tools/fe_bench.py [n] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import pararealml_b200 as ns  # noqa: E402
from pararealml_b200.operators.fdm import (  # noqa: E402
    FDMOperator, ForwardEulerMethod, ThreePointCentralDifferenceMethod,
)
from pararealml_b200.operators.fdm import device as dv  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ivp, d_t = bench.burgers_problem(ns, n, steps + 2)
op = FDMOperator(ForwardEulerMethod(), ThreePointCentralDifferenceMethod(), 4 * d_t)
cp, t, y0, low, plan = op.prepare(ivp)
y = op.initial_planes(ivp, low, plan, y0)
traj = torch.empty((steps, 3 * n**3), dtype=torch.float64, device="cuda")
import numpy as np  # noqa: E402
tt = np.arange(steps + 1) * 4 * d_t
op.integrate_on_device(cp, plan, y, tt[:3], traj[:2])
torch.cuda.synchronize()
l0 = dv.total_launches()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
op.integrate_on_device(cp, plan, y, tt, traj)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
print(f"forward Euler {n}^3: {ms:.3f} ms/step, {n**3 / ms / 1e6:.1f} Gcell-steps/s, "
      f"{dv.total_launches() - l0} launches for {steps} steps, fused={plan.fused is not None}")
