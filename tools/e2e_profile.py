import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
import pararealml_b200 as ns
from pararealml_b200.operators.fdm import RK4, FDMOperator, ThreePointCentralDifferenceMethod
from pararealml_b200.operators.fdm import device as dv
ivp, d_t = bench.burgers_problem(ns, 512, 8)
op = FDMOperator(RK4(), ThreePointCentralDifferenceMethod(), d_t)
op.solve(ivp)
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    cp, t, y0, low, plan = op.prepare(ivp)
    t1 = time.perf_counter()
    yd = dv.upload_state(y0, low.n_cells, low.y_dim); torch.cuda.synchronize()
    t2 = time.perf_counter()
    host = torch.empty((8, low.y_dim*low.n_cells), dtype=torch.float64, pin_memory=True)
    t3 = time.perf_counter()
    print(f"prepare {t1-t0:.3f} upload {t2-t1:.3f} pinned alloc {t3-t2:.3f}")
    del host
    t0 = time.perf_counter(); sol = op.solve(ivp); torch.cuda.synchronize(); print(f"solve total {time.perf_counter()-t0:.3f}")
    del sol
