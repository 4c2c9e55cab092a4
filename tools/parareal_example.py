"""examples/diffusion_2d_parareal.py of the reference, as shipped (21x21 mesh,
T = 40, fine RK4 d_t = 1e-3, coarse RK4 d_t = 1e-2, tol 0.0025), on the B200
operators.  Run under torchrun with one rank per GPU:

    torchrun --nproc-per-node P --master-addr 127.0.0.1 tools/parareal_example.py

Prints (rank 0) the wall times of the fine, coarse and Parareal solves like the
reference's ``mpi_time`` decorator (barrier - timer - barrier)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

import pararealml_b200 as ns
from golden import cases
from pararealml_b200.operators.fdm import RK4, FDMOperator, ThreePointCentralDifferenceMethod
from pararealml_b200.operators.parareal import PararealOperator


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    ivp = cases.diffusion_2d(ns, 40.0)
    f = FDMOperator(RK4(), ThreePointCentralDifferenceMethod(), 0.001)
    g = FDMOperator(RK4(), ThreePointCentralDifferenceMethod(), 0.01)
    p = PararealOperator(f, g, 0.0025)

    def timed(name, fn, repeat=3):
        fn()  # warm-up (plans, pinned buffers)
        best = float("inf")
        for _ in range(repeat):
            barrier()
            t0 = time.perf_counter()
            fn()
            barrier()
            best = min(best, time.perf_counter() - t0)
        if rank == 0:
            print(f"{name}: {best * 1e3:.1f} ms", flush=True)

    timed("fine", lambda: f.solve(ivp))
    timed("coarse", lambda: g.solve(ivp))
    timed(f"parareal (P={world})", lambda: p.solve(ivp))
    if rank == 0:
        print("parareal iterations:", p.last_iterations)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
