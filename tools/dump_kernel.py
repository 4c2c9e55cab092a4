"""Writes the generated CUDA source and the NVRTC cubin of a benchmark workload
to a directory (no GPU needed): python tools/dump_kernel.py burgers_3d 512 /tmp/k
Then: cuobjdump -sass /tmp/k/kernel.cubin"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import pararealml_b200 as ns  # noqa: E402
from pararealml_b200 import _native  # noqa: E402
from pararealml_b200.operators.fdm import codegen  # noqa: E402
from pararealml_b200.operators.fdm.fdm_operator import plan_overrides  # noqa: E402
from pararealml_b200.operators.fdm.lowering import lower_problem  # noqa: E402


def main():
    workload, n, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    os.makedirs(out, exist_ok=True)
    builder = bench.WORKLOADS[workload][0]
    # the generated code depends on the mesh only: skip the (multi-GB) state
    bench.gaussian_y0 = lambda n, y_dim=3: __import__("numpy").zeros((n, n, n, y_dim))
    ivp, _ = builder(ns, n, 1)
    cp = ivp.constrained_problem
    low = lower_problem(cp)
    src = codegen.generate_source(low.spec(**plan_overrides(cp, low, None)))
    with open(os.path.join(out, "kernel.cu"), "w") as fh:
        fh.write(src)
    _native.compile_to_cubin(src, os.path.join(out, "kernel.cubin"))
    print("wrote", out)


if __name__ == "__main__":
    main()
