"""Generates the golden fixtures in this directory by running the UNMODIFIED
reference (``/root/reference``, imported through ``tests/refshim.py``) on the
cases of ``cases.py``.  Run in the build container only:

    python tests/golden/generate_golden.py [case-name ...]

Each FDM case stores the strided trajectory (plus the last step) of
``FDMOperator.solve``; each Parareal case stores, per emulated world size, the
strided trajectory, the number of executed iterations and whether all ranks
returned identical arrays.  Parareal world sizes > 1 run on one thread per rank
behind a fake ``mpi4py`` communicator (no MPI runtime exists in this image).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import refshim  # noqa: E402

ref = refshim.install()
from pararealml.operators.fdm import (  # noqa: E402
    RK4,
    ExplicitMidpointMethod,
    FDMOperator,
    ForwardEulerMethod,
    ThreePointCentralDifferenceMethod,
)
from pararealml.operators.parareal import PararealOperator  # noqa: E402

from golden import cases  # noqa: E402

INTEGRATORS = {
    "rk4": RK4,
    "explicit_midpoint": ExplicitMidpointMethod,
    "forward_euler": ForwardEulerMethod,
}


def strided(y, stride):
    idx = sorted(set(range(stride - 1, len(y), stride)) | {len(y) - 1})
    return np.array(idx), y[idx]


def ref_fdm_operator(integrator, d_t, tol=1e-3):
    return FDMOperator(
        INTEGRATORS[integrator](), ThreePointCentralDifferenceMethod(tol), d_t
    )


def generate_fdm(case):
    ivp = case.build(ref)
    if case.seed is not None:
        np.random.seed(case.seed)
    sol = ref_fdm_operator(case.integrator, case.d_t, case.tol).solve(ivp)
    y = sol.discrete_y()
    idx, ys = strided(y, case.stride)
    np.savez_compressed(
        os.path.join(HERE, case.name + ".npz"),
        steps=idx,
        y=ys,
        t=sol.t_coordinates[idx],
        n_steps=len(y),
        y0=ivp.initial_condition.discrete_y_0(True),
    )
    print(f"{case.name}: {y.shape} -> {ys.shape}")


def generate_parareal(case):
    out = {}
    for size in case.sizes:
        counts = [0] * size

        def run(rank):
            ivp = case.build(ref)
            f = ref_fdm_operator(*case.f)
            g = ref_fdm_operator(*case.g)
            p = PararealOperator(f, g, case.tol)
            inner = p._should_terminate

            def counting(old, new):
                counts[rank] += 1
                return inner(old, new)

            p._should_terminate = counting
            return p.solve(ivp).discrete_y()

        ys = refshim.run_ranks(size, run)
        same = all(np.array_equal(ys[0], y) for y in ys[1:])
        assert same and len(set(counts)) == 1
        idx, yst = strided(ys[0], case.stride)
        out[f"steps_{size}"] = idx
        out[f"y_{size}"] = yst
        out[f"iterations_{size}"] = counts[0]
        out[f"n_steps_{size}"] = len(ys[0])
        print(f"{case.name} P={size}: iterations={counts[0]} {ys[0].shape}")
    np.savez_compressed(os.path.join(HERE, case.name + ".npz"), **out)


if __name__ == "__main__":
    wanted = set(sys.argv[1:])
    for c in cases.FDM_CASES:
        if not wanted or c.name in wanted:
            generate_fdm(c)
    for c in cases.PARAREAL_CASES:
        if not wanted or c.name in wanted:
            generate_parareal(c)
