"""CPU oracle for the FDM + Parareal hot path -- TEST INFRASTRUCTURE ONLY.

A plain NumPy restatement of the reference algorithm (ViktorC/PararealML
v0.3.0, ``pararealml/operators/fdm`` and ``pararealml/operators/parareal``),
every function citing the reference file:line it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import this package; the product path
(``pararealml_b200``) never does and fails loudly without its CUDA library.

Pinning: the oracle is checked against (a) the literal expected arrays of the
reference's own unit tests (``tests/operators/fdm/test_numerical_
differentiator.py``, ``test_numerical_integrator.py``) re-stated in
``tests/golden/``, and (b) trajectories produced by importing the unmodified
reference in the build container (``tests/golden/generate_golden.py``).
"""
from oracle.differentiator import (  # noqa: F401
    anti_laplacian,
    curl,
    derivative,
    divergence,
    gradient,
    hessian,
    jacobi_step,
    laplacian,
    second_derivative,
    vector_laplacian,
)
from oracle.fdm import OracleFDMOperator, fdm_solve  # noqa: F401
from oracle.integrator import (  # noqa: F401
    explicit_midpoint_step,
    forward_euler_step,
    rk4_step,
)
from oracle.parareal import parareal_solve  # noqa: F401
