from pararealml_b200.operators.fdm.fdm_operator import FDMOperator  # noqa: F401
from pararealml_b200.operators.fdm.numerical_differentiator import (  # noqa: F401
    NumericalDifferentiator,
    ThreePointCentralDifferenceMethod,
)
from pararealml_b200.operators.fdm.numerical_integrator import (  # noqa: F401
    RK4,
    BackwardEulerMethod,
    CrankNicolsonMethod,
    ExplicitMidpointMethod,
    ForwardEulerMethod,
    NumericalIntegrator,
)
