"""B200-native FDM + Parareal hot path behind the PararealML operator API."""
from pararealml_b200.boundary_condition import (  # noqa: F401
    BoundaryCondition,
    CauchyBoundaryCondition,
    ConstantBoundaryCondition,
    ConstantFluxBoundaryCondition,
    ConstantValueBoundaryCondition,
    DirichletBoundaryCondition,
    NeumannBoundaryCondition,
    VectorizedBoundaryConditionFunction,
    vectorize_bc_function,
)
from pararealml_b200.constrained_problem import ConstrainedProblem  # noqa: F401
from pararealml_b200.constraint import (  # noqa: F401
    Constraint,
    apply_constraints_along_last_axis,
)
from pararealml_b200.differential_equation import (  # noqa: F401
    LHS,
    BurgersEquation,
    CahnHilliardEquation,
    ConvectionDiffusionEquation,
    DifferentialEquation,
    DiffusionEquation,
    LorenzEquation,
    LotkaVolterraEquation,
    NavierStokesEquation,
    NBodyGravitationalEquation,
    PopulationGrowthEquation,
    ShallowWaterEquation,
    SIREquation,
    SymbolicEquationSystem,
    Symbols,
    VanDerPolEquation,
    WaveEquation,
)
from pararealml_b200.initial_condition import (  # noqa: F401
    ConstantInitialCondition,
    ContinuousInitialCondition,
    DiscreteInitialCondition,
    GaussianInitialCondition,
    InitialCondition,
    MarginalBetaProductInitialCondition,
    VectorizedInitialConditionFunction,
    vectorize_ic_function,
)
from pararealml_b200.initial_value_problem import InitialValueProblem  # noqa: F401
from pararealml_b200.mesh import (  # noqa: F401
    CoordinateSystem,
    Mesh,
    from_cartesian_coordinates,
    to_cartesian_coordinates,
    unit_vectors_at,
)
from pararealml_b200.operator import Operator, discretize_time_domain  # noqa: F401
from pararealml_b200.solution import Diffs, Solution  # noqa: F401
