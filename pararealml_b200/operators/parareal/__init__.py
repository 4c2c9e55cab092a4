from pararealml_b200.operators.parareal.parareal_operator import (  # noqa: F401
    PararealOperator,
)
