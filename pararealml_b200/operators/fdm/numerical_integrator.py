"""Explicit time integrators of the FDM operator.

Mirrors the class names and constructor signatures of the reference's
``pararealml/operators/fdm/numerical_integrator.py``.  Inside ``FDMOperator``
an integrator is only a *selector* for the fused stage kernels
(csrc/fdm_template.cuh: ``PML_FE``, ``PML_MID1/2``, ``PML_RK4_1..4``), which
keep the y/k buffers on the device for all time steps.  The implicit methods
(``BackwardEulerMethod``, ``CrankNicolsonMethod``; reference :135-270) are out
of scope and raise instead of falling back to the CPU.

``integral`` is kept for API compatibility with user code that passes Python
callables (reference :15-41); such callables can only run on the host, so this
method is host glue and is never used by ``FDMOperator``.
"""
from abc import ABC, abstractmethod
from typing import Callable, Optional, Sequence, Union

import numpy as np

from pararealml_b200.constraint import (
    Constraint,
    apply_constraints_along_last_axis,
)

ConstraintFunction = Callable[
    [Optional[float]], Optional[Union[Sequence[Constraint], np.ndarray]]
]


class NumericalIntegrator(ABC):
    #: name of the stage-kernel family in the C ABI
    #: (``PML_INTEGRATOR_*`` in include/pararealml_b200.h)
    kernel_family: str = ""

    @abstractmethod
    def integral(
        self,
        y: np.ndarray,
        t: float,
        d_t: float,
        d_y_over_d_t: Callable[[float, np.ndarray], np.ndarray],
        y_constraint_function: ConstraintFunction,
    ) -> np.ndarray:
        """Estimate of y(t + d_t) for host callables."""


class ForwardEulerMethod(NumericalIntegrator):
    kernel_family = "forward_euler"

    def integral(self, y, t, d_t, d_y_over_d_t, y_constraint_function):
        full = y_constraint_function(t + d_t)
        return apply_constraints_along_last_axis(
            full, y + d_t * d_y_over_d_t(t, y)
        )


class ExplicitMidpointMethod(NumericalIntegrator):
    kernel_family = "explicit_midpoint"

    def integral(self, y, t, d_t, d_y_over_d_t, y_constraint_function):
        h = d_t / 2.0
        mid_c = y_constraint_function(t + h)
        full = y_constraint_function(t + d_t)
        y_mid = apply_constraints_along_last_axis(
            mid_c, y + h * d_y_over_d_t(t, y)
        )
        return apply_constraints_along_last_axis(
            full, y + d_t * d_y_over_d_t(t + h, y_mid)
        )


class RK4(NumericalIntegrator):
    kernel_family = "rk4"

    def integral(self, y, t, d_t, d_y_over_d_t, y_constraint_function):
        h = d_t / 2.0
        mid_c = y_constraint_function(t + h)
        full = y_constraint_function(t + d_t)
        con = apply_constraints_along_last_axis
        k1 = d_t * d_y_over_d_t(t, y)
        k2 = d_t * d_y_over_d_t(t + h, con(mid_c, y + k1 / 2.0))
        k3 = d_t * d_y_over_d_t(t + h, con(mid_c, y + k2 / 2.0))
        k4 = d_t * d_y_over_d_t(t + d_t, con(full, y + k3))
        return con(full, y + (k1 + 2.0 * k2 + 2.0 * k3 + k4) / 6.0)


class ImplicitMethod(NumericalIntegrator, ABC):
    """Out of scope for the B200 path: constructing one raises."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            f"{type(self).__name__} is an implicit integrator; the B200 FDM "
            "path only provides ForwardEulerMethod, ExplicitMidpointMethod "
            "and RK4 and does not fall back to the CPU"
        )

    def integral(self, y, t, d_t, d_y_over_d_t, y_constraint_function):
        raise NotImplementedError


class BackwardEulerMethod(ImplicitMethod):
    pass


class CrankNicolsonMethod(ImplicitMethod):
    pass
