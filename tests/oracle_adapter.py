"""Class-shaped adapters around the function-based oracle so that the
reference's own unit tests can be executed against it (see
``test_reference_suite.py``)."""
import numpy as np

import oracle
from oracle import differentiator as od
from oracle import integrator as oi


class OracleThreePointCentralDifferenceMethod:
    def __init__(self, tol: float = 1e-3):
        if tol < 0.0:
            raise ValueError("tolerance must be non-negative")
        self._tol = tol

    def gradient(self, y, mesh, x_axis, derivative_boundary_constraints=None):
        return od.gradient(y, mesh, x_axis, derivative_boundary_constraints)

    def hessian(self, y, mesh, x_axis1, x_axis2, derivative_boundary_constraints=None):
        return od.hessian(y, mesh, x_axis1, x_axis2, derivative_boundary_constraints)

    def divergence(self, y, mesh, derivative_boundary_constraints=None):
        return od.divergence(y, mesh, derivative_boundary_constraints)

    def curl(self, y, mesh, curl_ind=0, derivative_boundary_constraints=None):
        return od.curl(y, mesh, curl_ind, derivative_boundary_constraints)

    def laplacian(self, y, mesh, derivative_boundary_constraints=None):
        return od.laplacian(y, mesh, derivative_boundary_constraints)

    def vector_laplacian(self, y, mesh, vector_laplacian_ind, derivative_boundary_constraints=None):
        return od.vector_laplacian(y, mesh, vector_laplacian_ind, derivative_boundary_constraints)

    def anti_laplacian(self, laplacian, mesh, y_constraints, derivative_boundary_constraints=None, y_init=None):
        return od.anti_laplacian(
            laplacian, mesh, y_constraints, derivative_boundary_constraints,
            y_init, tol=self._tol,
        )


class _OracleIntegrator:
    step = None

    def integral(self, y, t, d_t, d_y_over_d_t, y_constraint_function):
        return type(self).step(y, t, d_t, d_y_over_d_t, y_constraint_function)


class OracleForwardEulerMethod(_OracleIntegrator):
    step = staticmethod(oi.forward_euler_step)


class OracleExplicitMidpointMethod(_OracleIntegrator):
    step = staticmethod(oi.explicit_midpoint_step)


class OracleRK4(_OracleIntegrator):
    step = staticmethod(oi.rk4_step)


class OracleFDMOperatorAdapter:
    """``FDMOperator(integrator, differentiator, d_t)`` signature on top of
    ``oracle.fdm_solve``; returns the reference's ``Solution``."""

    def __init__(self, integrator, differentiator, d_t):
        if d_t <= 0.0:
            raise ValueError("time step size must be greater than 0")
        self._name = {
            OracleForwardEulerMethod: "forward_euler",
            OracleExplicitMidpointMethod: "explicit_midpoint",
            OracleRK4: "rk4",
        }[type(integrator)]
        self._tol = differentiator._tol
        self._d_t = d_t

    d_t = property(lambda self: self._d_t)
    vertex_oriented = property(lambda self: True)

    def solve(self, ivp, parallel_enabled=True):
        from pararealml.solution import Solution

        t, y = oracle.fdm_solve(ivp, self._name, self._d_t, self._tol)
        return Solution(ivp, t, y, vertex_oriented=True, d_t=self._d_t)
