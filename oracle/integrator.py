"""NumPy oracle of the explicit integrators.

Restates ``pararealml/operators/fdm/numerical_integrator.py`` of the
reference: ForwardEuler :47-62, ExplicitMidpoint :70-90, RK4 :98-132.
``f(t, y)`` is the time derivative, ``constrain(t)`` returns the per-component
y constraints of time ``t`` (or None).  Test infrastructure only.
"""
import numpy as np


def _constrained(constraints, y):
    if constraints is not None:
        if y.ndim <= 1:
            raise ValueError("array must have at least 2 dimensions")
        if len(constraints) != y.shape[-1]:
            raise ValueError("constraint count must match last axis")
        for i, c in enumerate(constraints):
            if c is not None:
                c.apply(y[..., i : i + 1])
    return y


def forward_euler_step(y, t, d_t, f, constrain):
    c_full = constrain(t + d_t)
    return _constrained(c_full, y + d_t * f(t, y))


def explicit_midpoint_step(y, t, d_t, f, constrain):
    half = d_t / 2.0
    c_half = constrain(t + half)
    c_full = constrain(t + d_t)
    y_mid = _constrained(c_half, y + half * f(t, y))
    return _constrained(c_full, y + d_t * f(t + half, y_mid))


def rk4_step(y, t, d_t, f, constrain):
    half = d_t / 2.0
    c_half = constrain(t + half)
    c_full = constrain(t + d_t)
    k1 = d_t * f(t, y)
    k2 = d_t * f(t + half, _constrained(c_half, y + k1 / 2.0))
    k3 = d_t * f(t + half, _constrained(c_half, y + k2 / 2.0))
    k4 = d_t * f(t + d_t, _constrained(c_full, y + k3))
    return _constrained(c_full, y + (k1 + 2.0 * k2 + 2.0 * k3 + k4) / 6.0)


STEPPERS = {
    "forward_euler": forward_euler_step,
    "explicit_midpoint": explicit_midpoint_step,
    "rk4": rk4_step,
}
