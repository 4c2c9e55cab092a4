"""The drop-in boundary: ``Operator.solve(ivp) -> Solution``.

Mirrors the reference's ``pararealml/operator.py`` (``Operator`` :13-57,
``discretize_time_domain`` :60-74).
"""
from abc import ABC, abstractmethod
from typing import Optional

import numpy as np

from pararealml_b200.solution import Solution


class Operator(ABC):
    def __init__(self, d_t: float, vertex_oriented: Optional[bool]):
        if d_t <= 0.0:
            raise ValueError("time step size must be greater than 0")
        self._d_t = d_t
        self._vertex_oriented = vertex_oriented

    @property
    def d_t(self) -> float:
        return self._d_t

    @property
    def vertex_oriented(self) -> Optional[bool]:
        return self._vertex_oriented

    @abstractmethod
    def solve(self, ivp, parallel_enabled: bool = True) -> Solution:
        """Solves the initial value problem."""


def discretize_time_domain(t, d_t: float) -> np.ndarray:
    """``steps = round((t1 - t0) / d_t)`` and a linspace over
    ``[t0, t0 + steps * d_t]``; must stay bit-identical to the reference
    because stage times are derived from it."""
    t_0 = t[0]
    steps = int(round((t[1] - t_0) / d_t))
    return np.linspace(t_0, t_0 + steps * d_t, steps + 1)
