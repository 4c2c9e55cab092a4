"""Initial value problem = constrained problem + time interval + initial
condition (API mirror of the reference's ``initial_value_problem.py``)."""
from typing import Callable, Optional, Tuple

import numpy as np

TemporalDomainInterval = Tuple[float, float]


class InitialValueProblem:
    def __init__(
        self,
        cp,
        t_interval: TemporalDomainInterval,
        initial_condition,
        exact_y: Optional[Callable] = None,
    ):
        if t_interval[0] > t_interval[1]:
            raise ValueError(
                f"inverted time interval ({t_interval[0]}, {t_interval[1]})"
            )
        self._cp = cp
        self._t_interval = t_interval
        self._ic = initial_condition
        self._exact_y = exact_y

    @property
    def constrained_problem(self):
        return self._cp

    @property
    def t_interval(self) -> TemporalDomainInterval:
        return self._t_interval

    @property
    def initial_condition(self):
        return self._ic

    @property
    def has_exact_solution(self) -> bool:
        return self._exact_y is not None

    def exact_y(self, t: float, x: Optional[np.ndarray] = None) -> np.ndarray:
        if self._exact_y is None:
            raise RuntimeError("no exact solution was provided")
        return self._exact_y(self, t, x)
