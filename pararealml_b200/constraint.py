"""Masked value constraints (host side).

API mirror of the reference's ``pararealml/constraint.py`` (``Constraint``
:6-101, ``apply_constraints_along_last_axis`` :104-131).  The B200 kernels do
not consume these objects directly; the lowering step turns them into NaN-coded
face tables (``NaN`` = unconstrained), see ``to_nan_table`` /
``from_nan_table``.
"""
from typing import Optional, Sequence, Union

import numpy as np


class Constraint:
    """``values`` are written to the positions where ``mask`` is True."""

    def __init__(self, values: np.ndarray, mask: np.ndarray):
        n_true = int(mask.sum())
        if values.size != n_true:
            raise ValueError(
                f"{values.size} constraint values for {n_true} masked elements"
            )
        self._values = np.array(values, copy=True)
        self._mask = np.array(mask, copy=True)
        self._values.setflags(write=False)
        self._mask.setflags(write=False)

    @property
    def values(self) -> np.ndarray:
        return self._values

    @property
    def mask(self) -> np.ndarray:
        return self._mask

    def _check(self, array: np.ndarray, what: str):
        if array.shape[-self._mask.ndim:] != self._mask.shape:
            raise ValueError(
                f"{what} shape {array.shape} does not end in mask shape "
                f"{self._mask.shape}"
            )

    def apply(self, array: np.ndarray) -> np.ndarray:
        """In-place masked overwrite; returns ``array``."""
        self._check(array, "array")
        array[..., self._mask] = self._values
        return array

    def multiply_and_add(
        self,
        addend: np.ndarray,
        multiplier: Union[float, np.ndarray],
        result: np.ndarray,
    ) -> np.ndarray:
        """``result[mask] = addend[mask] + multiplier * values`` in place."""
        if addend.shape != result.shape:
            raise ValueError(
                f"addend shape {addend.shape} != result shape {result.shape}"
            )
        self._check(result, "result")
        if not isinstance(multiplier, float) and (
            multiplier.shape != self._values.shape
        ):
            raise ValueError(
                f"multiplier shape {multiplier.shape} != values shape "
                f"{self._values.shape}"
            )
        result[..., self._mask] = (
            addend[..., self._mask] + multiplier * self._values
        )
        return result


def apply_constraints_along_last_axis(
    constraints: Optional[Union[Sequence[Optional[Constraint]], np.ndarray]],
    array: np.ndarray,
) -> np.ndarray:
    """Applies ``constraints[i]`` to ``array[..., i:i+1]`` in place."""
    if constraints is None:
        return array
    if array.ndim <= 1:
        raise ValueError("array must have at least 2 dimensions")
    if len(constraints) != array.shape[-1]:
        raise ValueError(
            f"{len(constraints)} constraints for last axis of size "
            f"{array.shape[-1]}"
        )
    for i, c in enumerate(constraints):
        if c is not None:
            c.apply(array[..., i : i + 1])
    return array


def to_nan_table(constraint: Optional[Constraint], shape) -> np.ndarray:
    """NaN-coded dense table of a constraint (NaN = unconstrained)."""
    table = np.full(shape, np.nan)
    if constraint is not None:
        table[constraint.mask.reshape(shape)] = constraint.values
    return table


def from_nan_table(table: np.ndarray) -> Constraint:
    mask = ~np.isnan(table)
    return Constraint(table[mask], mask)
