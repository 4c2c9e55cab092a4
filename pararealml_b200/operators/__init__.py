"""Operators of the B200 hot path: ``fdm`` (finite differences) and
``parareal`` (time-parallel driver)."""
