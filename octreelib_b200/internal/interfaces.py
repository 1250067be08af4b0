"""Mixin for objects that carry an integer id (reference: octreelib/internal/interfaces.py:12-32)."""
from abc import ABC
from typing import Optional

__all__ = ["WithID"]


class WithID(ABC):
    """`.id` is either the id handed to the constructor or the next value of a process-wide counter
    (`WithID._id_static_counter`, shared by every subclass exactly like in the reference)."""

    _id_static_counter = 0

    @staticmethod
    def _draw_id() -> int:
        drawn = WithID._id_static_counter
        WithID._id_static_counter = drawn + 1
        return drawn

    def __init__(self, _id: Optional[int] = None):
        self._id = WithID._draw_id() if _id is None else _id

    @property
    def id(self):
        return self._id
