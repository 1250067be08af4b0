"""Mixin for objects that carry an integer id (reference: octreelib/internal/interfaces.py:12-32)."""
from abc import ABC
from typing import Optional

__all__ = ["WithID"]


class WithID(ABC):
    """`.id` is either the id handed to the constructor or the next value of a process-wide counter."""

    _id_static_counter = 0

    def __init__(self, _id: Optional[int] = None):
        if _id is None:
            _id = WithID._id_static_counter
            WithID._id_static_counter += 1
        self._id = _id

    @property
    def id(self):
        return self._id
