"""Generic type variable used by the container classes (reference: octreelib/internal/typing.py:5)."""
from typing import TypeVar

__all__ = ["T"]

T = TypeVar("T")
