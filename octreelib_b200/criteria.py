"""
Subdivision / filtering criteria for the native path.

The reference's `subdivide` / `filter` take lists of opaque Python callables
`criterion(points: (n,3) ndarray) -> bool` (grid/grid_base.py:126-147): a node splits if ANY
subdivision criterion is true (octree/octree.py:26); a leaf keeps its points only if ALL filtering
criteria are true (octree/octree.py:111).  A GPU cannot call Python per node, so the host folds the
list into a function of the point COUNT, which is what every criterion in the reference's tests
and docs is (`lambda points: len(points) > 100`):

  * `MaxPoints(n)` / `MinPoints(n)` are declarative callables (usable with the reference too);
  * any other callable is probed with zero-stride (n,3) arrays of several sizes and two different
    fill values; if its answers depend on the size only, the resulting truth table is used.
    Criteria that look at coordinates are rejected (NotImplementedError) -- no CPU fallback.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np

__all__ = ["MaxPoints", "MinPoints", "CountCriterion", "fold_count_criteria", "as_threshold"]

_LARGE_PROBES = [1 << k for k in range(11, 31, 2)] + [(1 << 31) - 1]


class CountCriterion:
    """`len(points) <op> n` as a callable object with inspectable parameters."""

    _OPS = {
        ">": lambda a, b: a > b,
        ">=": lambda a, b: a >= b,
        "<": lambda a, b: a < b,
        "<=": lambda a, b: a <= b,
        "==": lambda a, b: a == b,
        "!=": lambda a, b: a != b,
    }

    def __init__(self, op: str, n: int):
        if op not in self._OPS:
            raise ValueError(f"unknown comparison {op!r}")
        self.op, self.n = op, int(n)

    def on_count(self, count: int) -> bool:
        return bool(self._OPS[self.op](count, self.n))

    def __call__(self, points) -> bool:
        return self.on_count(len(points))

    def __repr__(self):
        return f"CountCriterion(len(points) {self.op} {self.n})"


class MaxPoints(CountCriterion):
    """Subdivide while a node holds more than `n` points: `lambda points: len(points) > n`."""

    def __init__(self, n: int):
        super().__init__(">", n)


class MinPoints(CountCriterion):
    """Keep a leaf only if it holds at least `n` points: `lambda points: len(points) >= n`."""

    def __init__(self, n: int):
        super().__init__(">=", n)


def _probe(n: int, fill: float) -> np.ndarray:
    """An (n, 3) float64 array that costs 8 bytes: zero strides over one value (read-only)."""
    a = np.lib.stride_tricks.as_strided(np.array([fill], dtype=np.float64), shape=(n, 3), strides=(0, 0), writeable=False)
    return a


def _eval(criterion: Callable, n: int) -> bool:
    if isinstance(criterion, CountCriterion):
        return criterion.on_count(n)
    try:
        a = bool(criterion(_probe(n, 0.0)))
        b = bool(criterion(_probe(n, 123456.789)))
    except Exception as exc:  # noqa: BLE001
        raise NotImplementedError(
            "octreelib_b200 evaluates subdivision / filtering criteria on the GPU as functions of the point count; "
            f"criterion {criterion!r} failed on a probe array of {n} points ({exc!r}). Use criteria such as "
            "`lambda points: len(points) > N` or octreelib_b200.criteria.MaxPoints / MinPoints.") from exc
    if a != b:
        raise NotImplementedError(
            f"criterion {criterion!r} depends on the point coordinates, not only on the point count; "
            "coordinate-dependent criteria are not supported by the GPU path (there is no CPU fallback)")
    return a


def fold_count_criteria(criteria: Sequence[Callable], mode: str, upto: int) -> Tuple[np.ndarray, bool]:
    """Truth table t[n], n = 0..upto, of `any(c(points))` (mode 'any') or `all(c(points))` (mode 'all')
    for clouds of n points, plus the value for 'very large' clouds (None if the criteria do not settle
    on one).  Raises NotImplementedError for coordinate-dependent criteria."""
    criteria = list(criteria)
    comb = any if mode == "any" else all
    table = np.zeros(upto + 1, dtype=np.uint8)
    for n in range(upto + 1):
        table[n] = comb([_eval(c, n) for c in criteria])
    beyond_vals = {bool(comb([_eval(c, n) for c in criteria])) for n in _LARGE_PROBES if n > upto}
    if len(beyond_vals) > 1:
        return table, None  # no single answer for "more points than the table covers"
    beyond = beyond_vals.pop() if beyond_vals else bool(table[-1])
    return table, beyond


def as_threshold(table: np.ndarray, beyond: bool):
    """If the table is the step `count > n`, return n; if it is never true return a huge n; else None."""
    t = table.astype(bool)
    if beyond is None:
        return None
    if not t.any() and not beyond:
        return (1 << 62)
    first = int(np.argmax(t)) if t.any() else len(t)
    if t[first:].all() and beyond and not t[:first].any():
        return first - 1
    return None
