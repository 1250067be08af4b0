"""Array aliases of the public signatures: one point is a float64 array of shape (3,), a cloud has shape (N, 3).
(The reference's octreelib/internal/point.py:15-16 spells the dtype with an alias that numpy 2 no longer has.)"""
from typing import Annotated, Literal

from numpy import float64
from numpy.typing import NDArray

_F64 = NDArray[float64]
PointCloud = Annotated[_F64, Literal["N", 3]]
Point = Annotated[_F64, Literal[3]]

__all__ = ["Point", "PointCloud"]
