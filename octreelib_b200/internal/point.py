"""
User-facing array aliases (reference: octreelib/internal/point.py:15-16).

The reference spells the dtype `np.float_`, which numpy 2 removed; `np.float64` is the same type.
"""
from typing import Annotated, Literal

import numpy as np
import numpy.typing as npt

__all__ = ["Point", "PointCloud"]

Point = Annotated[npt.NDArray[np.float64], Literal[3]]
PointCloud = Annotated[npt.NDArray[np.float64], Literal["N", 3]]
