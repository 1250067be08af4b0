"""Per-leaf RANSAC (reference: octreelib/ransac/)."""
from . import cuda_ransac as _cuda_ransac
from .cuda_ransac import *  # noqa: F401,F403

__all__ = _cuda_ransac.__all__
