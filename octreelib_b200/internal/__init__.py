"""Internal helper types: voxel identity, ids, array aliases (reference: octreelib/internal/)."""
from . import interfaces as _interfaces, point as _point, typing as _typing, voxel as _voxel
from .interfaces import *  # noqa: F401,F403
from .point import *  # noqa: F401,F403
from .typing import *  # noqa: F401,F403
from .voxel import *  # noqa: F401,F403

__all__ = _typing.__all__ + _voxel.__all__ + _point.__all__ + _interfaces.__all__
