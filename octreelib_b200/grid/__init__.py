"""Grid layer (reference: octreelib/grid/)."""
from . import grid as _grid, grid_base as _grid_base
from .grid import *  # noqa: F401,F403
from .grid_base import *  # noqa: F401,F403

__all__ = _grid_base.__all__ + _grid.__all__
