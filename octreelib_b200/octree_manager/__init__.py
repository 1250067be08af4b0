"""Per-cell multi-pose container (reference: octreelib/octree_manager/)."""
from . import octree_manager as _octree_manager
from .octree_manager import *  # noqa: F401,F403

__all__ = _octree_manager.__all__
