"""Octree layer (reference: octreelib/octree/)."""
from . import octree as _octree, octree_base as _octree_base
from .octree import *  # noqa: F401,F403
from .octree_base import *  # noqa: F401,F403

__all__ = _octree_base.__all__ + _octree.__all__
