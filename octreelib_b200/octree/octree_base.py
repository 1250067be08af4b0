"""
Abstract contracts of the octree layer (reference: octreelib/octree/octree_base.py:13-242).
Kept so that user code that subclasses or type-checks against them keeps working.
"""
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Callable, List

import numpy as np

from ..internal.point import Point, PointCloud
from ..internal.voxel import Voxel

__all__ = ["OctreeConfigBase", "OctreeBase", "OctreeNodeBase"]

Criteria = List[Callable[[PointCloud], bool]]


@dataclass
class OctreeConfigBase(ABC):
    """debug: kept for API parity; the reference never reads it (octree_base.py:21)."""

    debug: bool = True


class _OctreeContract(Voxel, ABC):
    """Operations shared by a node and a whole octree."""

    @property
    @abstractmethod
    def n_nodes(self):
        """Number of nodes (leaves and internal)."""

    @property
    @abstractmethod
    def n_leaves(self):
        """Number of leaves that hold at least one point."""

    @property
    @abstractmethod
    def n_points(self):
        """Number of stored points."""

    @abstractmethod
    def filter(self, filtering_criteria: Criteria):
        """Empty every leaf for which not all criteria hold."""

    @abstractmethod
    def map_leaf_points(self, function: Callable[[PointCloud], PointCloud]):
        """Replace every non-empty leaf's cloud by function(cloud)."""

    @abstractmethod
    def subdivide(self, subdivision_criteria: Criteria):
        """Split while any criterion holds."""

    @abstractmethod
    def subdivide_as(self, other):
        """Copy the shape of another tree."""

    @abstractmethod
    def get_points(self) -> PointCloud:
        """All stored points."""

    @abstractmethod
    def apply_mask(self, mask: np.ndarray):
        """Keep the points whose mask entry is True."""


class OctreeNodeBase(_OctreeContract):
    """A node that registers itself in its octree's leaf cache (octree_base.py:24-49)."""

    def __init__(self, corner_min: Point, edge_length: float, octree_cached_leaves: List["OctreeNodeBase"]):
        super().__init__(corner_min, edge_length)
        self._children = []
        self._has_children = False
        self._cached_leaves = octree_cached_leaves
        self._cached_leaves.append(self)

    @abstractmethod
    def get_leaf_points(self) -> List[Voxel]:
        """Non-empty leaves below this node."""


class OctreeBase(_OctreeContract):
    """One pose's octree inside one cell (octree_base.py:132-158)."""

    _node_type = OctreeNodeBase

    def __init__(self, octree_config: OctreeConfigBase, corner_min: Point, edge_length: float):
        super().__init__(corner_min, edge_length)
        self._config = octree_config

    @abstractmethod
    def get_leaf_points(self, non_empty: bool) -> List[Voxel]:
        """Leaves in the cache order of the reference."""

    @abstractmethod
    def insert_points(self, points: PointCloud):
        """Add points."""
