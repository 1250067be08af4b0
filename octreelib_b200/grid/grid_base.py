"""
Configuration dataclasses and the abstract `GridBase` contract
(reference: octreelib/grid/grid_base.py:17-211).
"""
from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from enum import Enum
from typing import Callable, Generic, List, Optional, Type

import numpy as np

from ..internal.point import Point, PointCloud
from ..internal.typing import T
from ..internal.voxel import Voxel
from ..octree import Octree, OctreeBase, OctreeConfig, OctreeConfigBase
from ..octree_manager import OctreeManager

__all__ = ["GridVisualizationType", "VisualizationConfig", "GridConfigBase", "GridBase"]


class GridVisualizationType(Enum):
    """POSE: one colour per pose; VOXEL: one colour per voxel id (grid_base.py:17-25)."""

    POSE = "pose"
    VOXEL = "voxel"


@dataclass
class VisualizationConfig:
    """Options of `Grid.visualize` (grid_base.py:28-48)."""

    type: GridVisualizationType = GridVisualizationType.VOXEL
    point_size: float = 0.1
    line_width_size: float = 0.01
    line_color: int = 0xFF0000
    filepath: str = "visualization.html"
    seed: int = 0
    unused_voxels: List[int] = field(default_factory=list)


@dataclass
class GridConfigBase(ABC):
    """
    octree_manager_type / octree_type / octree_config: plug points of the reference (validated the
        same way; the native pipeline is used for the stock types).
    debug: unused, kept for parity.   voxel_edge_length: cell size.   corner: grid origin.
    """

    octree_manager_type: Type[OctreeManager] = OctreeManager
    octree_type: Type[OctreeBase] = Octree
    octree_config: OctreeConfigBase = field(default_factory=OctreeConfig)
    debug: bool = False
    voxel_edge_length: float = 1
    corner: Point = field(default_factory=lambda: np.array(([0.0, 0.0, 0.0])))

    def __post_init__(self):
        # same checks and messages as grid_base.py:73-87 (the messages are pinned by test/grid/test_grid.py:148-182)
        plug_points = (
            (self.octree_manager_type, OctreeManager, "octree manager type", "octree_manager.OctreeManager"),
            (self.octree_type, OctreeBase, "octree type", "octree.OctreeBase"),
        )
        for given, required, what, where in plug_points:
            if not issubclass(given, required):
                raise TypeError(f"Cannot use the provided {what} {given.__name__}. It has to be a subclass of {where}.")


class GridBase(ABC, Generic[T]):
    """What a grid must offer (grid_base.py:90-211)."""

    def __init__(self, grid_config: GridConfigBase):
        self._grid_config = grid_config

    @abstractmethod
    def insert_points(self, pose_number: int, points: List[Point]) -> None:
        """Add a pose's cloud."""

    @abstractmethod
    def get_points(self, pose_number: int) -> List[Point]:
        """Points stored for a pose."""

    @abstractmethod
    def subdivide(self, subdivision_criteria: List[Callable[[PointCloud], bool]],
                  pose_numbers: Optional[List[int]] = None):
        """Split every cell's octree while any criterion holds."""

    @abstractmethod
    def filter(self, filtering_criteria: List[Callable[[PointCloud], bool]]):
        """Empty the leaves for which not all criteria hold."""

    @abstractmethod
    def map_leaf_points(self, function: Callable[[PointCloud], PointCloud]):
        """Transform every leaf's cloud."""

    @abstractmethod
    def map_leaf_points_cuda_ransac(self, poses_per_batch: int = 1, threshold: float = 0.01,
                                    hypotheses_number: int = 1024):
        """Keep, in every leaf, the inliers of its best RANSAC plane."""

    @abstractmethod
    def get_leaf_points(self, pose_number: int) -> List[Voxel]:
        """Leaves of a pose as voxels."""

    @abstractmethod
    def visualize(self, config: VisualizationConfig) -> None:
        """Write a k3d HTML snapshot."""

    @abstractmethod
    def n_nodes(self, pose_number: int):
        """Node count of a pose's octrees."""

    @abstractmethod
    def n_points(self, pose_number: int):
        """Point count of a pose."""

    @abstractmethod
    def n_leaves(self, pose_number: int):
        """Non-empty leaf count of a pose."""
