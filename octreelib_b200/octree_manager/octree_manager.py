"""
`OctreeManager`: all poses of ONE grid cell, subdivided in lock-step
(reference: octreelib/octree_manager/octree_manager.py:12-180).

Inside a `Grid` the native forest handles every cell at once and no manager objects exist; this
class is the same machinery restricted to a single fixed cell, for code (and the reference's
tests, test/octree/test_multi_pose.py) that uses a manager directly.
"""
from typing import Callable, List, Optional, Type

import numpy as np

from .._host import ForestHost
from ..internal.point import Point, PointCloud
from ..internal.voxel import Voxel, VoxelBase
from ..octree.octree_base import OctreeBase, OctreeConfigBase

__all__ = ["OctreeManager"]


class OctreeManager(VoxelBase):
    def __init__(self, octree_type: Type[OctreeBase], octree_config: OctreeConfigBase, corner_min: Point,
                 edge_length: float):
        super().__init__(corner_min, edge_length)
        self._octree_type = octree_type
        self._octree_config = octree_config
        self._host = ForestHost(edge_length, corner_min, single_cell=True)

    def _root_corner(self, _cell):
        return self._corner_min

    def subdivide(self, subdivision_criteria: List[Callable[[PointCloud], bool]],
                  pose_numbers: Optional[List[int]] = None):
        """Build the shape from the union of the listed poses' points and impose it on every pose
        (octree_manager.py:36-66).  A listed pose without points here raises KeyError."""
        self._host.subdivide(subdivision_criteria, pose_numbers)

    def map_leaf_points(self, function, pose_numbers: Optional[List[int]] = None):
        """octree_manager.py:68-83 (host-callback compatibility path, see ForestHost.map_leaf_points)."""
        self._host.map_leaf_points(function, pose_numbers)

    def filter(self, filtering_criteria: List[Callable[[PointCloud], bool]], pose_numbers: Optional[List[int]] = None):
        """octree_manager.py:85-99"""
        self._host.filter(filtering_criteria, pose_numbers)

    def get_leaf_points(self, non_empty: bool = True, pose_number: Optional[int] = None) -> List[Voxel]:
        """octree_manager.py:101-119"""
        if pose_number is None:
            out = []
            for p in self._host.pose_numbers:
                out += self._host.leaf_voxels(p, non_empty, self._root_corner, self._edge_length)
            return out
        if pose_number in self._host.pose_index:
            return self._host.leaf_voxels(pose_number, non_empty, self._root_corner, self._edge_length)
        return []

    def get_points(self, pose_number: Optional[int] = None) -> PointCloud:
        """octree_manager.py:121-130"""
        if pose_number is None:
            return np.vstack([self._host.points_dfs(p) for p in self._host.pose_numbers])
        return self._host.points_dfs(pose_number)

    def n_points(self, pose_number: Optional[int] = None) -> int:
        if pose_number is None:
            return sum(self._host.count(p, 1) for p in self._host.pose_numbers)
        return self._host.count(pose_number, 1)

    def n_leaves(self, pose_number: int) -> int:
        return self._host.count(pose_number, 0)

    def n_nodes(self, pose_number: int) -> int:
        return self._host.count(pose_number, 2)

    def insert_points(self, pose_number: int, points: PointCloud):
        """octree_manager.py:161-171 (appends when the pose already exists)."""
        self._host.insert(pose_number, np.asarray(points), allow_append=True)

    def apply_mask(self, mask: np.ndarray, pose_number: int):
        """octree_manager.py:173-180"""
        if pose_number in self._host.pose_index:
            self._host.forest.apply_pose_mask(self._host.pose_index[pose_number], mask)
