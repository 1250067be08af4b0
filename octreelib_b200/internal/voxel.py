"""
Voxel identity (reference: octreelib/internal/voxel.py:12-95).

A voxel is identified by `(tuple(corner_min), edge_length)`; ids are handed out first-come from a
process-global map, so that two voxels built anywhere with equal corner and edge share one id
(pinned by test/octree/test_multi_pose.py:167-182).  Leaves returned by the native grid are
`LeafVoxel` views: they register in the same map, but lazily, on first access of `.id`.
"""
import itertools
from typing import Optional

import numpy as np

from .interfaces import WithID
from .point import Point, PointCloud

__all__ = ["Voxel", "VoxelBase"]


def _identity(corner_min, edge_length):
    return (tuple(corner_min), edge_length)


class VoxelBase(WithID):
    """A cube given by its minimal corner and edge length, with a shared id."""

    _static_voxel_id_map = {}

    def __init__(self, corner_min: Point, edge_length: float):
        self._corner_min = corner_min
        self._edge_length = edge_length
        WithID.__init__(self, self._lookup_id())

    def _lookup_id(self) -> int:
        table = VoxelBase._static_voxel_id_map
        return table.setdefault(_identity(self._corner_min, self._edge_length), len(table))

    def __hash__(self):
        return hash(_identity(self._corner_min, self._edge_length))

    def __eq__(self, other: "VoxelBase"):
        return bool(np.all(self.corner_min == other.corner_min)) and self.edge_length == other.edge_length

    @property
    def corner_min(self):
        return self._corner_min

    @property
    def edge_length(self):
        return self._edge_length

    @property
    def corner_max(self):
        return self.corner_min + self.edge_length

    @property
    def all_corners(self):
        """The 8 corner points, x varying slowest."""
        e = self._edge_length
        return [self._corner_min + offset for offset in itertools.product([0, e], repeat=3)]


class Voxel(VoxelBase):
    """A voxel that owns a point cloud."""

    def __init__(self, corner_min: Point, edge_length: float, points: Optional[PointCloud] = None):
        super().__init__(corner_min, edge_length)
        self._points: PointCloud = np.empty((0, 3), dtype=float) if points is None else points

    def get_points(self) -> PointCloud:
        return self._points.copy()

    def insert_points(self, points: PointCloud):
        self._points = np.vstack([self._points, points])


class LeafVoxel(Voxel):
    """Host view of one (pose, leaf) block of the native grid: same surface as the live leaf nodes
    the reference returns from `get_leaf_points` (corner_min, edge_length, id, get_points,
    n_points, all_corners), with the id registered on first use."""

    def __init__(self, corner_min, edge_length, points):
        self._corner_min = corner_min
        self._edge_length = edge_length
        self._points = points
        self._id = None

    @property
    def id(self):
        if self._id is None:
            self._id = self._lookup_id()
        return self._id

    @property
    def n_points(self):
        return len(self._points)
