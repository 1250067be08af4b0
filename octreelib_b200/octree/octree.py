"""
`Octree`, `OctreeNode`, `OctreeConfig` with the reference's surface
(octreelib/octree/octree.py:14-16, 19-200, 203-295), backed by a single-cell native forest.

The reference keeps one Python `OctreeNode` object per node; here the tree lives on the GPU and
these classes are handles: every method forwards to the forest and the returned leaves are host
views (`LeafVoxel`).
"""
from dataclasses import dataclass
from typing import Callable, Generic, List

import numpy as np

from .._host import ForestHost
from ..internal import PointCloud, T, Voxel
from .octree_base import OctreeBase, OctreeConfigBase, OctreeNodeBase

__all__ = ["OctreeNode", "Octree", "OctreeConfig"]


@dataclass
class OctreeConfig(OctreeConfigBase):
    pass


class _SinglePose:
    """One cell, one pose (pose number 0) on a native forest."""

    def _init_native(self, corner_min, edge_length):
        self._host = ForestHost(edge_length, corner_min, single_cell=True)

    def _root_corner(self, _cell):
        return self._corner_min

    def subdivide(self, subdivision_criteria: List[Callable[[PointCloud], bool]]):
        """Split while any criterion holds (octree.py:20-32, 214-220).

        On a tree that is ALREADY split the reference evaluates the criteria on the root's own (empty) point array
        (octree.py:26) and does not descend: with count criteria that are false for an empty cloud the call is a no-op,
        it never deepens or coarsens the tree.  Only `OctreeManager` / `Grid` rebuild the shape from a fresh scheme."""
        if not self._host.empty and self._host.forest.stats(light=True)["n_internal"] > 0:
            from ..criteria import CountCriterion, fold_count_criteria
            plain = [CountCriterion(c.op, c.n) if isinstance(c, CountCriterion) else c for c in subdivision_criteria]
            table, _ = fold_count_criteria(plain, "any", 0)
            if not table[0]:
                return
            raise NotImplementedError("a subdivision criterion that is true for an empty point cloud, applied to an octree "
                                      "that is already split (the reference would split the root again over its children)")
        self._host.subdivide(subdivision_criteria)
        self._sync_cache()

    def subdivide_as(self, other):
        """Copy the subdivision scheme of another octree / octree node (octree.py:34-53, 222-227): nodes that `other`
        splits and this tree does not are split, nodes this tree splits and `other` does not are collapsed, the points
        are re-routed.  Both trees are taken by their own root (the reference does not compare the roots either; the
        scheme is copied node for node).  Difference by design: the reference forgets to put a COLLAPSED node back into
        its leaf list (octree.py:48-53), so its points vanish from `get_leaf_points`; here the node stays a leaf."""
        if not isinstance(other, _SinglePose):
            raise TypeError("subdivide_as expects an Octree / OctreeNode of this package")
        if other._host.empty:
            shape = dict(q=np.zeros((0, 3), np.int64), depth=np.zeros(0, np.uint32), path=np.zeros(0, np.uint64))
        else:
            shape = other._host.forest.export_shape()
        if self._host.empty:
            return  # nothing stored yet: the reference would create empty children, which hold no points either
        self._host.forest.impose_shape(shape)
        self._host._counts_cache = None
        self._sync_cache()

    def get_points(self) -> PointCloud:
        """Stored points in depth-first leaf order (octree.py:55-65); input order while unsplit."""
        return self._host.points_dfs(0)

    def insert_points(self, points: PointCloud):
        """Append points (octree.py:235-239)."""
        self._host.insert(0, np.asarray(points), allow_append=True)
        self._sync_cache()

    def filter(self, filtering_criteria: List[Callable[[PointCloud], bool]]):
        """Empty every leaf for which not all criteria hold (octree.py:102-112)."""
        self._host.filter(filtering_criteria)

    def map_leaf_points(self, function: Callable[[PointCloud], PointCloud]):
        """octree.py:249-254 (host-callback compatibility path, see ForestHost.map_leaf_points)."""
        self._host.map_leaf_points(function)

    def apply_mask(self, mask: np.ndarray):
        """Keep the points whose mask entry is True; mask in leaf order (octree.py:265-274)."""
        if not self._host.empty:
            self._host.forest.apply_pose_mask(0, mask)

    @property
    def n_points(self):
        return self._host.count(0, 1)

    @property
    def n_leaves(self):
        return self._host.count(0, 0)

    @property
    def n_nodes(self):
        return self._host.count(0, 2) if not self._host.empty else 1

    def _leaves(self, non_empty: bool) -> List[Voxel]:
        if self._host.empty:
            return [] if non_empty else [self]
        return self._host.leaf_voxels(0, non_empty, self._root_corner, self._edge_length)

    def _sync_cache(self):
        pass


class OctreeNode(_SinglePose, OctreeNodeBase):
    """A root node used directly (test/octree/test_octree.py:8-30).  `octree_cached_leaves` is kept in
    sync with the native tree: after every structural change it lists all leaves, empty ones
    included, in the reference's cache order."""

    def __init__(self, corner_min, edge_length, octree_cached_leaves: List):
        OctreeNodeBase.__init__(self, corner_min, edge_length, octree_cached_leaves)
        self._init_native(corner_min, edge_length)

    def _sync_cache(self):
        self._cached_leaves[:] = self._leaves(non_empty=False)

    def get_leaf_points(self) -> List[Voxel]:
        return self._leaves(non_empty=True)


class Octree(_SinglePose, OctreeBase, Generic[T]):
    """One pose's points in one cell as an octree (octree.py:203-295)."""

    _node_type = OctreeNode

    def __init__(self, octree_config, corner_min, edge_length):
        OctreeBase.__init__(self, octree_config, corner_min, edge_length)
        self._init_native(corner_min, edge_length)

    def get_leaf_points(self, non_empty: bool = True) -> List[Voxel]:
        return self._leaves(non_empty)
