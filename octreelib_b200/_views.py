"""
Host-side views over a native forest: turns the exported tables into the objects the reference's
API returns (`List[Voxel]`, `(n, 3)` arrays).  Pure bookkeeping on numpy arrays that were computed
on the GPU; no geometry or grouping is recomputed here.
"""
from __future__ import annotations

import numpy as np

from .internal.voxel import LeafVoxel

__all__ = ["tables", "leaf_voxels", "points_dict_order"]


def tables(forest) -> dict:
    """Cell / leaf / (cell, pose) tables of the current tree shape, cached per forest version."""
    cache = getattr(forest, "_table_cache", None)
    if cache is None or cache["version"] != forest.version:
        cache = dict(version=forest.version, cells=forest.export_cells(), leaves=forest.export_leaves(),
                     cell_poses=forest.export_cell_poses(), blocks=forest.export_blocks())
        forest._table_cache = cache
    return cache


def leaf_voxels(forest, pose_index: int, non_empty: bool, root_corner, root_edge, history: bool = False, pose_epoch: int = 0):
    """`get_leaf_points` of one pose (grid/grid.py:217-232 -> octree/octree.py:256-263).

    The forest's block table is already in the reference's order for this pose.  history = the grid has been subdivided
    more than once: the reference's leaf lists grow across calls, so the order of a pose's leaves inside a cell is
    (max(subdivide call that split the leaf's parent, calls made before the pose arrived), one-call order); that only
    matters here for `non_empty=False`, where the EMPTY leaves are listed from the leaf table as well.

    root_corner(cell_index) -> corner object of an unsplit cell root (an int64 array for grid cells,
    grid.py:96-105; the user's array for a stand-alone OctreeManager); root_edge: its edge object.
    Leaves below the root carry float64 corners and np.float64 edges (octree.py:181-187).
    """
    t = tables(forest)
    leaves, blocks = t["leaves"], t["blocks"]
    sel = np.flatnonzero(blocks["pose"] == pose_index)
    blk_leaf = blocks["leaf"][sel]
    blk_size = blocks["size"][sel].astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(blk_size)])
    xyz = forest.export_points(pose_index, order=0, n_hint=int(offs[-1]))["xyz"]

    def make(leaf: int, pts):
        if leaves["depth"][leaf] == 0:
            return LeafVoxel(root_corner(int(leaves["cell"][leaf])), root_edge, pts)
        return LeafVoxel(leaves["corner"][leaf].copy(), np.float64(leaves["edge"][leaf]), pts)

    if non_empty:
        return [make(int(blk_leaf[j]), xyz[offs[j]:offs[j + 1]]) for j in range(len(blk_leaf))]
    # every leaf (empty ones included) of the cells in which this pose owns an octree
    cp = t["cell_poses"]
    cells_of_pose = cp["cell"][cp["pose"] == pose_index]
    begin = t["cells"]["leaf_begin"]
    slot = {int(l): j for j, l in enumerate(blk_leaf)}
    out = []
    empty = np.empty((0, 3), dtype=float)
    for c in cells_of_pose:
        ids = np.arange(int(begin[c]), int(begin[c + 1]))
        if history:
            eff = np.maximum(np.asarray(leaves["parent_epoch"], dtype=np.int64)[ids], int(pose_epoch))
            ids = ids[np.argsort(eff, kind="stable")]
        for leaf in ids.tolist():
            j = slot.get(leaf)
            out.append(make(leaf, empty if j is None else xyz[offs[j]:offs[j + 1]]))
    return out


def points_dict_order(forest, pose_index: int) -> np.ndarray:
    """`Grid.get_points` (grid/grid.py:234-242): cells in creation order (the order of the reference's
    dict: by the pose that first touched the cell, lexicographic inside one insert), depth-first
    leaf order inside a cell (octree/octree.py:55-65)."""
    t = tables(forest)
    res = forest.export_points(pose_index, order=1)
    xyz, cell = res["xyz"], res["cell"]
    if len(xyz) == 0:
        return np.empty((0, 3), dtype=float)
    first_pose = t["cells"]["first_pose"]
    n_cells = len(first_pose)
    creation_rank = np.empty(n_cells, dtype=np.int64)
    creation_rank[np.lexsort((np.arange(n_cells), first_pose))] = np.arange(n_cells)
    order = np.argsort(creation_rank[cell], kind="stable")
    return xyz[order]
