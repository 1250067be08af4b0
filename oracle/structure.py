"""
CPU ORACLE (test infrastructure, NOT product code) -- structure half.

A numpy restatement of the reference's grid / octree algorithm for the hot path
`Grid.insert_points -> subdivide -> filter / get_leaf_points`.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import
this module; the product package `octreelib_b200` never does.

The data model restates the reference's (one tree per (cell, pose), one leaf cache per
tree) but with index arrays instead of point copies, so that the canonical point order
("original input index order inside every leaf") is directly observable.
Pinned against the real reference by `tests/golden/make_golden.py` (run in the build
container where `/root/reference` is mounted) -> `tests/golden/*.npz`.

Reference citations are `path:line` under /root/reference/.
"""

from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

__all__ = ["OracleGrid", "OracleLeaf", "max_points_criterion", "min_points_criterion"]


def max_points_criterion(n: int) -> Callable[[np.ndarray], bool]:
    """`lambda points: len(points) > n` -- the subdivision criterion every reference test uses
    (test/grid/test_grid.py:47, test/octree/test_multi_pose.py:48-49)."""
    return lambda points: len(points) > n


def min_points_criterion(n: int) -> Callable[[np.ndarray], bool]:
    """`lambda points: len(points) >= n` -- the filtering criterion of test/octree/test_octree.py:28."""
    return lambda points: len(points) >= n


class _Node:
    """One octree node of one pose (octree/octree_base.py:24-49)."""

    __slots__ = ("corner", "edge", "children", "idx", "pts")

    def __init__(self, corner: np.ndarray, edge):
        self.corner = corner  # int64[3] for a grid-cell root, float64[3] below (octree.py:181-187)
        self.edge = edge
        self.children: Optional[List["_Node"]] = None
        self.idx = np.empty((0,), dtype=np.int64)  # indices into the pose's input cloud
        self.pts = np.empty((0, 3), dtype=np.float64)  # the points themselves (voxel.py:81-83)


class _Tree:
    """One pose's octree inside one cell (octree/octree_base.py:132-158).

    `cache` restates `_cached_leaves` (octree_base.py:155-158): an insertion-ordered dict
    keyed by id(node); `del` + re-insert reproduces `list.remove` + `list.append`
    (octree.py:183-191) without the O(L) scan.
    """

    __slots__ = ("root", "cache")

    def __init__(self, corner: np.ndarray, edge):
        self.cache: Dict[int, _Node] = {}
        self.root = self._new_node(corner, edge)

    def _new_node(self, corner, edge) -> _Node:
        node = _Node(corner, edge)
        self.cache[id(node)] = node  # octree_base.py:48-49
        return node

    # -- octree.py:177-191 ------------------------------------------------------------
    def _generate_children(self, node: _Node):
        half = node.edge / np.float64(2)
        del self.cache[id(node)]
        kids = []
        for cid in range(8):
            # itertools.product([0, half], repeat=3): x is the slowest axis -> cid = 4ix+2iy+iz
            off = np.array([(cid >> 2) & 1, (cid >> 1) & 1, cid & 1]) * half
            kids.append(self._new_node(node.corner + off, half))
        node.children = kids

    # -- octree.py:67-100 -------------------------------------------------------------
    def insert(self, node: _Node, idx: np.ndarray, pts: np.ndarray):
        if node.children is None:
            node.idx = np.concatenate([node.idx, idx])
            node.pts = np.vstack([node.pts, pts])
            return
        if len(idx) == 0:
            return
        half = node.edge / 2
        sub = ((pts - node.corner) // half).astype(int)
        # child id = sum 2**i * idx[::-1][i]  (octree.py:94-97); an index outside {0,1} makes the
        # reference pick a wrong child or raise IndexError -> we raise.
        if ((sub < 0) | (sub > 1)).any():
            raise IndexError("point outside of its octree node (reference: octree.py:98)")
        cid = sub[:, 0] * 4 + sub[:, 1] * 2 + sub[:, 2]
        order = np.argsort(cid, kind="stable")  # canonical (stable) order, SURVEY 8(c)
        cid_sorted = cid[order]
        bounds = np.searchsorted(cid_sorted, np.arange(9))
        for c in range(8):
            lo, hi = bounds[c], bounds[c + 1]
            if hi > lo:
                sel = order[lo:hi]
                self.insert(node.children[c], idx[sel], pts[sel])

    def _split(self, node: _Node):
        self._generate_children(node)
        idx, pts = node.idx, node.pts
        node.idx = np.empty((0,), dtype=np.int64)
        node.pts = np.empty((0, 3), dtype=np.float64)
        self.insert(node, idx, pts)

    # -- octree.py:20-32 --------------------------------------------------------------
    def subdivide(self, node: _Node, criteria: Sequence[Callable], depth=0, max_depth=64):
        if any([c(node.pts) for c in criteria]):
            if depth >= max_depth:
                raise RecursionError("oracle depth cap reached (reference recurses without limit)")
            self._split(node)
            for ch in node.children:
                self.subdivide(ch, criteria, depth + 1, max_depth)

    # -- octree.py:34-53 --------------------------------------------------------------
    def subdivide_as(self, node: _Node, other: _Node):
        if other.children is not None and node.children is None:
            self._split(node)
        if other.children is not None:
            for a, b in zip(node.children, other.children):
                self.subdivide_as(a, b)
        elif node.children is not None:
            # collapse.  NOTE: the reference forgets to put the collapsed node back into the
            # leaf cache (octree.py:48-53); the oracle (and the build) re-add it.
            idx, pts = self._collect(node)
            self._drop(node)
            node.children = None
            node.idx, node.pts = idx, pts
            self.cache[id(node)] = node

    def _collect(self, node: _Node):
        if node.children is None:
            return node.idx, node.pts
        parts = [self._collect(ch) for ch in node.children]
        return np.concatenate([p[0] for p in parts]), np.vstack([p[1] for p in parts])

    def _drop(self, node: _Node):
        for ch in node.children:
            if ch.children is not None:
                self._drop(ch)
            self.cache.pop(id(ch), None)

    # -- counters: octree.py:144-175 --------------------------------------------------
    def n_nodes(self, node=None) -> int:
        node = node or self.root
        if node.children is None:
            return 1
        return 1 + sum(self.n_nodes(ch) for ch in node.children)

    def leaves_dfs(self, node=None):
        node = node or self.root
        if node.children is None:
            return [node]
        out = []
        for ch in node.children:
            out += self.leaves_dfs(ch)
        return out


class OracleLeaf:
    """What `get_leaf_points` returns per leaf: corner, edge, points, source indices."""

    __slots__ = ("corner", "edge", "points", "idx", "cell")

    def __init__(self, corner, edge, points, idx, cell):
        self.corner, self.edge, self.points, self.idx, self.cell = corner, edge, points, idx, cell


class _Cell:
    """One grid cell = one OctreeManager (octree_manager/octree_manager.py:12-34)."""

    __slots__ = ("key", "edge", "trees", "scheme")

    def __init__(self, key: np.ndarray, edge):
        self.key = key
        self.edge = edge
        self.trees: Dict[int, _Tree] = {}
        self.scheme = _Tree(key, edge)


class OracleGrid:
    """Restates `Grid` (grid/grid.py:39-362) on top of `_Cell` / `_Tree`."""

    def __init__(self, voxel_edge_length=1, corner=(0.0, 0.0, 0.0)):
        self.edge = voxel_edge_length
        self.corner = np.asarray(corner, dtype=np.float64)
        self.pose_cells: Dict[int, List[tuple]] = {}  # grid.py:53  (lexicographic per pose)
        self.cells: Dict[tuple, _Cell] = {}  # grid.py:56  (first-appearance order)

    # -- grid.py:58-109 ---------------------------------------------------------------
    def insert_points(self, pose: int, points: np.ndarray):
        if pose in self.pose_cells:
            raise ValueError(f"Cannot insert points to existing pose {pose}")
        self.pose_cells[pose] = []
        points = np.asarray(points)
        keys = ((points - self.corner) // self.edge * self.edge).astype(int)  # grid.py:72-76
        if len(points) == 0:
            return
        # lexicographic (x, y, z) order of the distinct keys, stable grouping of the points
        order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
        ks = keys[order]
        head = np.ones(len(ks), dtype=bool)
        head[1:] = (ks[1:] != ks[:-1]).any(axis=1)
        starts = np.flatnonzero(head)
        ends = np.append(starts[1:], len(ks))
        pts64 = np.vstack([np.empty((0, 3), dtype=float), points])  # float64 upcast (voxel.py:81-83)
        for s, e in zip(starts, ends):
            key = tuple(int(v) for v in ks[s])
            cell = self.cells.get(key)
            if cell is None:
                cell = self.cells[key] = _Cell(np.array(ks[s]), self.edge)
            self.pose_cells[pose].append(key)
            sel = order[s:e]  # np.lexsort is stable -> ascending input index
            tree = cell.trees.get(pose)
            if tree is None:
                tree = cell.trees[pose] = _Tree(cell.key, cell.edge)  # octree_manager.py:166-169
            tree.insert(tree.root, sel.astype(np.int64), pts64[sel])
            tree.subdivide_as(tree.root, cell.scheme.root)  # octree_manager.py:171

    # -- grid.py:244-258 -> octree_manager.py:36-66 -----------------------------------
    def subdivide(self, criteria: Sequence[Callable], pose_numbers: Optional[Sequence[int]] = None,
                  max_depth: int = 64):
        for cell in self.cells.values():
            poses = list(cell.trees.keys()) if pose_numbers is None else list(pose_numbers)
            scheme = _Tree(cell.key, cell.edge)
            parts = [cell.trees[p]._collect(cell.trees[p].root) for p in poses]  # KeyError if absent
            if parts:
                scheme.insert(scheme.root, np.concatenate([p[0] for p in parts]),
                              np.vstack([np.empty((0, 3))] + [p[1] for p in parts]))
            scheme.subdivide(scheme.root, criteria, 0, max_depth)
            for leaf in scheme.cache.values():  # octree_manager.py:63: shape only
                leaf.idx = np.empty((0,), dtype=np.int64)
                leaf.pts = np.empty((0, 3), dtype=np.float64)
            cell.scheme = scheme
            for tree in cell.trees.values():  # octree_manager.py:65-66: EVERY pose of the cell
                tree.subdivide_as(tree.root, scheme.root)

    # -- grid.py:260-267 -> octree.py:102-112 -----------------------------------------
    def filter(self, criteria: Sequence[Callable], pose_numbers: Optional[Sequence[int]] = None):
        for cell in self.cells.values():
            poses = cell.trees.keys() if pose_numbers is None else pose_numbers
            for p in poses:
                for leaf in cell.trees[p].cache.values():
                    if not all([c(leaf.pts) for c in criteria]):
                        leaf.idx = np.empty((0,), dtype=np.int64)
                        leaf.pts = np.empty((0, 3), dtype=np.float64)

    # -- grid.py:217-232 -> octree.py:256-263 -----------------------------------------
    def get_leaf_points(self, pose: int, non_empty: bool = True) -> List[OracleLeaf]:
        out = []
        for key in self.pose_cells[pose]:
            tree = self.cells[key].trees[pose]
            for leaf in tree.cache.values():
                if non_empty and len(leaf.idx) == 0:
                    continue
                out.append(OracleLeaf(leaf.corner, leaf.edge, leaf.pts, leaf.idx, key))
        return out

    # -- grid.py:234-242 -> octree.py:55-65 -------------------------------------------
    def get_points(self, pose: int) -> np.ndarray:
        parts = [np.empty((0, 3), dtype=float)]
        for cell in self.cells.values():
            tree = cell.trees.get(pose)
            if tree is not None:
                parts.append(tree._collect(tree.root)[1])
        return np.vstack(parts)

    def get_point_indices(self, pose: int) -> np.ndarray:
        parts = [np.empty((0,), dtype=np.int64)]
        for cell in self.cells.values():
            tree = cell.trees.get(pose)
            if tree is not None:
                parts.append(tree._collect(tree.root)[0])
        return np.concatenate(parts)

    # -- grid.py:343-362 --------------------------------------------------------------
    def n_leaves(self, pose: int) -> int:
        return sum(
            sum(1 for leaf in c.trees[pose].cache.values() if len(leaf.idx))
            for c in self.cells.values() if pose in c.trees
        )

    def n_points(self, pose: int) -> int:
        return sum(
            sum(len(leaf.idx) for leaf in c.trees[pose].cache.values())
            for c in self.cells.values() if pose in c.trees
        )

    def n_nodes(self, pose: int) -> int:
        return sum(c.trees[pose].n_nodes() for c in self.cells.values() if pose in c.trees)

    # -- grid.py:203-215 -> octree.py:265-274 -> 137-142 ------------------------------
    def apply_mask(self, pose: int, mask: np.ndarray):
        """`mask` covers the pose's points in get_leaf_points order."""
        pos = 0
        for key in self.pose_cells[pose]:
            tree = self.cells[key].trees[pose]
            for leaf in tree.cache.values():
                n = len(leaf.idx)
                if n == 0:
                    continue
                m = mask[pos:pos + n]
                leaf.idx = leaf.idx[m]
                leaf.pts = leaf.pts[m]
                pos += n
        assert pos == len(mask)

    # -- grid.py:124-215 --------------------------------------------------------------
    def ransac_batches(self, poses_per_batch: int):
        n = len(self.pose_cells)
        return [list(range(i, min(i + poses_per_batch, n))) for i in range(0, n, poses_per_batch)]

    def map_leaf_points_ransac(self, table: np.ndarray, threshold: float = 0.01, poses_per_batch: int = 10,
                               evaluate=None):
        """Restates grid.py:124-215 with the hypothesis table passed in (the reference draws it
        from the global numpy RNG, ransac/cuda_ransac.py:39-41).  `evaluate(points, block_sizes,
        table, threshold) -> dict(mask=..., ...)` defaults to the C oracle.  Returns the per-batch
        results for inspection."""
        if threshold <= 0:
            raise ValueError("Threshold must be positive")
        if evaluate is None:
            from oracle.ransac import ransac_evaluate as evaluate
        results = []
        for batch in self.ransac_batches(poses_per_batch):
            clouds, sizes = [], []
            for p in batch:
                leaves = self.get_leaf_points(p)
                clouds.append(np.vstack([np.empty((0, 3))] + [l.points for l in leaves]))
                sizes.append(np.array([len(l.points) for l in leaves], dtype=np.int32))
            cloud = np.vstack(clouds)
            block_sizes = np.concatenate(sizes)
            res = evaluate(cloud, block_sizes, table, threshold)
            res["batch"] = batch
            res["block_sizes"] = block_sizes
            res["points"] = cloud
            results.append(res)
            pos = 0
            for p, s in zip(batch, sizes):
                n = int(s.sum())
                self.apply_mask(p, res["mask"][pos:pos + n].astype(bool))
                pos += n
        return results
