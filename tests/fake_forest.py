"""CPU stand-in for `octreelib_b200.forest.Forest`, backed by the oracle (test infrastructure only).

It answers the calls `ForestHost` / `_views` make with tables in the NATIVE forest's format and order: cells
lexicographic, the leaf table of a cell in the one-call order (depth-first rank of the parent, child id) plus the epoch of
every leaf's parent, and the (pose, leaf) blocks ordered by the SAME rule the native forest implements on the device
(csrc/forest.cuh): inside (pose, cell) by (max(epoch of the leaf's parent, epoch of the pose), one-call order).  The
rule is restated here from the epochs alone - NOT read off the oracle's leaf lists - so the CPU tests compare the rule
with the oracle's history-keeping lists.  That lets the host-side logic (`_host.py`, `_views.py`, `grid/grid.py`) run in
the `-m "not gpu"` suite.
"""
import numpy as np

from oracle.structure import OracleGrid, max_points_criterion


def _shape_leaves(node, path=()):
    """(parent path or None, child id, path) of the leaves below `node`, depth-first"""
    if node.children is None:
        return [(None, 0, path)] if path == () else []
    out = []
    for cid, ch in enumerate(node.children):
        if ch.children is None:
            out.append((path, cid, path + (cid,)))
        else:
            out += _shape_leaves(ch, path + (cid,))
    return out


def _internal_paths(node, path=()):
    if node.children is None:
        return []
    out = [path]
    for cid, ch in enumerate(node.children):
        out += _internal_paths(ch, path + (cid,))
    return out


def _node_at(root, path):
    node = root
    for cid in path:
        if node.children is None:
            return None
        node = node.children[cid]
    return node


class FakeForest:
    def __init__(self, edge):
        self.og = OracleGrid(edge)
        self.edge = edge
        self.version = 0
        self.clouds = []
        self.n_subdivides = 0
        self.node_epoch = {}   # (cell key, path) of an internal node -> subdivide call that first split it
        self.pose_epoch = {}   # pose index -> subdivide calls made before the pose was created

    # ---- mutations -------------------------------------------------------------------------------
    def insert(self, points):
        idx = len(self.clouds)
        pts = np.asarray(points, dtype=np.float64)
        self.last_insert_rows = len(pts)
        self.clouds.append(pts)
        self.pose_epoch[idx] = self.n_subdivides
        self.og.insert_points(idx, pts)
        self.version += 1
        return idx

    def _after_subdivide(self):
        """epochs of the rebuilt shape: carried over where the node existed before, the current call where it is new"""
        self.n_subdivides += 1
        old, self.node_epoch = self.node_epoch, {}
        for key in self.og.cells:
            root, _ = self._shape(key)
            for path in _internal_paths(root):
                self.node_epoch[(key, path)] = old.get((key, path), self.n_subdivides)
        self.version += 1

    def subdivide(self, max_points, pose_indices=None):
        self.og.subdivide([max_points_criterion(int(max_points))], pose_indices)
        self._after_subdivide()

    def subdivide_table(self, table, beyond, pose_indices=None):
        table = np.asarray(table)
        self.og.subdivide([lambda pts: bool(table[len(pts)]) if len(pts) < len(table) else bool(beyond)], pose_indices)
        self._after_subdivide()

    def subdivide_levels(self, first_levels, thresholds=None, tables=None, beyonds=None, pose_indices=None):
        """level-dependent rule (node-size thresholds): the level of the node under test is recovered from its edge"""
        from octreelib_b200.criteria import _edge_of_node_under_test, halvings

        firsts = [int(f) for f in first_levels]

        def crit(pts):
            level = halvings(self.edge, _edge_of_node_under_test())
            e = max(i for i, f in enumerate(firsts) if f <= level)
            if tables is None:
                return len(pts) > int(thresholds[e])
            table = np.asarray(tables[e])
            return bool(table[len(pts)]) if len(pts) < len(table) else bool(beyonds[e])

        self.og.subdivide([crit], pose_indices)
        self._after_subdivide()

    def impose_shape(self, shape):
        """ol_forest_impose_shape: exactly the listed nodes (cell coordinates, depth, Morton path) become the split nodes"""
        from oracle.structure import _Tree

        keys, cells, _, _ = self._tables()
        by_q = {tuple(int(v) for v in cells["q"][i]): keys[i] for i in range(len(keys))}
        wanted = {key: [] for key in keys}
        for q, depth, path in zip(np.asarray(shape["q"]).reshape(-1, 3).tolist(), shape["depth"].tolist(), shape["path"].tolist()):
            key = by_q.get(tuple(int(v) for v in q))
            if key is not None:
                wanted[key].append((int(depth), int(path)))
        for key in keys:
            cell = self.og.cells[key]
            skeleton = _Tree(cell.key, cell.edge)
            for depth, path in sorted(wanted[key]):
                node = skeleton.root
                for level in range(depth):
                    node = None if node is None or node.children is None else node.children[(path >> (3 * (depth - 1 - level))) & 7]
                if node is not None and node.children is None:
                    skeleton._generate_children(node)
            cell.scheme = skeleton
            for tree in cell.trees.values():
                tree.subdivide_as(tree.root, skeleton.root)
        old, self.node_epoch = self.node_epoch, {}
        for key in keys:
            root, _ = self._shape(key)
            for path in _internal_paths(root):
                self.node_epoch[(key, path)] = old.get((key, path), max(self.n_subdivides, 1))
        self.version += 1

    def filter(self, keep_table, pose_indices=None):
        table = np.asarray(keep_table)
        self.og.filter([lambda pts: bool(table[min(len(pts), len(table) - 1)])], pose_indices)
        self.version += 1

    def apply_pose_mask(self, pose_index, mask):
        """mask over the pose's points in block order (reference order of its non-empty leaves)"""
        mask = np.asarray(mask, dtype=bool)
        _, _, _, leaf_path = self._tables()
        start = 0
        for pose, lid, idx in self._blocks():
            if pose != pose_index:
                continue
            key, path = leaf_path[lid]
            node = _node_at(self.og.cells[key].trees[pose].root, path)
            keep = mask[start:start + len(idx)]
            node.idx, node.pts = node.idx[keep], node.pts[keep]
            start += len(idx)
        assert start == len(mask)
        self.version += 1

    # ---- tables in the native order --------------------------------------------------------------------
    def _cell_keys(self):
        return sorted(self.og.cells.keys())

    def _shape(self, key):
        """leaf paths of a cell in the forest's one-call order; the shape is the one every pose tree of the cell has"""
        cell = self.og.cells[key]
        root = next(iter(cell.trees.values())).root if cell.trees else cell.scheme.root
        rank = {p: r for r, p in enumerate(_internal_paths(root))}
        leaves = _shape_leaves(root)
        leaves.sort(key=lambda t: (0, 0) if t[0] is None else (rank[t[0]], t[1]))
        return root, [t[2] for t in leaves]

    def _tables(self):
        keys = self._cell_keys()
        cells_q, cells_corner, first_pose, n_nodes, leaf_begin = [], [], [], [], [0]
        leaf_corner, leaf_edge, leaf_cell, leaf_depth, leaf_path, leaf_epoch = [], [], [], [], [], []
        for ci, key in enumerate(keys):
            cell = self.og.cells[key]
            root, paths = self._shape(key)
            cells_q.append([int(round(k / self.edge)) for k in key])
            cells_corner.append([float(k) for k in key])
            first_pose.append(min(cell.trees.keys()))
            n_nodes.append(len(paths) + (len(paths) - 1) // 7)
            for path in paths:
                node = _node_at(root, path)
                leaf_corner.append(np.asarray(node.corner, dtype=np.float64))
                leaf_edge.append(float(node.edge))
                leaf_cell.append(ci)
                leaf_depth.append(len(path))
                leaf_path.append((key, path))
                leaf_epoch.append(self.node_epoch.get((key, path[:-1]), 1) if path else 0)
            leaf_begin.append(len(leaf_cell))
        cells = dict(q=np.array(cells_q, dtype=np.int64).reshape(-1, 3), corner=np.array(cells_corner, dtype=np.float64).reshape(-1, 3),
                     first_pose=np.array(first_pose, dtype=np.int32), n_nodes=np.array(n_nodes, dtype=np.int64),
                     leaf_begin=np.array(leaf_begin, dtype=np.int64))
        leaves = dict(corner=np.array(leaf_corner, dtype=np.float64).reshape(-1, 3), edge=np.array(leaf_edge, dtype=np.float64),
                      cell=np.array(leaf_cell, dtype=np.int32), depth=np.array(leaf_depth, dtype=np.int32),
                      parent_epoch=np.array(leaf_epoch, dtype=np.int32))
        return keys, cells, leaves, leaf_path

    def export_cells(self):
        return self._tables()[1]

    def export_leaves(self):
        return self._tables()[2]

    def export_cell_poses(self):
        keys = self._cell_keys()
        pairs = [(ci, p) for ci, key in enumerate(keys) for p in sorted(self.og.cells[key].trees.keys())]
        return dict(cell=np.array([c for c, _ in pairs], dtype=np.int32), pose=np.array([p for _, p in pairs], dtype=np.int32))

    def _blocks(self):
        _, cells, leaves, leaf_path = self._tables()
        out = []  # (pose, leaf id, idx array)
        begin = cells["leaf_begin"]
        for pose in range(len(self.clouds)):
            for ci in range(len(begin) - 1):
                ids = list(range(int(begin[ci]), int(begin[ci + 1])))
                if self.n_subdivides >= 2:  # the device rule: (max(parent epoch, pose epoch), one-call order), stable
                    ids.sort(key=lambda lid: max(int(leaves["parent_epoch"][lid]), self.pose_epoch.get(pose, 0)))
                for lid in ids:
                    key, path = leaf_path[lid]
                    tree = self.og.cells[key].trees.get(pose)
                    if tree is None:
                        continue
                    node = _node_at(tree.root, path)
                    if node is not None and len(node.idx):
                        out.append((pose, lid, node.idx))
        return out

    def export_blocks(self, pose_rank=None):
        blocks = self._blocks()
        return dict(pose=np.array([b[0] for b in blocks], dtype=np.int32), leaf=np.array([b[1] for b in blocks], dtype=np.int32),
                    size=np.array([len(b[2]) for b in blocks], dtype=np.int32))

    def export_points(self, pose_index=-1, order=0, pose_rank=None, n_hint=None, want_mask=False):
        assert pose_index >= 0, "the fake only serves per-pose exports"
        if order == 1:  # cells lexicographic x depth-first leaves
            idx, cell = [], []
            for ci, key in enumerate(self._cell_keys()):
                tree = self.og.cells[key].trees.get(pose_index)
                if tree is None:
                    continue
                for leaf in tree.leaves_dfs():
                    idx.append(leaf.idx)
                    cell.append(np.full(len(leaf.idx), ci, dtype=np.int32))
            idx = np.concatenate(idx) if idx else np.empty(0, dtype=np.int64)
            return dict(xyz=self.clouds[pose_index][idx], idx=idx, cell=np.concatenate(cell) if cell else np.empty(0, dtype=np.int32))
        _, _, leaves, _ = self._tables()
        idx = [b[2] for b in self._blocks() if b[0] == pose_index]
        cell = [np.full(len(b[2]), leaves["cell"][b[1]], dtype=np.int32) for b in self._blocks() if b[0] == pose_index]
        idx = np.concatenate(idx) if idx else np.empty(0, dtype=np.int64)
        return dict(xyz=self.clouds[pose_index][idx], idx=idx, cell=np.concatenate(cell) if cell else np.empty(0, dtype=np.int32))

    def stats(self, light=False):
        _, cells, leaves, _ = self._tables()
        blocks = self._blocks()
        return dict(n_points_inserted=sum(len(c) for c in self.clouds), n_points_alive=sum(len(b[2]) for b in blocks),
                    n_poses=len(self.clouds), n_cells=len(cells["q"]), n_leaves=len(leaves["edge"]), n_blocks=len(blocks),
                    n_internal=sum(len(_internal_paths(self._shape(k)[0])) for k in self.og.cells), sample_oob_seen=0,
                    max_block_size=max([len(b[2]) for b in blocks], default=0))

    def pose_counts(self, n_poses):
        return np.array([[self.og.n_leaves(p), self.og.n_points(p), self.og.n_nodes(p)] for p in range(n_poses)], dtype=np.int64)


class FakeSingleCellForest(FakeForest):
    """The forest behind a stand-alone `OctreeManager` / `Octree` / `OctreeNode`: one fixed cell, poses may be appended to."""

    def __init__(self, edge, corner):
        from oracle.structure import _Cell

        super().__init__(edge)
        self.key = tuple(float(c) for c in np.asarray(corner).reshape(3))
        self.og.cells[self.key] = _Cell(np.asarray(corner), edge)

    def _insert_into(self, pose, idx, pts):
        from oracle.structure import _Tree

        cell = self.og.cells[self.key]
        self.og.pose_cells.setdefault(pose, [self.key])
        tree = cell.trees.get(pose)
        if tree is None:
            tree = cell.trees[pose] = _Tree(cell.key, cell.edge)
        tree.insert(tree.root, idx.astype(np.int64), pts)
        tree.subdivide_as(tree.root, cell.scheme.root)
        self.version += 1

    # ---- Octree.subdivide_as: the scheme as data (ol_forest_export_shape / ol_forest_impose_shape) ----
    def export_shape(self):
        out = []

        def walk(node, depth, path):
            if node.children is None:
                return
            out.append((depth, path))
            for c, ch in enumerate(node.children):
                walk(ch, depth + 1, (path << 3) | c)

        walk(self.og.cells[self.key].scheme.root, 0, 0)
        return dict(q=np.zeros((len(out), 3), np.int64), depth=np.array([d for d, _ in out], np.uint32),
                    path=np.array([p for _, p in out], np.uint64))

    def impose_shape(self, shape):
        from oracle.structure import _Tree

        cell = self.og.cells[self.key]
        skeleton = _Tree(cell.key, cell.edge)
        for depth, path in sorted(zip(shape["depth"].tolist(), shape["path"].tolist())):
            node = skeleton.root
            for level in range(depth):
                node = node.children[(path >> (3 * (depth - 1 - level))) & 7]
            if node.children is None:
                skeleton._generate_children(node)
        cell.scheme = skeleton
        for tree in cell.trees.values():
            tree.subdivide_as(tree.root, skeleton.root)
        self.version += 1

    def insert(self, points):
        pose = len(self.clouds)
        pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
        self.last_insert_rows = len(pts)
        self.clouds.append(pts)
        self.pose_epoch[pose] = self.n_subdivides
        self._insert_into(pose, np.arange(len(pts)), pts)
        return pose

    def insert_segments(self, points, seg_sizes, seg_pose, seg_first, n_poses_total):
        pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
        assert len(seg_sizes) == 1 and seg_sizes[0] == len(pts)
        pose, first = int(seg_pose[0]), int(seg_first[0])
        assert first == len(self.clouds[pose])
        self.clouds[pose] = np.vstack([self.clouds[pose], pts])
        self._insert_into(pose, first + np.arange(len(pts)), pts)

    def _tables(self):
        keys, cells, leaves, leaf_path = super()._tables()
        cells["q"][:] = 0  # a single-cell forest has no cell coordinates
        return keys, cells, leaves, leaf_path
