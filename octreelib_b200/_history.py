"""
Leaf enumeration order after MORE than one `subdivide` call (host side, numpy on exported tables).

Every pose octree of the reference keeps its `_cached_leaves` list across calls (octree_base.py:48-49): a node that
splits leaves the list and its 8 children are appended, depth-first within one pass (octree.py:183-191).  A pass is a
`subdivide` call (`subdivide_as` over every pose octree, octree_manager.py:65-66) or the creation of a pose's octree,
which replays the then-current scheme in one go (octree_manager.py:166-171).  With

    node_epoch(v) = index of the first subdivide call whose scheme has v as an internal node   (1, 2, ...)
    pose_epoch(p) = number of subdivide calls made before pose p was inserted                  (0, 1, ...)

the leaves of pose p inside a cell are enumerated by
    (max(node_epoch(parent), pose_epoch(p)), depth-first pre-order rank of the parent, child id).
The forest always produces the one-call order (depth-first rank of the parent, child id) - identical as long as there
was a single call, which is the hot path.  From the second call on `ForestHost` keeps a `LeafHistory`: the epoch of
every internal node, keyed by (cell coordinates, depth, node index inside the cell), so that `get_leaf_points` can
put a pose's leaves into the reference's order with one stable sort.  Nothing here runs for a single subdivide.
(Checked against the real reference's order by tests/test_cpu_leaf_order_spec.py and tests/test_cpu_history_order.py.)
"""
from __future__ import annotations

from typing import Optional

import numpy as np

__all__ = ["LeafHistory", "parent_keys", "internal_node_keys"]

_NO_PARENT = -1


def _leaf_indices(leaves: dict, cells: dict):
    """(q[L,3], depth[L], rel[L,3]): cell coordinates, depth and integer position of every leaf inside its cell
    (rel = (leaf corner - cell corner) / leaf edge, in [0, 2^depth))."""
    cell = np.asarray(leaves["cell"], dtype=np.int64)
    depth = np.asarray(leaves["depth"], dtype=np.int64)
    q = np.asarray(cells["q"], dtype=np.int64).reshape(-1, 3)[cell]
    origin = np.asarray(cells["corner"], dtype=np.float64).reshape(-1, 3)[cell]
    corner = np.asarray(leaves["corner"], dtype=np.float64).reshape(-1, 3)
    edge = np.asarray(leaves["edge"], dtype=np.float64).reshape(-1, 1)
    rel = np.rint((corner - origin) / edge).astype(np.int64)
    rel[depth == 0] = 0
    return q, depth, rel


def parent_keys(leaves: dict, cells: dict) -> np.ndarray:
    """[L, 7] int64 key of every leaf's parent node: (qx, qy, qz, depth, ix, iy, iz); depth = -1 for an unsplit root."""
    q, depth, rel = _leaf_indices(leaves, cells)
    out = np.empty((len(depth), 7), dtype=np.int64)
    out[:, 0:3] = q
    out[:, 3] = np.where(depth > 0, depth - 1, _NO_PARENT)
    out[:, 4:7] = np.where((depth > 0)[:, None], rel >> 1, 0)
    return out


def internal_node_keys(leaves: dict, cells: dict) -> np.ndarray:
    """[n, 7] int64, unique: every internal node of the current shape = every proper ancestor of a leaf."""
    q, depth, rel = _leaf_indices(leaves, cells)
    rows = []
    max_depth = int(depth.max()) if len(depth) else 0
    for up in range(1, max_depth + 1):  # the ancestor `up` levels above the leaf
        sel = depth >= up
        if not sel.any():
            break
        block = np.empty((int(sel.sum()), 7), dtype=np.int64)
        block[:, 0:3] = q[sel]
        block[:, 3] = depth[sel] - up
        block[:, 4:7] = rel[sel] >> up
        rows.append(block)
    if not rows:
        return np.empty((0, 7), dtype=np.int64)
    return np.unique(np.vstack(rows), axis=0)


def _lookup(table_keys: np.ndarray, table_vals: np.ndarray, query: np.ndarray, missing: int) -> np.ndarray:
    """values of `query` rows in the (unique-row) table, `missing` where a row is absent"""
    if len(query) == 0:
        return np.empty(0, dtype=np.int64)
    if len(table_keys) == 0:
        return np.full(len(query), missing, dtype=np.int64)
    both = np.vstack([table_keys, query])
    _, inv = np.unique(both, axis=0, return_inverse=True)
    inv = np.asarray(inv).reshape(-1)
    by_id = np.full(int(inv.max()) + 1, missing, dtype=np.int64)
    by_id[inv[:len(table_keys)]] = table_vals
    return by_id[inv[len(table_keys):]]


class LeafHistory:
    """Epoch of every internal node of the current shape (see the module docstring)."""

    def __init__(self):
        self.keys = np.empty((0, 7), dtype=np.int64)
        self.epochs = np.empty(0, dtype=np.int64)
        self._parent_epoch_cache: Optional[tuple] = None

    def record(self, leaves: dict, cells: dict, epoch: int) -> None:
        """The shape now is `leaves`: nodes seen before keep their epoch, the others were split by call `epoch`."""
        nodes = internal_node_keys(leaves, cells)
        known = _lookup(self.keys, self.epochs, nodes, -1)
        self.keys = nodes
        self.epochs = np.where(known >= 0, known, int(epoch))
        self._parent_epoch_cache = None

    @property
    def trivial(self) -> bool:
        """one epoch only: the one-call order is the reference's order for every pose"""
        return len(self.epochs) == 0 or int(self.epochs.min()) == int(self.epochs.max())

    def parent_epochs(self, leaves: dict, cells: dict, version) -> np.ndarray:
        """[L] epoch of every leaf's parent (0 for an unsplit root); cached per table version"""
        if self._parent_epoch_cache is None or self._parent_epoch_cache[0] != version:
            pk = parent_keys(leaves, cells)
            ep = _lookup(self.keys, self.epochs, pk, 0)
            ep[pk[:, 3] == _NO_PARENT] = 0
            self._parent_epoch_cache = (version, ep)
        return self._parent_epoch_cache[1]

    def order(self, leaf_ids: np.ndarray, leaves: dict, cells: dict, version, pose_epoch: int) -> np.ndarray:
        """Permutation that puts `leaf_ids` (leaves of one pose, in the forest's one-call order: cell-major, then
        (depth-first rank of the parent, child id)) into the reference's order for a pose created at `pose_epoch`."""
        leaf_ids = np.asarray(leaf_ids, dtype=np.int64)
        if len(leaf_ids) == 0:
            return np.empty(0, dtype=np.int64)
        eff = np.maximum(self.parent_epochs(leaves, cells, version)[leaf_ids], int(pose_epoch))
        cell = np.asarray(leaves["cell"], dtype=np.int64)[leaf_ids]
        return np.lexsort((np.arange(len(leaf_ids)), eff, cell))
