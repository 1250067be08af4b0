"""Specification check (CPU) for the leaf order the forest has to produce after MORE than one subdivide.

The reference keeps one `_cached_leaves` list per pose octree across calls (octree_base.py:48-49, octree.py:183-191):
a node that splits is removed from the list and its 8 children are appended, depth-first within one pass.  A pass is a
`subdivide` call (subdivide_as over every pose octree, octree_manager.py:65-66) or the creation of a pose's octree,
which replays the then-current scheme in one go (octree_manager.py:166-171).

Closed form used by DESIGN.md section 8 (planned device implementation): with
    node_epoch(v) = index of the first subdivide call whose scheme has v as an internal node   (1, 2, ...)
    pose_epoch(p) = number of subdivide calls made before pose p was inserted                  (0, 1, ...)
the leaves of pose p inside a cell are ordered by
    (max(node_epoch(parent), pose_epoch(p)), depth-first pre-order rank of the parent, child id),
an unsplit root being the cell's only leaf.  This test derives node epochs by replaying the fixture's call sequence on
the oracle and checks the closed form against the REAL reference's order stored in the golden fixture.
"""
import numpy as np

from conftest import golden
from oracle.structure import OracleGrid, max_points_criterion


def _internal_paths(node, path=()):
    """paths (tuples of child ids) of the internal nodes below `node`, depth-first pre-order"""
    if node.children is None:
        return []
    out = [path]
    for cid, ch in enumerate(node.children):
        out += _internal_paths(ch, path + (cid,))
    return out


def _leaves_with_parent(node, path=()):
    """(parent path or None, child id, leaf node) of the leaves below `node`, depth-first"""
    if node.children is None:
        return [(None, 0, node)] if path == () else []
    out = []
    for cid, ch in enumerate(node.children):
        if ch.children is None:
            out.append((path, cid, ch))
        else:
            out += _leaves_with_parent(ch, path + (cid,))
    return out


def test_epoch_keyed_order_reproduces_the_reference_after_two_subdivides():
    g = golden("resubdivide_deepen_edge4")
    late = [int(p) for p in g["late"]]
    poses = [int(p) for p in g["poses"]]
    early = [p for p in poses if p not in late]
    og = OracleGrid(int(g["edge"]))
    node_epoch = {}  # (cell key, path) -> first subdivide call that made the node internal
    pose_epoch = {}
    calls = 0

    def record():
        for key, cell in og.cells.items():
            for path in _internal_paths(cell.scheme.root):
                node_epoch.setdefault((key, path), calls)

    for p in early:
        og.insert_points(p, g[f"cloud{p}"])
        pose_epoch[p] = calls
    og.subdivide([max_points_criterion(int(g["first_max"]))])
    calls += 1
    record()
    for p in late:
        og.insert_points(p, g[f"cloud{p}"])
        pose_epoch[p] = calls
    og.subdivide([max_points_criterion(int(g["second_max"]))])
    calls += 1
    record()

    differs_from_one_shot = 0
    for p in poses:
        corners, edges = [], []
        one_shot = []
        for key in og.pose_cells[p]:
            tree = og.cells[key].trees[p]
            rank = {path: r for r, path in enumerate(_internal_paths(tree.root))}
            keyed = []
            for parent, cid, leaf in _leaves_with_parent(tree.root):
                if len(leaf.idx) == 0:
                    continue  # get_leaf_points(non_empty=True)
                if parent is None:
                    keyed.append(((0, 0, 0), (0, 0), leaf))
                    continue
                epoch = max(node_epoch[(key, parent)], pose_epoch[p])
                keyed.append(((epoch, rank[parent], cid), (rank[parent], cid), leaf))
            for _, _, leaf in sorted(keyed, key=lambda t: t[0]):
                corners.append(np.asarray(leaf.corner, dtype=np.float64))
                edges.append(float(leaf.edge))
            one_shot += [np.asarray(leaf.corner, dtype=np.float64) for _, _, leaf in sorted(keyed, key=lambda t: t[1])]
        corners = np.array(corners).reshape(-1, 3)
        assert (corners == g[f"p{p}_corner"]).all(), f"pose {p}: the epoch-keyed order is not the reference's order"
        assert (np.array(edges) == g[f"p{p}_edge"]).all()
        differs_from_one_shot += int((np.array(one_shot).reshape(-1, 3) != g[f"p{p}_corner"]).any())
    assert differs_from_one_shot > 0  # the plain (depth-first rank, child id) order is NOT enough for this fixture
