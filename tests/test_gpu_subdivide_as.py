"""`Octree.subdivide_as` / `OctreeNode.subdivide_as` (octree/octree.py:34-53, 222-227) on the native forest, pinned against
tests/golden/subdivide_as_edge16.npz - recorded from the REAL reference by tests/golden/make_subdivide_as.py."""
import numpy as np
import pytest

from conftest import golden
from octreelib_b200.criteria import MaxPoints
from octreelib_b200.octree import Octree, OctreeConfig, OctreeNode

pytestmark = pytest.mark.gpu


def _tree(g, cloud, cls=Octree):
    if cls is Octree:
        t = Octree(OctreeConfig(), g["corner"], np.float64(g["edge"]))
    else:
        t = OctreeNode(g["corner"], np.float64(g["edge"]), [])
    t.insert_points(cloud)
    return t


def _leaf_table(tree):
    leaves = tree.get_leaf_points()
    return (np.array([v.corner_min for v in leaves], dtype=np.float64).reshape(-1, 3),
            np.array([v.edge_length for v in leaves], dtype=np.float64),
            [np.asarray(v.get_points(), dtype=np.float64).reshape(-1, 3) for v in leaves])


@pytest.mark.parametrize("cls", [Octree, OctreeNode])
def test_subdivide_as_matches_the_reference(cls):
    g = golden("subdivide_as_edge16")
    a = _tree(g, g["a"], cls)
    a.subdivide([MaxPoints(int(g["max_points"]))])
    b = _tree(g, g["b"], cls)
    b.subdivide_as(a)
    corner, edge, pts = _leaf_table(b)
    assert (corner == g["b_corner"]).all() and (edge == g["b_edge"]).all()
    assert [len(p) for p in pts] == g["b_sizes"].tolist()
    assert (np.vstack(pts) == g["b_points"]).all()          # same points, same order inside every leaf
    assert (b.n_leaves, b.n_nodes, b.n_points) == (int(g["b_n_leaves"]), int(g["b_n_nodes"]), int(g["b_n_points"]))
    # the scheme really is A's: same internal nodes
    sa, sb = a._host.forest.export_shape(), b._host.forest.export_shape()
    assert sorted(zip(sa["depth"].tolist(), sa["path"].tolist())) == sorted(zip(sb["depth"].tolist(), sb["path"].tolist()))


def test_subdivide_as_collapses_where_the_other_tree_is_coarser():
    """The reference raises ValueError when a collapsed node has grandchildren (`collapse_deep_reference_error` in the
    fixture: list.remove of a node that is not a leaf) and drops a collapsed node from its leaf list otherwise
    (`collapse_one_level_listed_leaves` == 0 although get_points keeps every point).  Here the result is simply the other
    tree's scheme with every point kept - the same tree a fresh `subdivide_as` builds."""
    g = golden("subdivide_as_edge16")
    assert "ValueError" in str(g["collapse_deep_reference_error"])
    a = _tree(g, g["a"])
    a.subdivide([MaxPoints(int(g["max_points"]))])
    fresh = _tree(g, g["b"])
    fresh.subdivide_as(a)
    b = _tree(g, g["b"])
    b.subdivide([MaxPoints(8)])                      # finer than A everywhere
    assert b.n_nodes > fresh.n_nodes
    b.subdivide_as(a)
    for got, want in zip(_leaf_table(b)[:2], _leaf_table(fresh)[:2]):
        assert (got == want).all()
    assert (np.vstack(_leaf_table(b)[2]) == np.vstack(_leaf_table(fresh)[2])).all()
    assert (b.n_leaves, b.n_nodes, b.n_points) == (fresh.n_leaves, fresh.n_nodes, fresh.n_points)
    # one level collapsed onto an unsplit tree
    unsplit = _tree(g, g["a"])
    c = _tree(g, g["b"])
    c.subdivide([MaxPoints(1500)])
    assert c.n_nodes == 9
    c.subdivide_as(unsplit)
    assert c.n_nodes == int(g["collapse_one_level_n_nodes"]) == 1
    assert c.n_points == int(g["collapse_one_level_n_points"]) and len(c.get_points()) == int(g["collapse_one_level_points"])
    assert len(c.get_leaf_points()) == 1             # (the reference lists none: octree.py:48-53)


def test_subdivide_as_of_an_empty_or_foreign_tree():
    g = golden("subdivide_as_edge16")
    a = _tree(g, g["a"])
    a.subdivide([MaxPoints(int(g["max_points"]))])
    empty = Octree(OctreeConfig(), g["corner"], np.float64(g["edge"]))
    empty.subdivide_as(a)                            # nothing stored: nothing to route
    assert empty.n_points == 0
    a.subdivide_as(empty)                            # the other tree is unsplit: collapse
    assert a.n_nodes == 1 and a.n_points == len(g["a"])
    with pytest.raises(TypeError):
        a.subdivide_as(object())
