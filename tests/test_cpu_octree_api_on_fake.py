"""The reference's own octree / octree-manager tests (test/octree/test_multi_pose.py, test/octree/test_octree.py),
re-pointed at this package's host classes running on the CPU stand-in for a single-cell forest.  The same tests run
against the CUDA forest in test_gpu_structure.py; the expectations beyond the reference's own asserts (append after a
subdivision, apply_mask, node counts) were read off the real reference while writing the tests."""
import numpy as np
import pytest

from fake_forest import FakeSingleCellForest
from octreelib_b200.internal import Voxel
from octreelib_b200.octree import Octree, OctreeConfig, OctreeNode
from octreelib_b200.octree_manager import OctreeManager


def _multi_pose():
    mp = OctreeManager(Octree, OctreeConfig(), np.array([0, 0, 0]), 5)
    mp._host._forest = FakeSingleCellForest(5, np.array([0, 0, 0]))
    c0 = np.array([[0, 0, 1], [0, 0, 2], [0, 0, 3]], dtype=float)
    c1 = np.array([[1, 0, 1], [4, 0, 2], [0, 2, 3]], dtype=float)
    mp.insert_points(0, c0)
    mp.insert_points(1, c1)
    return mp, {0: c0, 1: c1}


@pytest.mark.parametrize("crit,poses,nodes,leaves", [(2, [0], [9, 9], [2, 3]), (1, None, [33, 33], [3, 3])])
def test_multi_pose_subdivide(crit, poses, nodes, leaves):
    """test/octree/test_multi_pose.py:36-68"""
    mp, _ = _multi_pose()
    assert [mp.n_nodes(0), mp.n_nodes(1), mp.n_leaves(0), mp.n_leaves(1)] == [1, 1, 1, 1]
    mp.subdivide([lambda points: len(points) > crit], poses)
    assert [mp.n_nodes(0), mp.n_nodes(1)] == nodes
    assert [mp.n_leaves(0), mp.n_leaves(1)] == leaves


def test_multi_pose_map_leaf_points():
    """test/octree/test_multi_pose.py:71-75"""
    mp, _ = _multi_pose()
    mp.map_leaf_points(lambda points: points[0].reshape((1, 3)), [0])
    assert mp.n_points(0) == 1 and mp.n_points(1) == 3


def test_multi_pose_leaf_voxels():
    """test/octree/test_multi_pose.py:78-130"""
    mp, _ = _multi_pose()
    mp.subdivide([lambda points: len(points) > 2], [0])
    exp0 = [Voxel(np.array([0, 0, 0]), 2.5), Voxel(np.array([0, 0, 2.5]), 2.5)]
    exp1 = exp0 + [Voxel(np.array([2.5, 0, 0]), 2.5)]
    assert {v.id for v in mp.get_leaf_points(pose_number=0)} == {v.id for v in exp0}
    assert {v.id for v in mp.get_leaf_points(pose_number=1)} == {v.id for v in exp1}
    mp2, _ = _multi_pose()
    mp2.subdivide([lambda points: len(points) > 1], None)
    e0 = [Voxel(np.array([0, 0, 0.625]), 0.625), Voxel(np.array([0, 0, 1.25]), 1.25), Voxel(np.array([0, 0, 2.5]), 1.25)]
    e1 = [Voxel(np.array([0.625, 0, 0.625]), 0.625), Voxel(np.array([0, 1.25, 2.5]), 1.25), Voxel(np.array([2.5, 0, 0]), 2.5)]
    assert {v.id for v in mp2.get_leaf_points(pose_number=0)} == {v.id for v in e0}
    assert {v.id for v in mp2.get_leaf_points(pose_number=1)} == {v.id for v in e1}
    assert len(mp2.get_leaf_points()) == 6 and mp2.get_leaf_points(pose_number=5) == []


def test_multi_pose_filter_and_points():
    """test/octree/test_multi_pose.py:133-160"""
    mp, clouds = _multi_pose()
    as_set = lambda a: set(map(str, a.tolist()))  # noqa: E731
    assert as_set(mp.get_points(0)) == as_set(clouds[0]) and as_set(mp.get_points(1)) == as_set(clouds[1])
    assert as_set(mp.get_points()) == as_set(np.vstack([clouds[0], clouds[1]]))
    assert mp.n_points(0) == 3 and mp.n_points(1) == 3
    assert mp.n_points() == 6  # the intended sum; the reference's own line (octree_manager.py:138) calls a property and raises
    mp.subdivide([lambda points: len(points) > 2], [0])
    mp.filter([lambda points: False], [0])
    mp.filter([lambda points: True], [1])
    assert mp.n_points(0) == 0 and mp.n_points(1) == 3


def test_multi_pose_append_follows_the_scheme_and_apply_mask():
    """octree_manager.py:161-180: points appended to an existing pose are routed through the existing shape"""
    mp, _ = _multi_pose()
    mp.subdivide([lambda points: len(points) > 2], [0])
    mp.insert_points(1, np.array([[0.5, 0.5, 0.5], [4.5, 4.5, 4.5]]))
    assert mp.n_points(1) == 5 and mp.n_leaves(1) == 4 and mp.n_nodes(1) == 9  # n_leaves counts non-empty leaves
    sizes = sorted(v.n_points for v in mp.get_leaf_points(pose_number=1))
    assert sizes == [1, 1, 1, 2]
    mask = np.zeros(5, dtype=bool)
    mask[0] = True
    mp.apply_mask(mask, 1)
    assert mp.n_points(1) == 1


_CLOUD = np.array([[0, 0, 1], [0, 0, 2], [0, 0, 3], [9, 9, 8], [9, 9, 9]], dtype=float)


def test_octree():
    """test/octree/test_octree.py:8-18"""
    octree = Octree(OctreeConfig(), np.array([0, 0, 0]), np.float64(10))
    octree._host._forest = FakeSingleCellForest(np.float64(10), np.array([0, 0, 0]))
    octree.insert_points(_CLOUD)
    assert (_CLOUD == octree.get_points()).all()
    octree.subdivide([lambda points: len(points) > 2])
    assert octree.n_leaves == 3 and octree.n_points == 5 and octree.n_nodes == 17
    octree.filter([lambda points: len(points) >= 2])
    assert octree.n_points == 4
    octree.subdivide_as(octree)  # its own scheme: nothing changes
    assert octree.n_leaves == 2 and octree.n_points == 4 and octree.n_nodes == 17
    other = Octree(OctreeConfig(), np.array([0, 0, 0]), np.float64(10))
    other._host._forest = FakeSingleCellForest(np.float64(10), np.array([0, 0, 0]))
    other.insert_points(_CLOUD)
    other.subdivide_as(octree)   # octree.py:34-53: the other tree's scheme, this tree's points
    assert other.n_nodes == 17 and other.n_points == 5 and other.n_leaves == 3
    with pytest.raises(TypeError):
        other.subdivide_as("not an octree")


def test_octree_node():
    """test/octree/test_octree.py:21-30"""
    cached = []
    node = OctreeNode(np.array([0, 0, 0]), np.float64(10), cached)
    node._host._forest = FakeSingleCellForest(np.float64(10), np.array([0, 0, 0]))
    node.insert_points(_CLOUD)
    node.subdivide([lambda points: len(points) > 2])
    assert node.n_leaves == 3 and node.n_points == 5
    node.filter([lambda points: len(points) >= 2])
    assert node.n_points == 4
    assert len(cached) == 15
    assert sorted(v.n_points for v in node.get_leaf_points()) == [2, 2]


def test_standalone_octree_second_subdivide_is_a_no_op_like_the_reference():
    """octree.py:26: on a split root the criteria see the root's own empty point array, so a stand-alone `Octree` /
    `OctreeNode` never deepens or coarsens on a repeated call (checked against the real reference: n_nodes stays 9)."""
    pts = np.array([[0, 0, 1], [0, 0, 2], [0, 0, 3], [4, 4, 4], [4, 4, 4.5]], dtype=float)
    tree = Octree(OctreeConfig(), np.array([0, 0, 0]), 5)
    tree._host._forest = FakeSingleCellForest(5, np.array([0, 0, 0]))
    tree.insert_points(pts)
    tree.subdivide([lambda points: len(points) > 2])
    before = (tree.n_nodes, tree.n_leaves, [v.n_points for v in tree.get_leaf_points()])
    tree.subdivide([lambda points: len(points) > 1])   # finer: would deepen if the shape were rebuilt
    tree.subdivide([lambda points: len(points) > 100])  # coarser: would collapse
    assert (tree.n_nodes, tree.n_leaves, [v.n_points for v in tree.get_leaf_points()]) == before
    with pytest.raises(NotImplementedError):
        tree.subdivide([lambda points: len(points) >= 0])


def test_map_leaf_points_with_new_coordinates_on_the_stand_in():
    """octree.py:114-123 with a function that changes coordinates (every leaf shrunk towards its centroid): the host layer
    masks the old points, appends the new ones to the pose and checks, leaf by leaf, that the rebuilt tree holds exactly
    what the function returned (ForestHost.map_leaf_points); a result that leaves its leaf is refused up front."""
    rng = np.random.default_rng(8)
    cloud = np.vstack([rng.normal([3, 3, 3], 0.8, (150, 3)), rng.uniform(0, 10, (100, 3))])
    cloud = cloud[((cloud >= 0) & (cloud < 10)).all(axis=1)]
    tree = Octree(OctreeConfig(), np.array([0, 0, 0]), np.float64(10))
    tree._host._forest = FakeSingleCellForest(np.float64(10), np.array([0, 0, 0]))
    tree.insert_points(cloud)
    tree.subdivide([lambda points: len(points) > 12])
    before = [v.get_points().copy() for v in tree.get_leaf_points()]
    n_nodes = tree.n_nodes

    def shrink(points):
        c = points.mean(axis=0)
        return c + 0.5 * (points - c)

    tree.map_leaf_points(shrink)
    after = tree.get_leaf_points()
    assert len(after) == len(before) and tree.n_nodes == n_nodes
    for v, b in zip(after, before):
        assert (v.get_points() == shrink(b)).all()
    with pytest.raises(NotImplementedError):
        tree.map_leaf_points(lambda points: points + 100.0)
    assert tree.n_points == sum(len(b) for b in before)
