"""`get_leaf_points` order after more than one subdivide, on the CPU.

The native forest orders a pose's blocks inside a cell by (max(epoch of the leaf's parent, epoch of the pose), one-call
order) on the device (csrc/forest.cuh, `block_arrange_history_kernel`); `fake_forest.FakeForest` restates exactly that
rule from the epochs, and the public API (`Grid`, `ForestHost`, `_views`) runs unchanged on top of it.  Expected values:
the golden fixture generated from the REAL reference, and the oracle's own leaf lists (which restate the reference's
`_cached_leaves` history).  The same scenarios run against the CUDA forest in test_gpu_structure.py.
"""
import numpy as np
import pytest

from conftest import golden
from fake_forest import FakeForest
from octreelib_b200.criteria import MaxPoints
from octreelib_b200.grid import Grid, GridConfig


def _grid(edge):
    grid = Grid(GridConfig(voxel_edge_length=edge))
    grid._host._forest = FakeForest(edge)
    return grid


def _leaf_table(grid, pose, non_empty=True):
    vox = grid.get_leaf_points(pose, non_empty)
    corner = np.array([np.asarray(v.corner_min, dtype=np.float64) for v in vox]).reshape(-1, 3)
    edge = np.array([float(v.edge_length) for v in vox])
    sizes = np.array([v.n_points for v in vox], dtype=np.int64)
    pts = np.vstack([np.empty((0, 3))] + [v.get_points() for v in vox])
    return corner, edge, sizes, pts


def _check_against_golden(grid, g, poses, prefix):
    for p in poses:
        corner, edge, sizes, pts = _leaf_table(grid, p)
        assert (corner == g[f"{prefix}p{p}_corner"]).all(), f"{prefix} pose {p}: leaf order differs from the reference"
        assert (edge == g[f"{prefix}p{p}_edge"]).all()
        assert (sizes == g[f"{prefix}p{p}_size"]).all()
        assert (pts == g[f"cloud{p}"][g[f"{prefix}p{p}_idx"]]).all()
        assert [grid.n_leaves(p), grid.n_points(p), grid.n_nodes(p)] == g[f"{prefix}p{p}_counts"].tolist()


def test_two_subdivides_with_a_pose_in_between_match_the_reference_fixture():
    g = golden("resubdivide_deepen_edge4")
    late = [int(p) for p in g["late"]]
    poses = [int(p) for p in g["poses"]]
    early = [p for p in poses if p not in late]
    grid = _grid(int(g["edge"]))
    for p in early:
        grid.insert_points(p, g[f"cloud{p}"])
    grid.subdivide([MaxPoints(int(g["first_max"]))])
    _check_against_golden(grid, g, early, "s1_")
    for p in late:
        grid.insert_points(p, g[f"cloud{p}"])
    _check_against_golden(grid, g, poses, "s2_")
    grid.subdivide([MaxPoints(int(g["second_max"]))])
    _check_against_golden(grid, g, poses, "")


def _oracle_leaf_table(og, pose, non_empty=True):
    leaves = og.get_leaf_points(pose, non_empty)
    corner = np.array([np.asarray(l.corner, dtype=np.float64) for l in leaves]).reshape(-1, 3)
    return corner, np.array([float(l.edge) for l in leaves]), np.array([len(l.idx) for l in leaves], dtype=np.int64)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_three_subdivides_and_poses_arriving_at_different_times_match_the_oracle(seed):
    """call 1 -> pose 2 -> call 2 -> pose 3 -> call 3, every call finer than the one before; also the `non_empty=False`
    enumeration.  The fake forest's oracle keeps the reference's leaf lists, the host reorders the native tables."""
    rng = np.random.default_rng(seed)
    centers = rng.random((8, 3)) * 8
    clouds = {}
    for p in range(4):
        pts = centers[rng.integers(0, 8, 700)] + rng.normal(0, 0.3, (700, 3))
        clouds[p] = np.clip(pts, 0.01, 7.99).astype(np.float32).astype(np.float64)
    grid = _grid(4)
    og = grid._host._forest.og
    for p in (0, 1):
        grid.insert_points(p, clouds[p])
    grid.subdivide([MaxPoints(80)])
    grid.insert_points(2, clouds[2])
    grid.subdivide([MaxPoints(30)])
    grid.insert_points(3, clouds[3])
    grid.subdivide([MaxPoints(10)])
    moved = 0
    for p in range(4):
        for non_empty in (True, False):
            corner, edge, sizes, _ = _leaf_table(grid, p, non_empty)
            w_corner, w_edge, w_sizes = _oracle_leaf_table(og, p, non_empty)
            assert corner.shape == w_corner.shape and (corner == w_corner).all(), (p, non_empty)
            assert (edge == w_edge).all() and (sizes == w_sizes).all()
        # the one-call order of the same leaves (the leaf table's own order), for reference: it must differ for a pose
        t = grid._host._forest.export_blocks()
        lf = t["leaf"][t["pose"] == p]
        moved += int((np.sort(lf) != lf).any())
    assert moved > 0


def test_counterexample_single_pose():
    """The smallest case in which the history matters: c3 splits in call 1, c1 and c3.2 in call 2."""
    rng = np.random.default_rng(0)

    def blob(lo, hi, n):
        return rng.uniform(lo, hi, (n, 3))

    c3, c1 = np.array([0, 4, 4.0]), np.array([0, 0, 4.0])
    cloud = np.vstack([blob(c3 + np.array([0, 2, 0.0]), c3 + np.array([2, 4, 2.0]), 14), blob(c3, c3 + 4, 16), blob(c1, c1 + 4, 18)])
    cloud = cloud.astype(np.float32).astype(np.float64)
    grid = _grid(8)
    grid.insert_points(0, cloud)
    grid.subdivide([MaxPoints(20)])
    grid.subdivide([MaxPoints(10)])
    for non_empty in (True, False):
        corner, edge, sizes, _ = _leaf_table(grid, 0, non_empty)
        w_corner, w_edge, w_sizes = _oracle_leaf_table(grid._host._forest.og, 0, non_empty)
        assert (corner == w_corner).all() and (edge == w_edge).all() and (sizes == w_sizes).all()


def test_same_criterion_twice_keeps_the_one_call_order():
    rng = np.random.default_rng(5)
    cloud = (rng.random((2000, 3)) * 8).astype(np.float32).astype(np.float64)
    grid = _grid(4)
    grid.insert_points(0, cloud)
    grid.subdivide([MaxPoints(25)])
    before = _leaf_table(grid, 0)
    grid.subdivide([MaxPoints(25)])
    after = _leaf_table(grid, 0)
    assert all((a == b).all() for a, b in zip(before, after))


@pytest.mark.parametrize("seed", range(8))
def test_random_call_sequences_match_the_oracle(seed):
    """poses arriving at random times between subdivide calls of decreasing capacity, several cells"""
    rng = np.random.default_rng(100 + seed)
    edge = int(rng.choice([2, 4]))
    grid = _grid(edge)
    og = grid._host._forest.og
    centers = rng.random((6, 3)) * 8
    n_poses, capacity = 0, int(rng.integers(40, 120))
    for _ in range(int(rng.integers(4, 8))):
        if n_poses == 0 or rng.random() < 0.5:
            n = int(rng.integers(100, 500))
            pts = centers[rng.integers(0, 6, n)] + rng.normal(0, 0.4, (n, 3)) + (rng.random(3) * 6 if rng.random() < 0.3 else 0)
            grid.insert_points(n_poses, pts.astype(np.float32).astype(np.float64))
            n_poses += 1
        else:
            grid.subdivide([MaxPoints(capacity)])
            capacity = max(3, int(capacity * rng.uniform(0.3, 0.9)))
    for p in range(n_poses):
        for non_empty in (True, False):
            corner, edge_l, sizes, _ = _leaf_table(grid, p, non_empty)
            w_corner, w_edge, w_sizes = _oracle_leaf_table(og, p, non_empty)
            assert corner.shape == w_corner.shape and (corner == w_corner).all(), (p, non_empty)
            assert (edge_l == w_edge).all() and (sizes == w_sizes).all()


def test_apply_mask_after_a_second_subdivide_hits_the_right_leaves():
    """a mask built against `get_leaf_points` (octree.py:265-274) after a second, finer subdivide must be
    applied in that same (history) order - the block order the forest uses for `apply_pose_mask`."""
    from oracle.structure import OracleGrid, max_points_criterion

    rng = np.random.default_rng(9)
    centers = rng.random((5, 3)) * 8
    clouds = {p: np.clip(centers[rng.integers(0, 5, 500)] + rng.normal(0, 0.35, (500, 3)), 0.01, 7.99).astype(np.float32).astype(np.float64)
              for p in range(2)}
    grid, og = _grid(8), OracleGrid(8)   # an independent oracle holds the expectation
    for g in (grid, og):
        g.insert_points(0, clouds[0])
    grid.subdivide([MaxPoints(120)])
    og.subdivide([max_points_criterion(120)])
    for g in (grid, og):
        g.insert_points(1, clouds[1])
    grid.subdivide([MaxPoints(15)])
    og.subdivide([max_points_criterion(15)])
    for p in (0, 1):
        vox = grid.get_leaf_points(p)
        want = og.get_leaf_points(p)
        assert [v.n_points for v in vox] == [len(l.idx) for l in want]
        mask = rng.random(sum(v.n_points for v in vox)) < 0.6
        grid._host.forest.apply_pose_mask(grid._host.pose_index[p], mask)  # what OctreeManager / Octree.apply_mask call
        og.apply_mask(p, mask)
        got = np.vstack([np.empty((0, 3))] + [v.get_points() for v in grid.get_leaf_points(p)])
        exp = np.vstack([np.empty((0, 3))] + [l.points for l in og.get_leaf_points(p)])
        assert got.shape == exp.shape and (got == exp).all()
