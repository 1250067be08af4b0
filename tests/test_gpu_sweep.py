"""GPU: BASELINE config 5 style sweep (cell size x max points per leaf) and full-size invariants.

Fractional cell sizes: the reference's int-truncated cell key (grid/grid.py:72-76) collapses cells for
voxel_edge_length < 1 (SURVEY hazard 2), so - as SURVEY 8(d) prescribes - the oracle is run on the exactly
rescaled problem `points * 2^k, edge 1.0, threshold * 2^k`; power-of-two scaling is exact in float64, so every
leaf corner / edge of the native grid times 2^k must equal the oracle's bit for bit, with identical point order.
"""
import numpy as np
import pytest

from octreelib_b200.criteria import MaxPoints
from octreelib_b200.grid import Grid, GridConfig
from octreelib_b200.synthetic import indoor_scene, lidar64_scan
from oracle import ransac as oransac
from oracle.structure import OracleGrid, max_points_criterion

pytestmark = pytest.mark.gpu


def _compare_scaled(grid, og, clouds, scale):
    host = grid._host
    forest = host.forest
    blocks = forest.export_blocks()
    leaves = forest.export_leaves()
    for p in clouds:
        pi = host.pose_index[p]
        want = og.get_leaf_points(p)
        sel = np.flatnonzero(blocks["pose"] == pi)
        assert len(sel) == len(want)
        lf = blocks["leaf"][sel]
        w_corner = np.array([np.asarray(l.corner, dtype=np.float64) for l in want]).reshape(-1, 3)
        w_edge = np.array([float(l.edge) for l in want])
        assert (leaves["corner"][lf] * scale == w_corner).all(), "leaf corners / order differ"
        assert (leaves["edge"][lf] * scale == w_edge).all()
        assert (blocks["size"][sel] == np.array([len(l.idx) for l in want])).all()
        got = forest.export_points(pi, order=0)
        assert (got["idx"] == np.concatenate([l.idx for l in want])).all(), "point order inside leaves differs"
        assert [grid.n_leaves(p), grid.n_points(p), grid.n_nodes(p)] == [og.n_leaves(p), og.n_points(p), og.n_nodes(p)]


@pytest.mark.parametrize("max_points", [50, 250, 1000])
@pytest.mark.parametrize("edge", [0.25, 0.5, 1.0, 2.0, 4.0])
def test_sweep_cell_size_and_leaf_capacity(edge, max_points):
    clouds = {0: indoor_scene(60000, seed=3)[:, :], 1: lidar64_scan(1, seed=2)[::4] * 0.25 + np.array([20.0, 20.0, 0.0])}
    clouds = {p: c.astype(np.float32).astype(np.float64) for p, c in clouds.items()}
    scale = 1.0 / edge if edge < 1.0 else 1.0
    o_edge = 1.0 if edge < 1.0 else edge
    grid, og = Grid(GridConfig(voxel_edge_length=edge)), OracleGrid(int(o_edge) if o_edge >= 1 else o_edge)
    for p, c in clouds.items():
        grid.insert_points(p, c)
        og.insert_points(p, c * scale)
    grid.subdivide([MaxPoints(max_points)])
    og.subdivide([max_points_criterion(max_points)])
    _compare_scaled(grid, og, clouds, scale)
    # RANSAC on the same structure: threshold scales with the coordinates
    thr = 0.01
    np.random.seed(11)
    table = oransac.make_table(128, 6)
    np.random.seed(11)
    grid.map_leaf_points_cuda_ransac(poses_per_batch=10, threshold=thr, hypotheses_number=128)
    og.map_leaf_points_ransac(table, threshold=thr * scale, poses_per_batch=10)
    # power-of-two scaling is exact through the whole float64 plane fit, the float32 rounding of the plane and the
    # distance test, so even the rescaled oracle must agree bit for bit on which points survive
    _compare_scaled(grid, og, clouds, scale)


def test_full_size_invariants_10M():
    """Properties that do not need the oracle, at BASELINE config 3's size (1e7 points, 0.5 m cells): every point
    is in exactly one leaf, inside that leaf's box, leaves are within capacity or at the depth cap, per-pose
    point order inside a block is increasing, cell keys are sorted, counters add up."""
    import torch

    from octreelib_b200 import synthetic

    dev = torch.device("cuda", 0)
    n = 10_000_000
    pts = synthetic.indoor_torch(n, seed=5, device=dev)
    half = n // 2
    grid = Grid(GridConfig(voxel_edge_length=0.5))
    grid.insert_points(0, pts[:half])
    grid.insert_points(1, pts[half:])
    grid.subdivide([MaxPoints(100)])
    f = grid._host.forest
    st = f.stats()
    assert st["n_points_alive"] == n and st["n_points_inserted"] == n
    cells = f.export_cells()
    q = cells["q"]
    order = np.lexsort((q[:, 2], q[:, 1], q[:, 0]))
    assert (order == np.arange(len(q))).all(), "cells are not in lexicographic order"
    assert (cells["leaf_begin"][1:] >= cells["leaf_begin"][:-1]).all() and cells["leaf_begin"][-1] == st["n_leaves"]
    leaves = f.export_leaves()
    blocks = f.export_blocks()
    assert int(blocks["size"].sum()) == n
    # union-of-poses leaf load <= capacity (the split criterion is evaluated on the union, octree_manager.py:50-61)
    load = np.bincount(blocks["leaf"], weights=blocks["size"], minlength=st["n_leaves"])
    assert load.max() <= 100
    for pose in (0, 1):
        out = f.export_points(pose, order=0)
        sel = blocks["pose"] == pose
        sizes = blocks["size"][sel]
        lf = np.repeat(blocks["leaf"][sel], sizes)
        assert len(out["idx"]) == sizes.sum() == (half if pose == 0 else n - half)
        lo = leaves["corner"][lf]
        hi = lo + leaves["edge"][lf][:, None]
        assert ((out["xyz"] >= lo) & (out["xyz"] < hi)).all(), "a point lies outside its leaf"
        starts = np.concatenate([[0], np.cumsum(sizes)[:-1]])
        d = np.diff(out["idx"])
        inner = np.ones(len(d), dtype=bool)
        inner[starts[1:] - 1] = False
        assert (d[inner] > 0).all(), "input order inside a block is not preserved"
        assert len(np.unique(out["idx"])) == len(out["idx"])
        assert grid.n_points(pose) == len(out["idx"]) and grid.n_leaves(pose) == int(sel.sum())
    total_nodes = cells["n_nodes"].sum()
    assert total_nodes == st["n_leaves"] + st["n_internal"]
