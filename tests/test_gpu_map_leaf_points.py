"""Coordinate-changing `Grid.map_leaf_points` (grid/grid.py:111-122 -> octree/octree.py:114-123) pinned against
tests/golden/map_leaf_points_edge4.npz, recorded from the REAL reference by tests/golden/make_map_leaf_points.py."""
import numpy as np
import pytest

from conftest import golden
from octreelib_b200.grid import Grid, GridConfig

pytestmark = pytest.mark.gpu


def shrink(cloud):
    c = cloud.mean(axis=0)
    return c + 0.5 * (cloud - c)


def centroid(cloud):
    return cloud.mean(axis=0, keepdims=True)


def _check(grid, g, step):
    for p in (0, 1):
        leaves = grid.get_leaf_points(p)
        corner = np.array([np.asarray(v.corner_min, dtype=np.float64) for v in leaves]).reshape(-1, 3)
        assert (corner == g[f"{step}_corner{p}"]).all()
        assert (np.array([float(v.edge_length) for v in leaves]) == g[f"{step}_edge{p}"]).all()
        assert [len(v.get_points()) for v in leaves] == g[f"{step}_sizes{p}"].tolist()
        pts = np.vstack([np.empty((0, 3))] + [np.asarray(v.get_points()).reshape(-1, 3) for v in leaves])
        assert (pts == g[f"{step}_points{p}"]).all()      # bit for bit: same function, same rows, same order
        assert [grid.n_leaves(p), grid.n_points(p), grid.n_nodes(p)] == g[f"{step}_counts{p}"].tolist()


def test_coordinate_changing_map_matches_the_reference():
    g = golden("map_leaf_points_edge4")
    grid = Grid(GridConfig(voxel_edge_length=float(g["edge"])))
    grid.insert_points(0, g["cloud0"])
    grid.insert_points(1, g["cloud1"])
    grid.subdivide([lambda pts: len(pts) > int(g["max_points"])])
    grid.map_leaf_points(shrink)
    _check(grid, g, "s1")
    grid.map_leaf_points(centroid, [1])
    _check(grid, g, "s2")
    # the grid is still a working grid: RANSAC on top of the replaced points runs and only removes points
    before = [grid.n_points(p) for p in (0, 1)]
    np.random.seed(3)
    grid.map_leaf_points_cuda_ransac(poses_per_batch=2, threshold=0.05, hypotheses_number=64)
    assert all(grid.n_points(p) <= b for p, b in zip((0, 1), before))


def test_a_map_that_leaves_its_leaf_is_refused_before_anything_changes():
    g = golden("map_leaf_points_edge4")
    grid = Grid(GridConfig(voxel_edge_length=float(g["edge"])))
    grid.insert_points(0, g["cloud0"])
    grid.subdivide([lambda pts: len(pts) > int(g["max_points"])])
    before = grid.get_points(0).copy()
    with pytest.raises(NotImplementedError):
        grid.map_leaf_points(lambda cloud: cloud + 100.0)
    assert (grid.get_points(0) == before).all()
