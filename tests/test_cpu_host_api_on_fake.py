"""The host side of the public API (`Grid` -> `ForestHost` -> `_views`) on the CPU: the native forest is replaced by
`fake_forest.FakeForest` (oracle-backed tables in the native format and order), the expected values are the golden
fixtures generated from the REAL reference.  The same scenarios run against the CUDA forest in test_gpu_structure.py."""
import numpy as np
import pytest

from conftest import golden
from fake_forest import FakeForest
from octreelib_b200.grid import Grid, GridConfig

STRUCTURE_CASES = ["ref_test_grid_gt2", "ref_test_grid_gt3", "random_3pose_edge2", "random_3pose_edge2_filter",
                   "clustered_2pose_edge4", "offset_poses_edge1", "subset_subdivide_edge2", "far_offset_edge1"]


def _grid(edge):
    edge = int(edge) if float(edge) == int(edge) else float(edge)
    grid = Grid(GridConfig(voxel_edge_length=edge))
    grid._host._forest = FakeForest(edge)
    return grid


def _check(grid, g, poses, prefix=""):
    for p in poses:
        vox = grid.get_leaf_points(p)
        assert (np.array([np.asarray(v.corner_min, dtype=np.float64) for v in vox]).reshape(-1, 3) == g[f"{prefix}p{p}_corner"]).all()
        assert (np.array([float(v.edge_length) for v in vox]) == g[f"{prefix}p{p}_edge"]).all()
        assert (np.array([v.n_points for v in vox], dtype=np.int64) == g[f"{prefix}p{p}_size"]).all()
        pts = np.vstack([np.empty((0, 3))] + [v.get_points() for v in vox])
        assert (pts == g[f"cloud{p}"][g[f"{prefix}p{p}_idx"]]).all()
        assert [grid.n_leaves(p), grid.n_points(p), grid.n_nodes(p)] == g[f"{prefix}p{p}_counts"].tolist()
        assert (grid.get_points(p) == g[f"cloud{p}"][g[f"{prefix}p{p}_getpoints_idx"]]).all()


@pytest.mark.parametrize("name", STRUCTURE_CASES)
def test_public_api_matches_reference_golden(name):
    g = golden(name)
    poses = [int(p) for p in g["poses"]]
    grid = _grid(float(g["edge"]))
    for p in poses:
        grid.insert_points(p, g[f"cloud{p}"])
    _check(grid, g, poses, "pre_")
    sub = [int(x) for x in g["subdivide_poses"]] or None
    grid.subdivide([lambda pts, n=int(g["max_points"]): len(pts) > n], sub)
    if int(g["filter_min"]) >= 0:
        grid.filter([lambda pts, n=int(g["filter_min"]): len(pts) >= n])
    _check(grid, g, poses)
    # root leaves of integer-edged grids carry int64 corners like the reference's (grid.py:96-105)
    with pytest.raises(ValueError):
        grid.insert_points(poses[0], g[f"cloud{poses[0]}"])
    with pytest.raises(KeyError):
        grid.get_leaf_points(max(poses) + 17)


def test_late_poses_follow_the_scheme_through_the_public_api():
    g = golden("late_poses_edge2")
    late = [int(p) for p in g["late"]]
    poses = [int(p) for p in g["poses"]]
    grid = _grid(float(g["edge"]))
    for p in poses:
        if p not in late:
            grid.insert_points(p, g[f"cloud{p}"])
    grid.subdivide([lambda pts, n=int(g["max_points"]): len(pts) > n])
    for p in late:
        grid.insert_points(p, g[f"cloud{p}"])
    _check(grid, g, poses)


def test_map_leaf_points_with_a_selection_callback():
    """test/grid/test_grid.py:96-103 style: keep the first point of every leaf"""
    g = golden("random_3pose_edge2")
    poses = [int(p) for p in g["poses"]]
    grid = _grid(float(g["edge"]))
    for p in poses:
        grid.insert_points(p, g[f"cloud{p}"])
    grid.subdivide([lambda pts: len(pts) > int(g["max_points"])])
    before = {p: [v.get_points() for v in grid.get_leaf_points(p)] for p in poses}
    grid.map_leaf_points(lambda cloud: [cloud[0]])
    for p in poses:
        after = grid.get_leaf_points(p)
        assert len(after) == len(before[p])
        for v, b in zip(after, before[p]):
            assert v.n_points == 1 and (v.get_points()[0] == b[0]).all()
        assert grid.n_points(p) == len(before[p])
    with pytest.raises(NotImplementedError):
        grid.map_leaf_points(lambda cloud: cloud + 1.0)


def test_all_leaves_including_empty_ones():
    """`get_leaf_points(pose, non_empty=False)` (grid.py:217-232) before and after a filter that empties leaves."""
    g = golden("all_leaves_edge2")
    grid = _grid(float(g["edge"]))
    for p in (0, 1):
        grid.insert_points(p, g[f"cloud{p}"])
    grid.subdivide([lambda pts, n=int(g["max_points"]): len(pts) > n])
    for stage in ("a", "b"):
        if stage == "b":
            grid.filter([lambda pts, n=int(g["filter_min"]): len(pts) >= n])
        for p in (0, 1):
            vox = grid.get_leaf_points(p, non_empty=False)
            assert (np.array([np.asarray(v.corner_min, dtype=np.float64) for v in vox]).reshape(-1, 3) == g[f"{stage}_p{p}_corner"]).all()
            assert (np.array([float(v.edge_length) for v in vox]) == g[f"{stage}_p{p}_edge"]).all()
            assert (np.array([v.n_points for v in vox], dtype=np.int64) == g[f"{stage}_p{p}_size"]).all()
