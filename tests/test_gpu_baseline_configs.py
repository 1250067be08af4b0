"""GPU: the BASELINE.json configurations at their EXACT sizes against the CPU oracle: structure (cells,
leaf corners / edges / order, point order, counters) bit for bit, then `map_leaf_points_cuda_ransac` with the reference's
default H = 1024, K = 6 - identical surviving points per leaf.

  C1  one synthetic cloud of 100 k points, edge 1.0, <= 100 points / leaf
  C2  64-beam LiDAR, 120 k points x 10 poses in one grid, synchronized subdivision
  C3  planar indoor scene, 0.5 m cells - a 1 M-point sample of the 10 M configuration (the oracle needs minutes for
      10 M; the 10 M invariants are in test_gpu_sweep.py); fractional edge -> exactly rescaled oracle
"""
import numpy as np
import pytest

from gpu_util import compare_grid_with_oracle
from octreelib_b200.criteria import MaxPoints
from octreelib_b200.grid import Grid, GridConfig
from octreelib_b200.synthetic import indoor_scene, lidar64_scan
from oracle import ransac as oransac
from oracle.structure import OracleGrid, max_points_criterion
from test_gpu_sweep import _compare_scaled

pytestmark = pytest.mark.gpu


def _ransac_both(grid, og, threshold, o_threshold, seed):
    np.random.seed(seed)
    table = oransac.make_table(1024, 6)
    np.random.seed(seed)
    grid.map_leaf_points_cuda_ransac(poses_per_batch=10, threshold=threshold, hypotheses_number=1024, initial_points_number=6)
    og.map_leaf_points_ransac(table, threshold=o_threshold, poses_per_batch=10,
                              evaluate=lambda pts, bs, tab, th: oransac.ransac_evaluate(pts, bs, tab, th, threads=16))
    assert grid._host.forest.stats()["sample_oob_seen"] == 0


def test_config1_100k_single_cloud():
    clouds = {0: lidar64_scan(0, seed=0)[:100_000]}
    assert len(clouds[0]) == 100_000
    grid, og = Grid(GridConfig(voxel_edge_length=1.0)), OracleGrid(1.0)
    grid.insert_points(0, clouds[0])
    og.insert_points(0, clouds[0])
    grid.subdivide([lambda points: len(points) > 100])
    og.subdivide([max_points_criterion(100)])
    compare_grid_with_oracle(grid, og, clouds)
    _ransac_both(grid, og, 0.02, 0.02, seed=1)
    compare_grid_with_oracle(grid, og, clouds)


def test_config2_10_poses_x_120k_synchronized():
    clouds = {p: lidar64_scan(p, seed=0) for p in range(10)}
    assert all(len(c) == 120_000 for c in clouds.values())
    grid, og = Grid(GridConfig(voxel_edge_length=1.0)), OracleGrid(1.0)
    for p, c in clouds.items():
        grid.insert_points(p, c)
        og.insert_points(p, c)
    grid.subdivide([MaxPoints(100)])
    og.subdivide([max_points_criterion(100)])
    compare_grid_with_oracle(grid, og, clouds)
    _ransac_both(grid, og, 0.02, 0.02, seed=2)
    compare_grid_with_oracle(grid, og, clouds)


def test_config3_indoor_1M_sample_half_metre_cells():
    clouds = {0: indoor_scene(1_000_000, seed=0)}
    grid, og = Grid(GridConfig(voxel_edge_length=0.5)), OracleGrid(1)
    grid.insert_points(0, clouds[0])
    og.insert_points(0, clouds[0] * 2.0)  # exact power-of-two rescale: the reference cannot do fractional edges
    grid.subdivide([MaxPoints(100)])
    og.subdivide([max_points_criterion(100)])
    _compare_scaled(grid, og, clouds, 2.0)
    _ransac_both(grid, og, 0.01, 0.02, seed=3)
    _compare_scaled(grid, og, clouds, 2.0)
