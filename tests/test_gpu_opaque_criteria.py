"""Coordinate-dependent (opaque) subdivision / filtering criteria: evaluated on the host node by node, the scheme imposed on
the device.  Pinned against tests/golden/opaque_criteria_edge4.npz (REAL reference, tests/golden/make_opaque_criteria.py)."""
import numpy as np
import pytest

from conftest import golden
from octreelib_b200.grid import Grid, GridConfig
from test_gpu_map_leaf_points import _check

pytestmark = pytest.mark.gpu


def extent(points):
    return len(points) > 10 and np.ptp(points, axis=0).max() > 1.0


def spread(points):
    return len(points) >= 3 and points.std(axis=0).max() > 0.2


def test_opaque_criteria_match_the_reference():
    g = golden("opaque_criteria_edge4")
    grid = Grid(GridConfig(voxel_edge_length=float(g["edge"])))
    grid.insert_points(0, g["cloud0"])
    grid.insert_points(1, g["cloud1"])
    grid.subdivide([extent])
    _check(grid, g, "s1")
    grid.filter([spread])
    _check(grid, g, "s2")


def test_opaque_criterion_that_never_stops_raises_like_the_reference():
    grid = Grid(GridConfig(voxel_edge_length=4.0))
    grid.insert_points(0, np.random.default_rng(1).random((200, 3)) * 3.9)
    with pytest.raises(RecursionError):
        grid.subdivide([lambda points: len(points) > 0 and points[:, 0].mean() > -1.0])
