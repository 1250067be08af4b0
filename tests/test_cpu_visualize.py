"""`Grid.visualize` (grid/grid.py:269-341) against what the REAL reference hands to k3d (tests/golden/visualize_edge4.npz,
recorded through a stub k3d module by tests/golden/make_golden.py): object order, colours drawn from `random.seed(seed)`
for both visualisation types, unused voxels, wire frames.  k3d itself is not installed in this image; the host code runs
on the CPU stand-in for the forest."""
import sys
import types

import numpy as np
import pytest

from conftest import golden
from fake_forest import FakeForest
from octreelib_b200.grid import Grid, GridConfig, GridVisualizationType, VisualizationConfig


class _Plot(list):
    def __iadd__(self, item):
        self.append(item)
        return self

    def get_snapshot(self):
        return f"<recorded {len(self)} objects>"


@pytest.fixture
def k3d_stub(monkeypatch):
    plots = []
    mod = types.ModuleType("k3d")

    def make_plot():
        plots.append(_Plot())
        return plots[-1]

    mod.Plot = make_plot
    mod.points = lambda positions, point_size, color: ("points", np.array(positions, dtype=np.float64).reshape(-1, 3),
                                                       float(point_size), int(color))
    mod.lines = lambda vertices, indices, width, color, indices_type: (
        "lines", np.array(vertices, dtype=np.float64).reshape(-1, 3), float(width), int(color), np.array(indices), indices_type)
    monkeypatch.setitem(sys.modules, "k3d", mod)
    return plots


@pytest.mark.parametrize("tag,vtype", [("pose", GridVisualizationType.POSE), ("voxel", GridVisualizationType.VOXEL),
                                       ("pose_unused", GridVisualizationType.POSE), ("voxel_unused", GridVisualizationType.VOXEL)])
def test_visualize_draws_what_the_reference_draws(tag, vtype, k3d_stub, tmp_path):
    g = golden("visualize_edge4")
    grid = Grid(GridConfig(voxel_edge_length=int(g["edge"])))
    grid._host._forest = FakeForest(int(g["edge"]))
    for p in range(3):
        grid.insert_points(p, g[f"cloud{p}"])
    grid.subdivide([lambda pts, n=int(g["max_points"]): len(pts) > n])
    unused = []
    if tag.endswith("unused"):
        leaves = grid.get_leaf_points(int(g["unused_from_pose"]))
        unused = [leaves[int(i)].id for i in g["unused_leaf_positions"]]
    cfg = VisualizationConfig(type=vtype, point_size=0.05, line_width_size=0.02, line_color=0x00FF00,
                              filepath=str(tmp_path / (tag + ".html")), seed=int(g[f"{tag}_seed"]), unused_voxels=unused)
    grid.visualize(cfg)
    plot = k3d_stub[-1]
    assert (tmp_path / (tag + ".html")).read_text().startswith("<recorded")
    kinds = np.array([0 if o[0] == "points" else 1 for o in plot])
    assert (kinds == g[f"{tag}_kinds"]).all()
    assert (np.array([len(o[1]) for o in plot]) == g[f"{tag}_sizes"]).all()
    assert (np.vstack([o[1] for o in plot]) == g[f"{tag}_data"]).all()
    assert (np.array([o[2] for o in plot]) == g[f"{tag}_scalar"]).all()
    assert (np.array([o[3] for o in plot]) == g[f"{tag}_color"]).all(), "colour sequence differs from the reference's"
    for o in plot:
        if o[0] == "lines":
            assert o[5] == "segment" and o[4].shape == (6, 8)
