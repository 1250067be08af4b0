import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from conftest import golden
from octreelib_b200.grid import Grid, GridConfig
g = golden("random_3pose_edge2_filter")
poses = [int(p) for p in g["poses"]]
grid = Grid(GridConfig(voxel_edge_length=2))
for p in poses:
    grid.insert_points(p, g[f"cloud{p}"])
grid.subdivide([lambda pts: len(pts) > 25])
f = grid._host.forest
b0 = f.export_blocks(); l0 = f.export_leaves()
print("before filter stats", f.stats())
grid.filter([lambda pts: len(pts) >= 4])
print("after filter stats", f.stats())
b1 = f.export_blocks(); l1 = f.export_leaves()
print("leaves equal", (l0["corner"] == l1["corner"]).all(), (l0["edge"] == l1["edge"]).all())
for p in poses:
    s0 = b0["pose"] == p; s1 = b1["pose"] == p
    exp0 = b0["leaf"][s0][b0["size"][s0] >= 4]
    print("pose", p, "blocks before", s0.sum(), "expected after", len(exp0), "got", s1.sum(), "golden", len(g[f"p{p}_size"]))
    print("  leaf ids equal:", np.array_equal(exp0, b1["leaf"][s1]), " sizes equal:", np.array_equal(b0["size"][s0][b0["size"][s0] >= 4], b1["size"][s1]))
    print("  golden sizes eq:", np.array_equal(g[f"p{p}_size"], b1["size"][s1]))
    gc = g[f"p{p}_corner"]; mc = l1["corner"][b1["leaf"][s1]]
    bad = np.flatnonzero((gc != mc).any(axis=1))
    print("  first bad rows", bad[:5], gc[bad[:3]], mc[bad[:3]], g[f"p{p}_edge"][bad[:3]], l1["edge"][b1["leaf"][s1]][bad[:3]])
