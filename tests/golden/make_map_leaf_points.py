"""Golden fixture for coordinate-changing `Grid.map_leaf_points` from the REAL reference (build container only):
    python tests/golden/make_map_leaf_points.py
Two poses, one subdivide, then (1) every leaf of both poses shrunk towards its centroid, (2) every leaf of pose 1 replaced by
its centroid (grid/grid.py:111-122 -> octree_manager.py:68-83 -> octree/octree.py:114-123).  Recorded after each step: the
leaves of every pose in the reference's order (corner, edge, points) and the counters."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np

if not hasattr(np, "float_"):
    np.float_ = np.float64
from oracle import ref_loader  # noqa: E402

ref_loader.load(cudasim=True)
from octreelib.grid import Grid, GridConfig  # noqa: E402  (the REFERENCE: oracle/_ref is first on sys.path)

from make_subdivide_as import stable_order  # noqa: E402


def shrink(cloud):
    c = cloud.mean(axis=0)
    return c + 0.5 * (cloud - c)


def centroid(cloud):
    return cloud.mean(axis=0, keepdims=True)


def dump(grid, poses):
    out = {}
    for p in poses:
        leaves = grid.get_leaf_points(p)
        out[f"corner{p}"] = np.array([np.asarray(v.corner_min, dtype=np.float64) for v in leaves]).reshape(-1, 3)
        out[f"edge{p}"] = np.array([float(v.edge_length) for v in leaves])
        out[f"sizes{p}"] = np.array([len(v.get_points()) for v in leaves])
        out[f"points{p}"] = np.vstack([np.empty((0, 3))] + [np.asarray(v.get_points(), dtype=np.float64).reshape(-1, 3) for v in leaves])
        out[f"counts{p}"] = np.array([grid.n_leaves(p), grid.n_points(p), grid.n_nodes(p)])
    return out


def main():
    rng = np.random.default_rng(77)
    clouds = [np.vstack([rng.normal([5, 6, 2], 1.2, (900, 3)), rng.uniform(0, 12, (600, 3))]) for _ in range(2)]
    grid = Grid(GridConfig(voxel_edge_length=4.0))
    for p, c in enumerate(clouds):
        grid.insert_points(p, c)
    grid.subdivide([lambda pts: len(pts) > 30])
    grid.map_leaf_points(shrink)
    step1 = dump(grid, [0, 1])
    grid.map_leaf_points(centroid, [1])
    step2 = dump(grid, [0, 1])
    data = dict(cloud0=clouds[0], cloud1=clouds[1], edge=4.0, max_points=30)
    data.update({f"s1_{k}": v for k, v in step1.items()})
    data.update({f"s2_{k}": v for k, v in step2.items()})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "map_leaf_points_edge4.npz"), **data)
    print("map_leaf_points fixture:", {k: v.tolist() for k, v in step2.items() if k.startswith("counts")})


if __name__ == "__main__":
    with stable_order():
        main()
