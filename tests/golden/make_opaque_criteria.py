"""Golden fixture for coordinate-dependent (opaque) criteria from the REAL reference (build container only):
    python tests/golden/make_opaque_criteria.py
Two poses; `subdivide` with an extent criterion (octree_manager.py:36-66 -> octree.py:20-32), then `filter` with a spread
criterion (octree.py:102-112).  Recorded after each step: the leaves of every pose in the reference's order + counters."""
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
from octreelib.grid import Grid, GridConfig  # noqa: E402  (the REFERENCE)

from make_map_leaf_points import dump  # noqa: E402
from make_subdivide_as import stable_order  # noqa: E402


def extent(points):
    return len(points) > 10 and np.ptp(points, axis=0).max() > 1.0


def spread(points):
    return len(points) >= 3 and points.std(axis=0).max() > 0.2


def main():
    rng = np.random.default_rng(4242)
    clouds = [np.vstack([rng.normal([5, 6, 2], 1.0, (700, 3)), rng.normal([9, 2, 6], 0.3, (300, 3)), rng.uniform(0, 12, (500, 3))])
              for _ in range(2)]
    grid = Grid(GridConfig(voxel_edge_length=4.0))
    for p, c in enumerate(clouds):
        grid.insert_points(p, c)
    grid.subdivide([extent])
    step1 = dump(grid, [0, 1])
    grid.filter([spread])
    step2 = dump(grid, [0, 1])
    data = dict(cloud0=clouds[0], cloud1=clouds[1], edge=4.0)
    data.update({f"s1_{k}": v for k, v in step1.items()})
    data.update({f"s2_{k}": v for k, v in step2.items()})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "opaque_criteria_edge4.npz"), **data)
    print("opaque criteria fixture:", {k: v.tolist() for k, v in step1.items() if k.startswith("counts")},
          {k: v.tolist() for k, v in step2.items() if k.startswith("counts")})


if __name__ == "__main__":
    with stable_order():
        main()
