"""Size distribution of the fitted (pose, leaf) blocks and of the winning hypothesis index on the bench workload
(design aid for the RANSAC kernel's sub-warp grouping; not the benchmark)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from octreelib_b200.criteria import MaxPoints
from octreelib_b200.grid import Grid, GridConfig

name = sys.argv[1] if len(sys.argv) > 1 else "c4_street_100M"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
dev = torch.device("cuda", 0)
w = bench.WORKLOADS[name]
clouds, numbers, P, total = bench.make_workload(name, 0, 1, dev, scale)
grid = Grid(GridConfig(voxel_edge_length=w["edge"]))
for n, c in zip(numbers, clouds):
    grid.insert_points(n, c)
grid.subdivide([MaxPoints(w["max_points"])])
f = grid._host.forest
np.random.seed(0)
from octreelib_b200.ransac import CudaRansac
r = CudaRansac(threshold=w["threshold"], hypotheses_number=1024, initial_points_number=6)
f.ransac(r.random_hypotheses, w["threshold"], list(range(P)), 10, apply=False)
res = f.export_ransac(scored_only=True)
sz, best, cnt = res["size"], res["best"], res["best_count"]
print("fitted blocks", len(sz), "mean size", sz.mean())
h = np.bincount(np.minimum(sz, 64))
cum = np.cumsum(h) / len(sz)
for n in (6, 7, 8, 10, 12, 16, 24, 32, 48, 63):
    print(f"  size <= {n}: {cum[min(n, len(cum) - 1)]:.4f}")
full = cnt == sz
print("share with a hypothesis keeping every point:", full.mean())
hb = np.bincount(np.minimum(best[full], 64))
cb = np.cumsum(hb) / full.sum()
for t in (0, 1, 2, 3, 7, 15, 31, 63):
    print(f"  winning index <= {t}: {cb[min(t, len(cb)-1)]:.4f}")
