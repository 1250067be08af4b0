"""Host wall-clock per phase of one step (debug aid, not the benchmark)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from octreelib_b200.criteria import MaxPoints
from octreelib_b200.grid import Grid, GridConfig

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
host = len(sys.argv) > 2 and sys.argv[2] == "host"
dev = torch.device("cuda", 0)
w = bench.WORKLOADS["c4_street_100M"]
clouds, numbers, P, total = bench.make_workload("c4_street_100M", 0, 1, dev, scale)
if host:
    hc = []
    for c in clouds:
        h = torch.empty(c.shape, dtype=torch.float64, pin_memory=True); h.copy_(c); hc.append(h.numpy())
    clouds = hc
def sync(): torch.cuda.synchronize()
for it in range(3):
    np.random.seed(0)
    sync(); t0 = time.perf_counter()
    grid = Grid(GridConfig(voxel_edge_length=1.0))
    f = grid._host.forest
    f.profile(True)
    for n, c in zip(numbers, clouds):
        grid.insert_points(n, c)
    sync(); t1 = time.perf_counter()
    grid.subdivide([MaxPoints(100)])
    sync(); t2 = time.perf_counter()
    st_mid = f.stats()
    sync(); t2b = time.perf_counter()
    grid.map_leaf_points_cuda_ransac(poses_per_batch=10, threshold=0.02, hypotheses_number=1024)
    sync(); t3 = time.perf_counter()
    r = f.export_ransac(); l = f.export_leaves()
    sync(); t4 = time.perf_counter()
    st = f.stats()
    sync(); t5 = time.perf_counter()
    prof = f.profile_read()
    print(f"iter {it}: insert {1e3*(t1-t0):.1f} ms, subdivide {1e3*(t2-t1):.1f}, stats(blocks) {1e3*(t2b-t2):.1f}, ransac {1e3*(t3-t2b):.1f}, "
          f"exports {1e3*(t4-t3):.1f}, stats {1e3*(t5-t4):.1f}, total {1e3*(t5-t0):.1f} ms; points {total}, poses {P}")
    print("   blocks before ransac", st_mid["n_blocks"], "max_block", st_mid["max_block_size"], "leaves", st_mid["n_leaves"], "cells", st_mid["n_cells"],
          "depth", st_mid["max_depth_reached"], "alive after", st["n_points_alive"], "ransac rows", len(r["best"]), "scored", int((r["best"]>=0).sum()),
          "peak GB", st["device_bytes_peak"]/1e9)
    print("   ", {k: (v[0], round(v[1],3)) for k, v in prof.items()})
    del grid, f
