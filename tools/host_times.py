"""Host wall-clock of the public calls of one step, WITHOUT extra synchronisation (debug aid, not the benchmark):
how long the Python / ctypes / launch side of each call takes, i.e. what the GPU may have to wait for.

    python tools/host_times.py [scale]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from octreelib_b200.criteria import MaxPoints
from octreelib_b200.grid import Grid, GridConfig

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
dev = torch.device("cuda", 0)
w = bench.WORKLOADS["c4_street_100M"]
clouds, numbers, P, total = bench.make_workload("c4_street_100M", 0, 1, dev, scale)
grid = None
for it in range(6):
    np.random.seed(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    grid = None
    t0b = time.perf_counter()
    grid = Grid(GridConfig(voxel_edge_length=1.0))
    f = grid._host.forest
    t1 = time.perf_counter()
    for n, c in zip(numbers, clouds):
        grid.insert_points(n, c)
    t2 = time.perf_counter()
    f._flush()
    t2b = time.perf_counter()
    grid.subdivide([MaxPoints(100)])
    t3 = time.perf_counter()
    grid.map_leaf_points_cuda_ransac(poses_per_batch=10, threshold=0.02, hypotheses_number=1024)
    t4 = time.perf_counter()
    st = f.stats(light=True)
    t5 = time.perf_counter()
    print(f"iter {it}: drop previous {1e3*(t0b-t0):.2f} ms, create {1e3*(t1-t0b):.2f}, {P} x insert_points {1e3*(t2-t1):.2f}, flush {1e3*(t2b-t2):.2f}, "
          f"subdivide {1e3*(t3-t2b):.2f}, ransac {1e3*(t4-t3):.2f}, stats {1e3*(t5-t4):.2f}, total {1e3*(t5-t0):.2f} ms", flush=True)
