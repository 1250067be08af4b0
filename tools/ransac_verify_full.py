"""Full-size check of the RANSAC FP32 pre-filter (csrc/ransac.cu): run the per-leaf RANSAC of a bench
workload three times - filtered (default), exact-only, verify - and require identical per-block tables
and masks.  Verify mode evaluates EVERY hypothesis exactly and checks that its inlier count lies in the
pre-filter's interval (the call raises AssertionError otherwise).  Prints the pre-filter statistics.

    python tools/ransac_verify_full.py [workload] [scale]
"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from octreelib_b200 import _native as N
from octreelib_b200.criteria import MaxPoints
from octreelib_b200.grid import Grid, GridConfig

workload = sys.argv[1] if len(sys.argv) > 1 else "c4_street_100M"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
dev = torch.device("cuda", 0)
w = bench.WORKLOADS[workload]
clouds, numbers, P, total = bench.make_workload(workload, 0, 1, dev, scale)
lib = N.lib()
grid = Grid(GridConfig(voxel_edge_length=w["edge"]))
for n_, c in zip(numbers, clouds):
    grid.insert_points(n_, c)
grid.subdivide([MaxPoints(w["max_points"])])
f = grid._host.forest
np.random.seed(0)
table = np.random.random((bench.H, bench.K))
rank = [int(x) for x in grid._host.pose_numbers]
out = (C.c_uint64 * 16)()
results = {}
for name, flags in (("filtered", N.RANSAC_FLAG_STATS), ("exact_only", N.RANSAC_FLAG_EXACT_ONLY), ("verify", N.RANSAC_FLAG_VERIFY)):
    N.check(lib.ol_ransac_stats_read(C.byref(out), 1))
    f.profile(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    f.ransac(table, w["threshold"], rank, 10, apply=False, flags=flags)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    prof = f.profile_read()
    N.check(lib.ol_ransac_stats_read(C.byref(out), 1))
    res = f.export_ransac(scored_only=True)
    pts = f.export_points(-1, order=0, pose_rank=rank, want_mask=True)
    results[name] = (res, pts["mask"].copy())
    st = list(out)
    print(f"{name:10s}: wall {dt * 1e3:8.1f} ms, ransac_kernel {prof['ransac_kernel'][1]:8.2f} ms, blocks scored {len(res['best'])}, "
          f"stats blocks={st[0]} filtered={st[1]} trivial={st[2]} exact={st[3]} early_exit={st[4]} violations={st[5]}", flush=True)
ref, ref_mask = results["exact_only"]
for name in ("filtered", "verify"):
    res, mask = results[name]
    for k in ("pose", "leaf", "size", "best", "best_count"):
        assert (res[k] == ref[k]).all(), (name, k)
    assert (res["plane"].view(np.uint32) == ref["plane"].view(np.uint32)).all(), name
    assert (mask == ref_mask).all(), name
print(f"OK: {workload} x{scale}: {total} points, {len(ref['best'])} fitted blocks, filtered == exact_only == verify "
      f"(best, best_count, plane bits, masks); mean block size {ref['size'].mean():.2f}, "
      f"share of blocks with best_count == size {(ref['best_count'] == ref['size']).mean():.3f}")
