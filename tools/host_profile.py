"""Where does the host time of one step go?  cProfile of a full step (debug aid, not the benchmark).

    python tools/host_profile.py [workload] [scale]
"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

workload = sys.argv[1] if len(sys.argv) > 1 else "c4_street_100M"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
dev = torch.device("cuda", 0)
w = bench.WORKLOADS[workload]
clouds, numbers, P, total = bench.make_workload(workload, 0, 1, dev, scale)
for _ in range(3):
    bench.run_step(clouds, numbers, P, w, 1)
torch.cuda.synchronize()
t0 = time.perf_counter()
bench.run_step(clouds, numbers, P, w, 1)
torch.cuda.synchronize()
print(f"plain step: {(time.perf_counter() - t0) * 1e3:.1f} ms for {total} points")
pr = cProfile.Profile()
pr.enable()
bench.run_step(clouds, numbers, P, w, 1)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
print(s.getvalue())
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25)
print(s.getvalue())
