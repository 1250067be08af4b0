"""GPU timeline of one END-TO-END step (pinned host inputs, result tables read back): tools/timeline.py for the e2e leg.

    python tools/timeline_e2e.py [--scale S]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--out", default="gpurun_out/timeline_e2e.json")
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
w = bench.WORKLOADS["c4_street_100M"]
clouds, numbers, P, total = bench.make_workload("c4_street_100M", 0, 1, dev, args.scale)
host = []
for c in clouds:
    h = torch.empty(c.shape, dtype=torch.float64, pin_memory=True)
    h.copy_(c)
    host.append(h.numpy())
del clouds
torch.cuda.synchronize()
for _ in range(3):
    bench.run_step(host, numbers, P, w, 1, read_tables=True)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    res = bench.run_step(host, numbers, P, w, 1, read_tables=True)
    del res
    torch.cuda.synchronize()
prof.export_chrome_trace(args.out)
ev = json.load(open(args.out))["traceEvents"]
os.remove(args.out)
gpu = sorted((e for e in ev if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")), key=lambda e: e["ts"])
t0 = gpu[0]["ts"]
h2d = [e for e in gpu if "HtoD" in e["name"] and e["dur"] > 20]
d2h = [e for e in gpu if "DtoH" in e["name"] and e["dur"] > 20]
kern = [e for e in gpu if e.get("cat") == "kernel"]
print(f"[e2e timeline] span {(max(e['ts'] + e['dur'] for e in gpu) - t0) / 1e3:.2f} ms; large H2D copies: {len(h2d)}, "
      f"{sum(e['dur'] for e in h2d) / 1e3:.2f} ms busy, first at {(h2d[0]['ts'] - t0) / 1e3:.2f} ms, last ends at "
      f"{(h2d[-1]['ts'] + h2d[-1]['dur'] - t0) / 1e3:.2f} ms; kernels {sum(e['dur'] for e in kern) / 1e3:.2f} ms, first kernel after the upload at "
      f"{(min(e['ts'] for e in kern if e['ts'] > h2d[-1]['ts']) - t0) / 1e3:.2f} ms; large D2H copies: {len(d2h)}, "
      f"{sum(e['dur'] for e in d2h) / 1e3:.2f} ms busy, from {(d2h[0]['ts'] - t0) / 1e3:.2f} to {(d2h[-1]['ts'] + d2h[-1]['dur'] - t0) / 1e3:.2f} ms")
bytes_h2d = sum(a.nbytes for a in host)
print(f"[e2e timeline] upload {bytes_h2d / 1e9:.2f} GB at {bytes_h2d / 1e9 / (sum(e['dur'] for e in h2d) / 1e6):.1f} GB/s while busy, "
      f"{bytes_h2d / 1e9 / ((h2d[-1]['ts'] + h2d[-1]['dur'] - h2d[0]['ts']) / 1e6):.1f} GB/s over its span")
for e in d2h:
    print(f"    D2H {(e['ts'] - t0) / 1e3:8.2f} ms  {e['dur'] / 1e3:6.2f} ms")
