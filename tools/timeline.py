"""GPU timeline of one step (debug aid, not the benchmark): every kernel / memcpy / memset of the step with its start
and duration (CUPTI through torch.profiler - nsys is not installed), the idle gaps of the GPU between them and the CUDA
runtime calls the host made inside each gap.  Works single-GPU and under torchrun.

    python tools/timeline.py [--scale S] [--steps K] [--out gpurun_out/timeline.json]
    python -m torch.distributed.run --nproc-per-node 2 ... tools/timeline.py --scale 0.25

`--scale 0.125` on one GPU reproduces the per-rank problem size of the 8-GPU run without the exchange.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--gap-us", type=float, default=8.0)
    ap.add_argument("--out", default="gpurun_out/timeline.json")
    ap.add_argument("--workload", default="c4_street_100M")
    ap.add_argument("--list", action="store_true", help="print every GPU activity in order (start, duration, gap before it)")
    args = ap.parse_args()
    rank, world, local = bench.dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = bench.init_dist(world, local)
    w = bench.WORKLOADS[args.workload]
    clouds, numbers, P, total = bench.make_workload(args.workload, rank, world, dev, args.scale)
    for _ in range(args.warmup):
        bench.run_step(clouds, numbers, P, w, world)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            res = bench.run_step(clouds, numbers, P, w, world)
            del res
        torch.cuda.synchronize()
    if rank != 0:
        if world > 1:
            dist.barrier()
        return
    path = args.out
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    gpu = sorted((e for e in ev if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")), key=lambda e: e["ts"])
    rt = sorted((e for e in ev if e.get("ph") == "X" and e.get("cat") in ("cuda_runtime", "cuda_driver")), key=lambda e: e["ts"])
    if not gpu:
        print("no GPU activity captured")
        return
    t0, t1 = gpu[0]["ts"], max(e["ts"] + e["dur"] for e in gpu)
    busy, end, gaps = 0.0, gpu[0]["ts"], []
    for i, e in enumerate(gpu):
        if e["ts"] > end:
            if e["ts"] - end >= args.gap_us:
                gaps.append((e["ts"] - end, end, e["ts"], gpu[i - 1]["name"][:60], e["name"][:60]))
            busy += 0.0
        s = max(e["ts"], end)
        if e["ts"] + e["dur"] > s:
            busy += e["ts"] + e["dur"] - s
            end = e["ts"] + e["dur"]
    span = t1 - t0
    print(f"[timeline] rank 0 of {world}, scale {args.scale}: {len(gpu)} GPU activities over {args.steps} step(s), span {span / 1e3:.3f} ms, "
          f"busy {busy / 1e3:.3f} ms ({100 * busy / span:.1f} %), {len(gaps)} gaps >= {args.gap_us} us totalling "
          f"{sum(g[0] for g in gaps) / 1e3:.3f} ms")
    if args.list:
        print("[timeline] sequence (start ms | dur us | gap before us | name | host calls in the gap):")
        prev_end = gpu[0]["ts"]
        for e in gpu:
            gap = e["ts"] - prev_end
            calls = {}
            if gap >= 4.0:
                for r in rt:
                    if r["ts"] + r["dur"] >= prev_end and r["ts"] <= e["ts"]:
                        calls[r["name"]] = calls.get(r["name"], 0) + 1
            cs = ", ".join(f"{k.replace('cuda', '')} x{v}" for k, v in sorted(calls.items(), key=lambda kv: -kv[1])[:5])
            print(f"    {(e['ts'] - t0) / 1e3:7.3f} | {e['dur']:7.1f} | {gap:7.1f} | {e['name'].replace('ol::', '').replace('(anonymous namespace)::', '')[:44]:44s} | {cs}")
            prev_end = max(prev_end, e["ts"] + e["dur"])
    byname = {}
    for e in gpu:
        k = e["name"].split("<")[0].split("(")[0][:48]
        c = byname.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += e["dur"]
    print("[timeline] GPU time by kernel:")
    for k, (c, d) in sorted(byname.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"    {d / 1e3:8.3f} ms  x{c:<4d} {k}")
    print("[timeline] gaps (us | after -> before | host runtime calls inside):")
    for g, a, b, prev, nxt in sorted(gaps, key=lambda g: -g[0])[:40]:
        calls = {}
        for r in rt:
            if r["ts"] + r["dur"] >= a and r["ts"] <= b:
                calls[r["name"]] = calls.get(r["name"], 0) + 1
        cs = ", ".join(f"{k} x{v}" for k, v in sorted(calls.items(), key=lambda kv: -kv[1])[:6])
        print(f"    {g:8.1f} us @ {(a - t0) / 1e3:7.3f} ms | {prev} -> {nxt} | {cs}")
    if world > 1:
        dist.barrier()


if __name__ == "__main__":
    main()
