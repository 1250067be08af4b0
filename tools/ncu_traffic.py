#!/usr/bin/env python
"""Turns an `ncu --set full` capture of the bench workload into profiles/r02_ncu_traffic.json - the per-launch DRAM traffic
(dram__bytes_read.sum + dram__bytes_write.sum) of the HBM-bound kernels of the CURRENT build that bench.py puts into
`roofline.traffic`, plus a markdown summary of the captured kernels for profiles/.

    ncu --set full --clock-control none --import-source on -k regex:'os_pass_kernel|part_move_kernel|...' -o gpurun_out/x python bench.py ...
    ncu -i gpurun_out/x.ncu-rep --page raw --csv > gpurun_out/x_raw.csv        (here: no GPU needed)
    python tools/ncu_traffic.py gpurun_out/x_raw.csv profiles/r02_ncu_traffic.json profiles/r02_ncu_summary_v1.md [points]

`points` = the number of points of the captured workload (default 100 000 000: c4_street_100M); every kernel listed below
walks all of them once per launch, which is what bench.py scales the per-launch traffic by.
"""
import csv
import json
import sys

STAGE_OF = [("os_pass_kernel", "sort_main_pass"), ("part_move_kernel", "part_move"), ("part_hist_kernel", "part_hist"),
            ("keygen_kernel", "keygen"), ("os_hist_kernel", "sort_main_hist"), ("runs_fused_kernel", "runs"),
            ("runs_flags_kernel", "runs_flags"), ("runs_emit2_kernel", "runs_emit"),
            ("ransac_lane_kernel", "ransac_lane"), ("ransac_small_kernel", "ransac_small"), ("gather_kernel", "gather_morton"),
            ("gather_blocks_kernel", "gather_points"), ("refstart_assign_kernel", "refstart_assign"),
            ("refstart_count_kernel", "refstart_count"), ("block_arrange2_kernel", "block_arrange"),
            ("leaf_block_span_kernel", "leaf_block_span"), ("compact_move_kernel", "compact_move"),
            ("insert_batch_kernel", "insert_batch")]
COLS = {"dram_r": "dram__bytes_read.sum", "dram_w": "dram__bytes_write.sum", "time": "gpu__time_duration.sum",
        "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "issue": "sm__inst_issued.avg.pct_of_peak_sustained_active", "warps": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "fp64": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
        "grid": "launch__grid_size", "block": "launch__block_size"}


def num(s):
    try:
        return float(str(s).replace(",", ""))
    except ValueError:
        return None


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return v * scale


def to_ns(v, unit):
    return v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)


def main():
    raw, out_json, out_md = sys.argv[1], sys.argv[2], sys.argv[3]
    points = float(sys.argv[4]) if len(sys.argv) > 4 else 1e8
    rows = list(csv.reader(open(raw, newline="")))
    header = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[header], rows[header + 1]
    col = {n: i for i, n in enumerate(names)}
    launches = []
    for r in rows[header + 2:]:
        if len(r) < len(names):
            continue
        k = r[col["Kernel Name"]]
        rec = {"kernel": k}
        for key, metric in COLS.items():
            if metric in col:
                v = num(r[col[metric]])
                if v is None:
                    continue
                u = units[col[metric]]
                if key in ("dram_r", "dram_w"):
                    v = to_bytes(v, u)
                elif key == "time":
                    v = to_ns(v, u)
                rec[key] = v
        launches.append(rec)
    table = {}
    md = ["| kernel | launches | grid x block | regs | time/launch (us) | DRAM read+write / launch (MB) | DRAM % | SM % | issue % | warps % | FP64 pipe % |",
          "|---|---|---|---|---|---|---|---|---|---|---|"]
    for needle, stage in STAGE_OF:
        sel = [l for l in launches if needle in l["kernel"]]
        if not sel:
            continue
        # the launches over the whole point set: the largest grids of that kernel
        gmax = max(l.get("grid", 0) for l in sel)
        big = [l for l in sel if l.get("grid", 0) >= 0.5 * gmax]
        traffic = sum(l.get("dram_r", 0) + l.get("dram_w", 0) for l in big) / len(big)
        t_us = sum(l.get("time", 0) for l in big) / len(big) / 1e3

        def avg(key):
            vals = [l[key] for l in big if key in l]
            return sum(vals) / len(vals) if vals else float("nan")

        table[stage] = {"kernel": big[0]["kernel"][:120], "launches_captured": len(big), "dram_bytes_per_launch": traffic,
                        "time_us_under_ncu": t_us, "grid": gmax, "elements_per_launch": points, "source": f"{raw} (ncu --set full, serialised launches)"}
        md.append(f"| `{big[0]['kernel'][:70]}` | {len(big)} | {int(gmax)} x {int(avg('block'))} | {int(avg('regs'))} | {t_us:.1f} | "
                  f"{traffic / 1e6:.1f} | {avg('dram_pct'):.1f} | {avg('sm_pct'):.1f} | {avg('issue'):.1f} | {avg('warps'):.1f} | {avg('fp64'):.1f} |")
    json.dump(table, open(out_json, "w"), indent=1)
    open(out_md, "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
