#!/usr/bin/env python
"""Summarise an .ncu-rep (one `--set full` capture of tools/profile_plan.py) as a markdown table plus a JSON of the DRAM
traffic per kernel launch (what bench.py reports as roofline.traffic).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_cfg3.md profiles/ncu_traffic_cfg3.json
"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time", 1.0),
    ("dram__bytes_read.sum", "DRAM rd", 1.0),
    ("dram__bytes_write.sum", "DRAM wr", 1.0),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1.0),
    ("lts__t_sector_hit_rate.pct", "L2 hit %", 1.0),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %", 1.0),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", 1.0),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1.0),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst", 1.0),
    ("smsp__inst_executed.sum", "warp inst", 1.0),
    ("launch__registers_per_thread", "regs", 1.0),
    ("launch__grid_size", "grid", 1.0),
    ("launch__block_size", "block", 1.0),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb", 1.0),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier", 1.0),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb", 1.0),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch", 1.0),
]
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def main():
    rep, md_path, json_path = sys.argv[1], sys.argv[2], sys.argv[3]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    kernels = []
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].split("::")[-1]
        k = {"kernel": name}
        for m, label, _ in METRICS:
            if m not in col:
                continue
            try:
                v = float(r[col[m]].replace(",", ""))
            except ValueError:
                continue
            u = units[col[m]]
            if u in TO_BYTES:
                v *= TO_BYTES[u]
            if label == "time":
                v *= TO_US.get(u, 1.0)
            k[label] = v
        kernels.append(k)
    total = sum(k.get("time", 0.0) for k in kernels)
    lines = ["| kernel | time µs | share | DRAM read MB | DRAM write MB | DRAM % of peak | L2 hit % | occupancy % | issue % | "
             "threads/inst | warp inst (M) | regs | grid × block | top stalls (warps per issue) |", "|" + "---|" * 14]
    traffic = {}
    for k in kernels:
        stalls = sorted(((k.get(s, 0.0), s.replace("stall ", "")) for s in ("stall long_sb", "stall barrier", "stall short_sb",
                                                                            "stall branch")), reverse=True)[:2]
        lines.append(f"| {k['kernel']} | {k.get('time', 0):.1f} | {100 * k.get('time', 0) / total:.1f} % | "
                     f"{k.get('DRAM rd', 0) / 1e6:.1f} | {k.get('DRAM wr', 0) / 1e6:.1f} | {k.get('DRAM %', 0):.1f} | "
                     f"{k.get('L2 hit %', 0):.1f} | {k.get('occupancy %', 0):.1f} | {k.get('issue %', 0):.1f} | "
                     f"{k.get('thr/inst', 0):.1f} | {k.get('warp inst', 0) / 1e6:.2f} | {int(k.get('regs', 0))} | "
                     f"{int(k.get('grid', 0))} × {int(k.get('block', 0))} | "
                     + ", ".join(f"{n} {v:.1f}" for v, n in stalls) + " |")
        traffic[k["kernel"]] = {"dram_bytes_per_launch": k.get("DRAM rd", 0) + k.get("DRAM wr", 0),
                                "ncu_time_us": k.get("time", 0)}
    with open(md_path, "w") as f:
        f.write("\n".join(lines) + f"\n\nSum of kernel times under ncu (cold cache, serialised): {total:.1f} µs.\n")
    with open(json_path, "w") as f:
        json.dump(traffic, f, indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
