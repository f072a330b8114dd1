#!/usr/bin/env python
"""Summarise `ncu --set full` raw-page CSV exports (gpurun_out/*_raw.csv) into a markdown table and merge the per-kernel DRAM
traffic into profiles/ncu_traffic.json (read by bench.py for `roofline.traffic`).

    python tools/ncu_summarize.py profiles/r02_ncu_step_kernels.md gpurun_out/r02_step_kernels_raw.csv [more.csv ...]
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB rd"), ("dram__bytes_write.sum", "MB wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def short(name):
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"^void ", "", name)
    m = re.match(r"([A-Za-z0-9_]+(?:<[^>]*>)?)", name)
    return m.group(1) if m else name[:60]


def mb(v, unit):
    v = float(v)
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)


def main():
    out_md, files = sys.argv[1], sys.argv[2:]
    lines = ["| kernel | " + " | ".join(c[1] for c in COLS) + " |", "|---|" + "---|" * len(COLS)]
    traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for f in files:
        rows = list(csv.reader(open(f)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        idx = {h: i for i, h in enumerate(hdr)}
        seen = {}
        for r in data:
            k = short(r[idx["Kernel Name"]])
            seen[k] = seen.get(k, 0) + 1
            vals = []
            for c, _ in COLS:
                if c not in idx:
                    vals.append("-")
                    continue
                v, u = r[idx[c]], units[idx[c]]
                if c.startswith("dram__bytes"):
                    vals.append("%.1f" % mb(v, u))
                elif c == "gpu__time_duration.sum":
                    vals.append("%.1f" % (float(v) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)))
                else:
                    try:
                        vals.append("%.1f" % float(v) if "." in v else v)
                    except ValueError:
                        vals.append(v)
            lines.append("| `%s` #%d | " % (k, seen[k]) + " | ".join(vals) + " |")
            rd = mb(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
            wr = mb(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            key = k if seen[k] == 1 else "%s#%d" % (k, seen[k])
            traffic[key] = {"dram_bytes": int((rd + wr) * 1e6), "source": os.path.relpath(out_md, ROOT) + " <- " + os.path.basename(f)}
    with open(out_md, "a") as fo:
        fo.write("\n".join(lines) + "\n")
    with open(traffic_path, "w") as fo:
        json.dump(traffic, fo, indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
