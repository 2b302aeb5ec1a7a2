#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into the text committed under profiles/."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]


def main(rep, top=25):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    for row in r[2:]:
        print("kernel:", row[hdr.index("Kernel Name")][:80])
        for i, h in enumerate(hdr):
            if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                print(f"  {h:85s} {row[i]:>14s} {units[i]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    # first kernel only
    h = rows[1]
    isrc, isamp = h.index("Source"), h.index("# Samples")
    data = []
    for x in rows[2:]:
        if len(x) < len(h) or x[0] == "Kernel Name":
            break
        data.append(x)
    tot = sum(int(x[isamp]) for x in data) or 1
    print(f"\nsource page (first kernel): {len(data)} SASS instructions, {tot} stall samples; top {top} by samples")
    order = sorted(range(len(data)), key=lambda k: -int(data[k][isamp]))[:top]
    for k in sorted(order):
        x = data[k]
        prev = data[k - 1][isrc].strip()[:60] if k else ""
        print(f"  #{k:5d} {100.0 * int(x[isamp]) / tot:5.1f}%  {x[isrc].strip()[:70]:70s} | prev: {prev}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
