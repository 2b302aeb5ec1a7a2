#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/*_launches_summary.txt)."""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = [r for r in csv.reader(open(path)) if r and r[0] != "" and not r[0].startswith("==")]
    hdr = rows[0]
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
        tot[r[ik]] += v
        cnt[r[ik]] += 1
    all_us = sum(tot.values())
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"{k[:72]:72s} n={cnt[k]:4d} total={tot[k]:10.1f}us avg={tot[k] / cnt[k]:8.2f}us share={100 * tot[k] / all_us:5.1f}%")
    print(f"all kernels: {sum(cnt.values())} launches, {all_us:.1f} us")


if __name__ == "__main__":
    main(sys.argv[1])
