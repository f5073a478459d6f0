#!/usr/bin/env python
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv <cmd>`) into a per-kernel
table: launches, total / mean device time and the SHARE of the profiled region (ncu times are cold-cache and serialised:
compare shares, not absolutes).   usage: python profiles/summarize_launches.py gpurun_out/launches.csv [skip_first_n]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "s": 1e9}.get(unit, 1.0)
        rows.append((name, ns))
    rows = rows[skip:]
    agg = OrderedDict()
    for name, ns in rows:
        short = re.sub(r"\(.*$", "", name)
        short = re.sub(r"^void ", "", short)
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values())
    print(f"launches profiled: {len(rows)} (skipped first {skip}); total device time {total / 1e6:.3f} ms\n")
    print("| kernel | launches | total ms | mean us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {ns / 1e6:.3f} | {ns / n / 1e3:.2f} | {100 * ns / total:.1f}% |")


if __name__ == "__main__":
    main()
