#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.

    python profiles/summarize_ncu.py gpurun_out/launches.csv [--skip N] > profiles/launches_rNN.md

Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md).
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*$", "", r["Kernel Name"]).strip()
        name = name.split("::")[-1]
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((name, val * scale))
    rows = rows[skip:]
    agg = OrderedDict()
    for name, us in rows:
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += us
        a[2] = max(a[2], us)
    total = sum(a[1] for a in agg.values()) or 1.0
    print(f"| kernel | launches | total us | share | mean us | max us |")
    print("|---|---:|---:|---:|---:|---:|")
    for name, (n, t, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {t:.1f} | {100 * t / total:.1f} % | {t / n:.1f} | {mx:.1f} |")
    print(f"\n{len(rows)} launches, {total:.1f} us total (ncu-serialised, cold cache)")


if __name__ == "__main__":
    main()
