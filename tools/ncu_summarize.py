#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.

    python tools/ncu_summarize.py gpurun_out/launches.csv > profiles/launches_rNN.md

ncu times are cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md)."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    n = 0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        name = re.sub(r"\(.*$", "", name)
        name = name.replace("cxrm::(anonymous namespace)::", "").replace("cxrm::", "")
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "s": 1e9, "second": 1e9}.get(unit, 1.0)
        agg[name][0] += 1
        agg[name][1] += ns
        total += ns
        n += 1
    print(f"# ncu launch list summary: {path}\n")
    print(f"{n} launches, {total / 1e6:.2f} ms of kernel time (cold-cache, serialised)\n")
    print("| kernel | launches | total ms | mean us | share |")
    print("|---|---:|---:|---:|---:|")
    for name, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {c} | {ns / 1e6:.3f} | {ns / c / 1e3:.2f} | {100 * ns / total:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
