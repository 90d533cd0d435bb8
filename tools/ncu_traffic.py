#!/usr/bin/env python
"""Per-kernel DRAM traffic and headline metrics from an `ncu --set full` capture (raw page as CSV).

    ncu -i gpurun_out/full_x.ncu-rep --page raw --csv > gpurun_out/full_x_raw.csv
    python tools/ncu_traffic.py gpurun_out/full_x_raw.csv [kernel-substring valid_images dtype out.json]

Prints one markdown row per captured launch; with the optional arguments writes the mean
dram__bytes_read.sum + dram__bytes_write.sum of the matching launches to out.json
(bench.py reads profiles/roofline_traffic.json for `roofline.traffic`)."""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
        "msecond": 1e3, "nsecond": 1e-3, "%": 1.0}
COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB_rd"), ("dram__bytes_write.sum", "MB_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return float("nan")


def main(argv):
    with open(argv[1], newline="") as f:
        rows = list(csv.reader(l for l in f if not l.startswith("==")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    kn = ix["Kernel Name"]
    print("| kernel | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|" + "---:|" * len(COLS))
    out = []
    for r in data:
        vals = {}
        for name, short in COLS:
            if name not in ix:
                vals[short] = float("nan")
                continue
            v = num(r[ix[name]])
            u = units[ix[name]]
            if short.startswith("MB"):
                v = v * UNIT.get(u, 1.0) / 1e6
            elif short == "us":
                v = v * UNIT.get(u, 1.0)
            vals[short] = v
        name = r[kn].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        out.append((name, vals))
        print(f"| `{name}` | " + " | ".join(f"{vals[c[1]]:.2f}" if vals[c[1]] < 1e6 else f"{vals[c[1]]:.0f}" for c in COLS) + " |")
    if len(argv) >= 6:
        sel = [v for n, v in out if argv[2] in n]
        if sel:
            b = sum((v["MB_rd"] + v["MB_wr"]) * 1e6 for v in sel) / len(sel)
            json.dump({"kernel": argv[2], "valid_images": int(argv[3]), "dtype": argv[4], "bytes_per_launch": b,
                       "launches": len(sel), "mean_us_under_ncu": sum(v["us"] for v in sel) / len(sel),
                       "source": argv[1]}, open(argv[5], "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv)
