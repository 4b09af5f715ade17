#!/usr/bin/env python
"""Summarises `ncu --set full` captures (exported on the GPU box with `ncu -i x.ncu-rep --page raw --csv`) into a table per capture and
regenerates profiles/ncu_traffic.json (read by bench.py for `roofline.traffic`).

    python tools/ncu_summary.py profiles/r2g_cfg2_cold_raw.csv profiles/r2g_cfg2_warm_raw.csv profiles/r2g_cfg4_warm_raw.csv \\
        profiles/r2g_pixels_raw.csv > profiles/r2g_ncu_summary.md
"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [("gpu__time_duration.sum", "us", 1e-3), ("dram__bytes_read.sum", "dram rd KB", None), ("dram__bytes_write.sum", "dram wr KB", None),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", 1.0),
        ("sm__inst_executed_pipe_tensor.sum", "tensor inst", 1.0),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1.0),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1.0),
        ("lts__t_sector_hit_rate.pct", "L2 hit %", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %", 1.0),
        ("launch__registers_per_thread", "regs", 1.0)]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m.get(unit, 1)


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("vt::", "")


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        d = {n: (v, u) for n, u, v in zip(names, units, r)}
        out.append(d)
    return out


def num(d, key):
    if key not in d or d[key][0] in ("", "n/a"):
        return None
    v, u = d[key]
    v = v.replace(",", "")
    if "byte" in u:
        return to_bytes(v, u)
    f = float(v)
    if u in ("usecond", "us"):
        f *= 1e3
    elif u in ("msecond", "ms"):
        f *= 1e6
    elif u in ("second", "s"):
        f *= 1e9
    return f


def table(path):
    rows = load(path)
    print(f"\n### {os.path.basename(path)} ({len(rows)} launches)\n")
    print("| kernel | grid | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    agg = OrderedDict()
    for d in rows:
        k = short(d["Kernel Name"][0])
        grid = d.get("Grid Size", ("", ""))[0]
        cells = []
        for key, label, scale in COLS:
            v = num(d, key)
            if v is None:
                cells.append("-")
            elif "KB" in label:
                cells.append(f"{v / 1e3:.0f}")
            elif label == "us":
                cells.append(f"{v * 1e-3:.2f}")
            else:
                cells.append(f"{v:.1f}" if v < 1000 else f"{v:.0f}")
        print(f"| `{k}` | {grid} | " + " | ".join(cells) + " |")
        a = agg.setdefault(k, {"n": 0, "dram": 0.0, "us": 0.0})
        a["n"] += 1
        a["dram"] += (num(d, "dram__bytes_read.sum") or 0) + (num(d, "dram__bytes_write.sum") or 0)
        a["us"] += (num(d, "gpu__time_duration.sum") or 0) * 1e-3
    return agg


def main():
    paths = sys.argv[1:]
    print("# ncu `--set full --clock-control none` captures of the final kernels (raw page, one row per profiled launch)\n")
    print("cold = default cache control (L2 flushed before every replay pass); warm = `--cache-control none`. "
          "Durations under ncu are serialised and (cold) cache-cold: compare shares, not absolutes.")
    traffic = {"source": ", ".join(os.path.relpath(p, ROOT) for p in paths) + " (ncu --set full --clock-control none; cold captures flush L2 before every replay)"}
    for p in paths:
        agg = table(p)
        tag = "cold" if "cold" in p or "pixels" in p else "warm"
        for k, a in agg.items():
            base = k.split("<")[0]
            e = traffic.setdefault(base, {})
            e[f"dram_bytes_per_launch_{tag}"] = int(a["dram"] / a["n"])
            e[f"launches_captured_{tag}"] = a["n"]
            if tag == "cold":
                e["dram_bytes_per_launch"] = int(a["dram"] / a["n"])
    # the pixel workload converts 32 frames per launch (more than L2): per-frame figure next to the algorithmic 9,331,200 B / 1080p frame
    if "nv12_to_rgb_vec4_kernel" in traffic:
        t = traffic["nv12_to_rgb_vec4_kernel"]
        t["frames_per_launch"] = 32
        t["dram_bytes_per_frame"] = t["dram_bytes_per_launch"] // 32
        t["algorithmic_bytes_per_frame"] = 1920 * 1080 * 9 // 2
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
