#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
per kernel launches, summed time, share of the summed kernel time and mean DRAM bytes per launch.
usage: summarize_launches.py launches.csv out.md out.json"""
import collections
import csv
import hashlib
import json
import os
import sys


def source_sha():
    """fingerprint of the kernel sources the capture belongs to (bench.py compares it with the current sources and flags a
    stale `roofline.traffic`)"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = hashlib.sha256()
    for f in ("conv_umma.cu", "planes.cu", "elementwise.cu"):
        h.update(open(os.path.join(root, "bodyct-dram_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def main(src, out_md, out_json):
    lines = [l for l in open(src) if not l.startswith("==")]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        key = (r["ID"], r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
        else:
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        per.setdefault(key, {})[m] = v
    agg = collections.OrderedDict()
    for (_, name), m in per.items():
        short = name.split("(")[0].replace("void ", "")
        a = agg.setdefault(short, {"launches": 0, "ms": 0.0, "dram_read": 0.0, "dram_write": 0.0})
        a["launches"] += 1
        a["ms"] += m.get("gpu__time_duration.sum", 0.0)
        a["dram_read"] += m.get("dram__bytes_read.sum", 0.0)
        a["dram_write"] += m.get("dram__bytes_write.sum", 0.0)
    total = sum(a["ms"] for a in agg.values())
    rows = sorted(agg.items(), key=lambda kv: -kv[1]["ms"])
    with open(out_md, "w") as f:
        f.write(f"# ncu launch list summary of `{src.split('/')[-1]}` ({len(per)} launches, {total:.2f} ms summed kernel time)\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES with the in-bench CUDA-event shares.\n\n")
        f.write("| kernel | launches | ms | share | DRAM read MB/launch | DRAM write MB/launch |\n|---|---|---|---|---|---|\n")
        for k, a in rows[:40]:
            n = a["launches"]
            f.write(f"| `{k[:70]}` | {n} | {a['ms']:.3f} | {100 * a['ms'] / total:.1f} % | {a['dram_read'] / n / 1e6:.1f} | {a['dram_write'] / n / 1e6:.1f} |\n")
    out = {k: {"launches": a["launches"], "ms": a["ms"], "share": a["ms"] / total,
               "dram_bytes_per_launch": (a["dram_read"] + a["dram_write"]) / a["launches"]} for k, a in rows}
    json.dump({"source": src.split("/")[-1], "source_sha": source_sha(), "total_ms": total, "kernels": out}, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:4])
