"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log into
profiles/<name>.json: per kernel (short name) launches, average duration and average DRAM bytes per launch.
bench.py reads profiles/r2_traffic.json to fill roofline.traffic.

  python scripts/ncu_traffic.py gpurun_out/r2_codec_metrics.csv profiles/r2_traffic.json "<command that was profiled>"
"""
import csv
import json
import re
import sys


def main(src, dst, source):
    rows = list(csv.reader(open(src, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    per = {}
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        per.setdefault(r[ii], {"name": r[ki]})[r[mi]] = v
    agg = {}
    for d in per.values():
        m = re.search(r"(?:rstnet::)?(\w+)(?:<[^>]*>)?\(", d["name"])
        short = m.group(1) if m else d["name"][:40]
        a = agg.setdefault(short, {"launches": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
        a["launches"] += 1
        a["ns"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    out = {"source": source, "kernels": {}}
    tot = sum(a["ns"] for a in agg.values()) or 1.0
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        n = a["launches"]
        out["kernels"][k] = {"launches": n, "avg_us": a["ns"] / n / 1e3, "share_of_time": a["ns"] / tot,
                             "dram_read_bytes_per_launch": a["rd"] / n, "dram_write_bytes_per_launch": a["wr"] / n,
                             "dram_bytes_per_launch": (a["rd"] + a["wr"]) / n}
    json.dump(out, open(dst, "w"), indent=1)
    for k, v in list(out["kernels"].items())[:12]:
        print(f"{k:40s} n={v['launches']:5d} avg {v['avg_us']:8.1f} us  share {v['share_of_time']:.3f}  dram/launch {v['dram_bytes_per_launch'] / 1e6:9.2f} MB")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
