"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: share, total us, launches.
Usage: python tools/launch_summary.py gpurun_out/launches.csv > profiles/<round>_launches_summary_table.md"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void\s+", "", name)
        name = re.sub(r"^at::", "", name)
        rows.append((name, us))
    tot = sum(u for _, u in rows)
    agg = defaultdict(lambda: [0.0, 0])
    for n, u in rows:
        agg[n][0] += u
        agg[n][1] += 1
    print(f"sum = {tot / 1e3:.1f} ms over {len(rows)} launches\n")
    print("| share | total us | launches | kernel |\n|---:|---:|---:|---|")
    for n, (u, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"| {100 * u / tot:.1f}% | {u:.1f} | {c} | `{n}` |")


if __name__ == "__main__":
    main(sys.argv[1])
