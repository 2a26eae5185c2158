"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel."""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        name = r["Kernel Name"]
        rows.append((name, v_us, r.get("Grid Size", ""), r.get("Block Size", "")))
    tot = sum(v for _, v, _, _ in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, v, _, _ in rows:
        key = n.split("(")[0]
        agg[key][0] += 1
        agg[key][1] += v
    print(f"# {path}: {len(rows)} launches, {tot / 1e3:.3f} ms total (serialised, cold-cache: compare shares)")
    print(f"{'kernel':70s} {'launches':>8s} {'ms':>9s} {'share':>7s}")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {c:8d} {v / 1e3:9.3f} {100 * v / tot:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
