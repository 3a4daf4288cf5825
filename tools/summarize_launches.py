"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list: per kernel launches, average and
total duration, share of the captured window.   python tools/summarize_launches.py X.csv "comment line" > X_summary.csv"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        unit = r.get("Metric Unit", "ns")
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        rows.append((name, v))
    agg = OrderedDict()
    for n, v in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    for c in sys.argv[2:]:
        print("# " + c)
    print("kernel,launches,avg_us,total_us,share_pct")
    for n, (k, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n},{k},{t / k:.1f},{t:.1f},{100 * t / tot:.2f}")


if __name__ == "__main__":
    main()
