"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share."""
import collections
import csv
import re
import sys


def main(path, skip_regex=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = row["Kernel Name"]
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Unit"] == "us":
            v *= 1e3
        elif row["Metric Unit"] == "ms":
            v *= 1e6
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("kabc::", "")
        if skip_regex and re.search(skip_regex, short):
            continue
        agg[short][0] += 1
        agg[short][1] += v
    tot = sum(v for _, v in agg.values())
    print(f"{'kernel':64s} {'n':>5s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:64]:64s} {n:5d} {v / 1e3:10.1f} {v / n / 1e3:9.2f} {v / tot * 100:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
