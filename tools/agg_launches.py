"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"total {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:25]:
    print(f"{t / 1e3:10.3f} ms {100 * t / tot:5.1f}%  n={n:5d} avg={t / n:9.1f} us  {k}")
