"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v_us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    tot[name][0] += 1
    tot[name][1] += v_us
total = sum(v[1] for v in tot.values())
print(f"{'kernel':60s} {'launches':>9s} {'total us':>12s} {'avg us':>9s} {'share':>7s}")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {n:9d} {t:12.1f} {t / n:9.2f} {100 * t / total:6.1f}%")
print(f"{'TOTAL':60s} {sum(v[0] for v in tot.values()):9d} {total:12.1f}")
