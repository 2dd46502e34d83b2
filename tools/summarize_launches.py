"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name (and per-launch top list)."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    name = r["Kernel Name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name)[:70]
    rows.append((name, us, r["Grid Size"], r["Block Size"]))
tot = sum(r[1] for r in rows)
agg = defaultdict(lambda: [0.0, 0])
for n, us, _, _ in rows:
    agg[n][0] += us
    agg[n][1] += 1
print("total %.1f us over %d launches" % (tot, len(rows)))
for n, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
    print("%8.1f us  %5.1f%%  x%-4d %s" % (us, 100 * us / tot, c, n))
if len(sys.argv) > 2:
    print("--- top single launches")
    for n, us, g, b in sorted(rows, key=lambda r: -r[1])[:int(sys.argv[2])]:
        print("%8.1f us  grid %s block %s  %s" % (us, g, b, n))
