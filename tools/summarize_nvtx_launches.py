"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`): total, per-kernel-name totals, the longest launches.
usage: summarize_nvtx_launches.py launches.csv [top_n_launches]"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    us = v / 1000 if r["Metric Unit"] in ("ns", "nsecond") else v
    name = r["Kernel Name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"at::native::|at::cuda::|cutlass::|cudnn::", "", name)[:80]
    rows.append((name, us, r["Grid Size"]))
tot = sum(r[1] for r in rows)
ours = sum(r[1] for r in rows if r[0].startswith("gwd_"))
print("%d launches, %.1f us serialised; gwd_* kernels: %d launches, %.1f us (%.1f%%)" % (len(rows), tot, sum(r[0].startswith("gwd_") for r in rows), ours, 100 * ours / tot))
agg = defaultdict(lambda: [0.0, 0])
for n, us, _ in rows:
    agg[n][0] += us
    agg[n][1] += 1
for k, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print("%9.1f us %5.1f%% x%-4d %s" % (us, 100 * us / tot, c, k))
if len(sys.argv) > 2:
    for n, us, g in sorted(rows, key=lambda r: -r[1])[:int(sys.argv[2])]:
        print("%8.1f us grid %s %s" % (us, g, n))
