"""Summarise the ncu launch list of tools/profile_train.py: the profiled step = everything after the warm-up forward."""
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
    name = re.sub(r"\(.*", "", r["Kernel Name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", ""))[:72]
    rows.append((name, us, r["Grid Size"]))
first_bwd = [i for i, r in enumerate(rows) if "act_bwd" in r[0] or "transpose" in r[0]][0]
F = first_bwd // 2          # two forwards (warm-up + step) precede the first backward kernel
step = rows[F:]
tot = sum(r[1] for r in step)
fwd = sum(r[1] for r in step[:F])
print("step: %d launches, %.1f us (forward %d launches %.1f us, backward + optimizer %.1f us)" % (len(step), tot, F, fwd, tot - fwd))
agg = defaultdict(lambda: [0.0, 0])
for n, us, _ in step:
    agg[n][0] += us
    agg[n][1] += 1
for k, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]:
    print("%9.1f us %5.1f%% x%-4d %s" % (us, 100 * us / tot, c, k))
if len(sys.argv) > 2:
    for n, us, g in sorted(step, key=lambda r: -r[1])[:int(sys.argv[2])]:
        print("%8.1f us grid %s %s" % (us, g, n))
