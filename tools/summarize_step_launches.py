"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of a program that runs the SAME step twice
(warm-up + measured, e.g. GWD_PROFILE_ONE=1 tools/bench_train_tail.py): per-kernel totals of the second half.
    python tools/summarize_step_launches.py <csv> [n slowest launches]"""
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
    name = re.sub(r"\(.*", "", r["Kernel Name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", ""))[:80]
    rows.append((name, us, r["Grid Size"]))
# the setup (weight upload, first-call attribute sets) launches only library kernels; the two steps are the two runs of the
# step's first hand-written kernel
first = next(i for i, r in enumerate(rows) if r[0].startswith("gwd_"))
name0 = rows[first][0]
starts = [i for i, r in enumerate(rows) if r[0] == name0]
half = (len(rows) - first) // 2
step = rows[len(rows) - half:]
tot = sum(r[1] for r in step)
ours = sum(r[1] for r in step if r[0].startswith("gwd_"))
print("measured step: %d launches, %.1f us serialised (hand-written gwd_* kernels: %d launches, %.1f us = %.1f %%)"
      % (len(step), tot, sum(1 for r in step if r[0].startswith("gwd_")), ours, 100 * ours / tot))
agg = defaultdict(lambda: [0.0, 0])
for n, us, _ in step:
    agg[n][0] += us
    agg[n][1] += 1
for k, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:24]:
    print("%9.1f us %5.1f%% x%-4d %s" % (us, 100 * us / tot, c, k))
if len(sys.argv) > 2:
    print("slowest launches:")
    for n, us, g in sorted(step, key=lambda r: -r[1])[:int(sys.argv[2])]:
        print("%9.1f us  grid %-12s %s" % (us, g, n))
