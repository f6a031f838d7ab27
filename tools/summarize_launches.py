"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((name, v * scale, r.get("Grid Size", ""), r.get("Block Size", "")))
tot = sum(t for _, t, _, _ in rows)
agg = defaultdict(lambda: [0.0, 0])
for n, t, _, _ in rows:
    agg[n][0] += t
    agg[n][1] += 1
print("total %.1f us over %d launches" % (tot, len(rows)))
for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%8.1f us %5.1f%% %5d x  %s" % (t, 100 * t / tot, c, n))
if len(sys.argv) > 2:
    print("\ntop launches:")
    for n, t, g, b in sorted(rows, key=lambda r: -r[1])[: int(sys.argv[2])]:
        print("%8.1f us  grid %s block %s  %s" % (t, g, b, n))
