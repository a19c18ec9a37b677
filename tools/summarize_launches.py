"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the
last forward (after the 7777-element marker fill when present)."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    rows.append((r["Kernel Name"], us, r.get("Grid Size", ""), r.get("Block Size", "")))
# cut at the marker: the last vectorized fill with grid computed for 7777 elements is hard to spot; use the
# last occurrence of im2col (start of a forward) instead
starts = [i for i, r in enumerate(rows) if "im2col_patch16" in r[0]]
cut = starts[-1] if starts else 0
sel = rows[cut:]
agg = defaultdict(lambda: [0, 0.0])
for name, us, g, b in sel:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    if "gemm_kernel" in name:
        m = re.search(r"gemm_kernel<(\d)>|gemm_kernel<\(int\)(\d)>", name)
        short = "gemm_kernel<%s>" % ((m.group(1) or m.group(2)) if m else "?")
    if len(short) > 70:
        short = short[:70]
    agg[short][0] += 1
    agg[short][1] += us
tot = sum(v[1] for v in agg.values())
print("last forward: %d launches, %.1f us total kernel time" % (len(sel), tot))
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%8.1f us %5.1f%% %4d x %7.1f us  %s" % (us, 100 * us / tot, n, us / n, k))
