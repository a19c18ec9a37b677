"""Top stalled SASS instructions per kernel of an .ncu-rep (needs `ncu` on PATH; reads the source page).

    python tools/ncu_hot.py gpurun_out/prof.ncu-rep [top_n] [kernel-substring]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
filt = sys.argv[3] if len(sys.argv) > 3 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in txt.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = [line]
        blocks.append(cur)
    elif cur is not None:
        cur.append(line)
for bi, b in enumerate(blocks):
    name = next(csv.reader([b[0]]))[1]
    if filt and filt not in name:
        continue
    rd = list(csv.DictReader(io.StringIO("\n".join(b[1:]))))
    tot = sum(int(r["# Samples"] or 0) for r in rd)
    print("=== launch %d: %s  (samples %d, %d SASS instr)" % (bi, name[:90], tot, len(rd)))
    stall_cols = [c for c in rd[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
    agg = {c: sum(int(r[c] or 0) for r in rd) for c in stall_cols}
    print("   stall mix: " + ", ".join("%s %.0f%%" % (c[6:], 100.0 * v / max(1, tot)) for c, v in
                                       sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
    order = sorted(range(len(rd)), key=lambda i: -int(rd[i]["# Samples"] or 0))[:top]
    for i in sorted(order):
        r = rd[i]
        s = int(r["# Samples"] or 0)
        dom = max(stall_cols, key=lambda c: int(r[c] or 0))
        print("   %5d %5.1f%%  %-14s %s" % (s, 100.0 * s / max(1, tot), dom[6:], r["Source"].strip()[:100]))
