"""profiles/ncu_traffic.json from an ncu launch list with DRAM byte counters.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/traffic.csv python tools/profile_step.py --config toc3d_fast --iters 2 --eager
    python tools/ncu_traffic.py gpurun_out/traffic.csv toc3d_fast "r01k"
Per kernel family of the LAST forward: launches, total us, DRAM MB read + written; the GEMM average per launch is
what bench.py reports as roofline.traffic.
"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

path, workload, tag = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(path) if not l.startswith("==")]
per = defaultdict(dict)
order = []
for r in csv.DictReader(lines):
    i = r["ID"]
    if i not in per:
        order.append(i)
        per[i]["name"] = r["Kernel Name"]
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        per[i]["us"] = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
    else:
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        per[i][m] = v * scale
starts = [n for n, i in enumerate(order) if "im2col_patch16" in per[i]["name"]]
sel = order[starts[-1]:] if starts else order
fam = defaultdict(lambda: [0, 0.0, 0.0])
for i in sel:
    d = per[i]
    name = re.sub(r"\(.*", "", d["name"]).replace("void ", "")
    name = re.sub(r"toc3d::(gemm::|attn_tc::|attn::)?", "", name)
    f = fam[name]
    f[0] += 1
    f[1] += d.get("us", 0.0)
    f[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
data = json.load(open(out_path)) if os.path.exists(out_path) else {}
g = [(k, v) for k, v in fam.items() if "gemm_kernel" in k]
gn = sum(v[0] for _, v in g)
data[workload] = {
    "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, capture %s, last forward of tools/profile_step.py --eager "
              "(cold caches between kernels, serialised)" % tag,
    "gemm_dram_bytes_per_launch": sum(v[2] for _, v in g) / max(1, gn),
    "gemm_launches": gn,
    "families": {k: {"launches": v[0], "us": round(v[1], 1), "dram_MB": round(v[2] / 1e6, 2)} for k, v in sorted(fam.items())},
}
json.dump(data, open(out_path, "w"), indent=1, sort_keys=True)
tot = sum(v[1] for v in fam.values())
print("%s: %d launches, %.1f us" % (workload, len(sel), tot))
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print("%8.1f us %5.1f%% %4d x  %8.2f MB DRAM/launch  %s" % (v[1], 100 * v[1] / tot, v[0], v[2] / v[0] / 1e6, k))
