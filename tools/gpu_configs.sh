#!/bin/bash
# One bench line per shipped config (headline metrics + marginal roofline, no CPU baseline).  Usage: tools/gpu_configs.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
for cfg in toc3d_fast toc3d_faster eva_vit_l toc3d_fast_1600 toc3d_faster_1600 eva_vit_l_1600; do
  timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs --no-batch4 > gpurun_out/bench_${cfg}_$tag.json 2> gpurun_out/bench_${cfg}_$tag.err
  echo "$cfg rc=$?"; python - <<PY
import json
d = json.load(open("gpurun_out/bench_${cfg}_$tag.json"))
r = d["roofline"]
print("  %s: %.1f samples/s  %.2f ms  e2e %.1f  GEMM %.0f TF/s (%.2f of sustained, %.3f ms)  attention %.3f ms  token kernels %.3f ms  clocks %s %s" % (
    "$cfg", d["value"], d["ms_per_step"], d["e2e"]["value"], r["achieved"], r["frac"], r["gemm_ms_per_step"], r["attention"]["ms_per_step"],
    r["token_kernels_ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
PY
done
