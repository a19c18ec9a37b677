#!/bin/bash
# One bench line per shipped config (no CPU baseline), plus batch sweeps of the headline config.
tag=${1:-x}
mkdir -p gpurun_out
for cfg in toc3d_fast toc3d_faster eva_vit_l toc3d_fast_1600 toc3d_faster_1600 eva_vit_l_1600; do
  timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${cfg}_$tag.json 2> gpurun_out/bench_${cfg}_$tag.err
  echo "$cfg rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench_${cfg}_$tag.json"))
r=d["roofline"]
print("  %s: %.1f samples/s  %.2f ms  e2e %.1f  gemm %.0f TF/s  whole-step %.0f TF/s" % ("$cfg", d["value"], d["ms_per_step"], d["e2e"]["value"], r["achieved"], r["whole_step_tflops"]))
print("  ", {k:(v["ms"], v.get("tflops")) for k,v in r["breakdown"].items()})
PY
done
for b in 2 4; do
  timeout 600 python bench.py --config toc3d_fast --batch $b --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_toc3d_fast_b${b}_$tag.json 2>> gpurun_out/bench_b_$tag.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_toc3d_fast_b${b}_$tag.json'))
print('  batch $b: %.1f samples/s %.2f ms gemm %.0f TF/s' % (d['value'], d['ms_per_step'], d['roofline']['achieved']))"
done
