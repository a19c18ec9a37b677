#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -p no:cacheprovider --timeout=300 -k "attention" 2>&1 | tail -6
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --timeout=900 -x 2>&1 | tail -4
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r02k.json 2> gpurun_out/bench_r02k.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r02k.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02k.json"))
r = d["roofline"]
print("%.1f samples/s %.3f ms e2e %.1f u8 %.1f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_u8_input"]["value"], d["gpu_launches_per_step"]), d["clocks"])
print("roofline achieved %.1f frac %.3f gemm ms %.3f | attention %.3f ms | token %.3f ms" % (r["achieved"], r["frac"], r["gemm_ms_per_step"], r["attention"]["ms_per_step"], r["token_kernels_ms_per_step"]))
print({k: (round(v["value"], 1), round(v["ms_per_step"], 3)) if isinstance(v, dict) else v for k, v in d["other_configs"].items()})
for k, v in sorted(r["eager_event_breakdown"]["kernels"].items(), key=lambda kv: -kv[1]["ms"]):
    if "attention" in k or "fill" in k: print("   ", k, v)
PY
