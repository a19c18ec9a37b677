#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -p no:cacheprovider --timeout=300 -k "attention" 2>&1 | tail -6
timeout 200 python tools/attn_bench.py 2>&1 | tee gpurun_out/attn_bench_r02g.txt
timeout 900 python -m pytest tests/test_backbone_gpu.py tests/test_parity_fullwidth_gpu.py -m gpu -q --no-header -p no:cacheprovider --timeout=900 -x 2>&1 | tail -4
timeout 900 python bench.py --no-cpu-baseline --no-other-configs --no-batch4 > gpurun_out/bench_r02g.json 2> gpurun_out/bench_r02g.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r02g.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02g.json"))
r = d["roofline"]
print("%.1f samples/s %.3f ms e2e %.1f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches_per_step"]), d["clocks"])
print("roofline achieved %.1f frac %.3f gemm ms %.3f | attention %.3f ms | token %.3f ms" % (r["achieved"], r["frac"], r["gemm_ms_per_step"], r["attention"]["ms_per_step"], r["token_kernels_ms_per_step"]))
for k, v in sorted(r["eager_event_breakdown"]["kernels"].items(), key=lambda kv: -kv[1]["ms"]):
    if "attention" in k: print("   ", k, v)
PY
