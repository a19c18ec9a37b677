#!/bin/bash
# Round 2, batch c (1 GPU): fused neck test, full default bench line (neck in the step, marginal roofline, other configs).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_neck.py tests/test_backbone_gpu.py -m gpu -q --no-header -p no:cacheprovider --timeout=600 2>&1 | tail -5
timeout 1500 python bench.py > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_r02c.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02c.json"))
r = d["roofline"]
print("%.1f samples/s %.3f ms e2e %.1f u8 %.1f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_u8_input"]["value"], d["gpu_launches_per_step"]))
print("clocks", d["clocks"])
print("roofline achieved %.1f frac %.3f (burst %.3f) gemm ms %.3f share %.2f | events: %.1f TF/s %.3f ms" % (
    r["achieved"], r["frac"], r["frac_of_burst_peak"], r["gemm_ms_per_step"], r["gemm_share_of_step"],
    r["eager_event_breakdown"]["achieved_gemm_tflops"], r["eager_event_breakdown"]["gemm_ms"]))
print("attention", r["attention"]); print("token kernels ms", r["token_kernels_ms_per_step"], "hbm", r["hbm_kernels"])
print("other", json.dumps(d.get("other_configs"), indent=1))
print("batch4", d.get("throughput_batch4")); print("gpu ref", d.get("reference_gpu_eager")); print("cpu", d.get("cpu_baseline"))
for k, v in sorted(r["eager_event_breakdown"]["kernels"].items(), key=lambda kv: -kv[1]["ms"]):
    print("   ", k, v)
PY
