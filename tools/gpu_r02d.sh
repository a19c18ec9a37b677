#!/bin/bash
# Round 2, batch d (1 GPU): wide-tile GEMM parity + timing, full suite, default bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -p no:cacheprovider --timeout=600 -k "gemm" 2>&1 | tail -6
timeout 300 python tools/gemm_bench.py --no-flush --ms 6000,4662,4578,3744 --tiles 0,256,320,352,384 > gpurun_out/gemm_bench_r02d_wide.txt 2>&1; grep -E "proj|w3|clocks" gpurun_out/gemm_bench_r02d_wide.txt
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --timeout=900 -x 2>&1 | tail -4
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r02d.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02d.json"))
r = d["roofline"]
print("%.1f samples/s %.3f ms e2e %.1f u8 %.1f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_u8_input"]["value"], d["gpu_launches_per_step"]))
print("clocks", d["clocks"])
print("roofline achieved %.1f frac %.3f gemm ms %.3f | attention %.3f ms | token %.3f ms" % (r["achieved"], r["frac"], r["gemm_ms_per_step"], r["attention"]["ms_per_step"], r["token_kernels_ms_per_step"]))
print({k: (v["value"], v["ms_per_step"]) if isinstance(v, dict) else v for k, v in d["other_configs"].items()})
for k, v in sorted(r["eager_event_breakdown"]["kernels"].items(), key=lambda kv: -kv[1]["ms"])[:8]:
    print("   ", k, v)
PY
