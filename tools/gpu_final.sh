#!/bin/bash
# Round-end verification, exactly what the driver runs: the whole GPU suite, smoke(), the reference arm and the default
# bench line (both with the driver's own flags).  Usage: tools/gpu_final.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider --timeout=900 -s > gpurun_out/gpu_suite_$tag.log 2>&1
echo "gpu suite rc=$?"; grep -E "passed|failed|error" gpurun_out/gpu_suite_$tag.log | tail -3 | cut -c1-300
grep -E "ISOLATED|score max-abs|last_feat|overlap|motion queries|first-frame" gpurun_out/gpu_suite_$tag.log | cut -c1-400 > gpurun_out/parity_table_$tag.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$tag.log | cut -c1-300
t0=$SECONDS
timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err
echo "reference rc=$? wall $((SECONDS - t0)) s"; cut -c1-300 gpurun_out/bench_${tag}_reference.json
t0=$SECONDS
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc=$? wall $((SECONDS - t0)) s"; tail -2 gpurun_out/bench_$tag.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$tag.json"))
r = d["roofline"]
print("%.1f samples/s %.3f ms e2e %.1f u8 %.1f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_u8_input"]["value"], d["gpu_launches_per_step"]), d["clocks"])
print("roofline: GEMM %.1f TF/s frac %.3f (burst %.3f) %.3f ms | attention %.3f ms %.0f TF/s | hbm kernels %s | token kernels %.3f ms" % (
    r["achieved"], r["frac"], r["frac_of_burst_peak"], r["gemm_ms_per_step"], r["attention"]["ms_per_step"], r["attention"]["achieved_tflops"],
    {k: r["hbm_kernels"][k] for k in ("achieved", "frac", "ms_per_step", "achieved_events")}, r["token_kernels_ms_per_step"]))
print({k: (round(v["value"], 1), round(v["ms_per_step"], 3)) if isinstance(v, dict) else v for k, v in d["other_configs"].items()})
print("batch4", d.get("throughput_batch4")); print("gpu ref", d.get("reference_gpu_eager")); print("cpu", d.get("cpu_baseline"))
PY
