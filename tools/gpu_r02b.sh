#!/bin/bash
# Round 2, batch b: full GPU suite (new full-width parity tests, a12 kernel), smoke, bench with the real reference legs,
# fold_norm2 breakdown, GEMM fixed-overhead sweep.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --timeout=900 -s > gpurun_out/gpu_suite_r02b.log 2>&1
echo "gpu suite rc=$?"; grep -E "passed|failed|error" gpurun_out/gpu_suite_r02b.log | tail -5 | cut -c1-300
grep -E "ISOLATED|score max-abs|last_feat|overlap|motion queries|first-frame" gpurun_out/gpu_suite_r02b.log | cut -c1-400 > gpurun_out/parity_table_r02b.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r02b.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r02b.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r02b.err
timeout 600 python bench.py --fold-norm2 --no-cpu-baseline --no-batch4 > gpurun_out/bench_r02b_fold.json 2> gpurun_out/bench_r02b_fold.err; echo "bench fold rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02b_ref.json 2> gpurun_out/bench_r02b_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_r02b_ref.json
timeout 300 python tools/gemm_bench.py --no-flush --ms 256,1024,2048,4096,4608,4662,6000 > gpurun_out/gemm_bench_r02b_msweep.txt 2>&1; cat gpurun_out/gemm_bench_r02b_msweep.txt
python - <<'PY'
import json
for n in ("r02b", "r02b_fold"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % n))
        print(n, "%.1f samples/s %.3f ms e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d.get("reference_gpu_eager"), d.get("cpu_baseline"))
        for k, v in sorted(d["roofline"]["breakdown"].items(), key=lambda kv: -kv[1]["ms"]):
            print("   ", k, v)
    except Exception as e:
        print(n, "failed:", e)
PY
