#!/bin/bash
# Round-2 opener: verify and time the opt-in kernels that were written at the end of round 1 without GPU time.
#   1. chained MLP launch (fuse_mlp), both publishing variants: gated parity tests + kernel-level timing
#   2. split-softmax attention (TOC3D_ATTN_SPLIT=1): the regular attention / backbone GPU tests run against it, then
#      per-shape timing next to the default kernel
#   3. whole-forward bench with each option
# Usage: tools/gpu_experiments.sh   (about 6 GPU-minutes)
mkdir -p gpurun_out
bash tools/gpu_chain_quick.sh 300
for split in 0 1; do
  echo "== TOC3D_ATTN_SPLIT=$split"
  TOC3D_ATTN_SPLIT=$split timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -rf -p no:cacheprovider --timeout=120 -k "attention" 2>&1 | tail -6 | tee gpurun_out/attn_tests_split$split.log
  TOC3D_ATTN_SPLIT=$split timeout 200 python tools/attn_bench.py 2>&1 | tee gpurun_out/attn_bench_split$split.txt
done
TOC3D_ATTN_SPLIT=1 timeout 600 python -m pytest tests/test_backbone_gpu.py -m gpu -x -q --no-header -p no:cacheprovider --timeout=600 2>&1 | tail -4 | tee gpurun_out/backbone_split1.log
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 400 python bench.py --fuse-mlp > gpurun_out/bench_fuse_mlp.json 2> gpurun_out/bench_fuse_mlp.err
timeout 400 python bench.py --fuse-block-tail > gpurun_out/bench_fuse_tail.json 2> gpurun_out/bench_fuse_tail.err
TOC3D_CHAIN_SIG=1 timeout 400 python bench.py --fuse-block-tail > gpurun_out/bench_fuse_tail_sig.json 2> gpurun_out/bench_fuse_tail_sig.err
TOC3D_CHAIN_SIG=1 timeout 400 python bench.py --fuse-mlp > gpurun_out/bench_fuse_mlp_sig.json 2> gpurun_out/bench_fuse_mlp_sig.err
TOC3D_ATTN_SPLIT=1 timeout 400 python bench.py > gpurun_out/bench_attn_split.json 2> gpurun_out/bench_attn_split.err
python - <<'PY'
import json
for n in ("default", "fuse_mlp", "fuse_mlp_sig", "fuse_tail", "fuse_tail_sig", "attn_split"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % n))
        b = d["roofline"]["breakdown"]
        print("%-14s %.1f samples/s  %.3f ms  attention %.3f ms  gemm share %.2f" % (
            n, d["value"], d["ms_per_step"], b.get("window_attention", {}).get("ms", 0), d["roofline"]["gemm_share_of_step"]))
    except Exception as e:
        print(n, "failed:", e)
PY
