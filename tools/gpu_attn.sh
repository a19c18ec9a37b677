#!/bin/bash
# Attention kernel check: parity tests + per-shape timing.  Usage: tools/gpu_attn.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -rf -p no:cacheprovider --timeout=120 -k "attention" > gpurun_out/attn_tests_$tag.log 2>&1
echo "attention tests rc=$?"; tail -6 gpurun_out/attn_tests_$tag.log | cut -c1-300
timeout 200 python tools/attn_bench.py > gpurun_out/attn_bench_$tag.txt 2>&1
echo "attn_bench rc=$?"; cat gpurun_out/attn_bench_$tag.txt
