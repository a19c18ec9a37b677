#!/bin/bash
# Standard GPU check: kernel tests, backbone tests, one bench line.  Usage: tools/gpu_check.sh <tag> [bench args]
tag=${1:-x}; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -rf -p no:cacheprovider --timeout=300 > gpurun_out/kernels_$tag.log 2>&1
echo "kernels rc=$?"; tail -3 gpurun_out/kernels_$tag.log | cut -c1-300
timeout 600 python -m pytest tests/test_backbone_gpu.py -m gpu -q --no-header -rf -s -p no:cacheprovider --timeout=600 > gpurun_out/backbone_$tag.log 2>&1
echo "backbone rc=$?"; grep -E "last_feat|passed|failed|Error" gpurun_out/backbone_$tag.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"; cat gpurun_out/bench_$tag.json; tail -5 gpurun_out/bench_$tag.err
