#!/bin/bash
# Quick GPU check of the kernel tests only (first thing to run after a kernel change).  Usage: tools/gpu_quick.sh <tag> [pytest -k expr]
tag=${1:-q}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -x -rf -p no:cacheprovider --timeout=120 "$@" > gpurun_out/kernels_$tag.log 2>&1
echo "kernels rc=$?"; tail -25 gpurun_out/kernels_$tag.log | cut -c1-400
