#!/bin/bash
# Attention check on one box: parity tests, in-step per-shape timing, clock64 timelines of the persistent kernel; with
# TOC3D_LIB=<other build of the library> a second timing pass for an A/B.  Usage: tools/gpu_attn_ab.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -x -rf -p no:cacheprovider --timeout=120 -k "attention" > gpurun_out/attn_tests_$tag.log 2>&1
echo "attention tests rc=$?"; tail -8 gpurun_out/attn_tests_$tag.log | cut -c1-300
echo "== product library"; timeout 200 python tools/attn_instep.py 2>&1 | tee gpurun_out/attn_instep_${tag}.txt
if [ -n "$TOC3D_AB_LIB" ]; then
  echo "== $TOC3D_AB_LIB"; TOC3D_LIB=$TOC3D_AB_LIB timeout 200 python tools/attn_instep.py 2>&1 | tee gpurun_out/attn_instep_${tag}_ab.txt
fi
if [ -f tools/probes/libtoc3d_trace.so ]; then
  for s in "48 129" "48 180" "48 256"; do
    timeout 100 python tools/attn_instep.py trace $s >> gpurun_out/attn_trace_$tag.txt 2>&1
  done
  cat gpurun_out/attn_trace_$tag.txt
fi
