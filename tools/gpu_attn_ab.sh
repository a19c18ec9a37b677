#!/bin/bash
# Attention A/B on one box: parity tests of the default kernels, in-step per-shape timing of the pipelined split-softmax
# kernel vs the ping-pong kernel (TOC3D_ATTN_PP=1), timelines of the default.  Usage: tools/gpu_attn_ab.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -x -rf -p no:cacheprovider --timeout=120 -k "attention" > gpurun_out/attn_tests_$tag.log 2>&1
echo "attention tests rc=$?"; tail -8 gpurun_out/attn_tests_$tag.log | cut -c1-300
echo "== default (pipelined split softmax)"; timeout 200 python tools/attn_instep.py 2>&1 | tee gpurun_out/attn_instep_${tag}_ps.txt
echo "== TOC3D_ATTN_PP=1 (ping-pong)"; TOC3D_ATTN_PP=1 timeout 200 python tools/attn_instep.py 2>&1 | tee gpurun_out/attn_instep_${tag}_pp.txt
if [ -f tools/probes/libtoc3d_trace.so ]; then
  for s in "48 129" "48 180" "48 256"; do
    timeout 100 python tools/attn_instep.py trace $s >> gpurun_out/attn_trace_$tag.txt 2>&1
  done
  cat gpurun_out/attn_trace_$tag.txt
fi
