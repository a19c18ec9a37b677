#!/bin/bash
# Attention diagnosis on one box: issue-rate probe, in-step per-shape timing, clock64 timelines.  Usage: tools/gpu_attn_diag.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm --format=csv,noheader
[ -x tools/probes/softmax_probe ] && timeout 120 tools/probes/softmax_probe > gpurun_out/softmax_probe_$tag.txt 2>&1; cat gpurun_out/softmax_probe_$tag.txt
timeout 200 python tools/attn_instep.py > gpurun_out/attn_instep_$tag.txt 2>&1; cat gpurun_out/attn_instep_$tag.txt
if [ -f tools/probes/libtoc3d_trace.so ]; then
  for s in "48 129" "48 180" "48 256" "18 201"; do
    timeout 100 python tools/attn_instep.py trace $s >> gpurun_out/attn_trace_$tag.txt 2>&1
  done
  cat gpurun_out/attn_trace_$tag.txt
fi
