#!/bin/bash
# One `--set full` capture (with source counters) of the ping-pong attention kernel.  Usage: tools/gpu_profile_attn.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"window_attention_pp" -s 3 -c 1 -f -o gpurun_out/prof_attn_$tag \
    python tools/attn_bench.py > gpurun_out/prof_attn_$tag.log 2>&1
echo "attn capture rc=$?"; tail -12 gpurun_out/prof_attn_$tag.log
ls -la gpurun_out | grep prof_attn_$tag
