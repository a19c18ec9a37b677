#!/bin/bash
# ncu evidence for one config: launch list of the last forward + one full capture of the GEMM kernels.
# Usage: tools/gpu_profile.sh <tag> [config]
tag=${1:-x}; cfg=${2:-toc3d_fast}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv \
    python tools/profile_step.py --config $cfg --iters 2 --eager > gpurun_out/launches_$tag.log 2>&1
echo "launch list rc=$?"; python tools/summarize_launches.py gpurun_out/launches_$tag.csv > gpurun_out/launch_summary_$tag.txt; head -30 gpurun_out/launch_summary_$tag.txt
# full capture: 4th..7th GEMM launches of the 2nd forward's accelerated blocks are representative; skip the first forward
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 150 -c 8 -f -o gpurun_out/prof_gemm_$tag \
    python tools/profile_step.py --config $cfg --iters 2 --eager > gpurun_out/prof_gemm_$tag.log 2>&1
echo "gemm capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"window_attention|layernorm_rows|merge_fast|fast_update|window_topk|score_tokens" -s 60 -c 12 -f -o gpurun_out/prof_tok_$tag \
    python tools/profile_step.py --config $cfg --iters 2 --eager > gpurun_out/prof_tok_$tag.log 2>&1
echo "token capture rc=$?"
ls -la gpurun_out | tail -12
