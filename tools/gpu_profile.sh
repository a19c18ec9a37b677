#!/bin/bash
# ncu evidence for one config.  Usage: tools/gpu_profile.sh <tag> [config]
#   1. launch list (gpu__time_duration + DRAM bytes) of the last eager forward  -> gpurun_out/launches_<tag>.csv
#   2. the same list for the CUDA-graph replay path                              -> gpurun_out/launches_graph_<tag>.csv
#   3. one `--set full` capture of a few GEMM / attention / token kernels        -> gpurun_out/prof_*_<tag>.ncu-rep
tag=${1:-x}; cfg=${2:-toc3d_fast}
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/launches_$tag.csv \
    python tools/profile_step.py --config $cfg --iters 2 --eager > gpurun_out/launches_$tag.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_graph_$tag.csv \
    python tools/profile_step.py --config $cfg --iters 3 > gpurun_out/launches_graph_$tag.log 2>&1
echo "graph launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel" -s 160 -c 8 -f -o gpurun_out/prof_gemm_$tag \
    python tools/profile_step.py --config $cfg --iters 2 --eager > gpurun_out/prof_gemm_$tag.log 2>&1
echo "gemm capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"window_attention|layernorm_rows|ln_gather_merge|fast_update|window_topk|score_tokens|topk_split" -s 70 -c 14 -f -o gpurun_out/prof_tok_$tag \
    python tools/profile_step.py --config $cfg --iters 2 --eager > gpurun_out/prof_tok_$tag.log 2>&1
echo "token capture rc=$?"
ls -la gpurun_out | grep $tag
