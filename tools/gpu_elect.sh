#!/bin/bash
tag=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -rf -p no:cacheprovider --timeout=300 > gpurun_out/kernels_$tag.log 2>&1
echo "kernels rc=$?"; tail -3 gpurun_out/kernels_$tag.log | cut -c1-300
timeout 200 python tools/attn_bench.py > gpurun_out/attn_bench_$tag.txt 2>&1; cat gpurun_out/attn_bench_$tag.txt
timeout 300 python tools/gemm_bench.py --ms 8640,4662,3744,6000 --tiles 0,96,128,160,192,256 --reps 10 > gpurun_out/gemm_bench_$tag.txt 2>&1; cat gpurun_out/gemm_bench_$tag.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('throughput_batch4'))"; tail -3 gpurun_out/bench_$tag.err
