#!/bin/bash
# Round-end verification: the whole GPU suite, smoke(), the default bench line.  Usage: tools/gpu_final.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider --timeout=600 > gpurun_out/gpu_suite_$tag.log 2>&1
echo "gpu suite rc=$?"; tail -4 gpurun_out/gpu_suite_$tag.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/smoke_$tag.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e'], d.get('e2e_u8_input'), d.get('throughput_batch4'), d['roofline']['frac'], d['cpu_baseline'])"; tail -3 gpurun_out/bench_$tag.err
