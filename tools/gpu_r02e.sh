#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_bench.py --no-flush --ms 6000,4662,4578 --tiles 0,256,288,320,352,384,512 > gpurun_out/gemm_bench_r02e_wide.txt 2>&1; grep -E "resid|clocks" gpurun_out/gemm_bench_r02e_wide.txt
timeout 300 python tools/gemm_bench.py --ms 6000,4662 --tiles 0,256,352 > gpurun_out/gemm_bench_r02e_wide_flush.txt 2>&1; grep -E "resid|clocks" gpurun_out/gemm_bench_r02e_wide_flush.txt
