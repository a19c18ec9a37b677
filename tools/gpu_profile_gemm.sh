#!/bin/bash
# ncu --set full of each GEMM epilogue variant on one M (default 8640), warm L2 (--cache-control none).
# gemm_bench launches 13 GEMMs per variant in the order qkv, proj, swiglu, w3, linear.
tag=${1:-x}; M=${2:-8640}
mkdir -p gpurun_out
i=0
for name in qkv proj swiglu w3 linear; do
  skip=$((i * 13 + 6)); i=$((i + 1))
  timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k gemm_kernel -s $skip -c 1 -f \
      -o gpurun_out/prof_gemm_${name}_$tag python tools/gemm_bench.py --ms $M --tiles 0 --reps 1 --no-flush > gpurun_out/prof_gemm_${name}_$tag.log 2>&1
  echo "$name rc=$?"
done
ls -la gpurun_out/*.ncu-rep | tail
