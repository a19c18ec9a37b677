#!/bin/bash
# ncu --set full of the chained launches (2 and 3 problems) at one M (default 8640), both publishing variants,
# warm L2 like tools/gpu_profile_gemm.sh.  chain_bench launches per M: 3 + 20 two-launch pairs, then 23 chain2
# (planned order), 23 chain2 (sequential order), 3 + 20 three-launch triples, 23 chain3.
# Usage: tools/gpu_profile_chain.sh <tag> [M]      -> gpurun_out/prof_chain{2,3}_sig{0,1}_<tag>.ncu-rep
tag=${1:-x}; M=${2:-8640}
mkdir -p gpurun_out
for sig in 0 1; do
  for np in 2 3; do
    # kernel name filter: gemm_chain_kernel<SIG, NPROB>; skip the warm-up launches of that instantiation
    TOC3D_CHAIN_SIG=$sig timeout 600 ncu --set full --clock-control none --cache-control none --import-source on \
        -k regex:"gemm_chain_kernel" -s $([ $np = 2 ] && echo 5 || echo 50) -c 1 -f -o gpurun_out/prof_chain${np}_sig${sig}_$tag \
        python tools/chain_bench.py $M > gpurun_out/prof_chain${np}_sig${sig}_$tag.log 2>&1
    echo "chain$np sig=$sig rc=$?"
  done
done
ls -la gpurun_out/*chain*.ncu-rep | tail
