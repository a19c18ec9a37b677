#!/bin/bash
# Opt-in chained MLP launch: gated parity tests + kernel-level timing, for both publishing variants
# (TOC3D_CHAIN_SIG unset: every epilogue warp publishes its tile; =1: publisher warp).
mkdir -p gpurun_out
for sig in 0 1; do
  echo "== TOC3D_CHAIN_SIG=$sig" | tee -a gpurun_out/experimental.log
  TOC3D_CHAIN_SIG=$sig TOC3D_EXPERIMENTAL=1 timeout ${1:-600} python -m pytest tests/test_experimental_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee -a gpurun_out/experimental.log
  TOC3D_CHAIN_SIG=$sig timeout 120 python tools/chain_bench.py 2>&1 | tee gpurun_out/chain_bench_sig$sig.txt
done
