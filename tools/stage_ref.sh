#!/bin/bash
# Stage the reference's own backbone / neck sources into the git-ignored baseline/_ref/ so that they travel to the
# GPU box with the gpurun snapshot (the box has no /root/reference).  Nothing under baseline/_ref/ is ever committed
# or imported by the product: it is what `bench.py --impl reference`, the `cpu_baseline` leg and the
# `reference_gpu_eager` leg time (the UNMODIFIED reference modules, loaded by tests/golden/ref_import.py).
#   usage: tools/stage_ref.sh [/root/reference]
set -e
SRC=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")/.." && pwd)
DST=$HERE/baseline/_ref
P=projects/mmdet3d_plugin/models
if [ ! -d "$SRC/$P" ]; then echo "stage_ref: $SRC/$P not present, nothing staged"; exit 0; fi
rm -rf "$DST"
for f in backbones/toc3d_eva_vit.py backbones/toc3d_utils.py backbones/eva_vit.py backbones/eva_utils.py \
         utils/misc.py utils/positional_encoding.py utils/gpu_timer.py necks/cp_fpn.py; do
  mkdir -p "$DST/$P/$(dirname $f)"
  cp "$SRC/$P/$f" "$DST/$P/$f"
done
( cd "$SRC" && git rev-parse HEAD 2>/dev/null || cat .SUBMODULES.json 2>/dev/null | head -5 ) > "$DST/STAGED_FROM" 2>/dev/null || true
echo "staged $(find "$DST" -name '*.py' | wc -l) reference files into $DST"
