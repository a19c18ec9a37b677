#!/bin/bash
# Round-end evidence on ONE box: whole GPU suite + smoke + the driver's two bench commands (tools/gpu_final.sh), then the
# ncu launch lists and --set full captures of the same forward (tools/gpu_profile.sh).  Usage: tools/gpu_evidence.sh <tag>
tag=${1:-x}
bash tools/gpu_final.sh $tag
bash tools/gpu_profile.sh $tag toc3d_fast
