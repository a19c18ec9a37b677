#!/bin/bash
# Multi-GPU bench lines the way the driver launches them.  Usage (under gpurun --gpus N): tools/gpu_scale.sh N <tag> [strong batch]
N=$1; tag=${2:-x}; sb=${3:-0}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err; echo "weak n$N rc=$?"; tail -2 gpurun_out/bench_${tag}_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_${tag}_ref_n$N.json 2> gpurun_out/bench_${tag}_ref_n$N.err; echo "reference arm n$N rc=$? lines $(wc -l < gpurun_out/bench_${tag}_ref_n$N.json)"
if [ "$sb" != 0 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --strong --batch $sb > gpurun_out/bench_${tag}_strong_n$N.json 2> gpurun_out/bench_${tag}_strong_n$N.err; echo "strong n$N rc=$?"
fi
python - <<PY
import json
for n in ("n$N", "strong_n$N"):
    try:
        d = json.load(open("gpurun_out/bench_${tag}_%s.json" % n))
        print(n, "%.1f samples/s %.3f ms" % (d["value"], d["ms_per_step"]), d.get("scaling"), "e2e", d.get("e2e", {}).get("value"), "u8", d.get("e2e_u8_input", {}).get("value"),
              "other", {k: round(v["value"], 1) for k, v in d.get("other_configs", {}).items() if isinstance(v, dict)}, d["clocks"])
    except Exception as e:
        print(n, "failed", e)
PY
