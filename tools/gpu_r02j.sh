#!/bin/bash
# Round 2, batch j (8 GPUs): weak-scaling bench line + strong scaling.
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --overlap-gather --no-roofline > gpurun_out/bench_r02j_n8.json 2> gpurun_out/bench_r02j_n8.err; echo "bench n8 rc=$?"; tail -2 gpurun_out/bench_r02j_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 --strong --batch 4 > gpurun_out/bench_r02j_strong_n8.json 2> gpurun_out/bench_r02j_strong_n8.err; echo "strong n8 rc=$?"; tail -2 gpurun_out/bench_r02j_strong_n8.err
python - <<'PY'
import json
for n in ("n8", "strong_n8"):
    try:
        d = json.load(open("gpurun_out/bench_r02j_%s.json" % n))
        print(n, "%.1f samples/s %.3f ms" % (d["value"], d["ms_per_step"]), d.get("scaling"), "e2e", d.get("e2e", {}).get("value"), "overlap", d.get("overlap_gather", {}).get("value"),
              "other", {k: round(v["value"], 1) for k, v in d.get("other_configs", {}).items() if isinstance(v, dict)}, d["clocks"])
    except Exception as e:
        print(n, "failed", e)
PY
