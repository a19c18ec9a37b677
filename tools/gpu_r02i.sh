#!/bin/bash
# Round 2, batch i (2 GPUs): ShardedBackbone over NCCL, weak-scaling bench, strong-scaling bench.
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_shard_nccl_gpu.py -m gpu -q --no-header -p no:cacheprovider --timeout=600 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --overlap-gather > gpurun_out/bench_r02i_n2.json 2> gpurun_out/bench_r02i_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/bench_r02i_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --strong --batch 2 > gpurun_out/bench_r02i_strong_n2.json 2> gpurun_out/bench_r02i_strong_n2.err; echo "strong n2 rc=$?"; tail -3 gpurun_out/bench_r02i_strong_n2.err
timeout 600 python bench.py --strong --batch 2 --steps 20 > gpurun_out/bench_r02i_strong_n1.json 2> gpurun_out/bench_r02i_strong_n1.err; echo "strong n1 rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-roofline --no-batch4 --no-other-configs > gpurun_out/bench_r02i_n1.json 2> gpurun_out/bench_r02i_n1.err; echo "n1 rc=$?"
python - <<'PY'
import json
for n in ("n1", "n2", "strong_n1", "strong_n2"):
    try:
        d = json.load(open("gpurun_out/bench_r02i_%s.json" % n))
        print(n, "%.1f samples/s %.3f ms" % (d["value"], d["ms_per_step"]), d.get("scaling"), "e2e", d.get("e2e", {}).get("value"), "overlap", d.get("overlap_gather", {}).get("value"),
              "other", {k: round(v["value"], 1) for k, v in d.get("other_configs", {}).items() if isinstance(v, dict)}, d["clocks"])
    except Exception as e:
        print(n, "failed", e)
PY
